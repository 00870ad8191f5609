"""adapt_render is asynchronous (a launch thread per handle): how long does the call take on the host, and what do small batches cost?

    python tools/async_probe.py [workload]
Prints the host time of adapt_render(32), and the throughput of 256 spp enqueued as 8 x 32, 64 x 4 and 256 x 1 spp calls back to back
(no synchronisation in between; one adapt_sync at the end), against one call of 256 spp."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ADAPT_QUIET", "1")
import bench                                                            # noqa: E402
from adapt_b200.renderer.vanilla_renderer import Renderer              # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "bunny90k"
e, a, o, c = bench.load_workload(workload)
r = Renderer(e, a, o, c, seed=0)
r.render_batch(8); r.synchronize()
t0 = time.perf_counter(); r.render_batch(32); t1 = time.perf_counter(); r.synchronize(); t2 = time.perf_counter()
print(f"{workload}: adapt_render(32) returned after {(t1 - t0) * 1e3:.3f} ms on the host; the samples were done {(t2 - t0) * 1e3:.1f} ms after the call")
total = 256
for per_call in (256, 32, 4, 1):
    r.stats(reset=True)
    t0 = time.perf_counter()
    for _ in range(total // per_call):
        r.render_batch(per_call)
    t_enq = time.perf_counter() - t0
    r.synchronize()
    dt = time.perf_counter() - t0
    st = r.stats()
    print(f"  {total // per_call:4d} x adapt_render({per_call:3d}): enqueued in {t_enq * 1e3:8.2f} ms, finished in {dt * 1e3:8.1f} ms, "
          f"{st['rays_closest'] / dt / 1e6:8.1f} Mrays/s, {st['iterations']} iterations")
