"""Does a k_logic launch of one path pool overlap with the k_trace launch of another?  Two Renderer handles on the same scene, each
driven by its own host thread on its own stream, against one handle with the same total pool (GPU session r02c).

    python tools/overlap_probe.py [workload] [spp]        (knobs through the environment: ADAPT_TRACE_BLOCKS_PER_SM, ADAPT_POOL)
"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ADAPT_QUIET", "1")
import torch                                                            # noqa: E402
import bench                                                            # noqa: E402
from adapt_b200.renderer.vanilla_renderer import Renderer              # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "bunny90k"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
e, a, o, c = bench.load_workload(workload)


def run(n_handles, pool):
    rs = [Renderer(e, a, o, c, seed=k, pool_size=pool) for k in range(n_handles)]
    for r in rs:
        r.render_batch(4); r.synchronize(); r.stats(reset=True)

    def work(r):
        r.render_batch(spp); r.synchronize()
    torch.cuda.synchronize()
    t0 = time.time()
    th = [threading.Thread(target=work, args=(r,)) for r in rs]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.time() - t0
    rays = sum(r.stats()["rays_closest"] for r in rs)
    st = rs[0].stats()
    for r in rs:
        r.close()
    return rays / dt / 1e6, dt, st["ms_logic"], st["ms_closest"] + st["ms_shadow"]


for n, pool in ((1, 1 << 22), (2, 1 << 21), (2, 1 << 22), (1, 1 << 21)):
    v, dt, ml, mt = run(n, pool)
    print(f"{workload} handles={n} pool={pool >> 20}Mi/handle blocks/SM={os.environ.get('ADAPT_TRACE_BLOCKS_PER_SM', 'max')}: {v:8.1f} Mrays/s aggregate, wall {dt * 1e3:7.1f} ms, "
          f"handle 0: logic {ml:7.1f} ms trace {mt:7.1f} ms", flush=True)
