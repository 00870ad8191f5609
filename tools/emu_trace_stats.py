"""Lane occupancy of the vote-scheduled traversal (pt_trace.cuh: trace_stream_vote), measured WITHOUT a GPU by running k_trace under
the SIMT emulator of tests/dev_host: of the 32 lanes x node steps a scheduling round offers, how many do a node step; how many lanes
take part when the leaf code runs.  ncu reports the same quantity on the B200 as "threads per instruction" of the node / leaf code
(16.8 / 13.2 on bunny90k with one node step per round, session r01f), so scheduling policies can be compared on the CPU before they
cost GPU time.  The numbers say nothing about latency or issue rate.

    python tools/emu_trace_stats.py [scene xml size spp]        # default: cbox bunny90k.xml 48 1
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("ADAPT_QUIET", "1")
import numpy as np                                                    # noqa: E402
from adapt_b200._lib import pack_scene                                # noqa: E402
from adapt_b200.scenes import DEFAULT_ROOT, ensure_big_meshes          # noqa: E402
from conftest import load_scene                                       # noqa: E402
import dev_host                                                       # noqa: E402

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
scene, name, size, spp = (argv + ["cbox", "bunny90k.xml", "48", "1"][len(argv):])[:4]
size, spp = int(size), int(spp)
if name in ("bunny90k.xml", "orb500k.xml", "car290k.xml"):
    ensure_big_meshes(DEFAULT_ROOT, (name[:-4],))
# a build with the node steps as a run-time value (the shipped default unrolls four at compile time)
lib_path = os.path.join(os.path.dirname(dev_host.WF_LIB), "libwavefront_host_rt.so")
deps = dev_host.WF_DEPS + dev_host.DEPS[3:]
if not os.path.exists(lib_path) or any(os.path.getmtime(d) > os.path.getmtime(lib_path) for d in deps):
    subprocess.check_call(["g++", "-O2", "-std=c++20", "-w", "-fPIC", "-ffp-contract=fast", "-march=x86-64-v3", "-DTRACE_NODE_STEPS_CT=0",
                           "-I" + dev_host.CUDA_INC, "-shared", "-o", lib_path, dev_host.WF_SRC, os.path.join(ROOT, "adapt_b200", "csrc", "bvh_build.cpp")])
e, a, o, c = load_scene(DEFAULT_ROOT, scene, name, size, size)
ps = pack_scene(e, a, o, c, seed=1)
print(f"{scene}/{name} {size}x{size} x {spp} spp, {a['primitives'].shape[0]} primitives, pool 2048 slots, 2 trace blocks")
print(f"{'refill':>6} {'leaf_t':>6} {'steps':>5} | {'node lanes/32':>13} {'leaf lanes/32':>13} {'prims/leaf lane':>15} {'rounds/ray':>10} {'node steps/ray':>14}")
ref = None
WIDE8 = "--cw8" in sys.argv                  # ADAPT_TRACE_MODE=3: the compressed 8-wide tree collapsed from the binary one
if WIDE8:
    print("tree: compressed 8-wide (80-byte nodes, cw8_step)")
for refill, leaf_t, steps in [(16, 12, 1), (16, 12, 2), (16, 12, 4), (16, 8, 4), (16, 8, 8), (16, 4, 4), (16, 16, 4), (8, 8, 4), (24, 8, 4)]:
    os.environ.update(ADAPT_REFILL=str(refill), ADAPT_LEAF_T=str(leaf_t), ADAPT_NODE_STEPS=str(steps), ADAPT_TRACE_MODE="3" if WIDE8 else "1")
    L = C.CDLL(lib_path)
    L.wavefront_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    acc = np.zeros((size, size, 3), np.float32); st = np.zeros(6, np.uint64); ts = np.zeros(8, np.uint64)
    L.wavefront_trace_stats(ts.ctypes.data_as(C.POINTER(C.c_uint64)))
    rc = L.wavefront_render(C.addressof(ps.desc), spp, 2048, 2, 0, acc.ctypes.data_as(C.POINTER(C.c_float)), st.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == 0
    L.wavefront_trace_stats(ts.ctypes.data_as(C.POINTER(C.c_uint64)))
    rounds, slots, nls, lr, ll, llp, _, rays = (float(x) for x in ts)
    if ref is None:
        ref = acc.copy()
    assert np.allclose(acc, ref, rtol=1e-5, atol=1e-6), "the scheduling policy must not change the image"
    print(f"{refill:6d} {leaf_t:6d} {steps:5d} | {32 * nls / slots:13.2f} {ll / max(lr, 1):13.2f} {llp / max(ll, 1):15.2f} {rounds / rays:10.2f} {nls / rays:14.2f}")
