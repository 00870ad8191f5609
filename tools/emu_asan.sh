#!/bin/bash
# The library's kernels under the SIMT emulator (tests/dev_host) built with AddressSanitizer: out-of-bounds accesses to the pool, the
# queues, the class lists or the scene tables -- silent corruption on the GPU -- become reports here.  CPU only; a minute or two.
#   bash tools/emu_asan.sh
set -e
cd "$(dirname "$0")/.."
OUT=/tmp/libwavefront_host_asan.so
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++20 -w -fPIC -ffp-contract=fast -march=x86-64-v3 -I${CUDA_INC:-/usr/local/cuda/include} \
    -shared -o $OUT tests/dev_host/wavefront_host.cpp adapt_b200/csrc/bvh_build.cpp
cat > /tmp/emu_asan.py <<'PY'
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
os.environ.setdefault("ADAPT_QUIET", "1")
import numpy as np
from conftest import load_scene
from adapt_b200.scenes import DEFAULT_ROOT, ensure_small_scenes
from adapt_b200._lib import pack_scene
L = C.CDLL(sys.argv[1])
L.wavefront_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
root = ensure_small_scenes(DEFAULT_ROOT)
for scene, name, integ, pool in [("cbox", "cbox.xml", "pt", 256), ("csphere", "balls-mono.xml", "pt", 256), ("test", "allbxdf.xml", "pt", 512),
                                 ("test", "textured.xml", "pt", 256), ("test", "media.xml", "vpt", 256), ("test", "allbxdf.xml", "vpt", 256)]:
    e, a, o, c = load_scene(root, scene, name, 12, 11)
    ps = pack_scene(e, a, o, c, seed=1, integrator=integ)
    acc = np.zeros((12, 11, 3), np.float32); st = np.zeros(5, np.uint64)
    rc = L.wavefront_render(C.addressof(ps.desc), 2, pool, 2, 0, acc.ctypes.data_as(C.POINTER(C.c_float)), st.ctypes.data_as(C.POINTER(C.c_uint64)))
    print(name, integ, "rc", rc, "paths", int(st[0]), "finite", bool(np.isfinite(acc).all()))
    assert rc == 0 and int(st[0]) == 12 * 11 * 2
print("no AddressSanitizer report")
PY
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python /tmp/emu_asan.py $OUT 2>&1 | grep -v "doesn't fully support makecontext"
