"""TEST INFRASTRUCTURE: CPU harness around the device BVH builder's per-element steps (adapt_b200/csrc/bvh_lbvh.h) plus a
validator / reference traversal for trees in the traversal layout.  Only tests/ and __graft_entry__.build() touch this."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
SRC = os.path.join(_HERE, "lbvh_host.cpp")
DEPS = [SRC, os.path.join(_ROOT, "adapt_b200", "csrc", "bvh_lbvh.h"), os.path.join(_ROOT, "adapt_b200", "csrc", "bvh_build.cpp"),
        os.path.join(_ROOT, "adapt_b200", "csrc", "bvh_build.h")]
LIB = os.path.join(_HERE, "_build", "liblbvh_host.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # -ffp-contract=off: nvcc cannot contract anything in these steps either (no a*b+c shapes), keep g++ from inventing one
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared", "-o", LIB, SRC, DEPS[2]])
    return LIB


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _norm(prims, sph, prim_obj, obj_class):
    prims = np.ascontiguousarray(prims, np.float32).reshape(-1, 9)
    n = prims.shape[0]
    sph = np.zeros(n, np.uint8) if sph is None else np.ascontiguousarray(sph, np.uint8)
    prim_obj = np.zeros(n, np.int32) if prim_obj is None else np.ascontiguousarray(prim_obj, np.int32)
    obj_class = np.zeros(int(prim_obj.max()) + 1, np.uint8) if obj_class is None else np.ascontiguousarray(obj_class, np.uint8)
    return prims, n, sph, prim_obj, obj_class


def build_tree(prims, sph=None, prim_obj=None, obj_class=None, max_leaf=4, builder="lbvh", order_seed=0, traverse_cost=1.0, eight=False):
    """-> dict(nodes (n_nodes,16), prims (n,12), depth, root_box) from the emulated device builders ("lbvh": linear BVH, "sah_device":
    level-synchronous binned SAH, its per-position steps run in an order permuted by order_seed) or the library's host SAH builder ("sah")."""
    L = load()
    prims, n, sph, prim_obj, obj_class = _norm(prims, sph, prim_obj, obj_class)
    nodes = np.zeros((max(1, n - 1), 16), np.float32)
    recs = np.zeros((n, 12), np.float32)
    nn, dp = C.c_int(), C.c_int()
    root = np.zeros(6, np.float32)
    nodes8 = None
    if builder == "sah_device":
        lv, nn8, dp8 = C.c_int(), C.c_int(), C.c_int()
        nodes8 = np.zeros((max(1, n - 1), 20), np.uint32) if eight else None
        rc = L.lbvh_host_build_sah(_p(prims), _p(sph), _p(prim_obj), _p(obj_class), n, max_leaf, _p(nodes), _p(recs), C.byref(nn), C.byref(dp), _p(root),
                                   C.c_uint(order_seed), C.byref(lv), C.c_float(traverse_cost), _p(nodes8) if eight else None, C.byref(nn8), C.byref(dp8))
        if eight:
            nodes8 = nodes8[:nn8.value].copy()
    elif builder == "lbvh":
        rc = L.lbvh_host_build(_p(prims), _p(sph), _p(prim_obj), _p(obj_class), n, max_leaf, _p(nodes), _p(recs), C.byref(nn), C.byref(dp), _p(root))
    else:
        rc = L.sah_host_build(_p(prims), _p(sph), _p(prim_obj), _p(obj_class), n, max_leaf, _p(nodes), _p(recs), C.byref(nn), C.byref(dp))
    if rc != 0:
        raise RuntimeError(f"{builder} host build failed: {rc}")
    out = dict(nodes=nodes[:nn.value].copy(), prims=recs, depth=dp.value, root_box=root)
    if nodes8 is not None:
        out["nodes8"] = nodes8; out["depth8"] = dp8.value
    return out


def validate(nodes, recs, prims, sph=None):
    """0 when the tree is structurally sound (see lbvh_validate), else a negative code; also returns the depth found."""
    L = load()
    prims, n, sph, _, _ = _norm(prims, sph, None, None)
    nodes = np.ascontiguousarray(nodes, np.float32); recs = np.ascontiguousarray(recs, np.float32)
    d = C.c_int()
    rc = L.lbvh_validate(_p(nodes), nodes.shape[0], _p(recs), n, _p(prims), _p(sph), C.byref(d))
    return rc, d.value


def trace_check(nodes, recs, rays_o, rays_d):
    """Closest hits through the tree and by brute force -> (prim, t, bf_prim, bf_t, nodes_per_ray)."""
    L = load()
    nodes = np.ascontiguousarray(nodes, np.float32); recs = np.ascontiguousarray(recs, np.float32)
    ro = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3); rd = np.ascontiguousarray(rays_d, np.float32).reshape(-1, 3)
    nr = ro.shape[0]
    op = np.zeros(nr, np.int32); ot = np.zeros(nr, np.float32); bp = np.zeros(nr, np.int32); bt = np.zeros(nr, np.float32)
    v = L.lbvh_trace_check(_p(nodes), _p(recs), recs.shape[0], _p(ro), _p(rd), nr, _p(op), _p(ot), _p(bp), _p(bt))
    return op, ot, bp, bt, v / 1000.0


def cw8_trace_check(nodes8, recs, prims, sph, rays_o, rays_d):
    """Closest hits through a compressed 8-wide tree with an independent decoder (+ structural checks) -> (rc, t, prim)."""
    L = load()
    prims, n, sph, _, _ = _norm(prims, sph, None, None)
    nodes8 = np.ascontiguousarray(nodes8, np.uint32); recs = np.ascontiguousarray(recs, np.float32)
    ro = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3); rd = np.ascontiguousarray(rays_d, np.float32).reshape(-1, 3)
    nr = ro.shape[0]
    ot = np.zeros(nr, np.float32); op = np.zeros(nr, np.int32)
    rc = L.cw8_trace_check(_p(nodes8), nodes8.shape[0], _p(recs), n, _p(prims), _p(sph), _p(ro), _p(rd), nr, _p(ot), _p(op))
    return rc, ot, op


def last_prims_tested():
    """Leaf primitives tested by the most recent trace_check call (all rays): the second tree-quality figure next to nodes per ray."""
    L = load()
    L.lbvh_last_prims_tested.restype = C.c_longlong
    return int(L.lbvh_last_prims_tested())


def refit_tree(tree, prims, order_seed=0):
    """Refit `tree` (dict from build_tree) IN PLACE over the new vertices `prims` (n,3,3) with the device refit's per-element steps."""
    L = load()
    prims = np.ascontiguousarray(prims, np.float32).reshape(-1, 9)
    rc = L.lbvh_host_refit(_p(tree["nodes"]), tree["nodes"].shape[0], _p(tree["prims"]), prims.shape[0], _p(prims), C.c_uint(order_seed))
    if rc != 0:
        raise RuntimeError(f"host refit failed: {rc}")
    return tree
