"""Build time of the acceleration structure: host binned-SAH builder vs device linear BVH, first build (inside adapt_create) and
rebuilds through adapt_update_geometry (device build: CUDA-event time of kernels + sort; host build: wall time of build + layout;
"call" = wall time of the whole update call incl. host packing of the primitive tables and the H2D copies).
    python tools/bvh_build_bench.py [bunny90k orb500k car290k]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ADAPT_QUIET", "1")

from adapt_b200.parsers.xml_parser import scene_parsing            # noqa: E402
from adapt_b200.renderer.vanilla_renderer import Renderer          # noqa: E402
from adapt_b200.scenes import DEFAULT_ROOT, ensure_big_meshes      # noqa: E402

for name in (sys.argv[1:] or ["bunny90k", "orb500k"]):
    ensure_big_meshes(DEFAULT_ROOT, (name,))
    e, a, o, c = scene_parsing(os.path.join(DEFAULT_ROOT, "cbox"), name + ".xml")
    c["film"]["width"] = c["film"]["height"] = 16
    for builder in (os.environ.get("BUILDERS", "sah,lbvh,sah_device").split(",")):
        r = Renderer(e, a, o, c, bvh_builder=builder)
        first = r.bvh_export(arrays=False)
        re_ms, call_ms = [], []
        for _ in range(4):
            t0 = time.perf_counter()
            r.update_geometry(a["primitives"], a["n_g"], a["n_s"])
            call_ms.append((time.perf_counter() - t0) * 1e3)
            re_ms.append(r.bvh_export(arrays=False)["build_ms"])
        print(json.dumps({"scene": name, "n_prims": first["n_prims"], "builder": builder, "n_nodes": first["n_nodes"], "depth": first["depth"],
                          "first_build_ms": round(first["build_ms"], 3), "rebuild_ms_min": round(min(re_ms), 3),
                          "update_call_ms_min": round(min(call_ms), 3), "host_threads": os.cpu_count()}))
        r.close()
