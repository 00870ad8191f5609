"""Stand-in for `pywavefront` (absent offline), written from its documented behaviour and independent of
adapt_b200/parsers/obj_loader.py: `Wavefront(path, collect_faces=True).materials` maps material names to objects
with `.vertex_format` (e.g. "T2F_N3F_V3F") and `.vertices`, a flat float list with one interleaved record per
face corner in file order (n-gons fan-triangulated); a default material is synthesised when no `usemtl` is seen.
TEST INFRASTRUCTURE ONLY (tests/golden/make_reference_golden.py)."""


class Material:
    def __init__(self, name):
        self.name = name
        self.vertex_format = ""
        self.vertices = []


class Wavefront:
    def __init__(self, path, collect_faces=False, **_):
        self.materials = {}
        self.vertices = []
        v, vt, vn = [], [], []
        mat = None
        for line in open(path):
            p = line.split()
            if not p or p[0].startswith("#"):
                continue
            if p[0] == "v":
                v.append(tuple(float(x) for x in p[1:4]))
                self.vertices.append(v[-1])
            elif p[0] == "vt":
                vt.append(tuple(float(x) for x in p[1:3]))
            elif p[0] == "vn":
                vn.append(tuple(float(x) for x in p[1:4]))
            elif p[0] == "usemtl":
                mat = self.materials.setdefault(p[1], Material(p[1]))
            elif p[0] == "f":
                if mat is None:
                    mat = self.materials.setdefault("default0", Material("default0"))
                corners = [c.split("/") for c in p[1:]]
                has_t = len(corners[0]) > 1 and corners[0][1] != ""
                has_n = len(corners[0]) > 2 and corners[0][2] != ""
                if not mat.vertex_format:
                    mat.vertex_format = ("T2F_" if has_t else "") + ("N3F_" if has_n else "") + "V3F"

                def emit(c):
                    def idx(s, n):
                        k = int(s)
                        return k - 1 if k > 0 else n + k
                    if has_t:
                        mat.vertices.extend(vt[idx(c[1], len(vt))])
                    if has_n:
                        mat.vertices.extend(vn[idx(c[2], len(vn))])
                    mat.vertices.extend(v[idx(c[0], len(v))])
                for k in range(1, len(corners) - 1):          # triangle fan
                    emit(corners[0]); emit(corners[k]); emit(corners[k + 1])
