"""Stand-in for `rectpack` (absent offline): just the calls parsers/texture_packing.py:84-92 makes -- newPacker(rotation=False),
add_bin, add_rect(w, h, rid), pack(), packer[0] -> rectangles with x / y / width / height / rid.  A plain shelf packer; where a
rectangle lands does not change what `Texture.query` returns (it never leaves its own rectangle, bxdf/texture.py:117-118).
TEST INFRASTRUCTURE ONLY (tests/golden/make_reference_golden.py)."""


class PackerBBF:      # only named in a type annotation (parsers/texture_packing.py:99)
    pass


class _Rect:
    def __init__(self, x, y, width, height, rid):
        self.x, self.y, self.width, self.height, self.rid = x, y, width, height, rid


class _Packer:
    def __init__(self):
        self.bin = None
        self.rects = []
        self.bins = [[]]

    def add_bin(self, w, h):
        self.bin = (w, h)

    def add_rect(self, w, h, rid=None):
        self.rects.append((w, h, rid))

    def pack(self):
        bw, bh = self.bin
        x = y = shelf = 0
        placed = []
        for w, h, rid in sorted(self.rects, key=lambda r: (-r[1], -r[0])):
            if x + w > bw:
                y += shelf
                x = shelf = 0
            if w > bw or y + h > bh:
                continue
            placed.append(_Rect(x, y, w, h, rid))
            x += w
            shelf = max(shelf, h)
        self.bins = [placed]

    def __getitem__(self, k):
        return self.bins[k]


def newPacker(rotation=False, **_):
    return _Packer()
