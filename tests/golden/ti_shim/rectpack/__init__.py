"""Placeholder for `rectpack` (absent offline); texture packing is outside the `pt` golden scenes."""


def newPacker(*a, **k):
    raise NotImplementedError("rectpack is not available in this container")


class PackerBBF:      # only named in a type annotation (parsers/texture_packing.py:99)
    pass
