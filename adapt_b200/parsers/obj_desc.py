"""Per-object descriptor + AABB (reference parsers/obj_desc.py:9-58)."""
import numpy as np

__all__ = ["get_aabb", "ObjDescriptor"]


def get_aabb(meshes: np.ndarray, _type: int = 0) -> np.ndarray:
    """(2,3) float32 AABB.  Planar meshes (extent <= 1e-3 on an axis) are padded by 2e-2 on that axis;
    spheres use center -/+ radius (obj_desc.py:9-25)."""
    if _type == 0:
        mini = meshes.min(axis=1).min(axis=0)
        maxi = meshes.max(axis=1).max(axis=0)
        large_diff = np.abs(maxi - mini) > 1e-3
        for i in range(3):
            if not large_diff[i]:
                mini[i] -= 2e-2
                maxi[i] += 2e-2
    else:
        mini = meshes[0, 0] - meshes[0, 1]
        maxi = meshes[0, 0] + meshes[0, 1]
    return np.float32((mini, maxi))


class ObjDescriptor:
    """Same fields as the reference class: tri_num, meshes, uv_coords, normals, vns, R, t, bsdf,
    texture_group, aabb, emitter_ref_id, type (0 mesh / 1 sphere)."""

    def __init__(self, meshes, normals, bsdf, vert_normal=None, uv_coords=None, texture_group=None,
                 R=None, t=None, emit_id=-1, _type=0):
        self.tri_num = meshes.shape[0]
        self.meshes = meshes
        self.uv_coords = uv_coords
        self.normals = normals
        self.vns = vert_normal
        self.R = R
        self.t = t
        self.bsdf = bsdf
        self.texture_group = texture_group
        self.aabb = get_aabb(meshes, _type)
        self.emitter_ref_id = emit_id
        self.type = _type

    def __repr__(self):
        centroid = (self.aabb[0] + self.aabb[1]) / 2
        if self.type == 0:
            return (f"<wavefront with {self.meshes.shape[0]} triangles centered at {centroid}.\n"
                    f" Transformed: {self.R is not None or self.t is not None}>")
        if self.type == 1:
            return f"<sphere centered at {self.meshes[0, 0]} with radius {self.meshes[0, 1]}>"
        raise NotImplementedError("Other object types are not supported yet")
