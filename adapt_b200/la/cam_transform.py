"""Host-side camera helpers (reference la/cam_transform.py:20-49).  The device-side frame helpers
(rotation_between, delocalize_rotate, convert_to_raw; :51-105) live in csrc/pt_device.cuh."""
import numpy as np
from scipy.spatial.transform import Rotation as Rot

__all__ = ["fov2focal", "np_rotation_between"]


def fov2focal(fov: float, img_size):
    fov = fov / 180.0 * np.pi
    return 0.5 * img_size / np.tan(0.5 * fov)


def np_rotation_between(fixed: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Rotation fixed -> target with the roll (first 'zxy' Euler angle) removed; +-I when (anti)parallel."""
    axis = np.cross(fixed, target)
    dot = np.dot(fixed, target)
    if abs(dot) > 1.0 - 1e-5:
        return np.sign(dot) * np.eye(3, dtype=np.float32)
    axis = axis / np.linalg.norm(axis)
    axis = axis * np.arccos(dot)
    euler_vec = Rot.from_rotvec(axis).as_euler("zxy")
    euler_vec[0] = 0
    return Rot.from_euler("zxy", euler_vec).as_matrix()
