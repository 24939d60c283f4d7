"""ORACLE — test infrastructure, NOT product code.

CPU restatement of the reference's per-detection LiDAR cropping:
point_cloud/preprocess.py:66-81 (per-box loop, empty box -> one zero point, reflectivity dropped at :92-95)
with the membership predicate of point_cloud/geometry.py:96-114 written as vectorised numpy float32.
Pinned by tests/golden/crop_*.npz, produced by the UNMODIFIED reference functions (numba) in
oracle/make_goldens.py.
"""
import numpy as np

from mmmot_b200.lidar_crop import box_planes


def points_in_boxes(points, boxes_lidar):
    """bool [P][n]: sign = x*nx + y*ny + z*nz + d evaluated left to right (geometry.py:108-113) in the precision numba
    gives it: float64 when the boxes are float64 (the real pipeline, box_np_ops.py:584-589), float32 otherwise."""
    b = np.asarray(boxes_lidar)
    dt = np.float64 if b.dtype == np.float64 else np.float32
    pl = box_planes(b.astype(dt))                                        # [n][6][4]
    p = np.asarray(points, dtype=np.float32).astype(dt)
    x, y, z = p[:, None, None, 0], p[:, None, None, 1], p[:, None, None, 2]
    s = (x * pl[None, :, :, 0] + y * pl[None, :, :, 1]) + z * pl[None, :, :, 2]
    s = s + pl[None, :, :, 3]
    return (s < 0).all(-1)


def crop_points_ref(points, boxes_lidar, without_reflectivity=True):
    points = np.asarray(points, dtype=np.float32)
    mask = points_in_boxes(points, boxes_lidar)
    out, split = [], [0]
    for b in range(mask.shape[1]):
        sel = points[mask[:, b]]
        if sel.shape[0] == 0:
            sel = np.zeros((1, points.shape[1]), dtype=np.float32)
        split.append(split[-1] + sel.shape[0])
        out.append(sel)
    out = np.concatenate(out, axis=0)
    if without_reflectivity:
        out = out[:, :3]
    return out, np.asarray(split, dtype=np.int64)
