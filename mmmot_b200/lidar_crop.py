"""Per-detection LiDAR cropping on the GPU (SURVEY.md §8f N1 — the step right before the hot path).

Mirrors the 3-D branch of reference point_cloud/preprocess.py:66-81 (``read_and_prep_points``): every
detection's rotated 3-D box (camera frame: location, dimensions l-h-w, rotation_y) is moved to the
LiDAR frame (box_np_ops.py:613-618), the scene points inside it are kept in scene order
(preprocess.py:39-42 -> box_np_ops.py:688-699 -> geometry.py:96-114), an empty box contributes one
all-zero point, and the result is the packed ``points`` + ``points_split`` pair that
``TrackingNet.forward`` / ``forward_batch`` take.

The box -> plane-equation preparation (n boxes x 6 planes, microseconds of host work) is done with the
same numpy operations, in the same dtype, as the reference so that the plane coefficients are identical
(float64 for boxes coming from ``box_camera_to_lidar`` — the reference's real pipeline, where ``np.ones``
promotes them — float32 for float32 boxes); the
O(P x n) membership test and the stable compaction run in libmmmot_sm100a.so (csrc/lidar_crop.cu).
There is no CPU path for the membership test.
"""
import ctypes

import numpy as np
import torch

from . import _lib

# corner order of a unit box (x0y0z0, x0y0z1, x0y1z1, x0y1z0, x1y0z0, x1y0z1, x1y1z1, x1y1z0) and the six faces
# listed so that their normals (cross of consecutive edges) point inward (box_np_ops.py:161-178, 702-720)
_CORNER_ORDER = [0, 1, 3, 2, 4, 5, 7, 6]
_FACES = [[0, 1, 2, 3], [7, 6, 5, 4], [0, 3, 7, 4], [1, 5, 6, 2], [0, 4, 5, 1], [3, 2, 6, 7]]


def box_camera_to_lidar(boxes_cam, r_rect, velo2cam):
    """[x, y, z, l, h, w, ry] camera frame -> [x, y, z, w, l, h, ry] LiDAR frame
    (reference box_np_ops.py:584-589, 613-618)."""
    xyz = boxes_cam[:, 0:3]
    l, h, w = boxes_cam[:, 3:4], boxes_cam[:, 4:5], boxes_cam[:, 5:6]
    r = boxes_cam[:, 6:7]
    hom = np.concatenate([xyz, np.ones(list(xyz.shape[:-1]) + [1])], axis=-1)
    xyz_lidar = (hom @ np.linalg.inv((r_rect @ velo2cam).T))[..., :3]
    return np.concatenate([xyz_lidar, w, l, h, r], axis=1)


def box_planes(boxes_lidar):
    """[n][7] LiDAR-frame boxes (x, y, z, w, l, h, yaw; origin (0.5, 0.5, 0), rotation about z) ->
    [n][6][4] inward plane equations (nx, ny, nz, d) in the boxes' dtype (float32 or float64).  Same numpy
    operations, in the same order, as reference box_np_ops.py:147-178 (corners), :236-254 (rotation), :312-337,
    :702-720 (faces) and geometry.py:84-93 (plane equations), so the results are identical."""
    rb = np.asarray(boxes_lidar)
    centers, dims, angles = rb[:, :3], rb[:, 3:6], rb[:, 6]
    unit = np.stack(np.unravel_index(np.arange(8), [2] * 3), axis=1).astype(dims.dtype)[_CORNER_ORDER]
    unit = unit - np.array([0.5, 0.5, 0], dtype=dims.dtype)
    corners = dims.reshape([-1, 1, 3]) * unit.reshape([1, 8, 3])
    rot_sin, rot_cos = np.sin(angles), np.cos(angles)
    ones, zeros = np.ones_like(rot_cos), np.zeros_like(rot_cos)
    rot_t = np.stack([[rot_cos, -rot_sin, zeros], [rot_sin, rot_cos, zeros], [zeros, zeros, ones]])
    corners = np.einsum('aij,jka->aik', corners, rot_t)
    corners += centers.reshape([-1, 1, 3])
    surf = np.array([[corners[:, i] for i in f] for f in _FACES]).transpose([2, 0, 1, 3])   # [n][6][4][3]
    vec = surf[:, :, :2, :] - surf[:, :, 1:3, :]
    normal = np.cross(vec[:, :, 0, :], vec[:, :, 1, :])
    d = -np.einsum('aij, aij->ai', normal, surf[:, :, 0, :])
    return np.concatenate([normal, d[..., None]], axis=-1).astype(rb.dtype if rb.dtype == np.float64 else np.float32)


def crop_points(points, boxes_lidar, without_reflectivity=True):
    """points: CUDA float32 [P][C>=3] scene cloud; boxes_lidar: [n][7] (numpy / CPU tensor, LiDAR frame).
    Returns (points_out CUDA [P_out][3 or C], points_split CPU int64 [n+1]) — the layout of
    ``det_info['points'][0]`` / ``det_info['points_split'][0]``."""
    lib = _lib.load()
    if points.device.type != "cuda":
        raise _lib.MmmotError("mmmot_b200.crop_points runs on CUDA only (no CPU fallback)")
    points = points.contiguous().float()
    P, C = points.shape
    boxes = np.asarray(boxes_lidar)
    f64 = boxes.dtype == np.float64         # the reference evaluates the predicate in the boxes' precision
    boxes = boxes.astype(np.float64 if f64 else np.float32).reshape(-1, 7)
    n = boxes.shape[0]
    dev = points.device
    with torch.cuda.device(dev):            # the library works on the CURRENT device
        planes = torch.from_numpy(np.ascontiguousarray(box_planes(boxes))).to(dev)
        ws = torch.empty(int(lib.mmmot_crop_workspace(P, n)), dtype=torch.uint8, device=dev)
        split = torch.empty(n + 1, dtype=torch.int32, device=dev)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.mmmot_crop_count(vp(points), P, C, vp(planes), int(f64), n, vp(split), vp(ws), ws.numel(), st), "mmmot_crop_count")
        split_h = split.cpu()                      # output size is data dependent: one sync, like the reference's host loop
        out_c = 3 if without_reflectivity else min(C, 4)
        out = torch.empty(int(split_h[-1]), out_c, device=dev)
        _lib.check(lib.mmmot_crop_scatter(vp(points), P, C, vp(planes), int(f64), n, vp(split), out_c, vp(out), vp(ws), ws.numel(), st),
                   "mmmot_crop_scatter")
    return out, split_h.long()
