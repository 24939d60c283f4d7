"""LiDAR cropping: oracle vs goldens of the UNMODIFIED reference (CPU), GPU kernel vs both."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR
from mmmot_b200.lidar_crop import box_camera_to_lidar, box_planes
from oracle.crop_ref import crop_points_ref

GOLD = sorted(glob.glob(os.path.join(GOLDEN_DIR, "crop_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    out, split = crop_points_ref(g["points"], g["boxes"])
    assert np.array_equal(split, g["split"]) and np.array_equal(out, g["out"])          # bit-exact
    lid = box_camera_to_lidar(g["cam"], g["rect"], g["v2c"])
    assert lid.dtype == np.float64 and np.array_equal(lid, g["lidar"])                  # promoted like the reference
    out, split = crop_points_ref(g["points"], g["boxes64"])                             # float64 boxes: float64 predicate
    assert np.array_equal(split, g["split64"]) and np.array_equal(out, g["out64"])


def test_plane_normals_point_inward():
    boxes = np.array([[1.0, 2.0, -0.5, 2.0, 4.0, 1.6, 0.7]], np.float32)
    pl = box_planes(boxes)[0]
    c = np.array([1.0, 2.0, -0.5 + 0.8], np.float32)                                   # box centre (origin z = 0)
    assert np.all(pl[:, :3] @ c + pl[:, 3] < 0)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_gpu_crop_matches_reference_golden(path):
    import mmmot_b200
    g = np.load(path)
    out, split = mmmot_b200.crop_points(torch.from_numpy(g["points"]).cuda(), g["boxes"])
    assert torch.equal(split, torch.from_numpy(g["split"]))
    assert torch.equal(out.cpu(), torch.from_numpy(g["out"]))                            # membership + order bit-exact
    out, split = mmmot_b200.crop_points(torch.from_numpy(g["points"]).cuda(), g["boxes64"])   # the float64 pipeline
    assert torch.equal(split, torch.from_numpy(g["split64"])) and torch.equal(out.cpu(), torch.from_numpy(g["out64"]))


@pytest.mark.gpu
def test_gpu_crop_scene_scale_and_feeds_pointnet():
    """KITTI-scale scene (120k points, 128 boxes): identical to the oracle, and the result is directly the
    (points, points_split) pair of the forward."""
    import mmmot_b200
    rng = np.random.default_rng(3)
    P, n = 120000, 128
    centers = rng.uniform([0, -30, -2], [70, 30, 0], size=(n, 3)).astype(np.float32)
    pts = np.concatenate([centers[rng.integers(0, n, P)] + rng.normal(size=(P, 3)) * [3.0, 2.0, 1.0],
                          rng.uniform(size=(P, 1))], 1).astype(np.float32)
    boxes = np.concatenate([centers, rng.uniform([1.2, 2.5, 1.2], [2.2, 5.0, 2.0], size=(n, 3)),
                            rng.uniform(-3.14, 3.14, size=(n, 1))], 1).astype(np.float32)
    out, split = mmmot_b200.crop_points(torch.from_numpy(pts).cuda(), boxes)
    ro, rs = crop_points_ref(pts, boxes)
    assert torch.equal(split, torch.from_numpy(rs)) and torch.equal(out.cpu(), torch.from_numpy(ro))
    assert out.shape[1] == 3 and int(split[-1]) == out.shape[0] and bool((split[1:] > split[:-1]).all())
