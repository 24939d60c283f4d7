"""Track-id assignment, stitching and KITTI text against goldens produced by the UNMODIFIED
reference TrackingModule / write_kitti_result (oracle/make_goldens.py::stitch_goldens)."""
import copy
import json
import os
import types

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, stitch_scenario


def _module(model=None):
    from mmmot_b200.tracking_model import TrackingModule
    return TrackingModule(model or types.SimpleNamespace(test_mode=0, eval=lambda: None), None, None, det_type="3D")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ids_and_kitti_text_match_reference(seed, tmp_path):
    from mmmot_b200.tracking_model import write_kitti_result
    gold = json.load(open(os.path.join(GOLDEN_DIR, "stitch.json")))[str(seed)]
    dets, samples = stitch_scenario(seed)
    tm = _module()
    for ((a, b), split, a_det, a_link, a_new, a_end), want in zip(samples, gold["steps"]):
        pair = [copy.deepcopy(dets[a]), copy.deepcopy(dets[b])]
        ids, boxes = tm.assign_det_id(a_det, a_link, a_new, a_end, split, pair)
        assert [[int(v) for v in x] for x in ids] == want["local"]
        for x, d in zip(ids, boxes):                      # kept rows only, ids mirrored into the dict
            assert d["id"].tolist() == [int(v) for v in x]
            assert d["bbox"].shape[0] == len(x) if len(x) else d["bbox"].numel() == 0
        aligned, adets, start = tm.align_id(ids, boxes)
        assert [[int(v) for v in x] for x in aligned] == want["aligned"]
        assert int(start) == want["frame_start"]
        assert [int(d["frame_idx"][0]) for d in adets] == want["frames"]
        assert int(tm.last_id) == want["last_id"]
    assert [[int(v) for v in x] for x in tm.frames_id] == gold["frames_id"]
    write_kitti_result(str(tmp_path), "0000", "step", tm.frames_id, copy.deepcopy(tm.frames_det), part="val")
    assert open(tmp_path / "step" / "val" / "0000.txt").read() == gold["kitti"]


def test_kitti_line_defaults():
    from mmmot_b200.tracking_model import kitti_result_line
    line = kitti_result_line({"frame": 3, "id": 7, "name": "Car", "bbox": [1, 2, 3, 4]})
    assert line == "3 7 Car -1 -1 -10 1.0000 2.0000 3.0000 4.0000 -1 -1 -1 -1000 -1000 -1000 -10 0.0"
    with pytest.raises(ValueError):
        kitti_result_line({"frame": 3, "id": 7, "name": "Car"})


def test_unlinked_kept_detection_is_an_error():
    dets, samples = stitch_scenario(0)
    (a, b), split, a_det, a_link, a_new, a_end = samples[0]
    kb = np.flatnonzero(a_det.numpy()[int(split[0]):] == 1)
    a_new = a_new.clone(); a_link = [a_link[0].clone()]
    a_new[int(split[0]) + kb[0]] = 0
    a_link[0][0][:, kb[0]] = 0
    with pytest.raises(AssertionError):
        _module().assign_det_id(a_det, a_link, a_new, a_end, split, [copy.deepcopy(dets[a]), copy.deepcopy(dets[b])])


@pytest.mark.gpu
def test_predict_end_to_end_gpu():
    """TrackingModule.predict over a synthetic 4-frame sequence: forward + LP on the GPU, ids stitched on the host;
    every kept detection gets exactly one id and ids never repeat inside a frame."""
    import mmmot_b200
    from mmmot_b200.synthetic import synthetic_pair, synthetic_state_dict
    net = mmmot_b200.TrackingNet(2, appear_skippool=True, score_arch="branch_cls", score_fusion_arch="C",
                                 affinity_op="multiply", softmax_mode="none", neg_threshold=0.2, test_mode=2, dropblock=0)
    net.load_state_dict(synthetic_state_dict("C", seed=0))
    net.cuda().eval()
    tm = mmmot_b200.TrackingModule(net, None, None, det_type="3D")
    tm.eval()
    n = 6
    frames = []
    for t in range(4):
        crops, det_info, _ = synthetic_pair(n // 2, n - n // 2, 16, 32, seed=10 + t)    # n detections of one frame
        g = torch.Generator().manual_seed(50 + t)
        frames.append((crops, det_info, {
            "name": torch.zeros(1, n).long(), "truncated": torch.zeros(1, n), "occluded": torch.zeros(1, n).long(),
            "alpha": torch.zeros(1, n), "bbox": torch.rand(1, n, 4, generator=g) * 100,
            "dimensions": torch.rand(1, n, 3, generator=g), "location": torch.rand(1, n, 3, generator=g) * 30,
            "rotation_y": torch.zeros(1, n), "frame_idx": torch.tensor([t])}))
    for t in range(3):
        ca, ia, da = frames[t]
        cb, ib, db = frames[t + 1]
        crops = torch.cat([ca, cb]).cuda()
        pts_a, pts_b = ia["points"][0], ib["points"][0]
        sa, sb = ia["points_split"][0], ib["points_split"][0]
        info = {"points": torch.cat([pts_a, pts_b])[None].cuda(),
                "points_split": torch.cat([sa, sb[1:] + sa[-1]])[None].cuda()}
        ids, out, start = tm.predict(crops, info, [copy.deepcopy(da), copy.deepcopy(db)], [torch.tensor([n]), torch.tensor([n])])
        assert start == (0 if t == 0 else 1)
        for x, d in zip(ids, out):
            assert len(set(int(v) for v in x)) == len(x)
            assert d["id"].tolist() == [int(v) for v in x]
    assert len(tm.frames_id) == len(tm.frames_det)
