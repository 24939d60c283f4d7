"""Per-detection image crop-and-resize: oracle vs goldens made by the reference's own PIL +
torchvision calls (CPU), CUDA path vs oracle / goldens bit-exact (GPU)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, RESIZE_BOXES, synthetic_image
from oracle import resize_ref


def _golden():
    return np.load(os.path.join(GOLDEN_DIR, "resize_kitti.npz"))


def test_resize_oracle_matches_golden():
    g = _golden()
    img = synthetic_image()
    out32 = resize_ref.crop_resize_ref(img, g["boxes"], 32)
    assert np.array_equal(out32, g["out32"])
    for i in (0, 9, 5, 15):      # 224: two stored crops + digests (boxes 5 and 15 reach outside the frame)
        u8 = resize_ref.crop_resize_u8(img, resize_ref.crop_box(g["boxes"][i]), 224)
        if i == 0:
            assert np.array_equal(u8, g["u8_224_first"])
        if i == 9:
            assert np.array_equal(u8, g["u8_224_full"])
        digest = hashlib.sha256(resize_ref.to_tensor_normalize(u8).tobytes()).hexdigest()
        assert digest == str(g["sha224"][i])


def test_resize_oracle_matches_pillow():
    """The restatement against Pillow itself (the reference's dependency), when it is installed."""
    Image = pytest.importorskip("PIL.Image")
    img = synthetic_image(200, 300, seed=11)
    pil = Image.fromarray(img, "RGB")
    rng = np.random.default_rng(3)
    for _ in range(12):
        w, h = rng.uniform(4, 280), rng.uniform(4, 190)
        x1, y1 = rng.uniform(-10, 300 - w + 10), rng.uniform(-10, 200 - h + 10)
        box = resize_ref.crop_box((x1, y1, x1 + w, y1 + h))
        for S in (224, 48):
            ref = np.asarray(pil.crop(box).resize((S, S), Image.BILINEAR))
            assert np.array_equal(resize_ref.crop_resize_u8(img, box, S), ref)


def test_crop_boxes_host():
    import mmmot_b200
    b = mmmot_b200.crop_boxes(np.asarray(RESIZE_BOXES))
    assert b.dtype == np.int32
    assert [tuple(x) for x in b[:2]] == [(712, 143, 811, 308), (599, 156, 630, 190)]
    assert tuple(b[5]) == (-13, 100, 61, 260)


@pytest.mark.gpu
@pytest.mark.parametrize("S", [224, 32, 64])
def test_crop_resize_gpu_bit_exact(S):
    import mmmot_b200
    g = _golden()
    img = synthetic_image()
    d_img = torch.from_numpy(img).cuda()
    out = mmmot_b200.crop_resize(d_img, g["boxes"], out_size=S).cpu().numpy()
    ref = resize_ref.crop_resize_ref(img, g["boxes"], S)
    assert out.shape == ref.shape and out.dtype == np.float32
    assert np.array_equal(out, ref)
    if S == 32:
        assert np.array_equal(out, g["out32"])
    if S == 224:
        for i, o in enumerate(out):
            assert hashlib.sha256(o.tobytes()).hexdigest() == str(g["sha224"][i])


@pytest.mark.gpu
def test_crop_resize_gpu_random_boxes():
    import mmmot_b200
    img = synthetic_image(240, 400, seed=2)
    rng = np.random.default_rng(9)
    boxes = []
    for _ in range(40):
        w, h = rng.uniform(1.5, 390), rng.uniform(1.5, 230)
        x1, y1 = rng.uniform(-15, 400 - w + 15), rng.uniform(-15, 240 - h + 15)
        boxes.append((x1, y1, x1 + w, y1 + h))
    boxes.append((5.0, 5.0, 6.0, 6.0))            # 1 x 1 crop
    boxes.append((0.0, 0.0, 400.0, 240.0))        # whole frame
    boxes = np.asarray(boxes)
    out = mmmot_b200.crop_resize(torch.from_numpy(img).cuda(), boxes, out_size=56).cpu().numpy()
    assert np.array_equal(out, resize_ref.crop_resize_ref(img, boxes, 56))


@pytest.mark.gpu
def test_crop_resize_feeds_forward():
    """The cropped tensor is directly the `dets` argument of the hot path (H, W multiples of 32)."""
    import mmmot_b200
    from mmmot_b200.synthetic import synthetic_pair, synthetic_state_dict
    img = synthetic_image()
    crops = mmmot_b200.crop_resize(torch.from_numpy(img).cuda(), np.asarray(RESIZE_BOXES[:8]), out_size=64)
    net = mmmot_b200.TrackingNet(2, appear_skippool=True, score_arch="branch_cls", score_fusion_arch="C",
                                 affinity_op="multiply", softmax_mode="none", neg_threshold=0.2, test_mode=2, dropblock=0)
    net.load_state_dict(synthetic_state_dict("C", seed=0))
    net.cuda().eval()
    _, det_info, dets_split = synthetic_pair(4, 4, 16, 64, seed=1)
    det_info = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in det_info.items()}
    det, link, new, end, _ = net(crops, det_info, dets_split)
    assert det.shape == (3, 8) and torch.isfinite(det).all() and torch.isfinite(link[0]).all()
