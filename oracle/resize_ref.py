"""ORACLE (test infrastructure, never imported by the product) — per-detection crop-and-resize.

Restates what the reference does on the host for every detection:

    dataset/test_seq_dataset.py:212-218   x1,y1 = floor(bbox[:2]); x2,y2 = ceil(bbox[2:])
                                          img.crop((x1, y1, x2, y2)).resize((224, 224), Image.BILINEAR)
    utils/build_util.py:137-142           valid_transform = Resize(test_resize) -> CenterCrop(input_size)
                                          -> ToTensor -> Normalize(mean, std)       (test_resize = input_size = 224
                                          in every shipped config, so Resize/CenterCrop are identities)

The resize is Pillow's two-pass (horizontal, then vertical) antialiased resampler on 8-bit data; Pillow is a
dependency of the reference, not part of it, so its published algorithm (src/libImaging/Resample.c, 8bpc path) is
restated here: double-precision triangle-filter coefficients normalised per output pixel, converted to 22-bit fixed
point, integer accumulation with a rounding half added, clip to 8 bits after each pass.  Pinned bit-exactly against
Pillow itself (the version in this image) by tests/test_oracle.py::test_resize_oracle_matches_pillow and against
the goldens in tests/golden/resize_*.npz, which oracle/make_goldens.py produces with the reference's own calls
(PIL crop/resize + torchvision transforms).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
MEAN = (0.485, 0.456, 0.406)          # utils/build_util.py:110-111
STD = (0.229, 0.224, 0.225)


def crop_box(bbox):
    """float bbox (x1, y1, x2, y2) -> integer crop box, reference test_seq_dataset.py:212-215."""
    b = np.asarray(bbox, dtype=np.float64)
    return int(np.floor(b[0])), int(np.floor(b[1])), int(np.ceil(b[2])), int(np.ceil(b[3]))


def coeffs(in_size, out_size):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the triangle filter over the full input range.
    Returns (bounds [out][2] = (first tap, tap count), kk [out][ksize] int)."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        w = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            w[:xmax] = w[:xmax] / ww
        for x in range(xmax):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(src, out_size, axis):
    """One resampling pass of uint8 src [h][w][c] along `axis` (1 = horizontal, 0 = vertical)."""
    if src.shape[axis] == out_size:
        return src                                          # Pillow skips a pass that does not change the size
    bounds, kk = coeffs(src.shape[axis], out_size)
    s = np.moveaxis(src, axis, 0).astype(np.int64)          # [in][...]
    out = np.empty((out_size,) + s.shape[1:], np.uint8)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = np.full(s.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += s[x0 + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def crop_resize_u8(image, box, out_size=224):
    """image uint8 [H][W][3]; box integer (x1, y1, x2, y2) -> uint8 [out][out][3].  Pixels outside the image are 0
    (PIL Image.crop pads with black)."""
    x1, y1, x2, y2 = box
    H, W = image.shape[:2]
    crop = np.zeros((y2 - y1, x2 - x1, image.shape[2]), np.uint8)
    sx1, sy1, sx2, sy2 = max(x1, 0), max(y1, 0), min(x2, W), min(y2, H)
    if sx2 > sx1 and sy2 > sy1:
        crop[sy1 - y1:sy2 - y1, sx1 - x1:sx2 - x1] = image[sy1:sy2, sx1:sx2]
    return _pass(_pass(crop, out_size, 1), out_size, 0)


def to_tensor_normalize(u8):
    """uint8 [S][S][3] -> float32 [3][S][S]: torchvision ToTensor (x/255 in fp32) then Normalize ((x-mean)/std in fp32)."""
    x = u8.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    mean = np.asarray(MEAN, np.float32)[:, None, None]
    std = np.asarray(STD, np.float32)[:, None, None]
    return ((x - mean) / std).astype(np.float32)


def crop_resize_ref(image, bboxes, out_size=224):
    """All detections of a frame: float32 [n][3][out][out] — what the reference feeds TrackingNet.forward as `dets`."""
    return np.stack([to_tensor_normalize(crop_resize_u8(image, crop_box(b), out_size)) for b in bboxes])
