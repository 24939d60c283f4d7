"""Per-detection image crop-and-resize on the GPU (SURVEY.md §8f N2 — the image-side step before the hot path).

Mirrors what the reference's dataset does for every detection of a frame on the host with PIL
(dataset/test_seq_dataset.py:212-218: integer crop box = floor/ceil of the detection box,
``img.crop(box).resize((224, 224), Image.BILINEAR)``) followed by the evaluation transform
(utils/build_util.py:108-111, 137-142: ToTensor + Normalize; Resize(224)/CenterCrop(224) are identities),
and returns the float32 ``[n][3][S][S]`` tensor ``TrackingNet.forward`` takes as ``dets`` — bit-identical to
the PIL + torchvision result (tests/test_image_crop.py).  The frame is uploaded once as uint8 (1.4 MB for a
KITTI frame instead of n x 600 KB of float crops); all resampling runs in libmmmot_sm100a.so
(csrc/crop_resize.cu).  There is no CPU path.
"""
import ctypes

import numpy as np
import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # utils/build_util.py:110
IMAGENET_STD = (0.229, 0.224, 0.225)       # utils/build_util.py:111


def crop_boxes(bboxes):
    """float (x1, y1, x2, y2) detection boxes -> int32 crop boxes, reference test_seq_dataset.py:212-215."""
    b = np.asarray(bboxes)
    if b.ndim != 2 or b.shape[1] != 4:
        raise _lib.MmmotError("bboxes must be n x 4 (x1, y1, x2, y2)")
    out = np.stack([np.floor(b[:, 0]), np.floor(b[:, 1]), np.ceil(b[:, 2]), np.ceil(b[:, 3])], axis=1)
    return out.astype(np.int32)


def crop_resize(image, bboxes, out_size=224, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """image: uint8 CUDA tensor [H][W][3] (RGB, as PIL decodes it); bboxes: n x 4 host array / CPU tensor of float
    detection boxes.  Returns float32 CUDA [n][3][out_size][out_size]."""
    lib = _lib.load()
    if not (isinstance(image, torch.Tensor) and image.is_cuda and image.dtype == torch.uint8 and image.dim() == 3
            and image.shape[2] == 3):
        raise _lib.MmmotError("image must be a uint8 CUDA tensor of shape H x W x 3")
    image = image.contiguous()
    if isinstance(bboxes, torch.Tensor):
        bboxes = bboxes.detach().cpu().numpy()
    boxes = crop_boxes(bboxes)
    n = boxes.shape[0]
    if n == 0:
        return torch.empty(0, 3, out_size, out_size, device=image.device)
    cw, ch = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    if (cw <= 0).any() or (ch <= 0).any():
        raise _lib.MmmotError("empty crop box (x2 <= x1 or y2 <= y1)")
    # filter-tap stride: Pillow's ksize = ceil(support) * 2 + 1 with support = max(crop side / out_size, 1)
    taps = int(np.ceil(max(float(max(cw.max(), ch.max())) / out_size, 1.0))) * 2 + 1
    if taps > lib.mmmot_crop_resize_max_taps():
        raise _lib.MmmotError(f"crop side {max(cw.max(), ch.max())} too large for out_size {out_size}")
    row_off = np.zeros(n + 1, np.int64)
    np.cumsum(ch, out=row_off[1:])
    dev = image.device
    d_boxes = torch.from_numpy(boxes).to(dev)
    d_off = torch.from_numpy(row_off).to(dev)
    if out is None:
        out = torch.empty(n, 3, out_size, out_size, device=dev)
    elif out.shape != (n, 3, out_size, out_size) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
        raise _lib.MmmotError("out must be a contiguous float32 CUDA tensor n x 3 x S x S on the image's device")
    total = int(row_off[-1])
    ws = torch.empty(lib.mmmot_crop_resize_workspace(n, total, out_size, taps), dtype=torch.uint8, device=dev)
    ms = (ctypes.c_float * 6)(*[float(np.float32(v)) for v in (*mean, *std)])
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.mmmot_crop_resize(vp(image), image.shape[0], image.shape[1], vp(d_boxes), vp(d_off), n, total,
                                         int(ch.max()), out_size, taps, ctypes.cast(ms, ctypes.c_void_p), vp(out), vp(ws),
                                         ws.numel(), st), "mmmot_crop_resize")
    return out
