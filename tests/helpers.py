"""Shared test helpers.  tests/ is the one place (with smoke() and bench's CPU legs) allowed to use oracle/."""
import glob
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerance stated by BASELINE.json north_star: 1e-4 relative for affinity/score tensors.
# Metric: max|a-b| / max|ref| (the metric SURVEY F8 / B-6 used to rule out TF32/BF16 operands).
TOL = 1e-4
# GroupNorm over a 2..4-element grid is ill-conditioned (two fp32 CPU implementations — the
# reference and the oracle restatement — already differ by 7e-4 there), so the degenerate
# N*M <= 4 cases get a looser bound.
TOL_DEGENERATE = 5e-3


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# Element-wise companion of the max-norm metric: |a-b| <= RTOL_EL*|ref| + ATOL_EL.
# The absolute term covers entries that are themselves at fp32 round-off of the tensor's scale
# (softmax tails, ReLU zeros): 1e-6 is ~1e-2 of the max-norm bound for tensors of scale 1e-2..1.
RTOL_EL, ATOL_EL = 1e-4, 1e-6
# No element may miss that bound by more than this factor (the few that exceed it at all are counted and reported).
WORST_EL = 4.0


def frac_outside(a, b, rtol=RTOL_EL, atol=ATOL_EL):
    """Fraction of elements violating |a-b| <= rtol*|ref| + atol', and the worst ratio |a-b| / (rtol*|ref| + atol').
    atol' = max(atol, 1e-5 * max|ref|): for tensors whose scale is far above 1 (raw link logits of softmax_mode 'none')
    the absolute floor follows the tensor's scale, at a tenth of the max-norm bound."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    atol = max(atol, 1e-5 * float(b.abs().max()))
    bound = rtol * b.abs() + atol
    ratio = (a - b).abs() / bound
    return float((ratio > 1).double().mean()), float(ratio.max())


def check_close(a, ref, tol, what, report=None, max_outside=0.0):
    """Both metrics: max-norm relative error < tol and at most `max_outside` of the elements outside the
    element-wise bound.  Appends (what, max-norm error, fraction outside, worst element ratio) to `report`."""
    e = relerr(a, ref)
    fo, worst = frac_outside(a, ref)
    if report is not None:
        report.append((what, e, fo, worst))
    assert e < tol, f"{what}: max-norm relative error {e:.3e} >= {tol:g}"
    assert fo <= max_outside, f"{what}: {fo:.3e} of the elements outside |a-b| <= {RTOL_EL:g}|ref| + {ATOL_EL:g} (worst x{worst:.2f})"
    assert worst < WORST_EL, f"{what}: an element misses the element-wise bound by x{worst:.2f}"


def golden_cases():
    """Eval-mode forward fixtures (training-mode ones are train_*.pt, see train_cases)."""
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.pt"))):
        if not os.path.basename(f).startswith("train_"):
            out.append(torch.load(f))
    return out


def train_cases():
    """Training-mode fixtures of the unmodified reference (oracle/make_goldens.py::train_goldens)."""
    return [torch.load(f) for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "train_*.pt")))]


def synthetic_gt(n, m, seed):
    """Seeded class flags / track ids in the DataLoader layout generate_gt reads (same generator as the golden script)."""
    g = torch.Generator().manual_seed(4000 + seed)
    cls = [(torch.rand(1, k, generator=g) < 0.75).long() for k in (n, m)]
    ids0 = torch.randperm(n + 3, generator=g)[:n]
    ids1 = torch.randperm(n + 3, generator=g)[:m]
    return cls, [ids0.unsqueeze(0), ids1.unsqueeze(0)]


# loss configuration of the shipped experiments (config.yaml:39-45 via utils/build_util.py:147-155)
LOSS_KW = dict(smooth_ratio=0, detloss_type="bce", det_ratio=1.5, trans_ratio=0.001, trans_last=True, linkloss_type="l2")


def case_tol(case):
    n, m = case[5], case[6]
    return TOL_DEGENERATE if n * m <= 4 else TOL


def det_close(a, ref, thr, tol):
    """det scores carry a hard step at neg_threshold (tracking_net.py:161-162): s - [s < thr].
    Compare on the pre-step value so a 1e-7 wobble at the threshold cannot flip a whole unit."""
    un = lambda s: torch.where(s < 0, s + 1, s)
    a, ref = un(a.detach().cpu().double()), un(ref.detach().cpu().double())
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30)) <= tol


def synthetic_image(h=375, w=1242, seed=5):
    """Deterministic uint8 RGB test frame [h][w][3] (KITTI-sized by default): smooth structure + LCG noise, pure
    integer numpy arithmetic so the golden script and the tests regenerate the same bytes anywhere."""
    import numpy as np
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    img = np.empty((h, w, 3), np.uint8)
    state = (x * 1103515245 + y * 12345 + seed * 2654435761) & 0x7FFFFFFF
    for c in range(3):
        state = (state * 1103515245 + 12345 + c) & 0x7FFFFFFF
        smooth = ((x * (3 + c)) // 7 + (y * (5 - c)) // 3 + 40 * c) % 256
        block = (((x // 16) * 37 + (y // 16) * 91 + c * 17) % 5) * 23
        noise = (state >> 16) % 41
        img[..., c] = np.clip((smooth * 2 + block * 2 + noise * 3) // 4, 0, 255).astype(np.uint8)
    return img


RESIZE_BOXES = [
    (712.40, 143.00, 810.73, 307.92), (599.41, 156.40, 629.75, 189.25), (387.63, 181.54, 423.81, 203.12),
    (0.00, 120.30, 180.20, 374.00), (1100.5, 150.2, 1241.0, 374.9), (-12.3, 100.0, 60.5, 260.0),
    (500.0, 160.0, 724.0, 384.0), (300.2, 170.9, 302.9, 175.1), (10.0, 10.0, 234.0, 234.0),
    (0.0, 0.0, 1242.0, 375.0), (640.7, 172.3, 657.1, 186.8), (800.0, -20.0, 1000.0, 150.0),
    (222.22, 111.11, 555.55, 333.33), (900.1, 180.2, 1010.9, 250.7), (50.5, 200.5, 274.5, 300.5),
    (1200.0, 300.0, 1260.0, 390.0),
]


def stitch_scenario(seed=0, frames=8):
    """Synthetic evaluation sequence for the id-stitching tests: per-frame detection dicts in the
    DataLoader layout of the reference (leading batch dimension 1) and, for every consecutive frame pair, a feasible
    set of assignment matrices (kept flags, one-to-one links, new/end flags).  Covers: dropped detections, births,
    deaths, a frame without kept detections, a kept set that differs between the two samples sharing a frame, and a
    break in the frame numbering.  Returns (dets per frame, list of samples); a sample is
    (frame ids (a, b), det_split, assign_det, [assign_link 1 x na x nb], assign_new, assign_end)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    counts = [int(c) for c in rng.integers(3, 7, size=frames)]
    frame_no = list(range(frames))
    for t in range(frames // 2 + 1, frames):
        frame_no[t] += 5                                   # numbering break inside the sequence
    dets = []
    for t, n in enumerate(counts):
        bbox = rng.uniform(0, 300, size=(n, 4)).astype(np.float32)
        dets.append({
            "name": torch.from_numpy(rng.integers(0, 4, size=(1, n))).long(),
            "truncated": torch.from_numpy(rng.uniform(0, 1, size=(1, n)).astype(np.float32)),
            "occluded": torch.from_numpy(rng.integers(0, 3, size=(1, n))).long(),
            "alpha": torch.from_numpy(rng.uniform(-3, 3, size=(1, n)).astype(np.float32)),
            "bbox": torch.from_numpy(bbox[None]),
            "dimensions": torch.from_numpy(rng.uniform(1, 4, size=(1, n, 3)).astype(np.float32)),
            "location": torch.from_numpy(rng.uniform(-20, 40, size=(1, n, 3)).astype(np.float32)),
            "rotation_y": torch.from_numpy(rng.uniform(-3, 3, size=(1, n)).astype(np.float32)),
            "frame_idx": torch.tensor([frame_no[t]]),
        })
    keep = [rng.uniform(size=n) < 0.8 for n in counts]
    keep[3][:] = False                                     # a frame where nothing is kept
    samples = []
    for a in range(frames - 1):
        b = a + 1
        if frame_no[b] != frame_no[a] + 1:
            continue                                       # the sequence is cut here: no sample spans the break
        ka, kb = keep[a].copy(), keep[b].copy()
        if a == 1:
            ka[np.flatnonzero(~ka)[:1]] = True             # this sample keeps one more detection of frame a than the last
        na, nb = counts[a], counts[b]
        link = np.zeros((na, nb), np.float32)
        new = np.zeros(na + nb, np.float32)
        end = np.zeros(na + nb, np.float32)
        free = list(rng.permutation(np.flatnonzero(ka)))
        for j in np.flatnonzero(kb):
            if free and rng.uniform() < 0.7:
                link[free.pop(), j] = 1
            else:
                new[na + j] = 1
        for k in np.flatnonzero(ka):
            if link[k].sum() == 0:
                end[k] = 1
        assign_det = np.concatenate([ka, kb]).astype(np.float32)
        samples.append(((a, b), [torch.tensor([na]), torch.tensor([nb])], torch.from_numpy(assign_det),
                        [torch.from_numpy(link[None])], torch.from_numpy(new), torch.from_numpy(end)))
    return dets, samples
