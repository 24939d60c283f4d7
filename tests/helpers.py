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


def golden_cases():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.pt"))):
        out.append(torch.load(f))
    return out


def case_tol(case):
    n, m = case[5], case[6]
    return TOL_DEGENERATE if n * m <= 4 else TOL


def det_close(a, ref, thr, tol):
    """det scores carry a hard step at neg_threshold (tracking_net.py:161-162): s - [s < thr].
    Compare on the pre-step value so a 1e-7 wobble at the threshold cannot flip a whole unit."""
    un = lambda s: torch.where(s < 0, s + 1, s)
    a, ref = un(a.detach().cpu().double()), un(ref.detach().cpu().double())
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30)) <= tol
