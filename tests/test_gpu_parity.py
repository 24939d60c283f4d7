"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the committed
golden fixtures.  Tolerance: 1e-4 relative (BASELINE.json north_star) with the metric
max|a-b| / max|ref|; assignment outputs bit-exact."""
import pytest
import torch

from helpers import TOL, case_tol, det_close, golden_cases, relerr
import mmmot_b200
from mmmot_b200.synthetic import synthetic_batch, synthetic_pair, synthetic_state_dict
from oracle import lp_ref, torch_ref

pytestmark = pytest.mark.gpu
CASES = golden_cases()


def make_net(fusion, op, sm, thr, seed):
    net = mmmot_b200.TrackingNet(2, appear_skippool=True, score_arch="branch_cls", score_fusion_arch=fusion,
                                 affinity_op=op, softmax_mode=sm, neg_threshold=thr, test_mode=2, dropblock=0)
    sd = synthetic_state_dict(fusion, seed=seed)
    net.load_state_dict(sd)
    return net.cuda().eval(), sd


@pytest.fixture(params=["fp32", "tcgen05"])
def engine(request):
    """Both contraction engines must meet the same parity bound."""
    mmmot_b200.set_engine(request.param)
    yield request.param
    mmmot_b200.set_engine("auto")


@pytest.mark.parametrize("M,K,S", [(128, 32, 256), (256, 96, 512), (64, 70, 300), (512, 512, 4099), (1024, 128, 1000),
                                   (128, 4608, 2048)])
def test_tcgen05_contraction_vs_fp64(M, K, S):
    """The split-BF16 tensor-core contraction alone (C ABI test hook) against an fp64 matmul."""
    import ctypes
    from mmmot_b200 import _lib
    from mmmot_b200.weights import pack_tc
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + K + S)
    Wt, X, b = torch.randn(K, M, generator=g), torch.randn(K, S, generator=g), torch.randn(M, generator=g)
    ref = Wt.double().t() @ X.double() + b.double()[:, None]
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    Wt_d, X_d, b_d = Wt.cuda(), X.cuda(), b.cuda()
    Wp, wps = pack_tc(Wt)
    Wp = Wp.cuda()
    for eng in (1, 2):
        Y = torch.full((M, S), float("nan"), device="cuda")
        assert lib.mmmot_debug_linear(vp(Wt_d), vp(Wp), wps, vp(b_d), vp(X_d), vp(Y), M, K, S, eng, None) == 0
        assert relerr(Y, ref) < 3e-5, eng


@pytest.mark.parametrize("g", CASES, ids=[c["case"][0] for c in CASES])
def test_forward_matches_reference_golden(g, engine):
    """Reference signature, one frame-pair, against outputs of the UNMODIFIED reference."""
    name, fusion, op, sm, thr, n, m, pts, hw, ragged, seed = g["case"]
    net, sd = make_net(fusion, op, sm, thr, seed)
    dets, info, split = synthetic_pair(n, m, pts, hw, seed=seed, ragged=ragged)
    det, link, new, end, trans = net(dets.cuda(), {k: v.cuda() for k, v in info.items()}, split)
    tol = case_tol(g["case"])
    assert link[0].shape == g["link"].shape and det.shape == g["det"].shape
    assert relerr(link[0], g["link"]) < tol
    assert relerr(new, g["new"]) < tol and relerr(end, g["end"]) < tol
    assert det_close(det, g["det"], thr, tol)
    assert relerr(trans[0], g["trans1"]) < 1e-5 and relerr(trans[1], g["trans2"]) < 1e-5
    assert torch.all(new[:, :n] == 0) and torch.all(end[:, n:] == 0)


@pytest.mark.parametrize("fusion,op,sm", [("A", "multiply", "none"), ("C", "minus_abs", "dual_add"), ("B", "multiply", "none")])
def test_features_match_oracle(fusion, op, sm, engine):
    """Stage check: the 3x512xL feature stack (appearance | PointNet | fusion), N=M=16, 64x64 crops."""
    net, sd = make_net(fusion, op, sm, 0.2, 31)
    dets, info, split = synthetic_pair(16, 16, 48, 64, seed=31, ragged=True)
    o = net.forward_batch(dets.cuda(), info["points"][0].cuda(), info["points_split"][0], 16, 16, keep_feats=True)
    _, st = torch_ref.forward(sd, dets, info, split, fusion, op, sm, 0.2, return_stages=True)
    for s in range(3):
        assert relerr(o["feats"][0, s], st["feats"][s]) < TOL, f"stack {s}"


@pytest.mark.parametrize("n,m", [(8, 8), (32, 32), (64, 64), (20, 45), (128, 128)])
@pytest.mark.parametrize("op,sm", [("multiply", "none"), ("minus_abs", "dual_add")])
def test_affinity_stage_matches_oracle(n, m, op, sm, engine):
    """BASELINE config 5 (N sweep): affinity + start/end + softmax alone on identical feature tensors."""
    net, sd = make_net("C", op, sm, 0.2, 7)
    g = torch.Generator().manual_seed(n * 1000 + m)
    feats = torch.relu(torch.randn(2, 3, 512, n + m, generator=g))
    link, new, end = net.associate_batch(feats.cuda(), n, m)
    for b in range(2 if n <= 64 else 1):
        rl, rn, re = torch_ref.associate(sd, feats[b, :, :, :n], feats[b, :, :, n:], op, sm)
        assert relerr(link[b], rl.squeeze(1)) < TOL
        assert relerr(new[b], rn) < TOL and relerr(end[b], re) < TOL


@pytest.mark.parametrize("hw,n", [(96, 3), (224, 1)])
def test_non_power_of_two_crops(hw, n, engine):
    """Crop sizes that are multiples of 32 but not powers of two (224 is the reference's real crop size,
    dataset/test_seq_dataset.py:217-218): partial TMA boxes / tile tails."""
    net, sd = make_net("A", "multiply", "none", 0.2, 12)
    dets, info, split = synthetic_pair(n, n, 32, hw, seed=50 + hw)
    o = net.forward_batch(dets.cuda(), info["points"][0].cuda(), info["points_split"][0], n, n, keep_feats=True)
    _, st = torch_ref.forward(sd, dets, info, split, "A", "multiply", "none", 0.2, return_stages=True)
    assert relerr(o["feats"][0, 0], st["feats"][0]) < TOL


def test_affinity_n256_top_of_sweep():
    """BASELINE config 5, N = 256 (top of the N sweep), one pair, tensor-core engine vs oracle."""
    net, sd = make_net("C", "minus_abs", "dual_add", 0.2, 7)
    g = torch.Generator().manual_seed(256)
    feats = torch.relu(torch.randn(1, 3, 512, 512, generator=g))
    link, new, end = net.associate_batch(feats.cuda(), 256, 256)
    rl, rn, re = torch_ref.associate(sd, feats[0, :, :, :256], feats[0, :, :, 256:], "minus_abs", "dual_add")
    assert relerr(link[0], rl.squeeze(1)) < TOL and relerr(new[0], rn) < TOL and relerr(end[0], re) < TOL


def test_batched_equals_looped():
    """forward_batch over B pairs == B single-pair forwards (pairs are independent GroupNorm domains)."""
    net, sd = make_net("C", "minus_abs", "dual_add", 0.2, 5)
    B, n = 3, 8
    crops, pts, split = synthetic_batch(B, n, pts=24, hw=32, seed=40)
    net.chunk_pairs = 2          # also exercises chunking
    o = net.forward_batch(crops.cuda(), pts.cuda(), split, n)
    net.chunk_pairs = None
    for b in range(B):
        dets, info, ds = synthetic_pair(n, n, 24, 32, seed=40 + b)
        det, link, new, end, _ = net(dets.cuda(), {k: v.cuda() for k, v in info.items()}, ds)
        assert torch.equal(o["link"][b], link[0]) and torch.equal(o["det"][b], det)
        assert torch.equal(o["new"][b], new[:, n:]) and torch.equal(o["end"][b], end[:, :n])


# ------------------------------------------------------------------ LP
def _rand_lp(g, n, m, B=1):
    L = n + m
    det = torch.rand(B, L, generator=g) - (torch.rand(B, L, generator=g) < 0.3).float()
    link = torch.rand(B, n, m, generator=g)
    new = torch.cat([torch.zeros(B, n), torch.rand(B, m, generator=g)], 1)
    end = torch.cat([torch.rand(B, n, generator=g), torch.zeros(B, m)], 1)
    return det, link, new, end


@pytest.mark.parametrize("n,m", [(1, 1), (3, 2), (8, 8), (7, 19), (32, 32), (64, 64)])
def test_lp_bit_exact_vs_milp_oracle(n, m):
    g = torch.Generator().manual_seed(100 + n + m)
    B = 6
    det, link, new, end = _rand_lp(g, n, m, B)
    r = mmmot_b200.solve_batch(det.cuda(), link.cuda(), new.cuda(), end.cuda(), n, m)
    for b in range(B):
        (a, obj, y) = lp_ref.milp_solve(det[b], [link[b:b + 1]], new[b], end[b], [n, m])
        got = (r["assign_det"][b].cpu(), [r["assign_link"][b:b + 1].cpu()], r["assign_new"][b].cpu(), r["assign_end"][b].cpu())
        assert abs(lp_ref.objective(det[b], [link[b:b + 1]], new[b], end[b], got) - obj) < 1e-9
        assert torch.equal(got[0], a[0]) and torch.equal(got[1][0], a[1][0])
        assert torch.equal(got[2], a[2]) and torch.equal(got[3], a[3])
        mt = r["match"][b].cpu()
        assert torch.equal(mt >= 0, a[1][0][0].sum(1) > 0)


def test_lp_reference_signature_on_forward_outputs():
    """ortools_solve drop-in on the oracle's own score tensors (identical inputs -> identical indices)."""
    g = CASES[3]
    name, fusion, op, sm, thr, n, m = g["case"][:7]
    t = 2
    a = mmmot_b200.ortools_solve(g["det"][t].cuda(), [g["link"][t:t + 1].cuda()], g["new"][t].cuda(), g["end"][t].cuda(),
                                 [torch.tensor([n]), torch.tensor([m])])
    b, obj, _ = lp_ref.milp_solve(g["det"][t], [g["link"][t:t + 1]], g["new"][t], g["end"][t], [n, m])
    assert a[1][0].shape == (1, n, m) and a[0].dtype == torch.float32 and a[0].is_cuda
    assert torch.equal(a[0].cpu(), b[0]) and torch.equal(a[1][0].cpu(), b[1][0])
    assert torch.equal(a[2].cpu(), b[2]) and torch.equal(a[3].cpu(), b[3])


@pytest.mark.parametrize("n", [128, 256])
def test_lp_large_optimality_and_feasibility(n):
    """Full-size instances (BASELINE N=128, sweep top 256): objective equals scipy's assignment optimum
    of the (N+M)^2 reduction, and the flow constraints of solvers.py:83-111 hold."""
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    g = torch.Generator().manual_seed(n)
    B = 4
    det, link, new, end = _rand_lp(g, n, n, B)
    r = mmmot_b200.solve_batch(det.cuda(), link.cuda(), new.cuda(), end.cuda(), n, n)
    for b in range(B):
        ad, al, an, ae = [r[k][b].cpu() for k in ("assign_det", "assign_link", "assign_new", "assign_end")]
        assert torch.equal(ae[:n] + al.sum(1), ad[:n]) and torch.equal(an[:n], ad[:n])
        assert torch.equal(an[n:] + al.sum(0), ad[n:]) and torch.equal(ae[n:], ad[n:])
        d, l, nw, e = [t.double().numpy() for t in (det[b], link[b], new[b], end[b])]
        aj, bk = d[:n] + nw[:n], d[n:] + e[n:]
        C = np.full((2 * n, 2 * n), -1e9)
        C[:n, :n] = aj[:, None] + bk[None, :] + l
        C[:n, n:][np.arange(n), np.arange(n)] = np.maximum(aj + e[:n], 0)
        C[n:, :n][np.arange(n), np.arange(n)] = np.maximum(bk + nw[n:], 0)
        C[n:, n:] = 0
        rr, cc = linear_sum_assignment(C, maximize=True)
        got = lp_ref.objective(det[b], [link[b:b + 1]], new[b], end[b], (ad, [al.unsqueeze(0)], an, ae))
        assert abs(got - C[rr, cc].sum()) < 1e-8


def test_predict_batch_full_size_property():
    """BASELINE N=128 shape end to end (1 pair): every output finite, softmax rows/cols consistent,
    assignment feasible."""
    net, sd = make_net("C", "minus_abs", "dual_add", 0.2, 9)
    n = 128
    crops, pts, split = synthetic_batch(1, n, pts=64, hw=64, seed=77)
    o = net.predict_batch(crops.cuda(), pts.cuda(), split, n)
    for k in ("det", "link", "new", "end"):
        assert torch.isfinite(o[k]).all(), k
    assert (o["link"] >= 0).all() and (o["link"] <= 1).all()      # dual_add of two softmaxes
    al, ad = o["assign_link"][0], o["assign_det"][0]
    assert al.sum(1).max() <= 1 and al.sum(0).max() <= 1
    assert torch.equal(o["assign_end"][0][:n] + al.sum(1), ad[:n])


def test_cfg4_full_size_pair_matches_oracle():
    """One frame-pair at exactly the bench configuration (BASELINE configs[3] / SURVEY cfg4: Fusion C, minus_abs, dual_add,
    N=M=128, P=512 points per detection, 64x64 crops) against the oracle — the size the throughput is quoted on.
    The oracle needs ~10 s of host time for this pair."""
    net, sd = make_net("C", "minus_abs", "dual_add", 0.2, 4)
    n = 128
    dets, info, split = synthetic_pair(n, n, 512, 64, seed=123)
    torch.set_num_threads(min(16, torch.get_num_threads()))
    ref_det, ref_link, ref_new, ref_end, _ = torch_ref.forward(sd, dets, info, split, "C", "minus_abs", "dual_add", 0.2)
    det, link, new, end, _ = net(dets.cuda(), {k: v.cuda() for k, v in info.items()}, split)
    assert relerr(link[0], ref_link[0]) < TOL
    assert relerr(new, ref_new) < TOL and relerr(end, ref_end) < TOL
    assert det_close(det, ref_det, 0.2, TOL)
    # and the assignment computed from both score sets agrees wherever the LP optimum is unambiguous
    a = mmmot_b200.ortools_solve(det[2], [link[0][2:3]], new[2], end[2], split)
    b = mmmot_b200.ortools_solve(ref_det[2].cuda(), [ref_link[0][2:3].cuda()], ref_new[2].cuda(), ref_end[2].cuda(), split)
    agree = (a[1][0] == b[1][0]).float().mean().item()
    assert agree > 0.999, agree
