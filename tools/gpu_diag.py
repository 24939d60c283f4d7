"""First-contact GPU diagnostic: per-stage errors vs the oracle, printed (not asserted), so one
gpurun call shows where a bug lives.  Writes gpurun_out/diag.txt."""
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmmot_b200  # noqa: E402
from mmmot_b200.synthetic import synthetic_batch, synthetic_pair, synthetic_state_dict  # noqa: E402
from oracle import lp_ref, torch_ref  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    out.write(s + "\n")
    out.flush()


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def stage(fusion, op, sm, n, m, pts, hw, ragged, seed):
    net = mmmot_b200.TrackingNet(2, appear_skippool=True, score_arch="branch_cls", score_fusion_arch=fusion,
                                 affinity_op=op, softmax_mode=sm, neg_threshold=0.2, test_mode=2, dropblock=0)
    sd = synthetic_state_dict(fusion, seed=seed)
    net.load_state_dict(sd)
    net.cuda().eval()
    dets, info, split = synthetic_pair(n, m, pts, hw, seed=seed, ragged=ragged)
    (rdet, rlink, rnew, rend, rtrans), st = torch_ref.forward(sd, dets, info, split, fusion, op, sm, 0.2, return_stages=True)
    o = net.forward_batch(dets.cuda(), info["points"][0].cuda(), info["points_split"][0], n, m, keep_feats=True)
    torch.cuda.synchronize()
    P(f"[{fusion} {op} {sm} n={n} m={m} pts={pts} hw={hw}]")
    P("   appear(stack0) %.2e  points(stack1) %.2e  fused(stack2) %.2e" % tuple(rel(o["feats"][0, s], st["feats"][s]) for s in range(3)))
    P("   det %.2e link %.2e new %.2e end %.2e" % (rel(o["det"][0], rdet), rel(o["link"][0], rlink[0]),
                                                    rel(o["new"][0], rnew[:, n:]), rel(o["end"][0], rend[:, :n])))
    # affinity in isolation on the oracle's features
    l2, n2, e2 = net.associate_batch(st["feats"].unsqueeze(0).cuda(), n, m)
    P("   affinity-only on oracle feats: link %.2e new %.2e end %.2e" % (rel(l2[0], rlink[0]), rel(n2[0], rnew[:, n:]), rel(e2[0], rend[:, :n])))
    return net, sd


try:
    P(torch.cuda.get_device_name(0))
    ENG = os.environ.get("MMMOT_DIAG_ENGINE", "auto")
    mmmot_b200.set_engine(ENG)
    P("engine:", ENG)
    if os.environ.get("MMMOT_KSEG"):
        mmmot_b200._lib.load().mmmot_set_kseg(int(os.environ["MMMOT_KSEG"]))
        P("kseg:", os.environ["MMMOT_KSEG"])
    for cfg in (("A", "multiply", "none", 8, 8, 32, 32, False, 1), ("C", "minus_abs", "dual_add", 6, 9, 24, 64, True, 2),
                ("B", "multiply", "single", 16, 16, 64, 64, True, 3), ("C", "minus_abs", "dual_add", 32, 32, 128, 64, True, 4)):
        try:
            stage(*cfg)
        except Exception:
            P("STAGE FAILED", cfg, traceback.format_exc())
    # LP
    g = torch.Generator().manual_seed(0)
    for n, m in ((3, 2), (8, 8), (32, 32), (64, 64)):
        L = n + m
        det = torch.rand(2, L, generator=g) - (torch.rand(2, L, generator=g) < 0.3).float()
        link = torch.rand(2, n, m, generator=g)
        new = torch.cat([torch.zeros(2, n), torch.rand(2, m, generator=g)], 1)
        end = torch.cat([torch.rand(2, n, generator=g), torch.zeros(2, m)], 1)
        r = mmmot_b200.solve_batch(det.cuda(), link.cuda(), new.cuda(), end.cuda(), n, m)
        torch.cuda.synchronize()
        for b in range(2):
            a, obj, _ = lp_ref.milp_solve(det[b], [link[b:b + 1]], new[b], end[b], [n, m])
            got = (r["assign_det"][b].cpu(), [r["assign_link"][b:b + 1].cpu()], r["assign_new"][b].cpu(), r["assign_end"][b].cpu())
            P(f"   LP n={n} m={m} b={b}: obj gpu {lp_ref.objective(det[b], [link[b:b+1]], new[b], end[b], got):.9f} milp {obj:.9f} same_link {torch.equal(got[1][0], a[1][0])} same_det {torch.equal(got[0], a[0])}")
    # timing at the BASELINE shape
    net = mmmot_b200.TrackingNet(2, appear_skippool=True, score_arch="branch_cls", score_fusion_arch="C", affinity_op="minus_abs",
                                 softmax_mode="dual_add", neg_threshold=0.2, test_mode=2, dropblock=0)
    net.load_state_dict(synthetic_state_dict("C", 0))
    net.cuda().eval()
    for B, n, pts in ((2, 32, 128), (2, 128, 512), (8, 128, 512)):
        crops, p, split = synthetic_batch(B, n, pts=pts, hw=64, seed=0)
        crops, p = crops.cuda(), p.cuda()
        net.predict_batch(crops, p, split, n)
        torch.cuda.synchronize()
        t = time.time()
        net.predict_batch(crops, p, split, n)
        torch.cuda.synchronize()
        dt = time.time() - t
        P(f"   timing B={B} n={n} pts={pts}: {dt*1e3:.1f} ms  -> {B/dt:.2f} pairs/s   launches so far {mmmot_b200._lib.load().mmmot_launch_count()}")
except Exception:
    P("DIAG FAILED", traceback.format_exc())
out.close()
