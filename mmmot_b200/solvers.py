"""Association solver with the reference's ``ortools_solve`` signature.

Mirrors reference solvers.py:9-138: same arguments (``link_score`` is a list holding one
``1 x N x M`` tensor, ``det_split`` a list of two ints / 1-element tensors), same outputs — four
fp32 0/1 tensors ``(assign_det (L,), [assign_link 1xNxM], assign_new (L,), assign_end (L,))`` on
the input device — so reference tracking_model.py:72-81 (``predict`` -> ``assign_det_id``) works on
them unchanged.  The programme is solved exactly on the GPU by ``mmmot_lp_assign``
(csrc/lp_assign.cu); there is no CPU solver in the product.
"""
import ctypes

import torch

from . import _lib


def solve_batch(det, link, new, end, n, m):
    """det B x L, link B x n x m, new/end B x L (zero-padded as the forward returns them); CUDA,
    arbitrary strides between pairs (views into the forward outputs are fine).  Returns a dict of
    assign_det/new/end (B x L), assign_link (B x n x m), match (B x n int32, -1 = no link)."""
    lib = _lib.load()
    if det.device.type != "cuda":
        raise _lib.MmmotError("mmmot_b200 solver runs on CUDA only (no CPU fallback)")
    B, L = det.shape[0], n + m

    def rowview(t, inner):
        # one pair's data must be contiguous; the stride between pairs is free
        if t.dtype != torch.float32 or t[0].numel() != inner or not t[0].is_contiguous():
            t = t.float().contiguous()
        return t, (t.stride(0) if B > 1 else inner)
    det, sd = rowview(det, L)
    link, sl = rowview(link, n * m)
    new, sn = rowview(new, L)
    end, se = rowview(end, L)
    dev = det.device
    a_det = torch.empty(B, L, device=dev)
    a_new = torch.empty(B, L, device=dev)
    a_end = torch.empty(B, L, device=dev)
    a_link = torch.empty(B, n, m, device=dev)
    match = torch.empty(B, n, dtype=torch.int32, device=dev)
    ws = torch.empty(int(lib.mmmot_lp_workspace(B, n, m)), dtype=torch.uint8, device=dev)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):            # the library works on the CURRENT device
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.mmmot_lp_assign(vp(det), sd, vp(link), sl, vp(new), sn, vp(end), se, B, n, m,
                                       vp(a_det), vp(a_link), vp(a_new), vp(a_end), vp(match),
                                       vp(ws), ws.numel(), st), "mmmot_lp_assign")
    return {"assign_det": a_det, "assign_link": a_link, "assign_new": a_new, "assign_end": a_end,
            "match": match}


def ortools_solve(det_score, link_score, new_score, end_score, det_split, gt=None):
    """Drop-in for reference solvers.py:9.  The loss-augmented ``gt`` branch (:50-81) is never used
    on the predict path and is not implemented."""
    if gt is not None:
        raise NotImplementedError("loss-augmented solve (gt != None) is training-only; not implemented")
    if len(det_split) != 2 or len(link_score) != 1:
        raise NotImplementedError("only 2-frame samples are supported (sample_max_len: 2)")
    n, m = int(det_split[0]), int(det_split[1])
    r = solve_batch(det_score.reshape(1, -1), link_score[0].reshape(1, n, m),
                    new_score.reshape(1, -1), end_score.reshape(1, -1), n, m)
    return r["assign_det"][0], [r["assign_link"]], r["assign_new"][0], r["assign_end"][0]
