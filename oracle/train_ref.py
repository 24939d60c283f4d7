"""ORACLE — test infrastructure, NOT product code.

CPU fp32 restatement of the reference's TRAINING-mode forward and loss:
``TrackingNet.forward`` with ``self.training`` (modules/tracking_net.py:149-193: BatchNorm layers use batch statistics,
det_scores stay raw logits, new/end scores are not zero-padded) and ``TrackingLoss`` (cost.py:134-185), plus the
running-average update a training-mode BatchNorm performs.  DropBlock (SkipPool heads 2, 3) and the PointNet head's Dropout
are drawn from torch's generators in the reference's order (rrc_pfv config: dropblock 5, use_dropout True).  Pinned by tests/golden/train_*.pt, generated from the UNMODIFIED reference
in .train() mode by oracle/make_goldens.py (tests/test_oracle.py).
"""
import torch
import torch.nn.functional as F

from mmmot_b200.schema import VGG_POOL_AFTER, VGG_STAGES
from oracle import torch_ref

EPS = 1e-5


def _bn_train(x, sd, p, stats):
    """BatchNorm in .train(): batch mean / biased variance over every dim but channels (torch.nn.BatchNorm{1,2}d);
    records (mean, biased var, count) for the running-average update."""
    dims = [d for d in range(x.dim()) if d != 1]
    mean = x.mean(dims)
    var = x.var(dims, unbiased=False)
    shape = (1, -1) + (1,) * (x.dim() - 2)
    stats[p] = (mean, var, x.numel() // x.shape[1])
    return (x - mean.reshape(shape)) * torch.rsqrt(var.reshape(shape) + EPS) * sd[p + ".weight"].reshape(shape) + \
        sd[p + ".bias"].reshape(shape)


def appearance_train(sd, dets, stats, dropblock=0):
    """modules/appear_net.py:166-190 with the VGG BatchNorm2d layers in training mode (modules/vgg.py:67-80)."""
    x, maps = dets, []
    for s, stage in enumerate(VGG_STAGES):
        p = f"appearance.layers.{s}"
        for idx, _, _ in stage:
            x = F.conv2d(x, sd[f"{p}.{idx}.weight"], sd[f"{p}.{idx}.bias"], padding=1)
            x = F.relu(_bn_train(x, sd, f"{p}.{idx + 1}", stats))
            if idx in VGG_POOL_AFTER[s]:
                x = F.max_pool2d(x, 2, 2)
        maps.append(x)
    # appear_net.py:143-152: the heads after the 4th and 5th max-pool get a DropBlock2D(block_size=dropblock)
    return torch.cat([torch_ref.skip_pool(sd, s, fmap, dropblock if s >= 2 else 0) for s, fmap in enumerate(maps)], dim=-1)


def determine_det_train(sd, feats, stats):
    """modules/tracking_net.py:149-152 in training: w_det with BatchNorm1d batch statistics, raw logits."""
    x = F.conv1d(feats, sd["w_det.0.weight"], sd["w_det.0.bias"])
    x = F.relu(_bn_train(x, sd, "w_det.1", stats))
    x = F.conv1d(x, sd["w_det.3.weight"], sd["w_det.3.bias"])
    x = F.relu(_bn_train(x, sd, "w_det.4", stats))
    return F.conv1d(x, sd["w_det.6.weight"], sd["w_det.6.bias"]).squeeze(1)


@torch.no_grad()
def forward_train(sd, dets, det_info, dets_split, fusion_arch="C", affinity_op="multiply", softmax_mode="single",
                  dropblock=0, use_dropout=False):
    """-> (det_scores 3xL raw, [link 3xNxM], new 3xM, end 3xN, trans), bn_stats {prefix: (mean, biased var, count)}."""
    stats = {}
    app = appearance_train(sd, dets, stats, dropblock)
    pts, trans = torch_ref.pointnet(sd, det_info["points"].transpose(-1, -2), det_info["points_split"].long().squeeze(0),
                                    dropout=use_dropout)
    feats = torch_ref.fusion(sd, fusion_arch, torch.cat([app, pts], dim=-1).t().unsqueeze(0))
    det = determine_det_train(sd, feats, stats)
    n, m = int(dets_split[0]), int(dets_split[1])
    link, new_s, end_s = torch_ref.associate(sd, feats[:, :, :n], feats[:, :, n:n + m], affinity_op, softmax_mode)
    return (det, [link.squeeze(1)], new_s, end_s, trans), stats


def running_after(sd, stats, momentum=0.1):
    """Running averages after one training-mode forward (torch BatchNorm: unbiased variance, momentum 0.1)."""
    out = {}
    for p, (mean, var, cnt) in stats.items():
        out[p + ".running_mean"] = sd[p + ".running_mean"] * (1 - momentum) + momentum * mean
        out[p + ".running_var"] = sd[p + ".running_var"] * (1 - momentum) + momentum * var * (cnt / max(cnt - 1, 1))
    return out


def tracking_loss(det_split, gt_det, gt_link, gt_new, gt_end, det_score, link_score, new_score, end_score, trans,
                  det_ratio=0.4, trans_ratio=0.4, trans_last=False):
    """cost.py:134-185 with the loss types of the shipped configs (det 'bce', new/end 'l2', link 'l2')."""
    def l2(score, gt):                                             # DetLoss 'l2' branch, cost.py:121-123
        gt = gt.unsqueeze(0).repeat(score.size(0), 1)
        return F.mse_loss(score * (gt != -1).float(), gt)
    first, last = int(det_split[0]), int(det_split[-1])
    loss = F.binary_cross_entropy_with_logits(det_score, gt_det.unsqueeze(0).repeat(det_score.size(0), 1)) * det_ratio
    loss = loss + l2(new_score, gt_new[first:]) * 0.4 + l2(end_score, gt_end[:-last]) * 0.4
    base = 0
    for i, link in enumerate(link_score):                          # LinkLoss, cost.py:80-98
        n, m = int(det_split[i]), int(det_split[i + 1])
        mask = (gt_det[base:base + n] == 1).float()[:, None] * (gt_det[base + n:base + n + m] == 1).float()[None, :]
        loss = loss + F.mse_loss(link * mask, gt_link[i].repeat(link.size(0), 1, 1))
    for t in (trans if trans_last else trans[-1:]):
        loss = loss + F.mse_loss(t * t.transpose(-1, -2), torch.eye(t.size(-1)).expand_as(t)) * trans_ratio
    return loss
