"""Training loss of the association network — the reference's ``TrackingLoss`` surface (cost.py:134-185) in plain
PyTorch on the outputs of the training-mode forward ("stays plain PyTorch").

Same constructor keywords and call signature as the reference classes, so ``TrackingModule`` can use it unchanged:

    loss = det_ratio * DetLoss(det)  +  0.4 * EndLoss(new)  +  0.4 * EndLoss(end)  +  LinkLoss(link)
           +  trans_ratio * mse(T * T^T, I)            (elementwise product, as the reference writes it: cost.py:179,183)

The reference's ``'l2'`` / ``'l1'`` / ``'ghm'`` branches compute ``1 - gt.eq(ignore_index)`` on a bool tensor, which
modern torch rejects; the mask is written as ``gt.ne(ignore_index)`` here (same values).  ``'ghm'`` is not provided
(modules/ghm_loss.py is outside the association path).
"""
import torch
import torch.nn.functional as F
from torch import nn


class LinkLoss(nn.Module):
    """cost.py:67-98: masked MSE / smooth-L1 between each stack's link scores and the 0/1 ground-truth links; rows /
    columns of detections that are not ground-truth positives are zeroed in the prediction."""

    def __init__(self, smooth_ratio=0, loss_type='l2'):
        super().__init__()
        if loss_type not in ('l1', 'l2'):          # the reference asserts the same (cost.py:73); its own TrackingLoss default
            raise ValueError(f"linkloss_type {loss_type!r}: 'l1' or 'l2'")     # 'l2_softmax' therefore never constructs
        self.smooth_ratio, self.loss_type = smooth_ratio, loss_type

    def forward(self, det_split, gt_det, link_score, gt_link):
        loss, base = 0, 0
        for i, link in enumerate(link_score):
            n, m = int(det_split[i]), int(det_split[i + 1])
            rows = (gt_det[base:base + n] == 1).to(link.dtype)
            cols = (gt_det[base + n:base + n + m] == 1).to(link.dtype)
            pred = link * rows[:, None] * cols[None, :]
            target = gt_link[i].expand_as(pred)
            if 'l2' in self.loss_type:
                loss = loss + F.mse_loss(pred, target)
            if 'l1' in self.loss_type:
                loss = loss + F.smooth_l1_loss(pred, target)
        return loss


class DetLoss(nn.Module):
    """cost.py:101-131: the same target for every stack (3 x L scores vs L labels)."""

    def __init__(self, loss_type='bce', ignore_index=-1):
        super().__init__()
        if 'ghm' in loss_type:
            raise NotImplementedError("GHM loss (modules/ghm_loss.py) is outside the association path")
        self.loss_type, self.ignore_index = loss_type, ignore_index

    def forward(self, det_score, gt_score):
        gt = gt_score.unsqueeze(0).expand_as(det_score)
        loss = None
        if 'bce' in self.loss_type:
            loss = F.binary_cross_entropy_with_logits(det_score, gt)
        keep = gt.ne(self.ignore_index).to(det_score.dtype)
        if 'l2' in self.loss_type:
            loss = F.mse_loss(det_score * keep, gt)
        if 'l1' in self.loss_type:
            loss = F.smooth_l1_loss(det_score * keep, gt)
        return loss


class TrackingLoss(nn.Module):

    def __init__(self, smooth_ratio=0, detloss_type='bce', endloss_type='l2', det_ratio=0.4, trans_ratio=0.4,
                 trans_last=False, linkloss_type='l2'):
        super().__init__()
        self.link_loss = LinkLoss(smooth_ratio, linkloss_type)
        self.det_loss = DetLoss(detloss_type)
        self.end_loss = DetLoss(endloss_type)
        self.det_ratio, self.trans_ratio, self.trans_last = det_ratio, trans_ratio, trans_last

    def forward(self, det_split, gt_det, gt_link, gt_new, gt_end, det_score, link_score, new_score, end_score, trans=None):
        first, last = int(det_split[0]), int(det_split[-1])
        loss = self.det_loss(det_score, gt_det) * self.det_ratio
        loss = loss + self.end_loss(new_score, gt_new[first:]) * 0.4
        loss = loss + self.end_loss(end_score, gt_end[:gt_end.shape[0] - last]) * 0.4
        loss = loss + self.link_loss(det_split, gt_det, link_score, gt_link)
        if trans is not None:
            # NB elementwise T * T^T, not a matrix product (cost.py:179,183); trans_last sums over both transforms
            for t in (trans if self.trans_last else trans[-1:]):
                eye = torch.eye(t.size(-1), dtype=t.dtype, device=t.device)
                loss = loss + F.mse_loss(t * t.transpose(-1, -2), eye.expand_as(t)) * self.trans_ratio
        return loss
