"""Losses of the SViT training step without host synchronisation.

Reference: slowfast/models/losses.py:50-168 (`boxes_loss_`, `VideoImageLoss`) and the GIoU of
slowfast/utils/box_ops.py:41-77.  The reference branches on `tar_mask.sum() > 0` / `mask.sum() > 0` (a device -> host
round trip per step) and builds the full N x N GIoU matrix only to take its diagonal; here every term is a masked mean
over paired boxes, which has the same value and the same gradients -- including the "no valid target" case, where the
reference substitutes a fresh zero and these give zero with zero gradient.  On CUDA tensors `haog_loss` is ONE kernel
launch (svit_haog_loss: values and gradients of the four terms); the tensor expressions below are the host-side statement
of the same definition, pinned to the reference by tests/test_losses.py and used to check the kernel.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def box_cxcywh_to_xyxy(x):
    """utils/box_ops.py:26-30."""
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def paired_generalized_box_iou(b1, b2):
    """diag(generalized_box_iou(b1, b2)) of utils/box_ops.py:41-77 for boxes paired row by row ([N, 4] xyxy each),
    without the N x N matrix and without the host-side validity asserts."""
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = (torch.min(b1[:, 2:], b2[:, 2:]) - torch.max(b1[:, :2], b2[:, :2])).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = area1 + area2 - inter
    iou = inter / union
    whc = (torch.max(b1[:, 2:], b2[:, 2:]) - torch.min(b1[:, :2], b2[:, :2])).clamp(min=0)
    area = whc[:, 0] * whc[:, 1]
    return iou - (area - union) / area


def _masked_mean(values, mask, per_item: int = 1):
    """sum(values where mask) / (per_item * count(mask)); 0 (with zero gradient) when the mask is empty."""
    m = mask.to(values.dtype)
    total = torch.where(mask, values, torch.zeros_like(values)).sum()
    return total / (m.sum() * per_item).clamp(min=1.0)


def boxes_loss(pred, tar):
    """losses.py:50-92.  pred [B, T, O, 5] = (score logit, cx, cy, w, h); tar [B, T, O, 4] (all-zero rows = no box) or
    [B, T, O, 5] with a leading soft mask.  Returns (l1, bce, giou)."""
    if tar.size(-1) == 4:
        tar_mask = 1 - torch.all(tar == 0, dim=-1).float()
        tar_mask_cont = tar_mask
    elif tar.size(-1) == 5:
        tar_mask_cont = tar[..., 0]
        tar_mask = (tar[..., 0] > 0.5).float()
        tar = tar[..., 1:]
    else:
        raise NotImplementedError("Boxes loss only supports 4 or 5 dimensional boxes")
    loss_mask = F.binary_cross_entropy_with_logits(pred[..., 0], tar_mask_cont, reduction="none").mean()
    mask = tar_mask.bool().reshape(-1)
    src, tgt = pred[..., 1:].reshape(-1, 4), tar.reshape(-1, 4)
    loss_l1 = _masked_mean((src - tgt).abs().sum(-1), mask, per_item=4)
    giou = paired_generalized_box_iou(box_cxcywh_to_xyxy(src), box_cxcywh_to_xyxy(tgt))
    loss_giou = _masked_mean(1 - giou, mask)
    return loss_l1, loss_mask, loss_giou


def haog_loss(extra_preds, metadata):
    """VideoImageLoss._haog_loss (losses.py:138-155): box L1 / BCE / GIoU + contact-state cross entropy over the
    annotated hands (target >= 0)."""
    if extra_preds["pred_bboxes"].is_cuda:  # one launch: values and gradients of the four terms (csrc/head_loss.cu)
        from . import ops
        l1, bce, giou, ce = ops.haog_loss(extra_preds["pred_bboxes"], metadata["haog_bboxes"],
                                          extra_preds["pred_contact_state"], metadata["contact_state"])
        return {"boxes_l1_loss": l1, "boxes_bce_loss": bce, "boxes_giou_loss": giou, "loss_contact_state": ce}
    # CPU tensors (host-side unit tests of the loss definition): the same masked means as plain tensor expressions
    l1, bce, giou = boxes_loss(extra_preds["pred_bboxes"], metadata["haog_bboxes"])
    pred = extra_preds["pred_contact_state"].flatten(0, 2)
    tar = metadata["contact_state"].flatten()
    mask = tar >= 0
    ce = F.cross_entropy(pred, tar.clamp(min=0), reduction="none")
    return {"boxes_l1_loss": l1, "boxes_bce_loss": bce, "boxes_giou_loss": giou,
            "loss_contact_state": _masked_mean(ce, mask)}


class VideoImageLoss(torch.nn.Module):
    """losses.py:100-168: video ranks get the class cross entropy (+ the consistency terms the lambda dict enables,
    see svit_b200.distributed.consistency_loss); image ranks the HAOG losses.  `is_video` replaces the reference's
    rank test (`local_rank not in cfg.IMAGE_TRAIN.GPU_IDS`)."""

    def __init__(self, cfg, lambdas: dict, is_video: bool = True, reduction: str = "mean"):
        super().__init__()
        self.cfg, self._lambda, self._is_vid, self.reduction = cfg, lambdas, is_video, reduction

    def is_vid(self):
        return self._is_vid or (not self.training)

    def forward(self, x, extra_preds, y, metadata):
        from .distributed import consistency_loss
        ret = {}
        if self.is_vid():
            ret["loss_ce"] = F.cross_entropy(x, y, reduction=self.reduction)
            if self.cfg.TRAIN.FORWARD_VIDEO_FRAMES:
                ret.update(consistency_loss(self._lambda, extra_preds, extra_preds["frames_output"]["extra_preds"]))
        else:
            ret.update(haog_loss(extra_preds, metadata))
        if "safety_loss" in extra_preds:
            ret["safety_loss"] = extra_preds["safety_loss"] * 0
        return ret

    def total(self, loss_dict):
        """tools/train_net.py:124: sum of lambda-weighted terms."""
        return sum(self._lambda[k] * v for k, v in loss_dict.items())
