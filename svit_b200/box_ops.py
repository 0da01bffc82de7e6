"""Host-side integer / ordering rules that produce the per-frame object slots (R4 of SURVEY.md 8a).

These run in the data pipeline on tiny [4, 4] tensors, exactly where the reference runs them
(datasets/ssv2_frames.py:474-529, 347-353; utils/box_ops.py:116-130, 140-194); results must be bit-exact,
including the reference's own quirks.  A batched device variant lives in ops.match_haog_device.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

HIGH_COST = 1e8
CONTACT_THRESHOLD = 0.1


def box_xyxy_to_cxcywh(x):
    x0, y0, x1, y1 = x.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0), (y1 - y0)], dim=-1)


def box_cxcywh_to_xyxy(x):
    xc, yc, w, h = x.unbind(-1)
    return torch.stack([xc - 0.5 * w, yc - 0.5 * h, xc + 0.5 * w, yc + 0.5 * h], dim=-1)


def assign_slots(labels: Sequence[Tuple[str, Sequence[float]]], num_boxes: int = 4) -> torch.Tensor:
    """Annotation order -> slots: 'hand' fills 0,1; every other category fills 2,3; at most two each,
    later ones dropped (ssv2_frames.py:503-517).  Returns [1, num_boxes, 4] xyxy float32."""
    out = torch.zeros((1, num_boxes, 4), dtype=torch.float32)
    filled = {"hand": 0, "obj": 0}
    base = {"hand": 0, "obj": 2}
    for category, box in labels:
        kind = "hand" if category == "hand" else "obj"
        if filled[kind] > 1:
            continue
        out[0, base[kind] + filled[kind]] = torch.as_tensor(list(box), dtype=torch.float32)
        filled[kind] += 1
    return out


def match_haog(haog: torch.Tensor, format: str = "xyxy"):
    """Pair hands with objects by distance and derive the contact state (utils/box_ops.py:140-194).

    The distance is taken between the first two coordinates of each box as stored (the reference overwrites
    its cxcywh conversion at :165), all-zero boxes cost 1e8 (both masks index cost COLUMNS, :168-169), and
    the crossed pairing re-orders the boxes to (0, 2, 3, 1) (:176-178)."""
    if format not in ("xyxy", "cxcywh"):
        raise NotImplementedError(format)
    squeeze = haog.ndim == 3
    if squeeze:
        assert haog.size(0) == 1, haog.size()
        haog = haog.squeeze(0)
    corners = haog.unsqueeze(0)[..., :2]
    cost = torch.cdist(corners[:, :2], corners[:, 2:], p=2).squeeze(0)
    cost[:, torch.all(haog[2:] == 0, dim=-1)] = HIGH_COST
    cost[:, torch.all(haog[:2] == 0, dim=-1)] = HIGH_COST
    straight = cost[0, 0] + cost[1, 1]
    crossed = cost[0, 1] + cost[1, 0]
    if crossed < straight:
        haog = torch.stack((haog[0], haog[2], haog[3], haog[1]), dim=0)
        dists = (cost[0, 1], cost[1, 0])
    else:
        dists = (cost[0, 0], cost[1, 1])
    state = [-1 if d == HIGH_COST else (3 if d < CONTACT_THRESHOLD else 0) for d in dists]
    if squeeze:
        haog = haog.unsqueeze(0)
    return haog, torch.tensor(state, dtype=torch.int64)


def zero_empty_boxes(boxes: torch.Tensor, mode: str = "cxcywh", eps: float = 0.05) -> torch.Tensor:
    """In place: zero every box with w <= eps or h <= eps (utils/box_ops.py:116-130)."""
    shape = boxes.shape
    flat = boxes.reshape(-1, 4)
    if mode == "xyxy":
        wh = flat[..., [2, 3]] - flat[..., [0, 1]]
    elif mode == "cxcywh":
        wh = flat[..., -2:]
    else:
        raise NotImplementedError(mode)
    assert torch.all(wh >= 0)
    flat[torch.any(wh <= eps, dim=-1)] = 0
    return flat.reshape(shape)


def normalise_boxes(boxes_xyxy: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """Crop-normalise, clip to [0,1], convert to cxcywh and zero empty boxes (ssv2_frames.py:347-353)."""
    b = boxes_xyxy.clone()
    b[..., [0, 2]] = b[..., [0, 2]] / w
    b[..., [1, 3]] = b[..., [1, 3]] / h
    return zero_empty_boxes(box_xyxy_to_cxcywh(b.clamp(0, 1)), mode="cxcywh")


def frame_to_slice(t: int, num_frames_in: int, patch_stride_t: int) -> int:
    """Input frame -> temporal slice of the patch grid used by the per-frame RoIAlign."""
    return t if num_frames_in == 1 else t // patch_stride_t


def object_token_index(thw: Sequence[int], t: int, o: int, O: int = 4) -> int:
    """Sequence index of object token (frame t, slot o): 1 + T'H'W' + t*O + o (video_model_builder.py:354-363)."""
    return 1 + thw[0] * thw[1] * thw[2] + t * O + o
