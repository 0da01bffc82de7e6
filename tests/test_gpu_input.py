"""GPU tests of SURVEY 8f N4 (input side on the device): crop + flip + normalise of uint8 frames and the box transform that
follows them, against fixtures produced by the reference's own functions (tests/golden/aug.pt: datasets/utils.py
tensor_normalize, datasets/transform.py random_crop / crop_clip_boxes / horizontal_flip, ssv2_frames.py:347-353,
utils/box_ops.py).  fp32 outputs and boxes are bit-exact; the bf16 clip is the rounding of the fp32 one."""
import os

import pytest
import torch

from svit_b200 import ops
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cases():
    return torch.load(os.path.join(ROOT, "tests", "golden", "aug.pt"), weights_only=False)


def test_crop_flip_normalize_bit_exact():
    for i, c in enumerate(_cases()):
        frames = c["frames"].unsqueeze(0).to(DEV)                    # [1, T, H, W, 3]
        got = ops.crop_flip_normalize_u8(frames, c["x_off"], c["y_off"], int(c["flip"]), c["crop"], c["mean"], c["std"],
                                         torch.float32)
        assert got.shape[1:] == c["clip"].shape
        assert torch.equal(got[0].cpu(), c["clip"]), i
        got16 = ops.crop_flip_normalize_u8(frames, c["x_off"], c["y_off"], int(c["flip"]), c["crop"], c["mean"], c["std"],
                                           torch.bfloat16)
        assert torch.equal(got16[0].float().cpu(), c["clip"].bfloat16().float()), i


def test_crop_flip_is_per_sample():
    cs = _cases()
    a, b = cs[0], cs[0]
    frames = torch.stack([a["frames"], b["frames"]]).to(DEV)
    got = ops.crop_flip_normalize_u8(frames, [a["x_off"], 0], [a["y_off"], 1], [1, 0], a["crop"], a["mean"], a["std"],
                                     torch.float32)
    assert torch.equal(got[0].cpu(), a["clip"])
    plain = ops.normalize_u8(frames[1:2].contiguous(), a["mean"], a["std"], torch.float32)   # [1, 3, T, H, W]
    assert torch.equal(got[1].cpu(), plain[0, :, :, 1:1 + a["crop"], 0:a["crop"]].cpu())
    with pytest.raises(ValueError):
        ops.crop_flip_normalize_u8(frames, 100, 0, None, a["crop"], a["mean"], a["std"])


def test_boxes_follow_the_frames_bit_exact():
    for i, c in enumerate(_cases()):
        boxes = c["boxes"].unsqueeze(0).to(DEV)                      # [1, 12, 4] xyxy pixels
        got = ops.boxes_crop_flip(boxes, c["x_off"], c["y_off"], int(c["flip"]), c["crop"], eps=0.05)
        assert torch.equal(got[0].cpu(), c["boxes_out"]), (i, got[0].cpu(), c["boxes_out"])
