"""GPU tests of SURVEY 8f N2: the one-launch HAOG loss kernel (svit_haog_loss) against fixtures produced by the UNMODIFIED
reference (models/losses.py:50-168; tests/golden/losses.pt: values and gradients, the 'no valid target' branches, the
5-column soft-mask targets) and the one-launch head kernel (svit_head_fwd) against the autograd path of the same module."""
import os

import pytest
import torch

import svit_b200
from svit_b200 import losses, ops
from svit_b200.config import tiny_cfg
from tests.conftest import ROOT, max_rel_err
from tests.golden.recipe import synth_input

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "losses.pt"), weights_only=False)


def _close(a, b, tol=5e-6):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


def test_haog_loss_kernel_boxes_terms_match_reference():
    for i, c in enumerate(_golden()["boxes"]):
        p = c["pred"].to(DEV).requires_grad_(True)
        B_, T_, O_, _ = c["pred"].shape
        contact = torch.zeros(B_, T_, 2, 5, device=DEV)
        ctar = torch.full((B_, T_ * 2), -1, dtype=torch.int64, device=DEV)
        l1, bce, giou, ce = ops.haog_loss(p, c["tar"].to(DEV), contact, ctar)
        assert _close(l1, c["l1"]) and _close(bce, c["bce"]) and _close(giou, c["giou"]), (i, l1, c["l1"], giou, c["giou"])
        assert float(ce) == 0.0
        (l1 + 2 * bce + 3 * giou).backward()
        assert _close(p.grad, c["dpred"], 1e-5), i


def test_haog_loss_kernel_matches_reference_haog_loss():
    for i, c in enumerate(_golden()["haog"]):
        a = c["pred_bboxes"].to(DEV).requires_grad_(True)
        b = c["pred_contact_state"].to(DEV).requires_grad_(True)
        out = losses.haog_loss({"pred_bboxes": a, "pred_contact_state": b},
                               {"haog_bboxes": c["haog_bboxes"].to(DEV), "contact_state": c["contact_state"].to(DEV)})
        assert set(out) == set(c["out"])
        for k in out:
            assert _close(out[k], c["out"][k]), (i, k, out[k], c["out"][k])
        sum(out.values()).backward()
        assert _close(a.grad, c["dboxes"], 1e-5), i
        assert _close(b.grad, c["dcontact"], 1e-5), i


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_head_kernel_matches_autograd_path(dtype):
    """no_grad -> svit_head_fwd (one launch); grad mode -> four linears + torch glue.  Same module, same weights."""
    cfg = tiny_cfg()
    torch.manual_seed(1)
    m = svit_b200.SViT(cfg, compute_dtype=torch.float32).to(DEV)
    for p in m.head.parameters():
        p.data.normal_(0, 0.3)
    x = synth_input("head.x", (3, 1 + 4 * 4, m.norm.weight.numel()), 3).to(DEV, dtype)
    for training in (False, True):
        m.head.train(training)
        if training:
            m.head.dropout.p = 0.0
        with torch.enable_grad():
            want_out, want = m.head(x.clone().requires_grad_(True), T=4)
        with torch.no_grad():
            got_out, got = m.head(x, T=4)
        tol = 2e-5
        assert max_rel_err(got_out.float().cpu(), want_out.detach().float().cpu()) < tol
        for k in ("logits", "obj_desc", "pred_bboxes", "pred_contact_state"):
            assert got[k].shape == want[k].shape, k
            assert max_rel_err(got[k].float().cpu(), want[k].detach().float().cpu()) < tol, (training, k)
