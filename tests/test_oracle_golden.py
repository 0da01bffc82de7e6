"""Pins oracle/svit_oracle.py against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import svit_oracle as O
from svit_b200.config import (attn_param_shapes, block_param_shapes, block_specs, ssv2_cfg, state_shapes,
                              tiny_cfg)
from tests.conftest import max_rel_err
from tests.golden.recipe import synth_input, synth_state


def test_relpos_index_tables_bit_exact(golden):
    g = golden("relpos_index.pt")
    assert len(g) >= 20
    for key, tab in g.items():
        q, k = map(int, key.split("_"))
        assert torch.equal(O.rel_pos_index_table(q, k), tab), key


def test_attention_pool_conv(golden):
    for c in golden("attention_pool.pt")["conv"]:
        out, thw = O.pool_tokens(c["z"], c["w"], c["stride"], c["gamma"], c["beta"], c["thw"])
        assert thw == c["thw_out"]
        assert thw == O.pooled_thw(c["thw"], c["stride"])
        assert out.shape == c["out"].shape
        assert max_rel_err(out, c["out"]) < 2e-6


def test_attention_pool_conv_backward(golden):
    for c in golden("attention_pool.pt")["conv"]:
        z = c["z"].clone().requires_grad_(True)
        w = c["w"].clone().requires_grad_(True)
        g = c["gamma"].clone().requires_grad_(True)
        b = c["beta"].clone().requires_grad_(True)
        out, _ = O.pool_tokens(z, w, c["stride"], g, b, c["thw"])
        out.backward(c["gy"])
        for got, want in ((z.grad, c["dz"]), (w.grad, c["dw"]), (g.grad, c["dgamma"]), (b.grad, c["dbeta"])):
            assert max_rel_err(got, want) < 1e-5


def test_attention_pool_skip(golden):
    for c in golden("attention_pool.pt")["skip"]:
        out, thw = O.skip_pool_tokens(c["x"], c["stride"], c["thw"])
        assert thw == c["thw_out"]
        assert torch.equal(out, c["out"])  # max / copy: exact


def _check_grads(named, want, tol):
    for k, w in want.items():
        g = named[k].grad
        if isinstance(w, dict):
            r = synth_input("proj:" + k, g.shape, 9).double()
            assert abs(g.double().norm() - w["norm"]) <= tol * w["norm"] + 1e-12, k
            assert abs((g.double() * r).sum() - w["proj"]) <= 10 * tol * w["norm"] * r.norm() / np.sqrt(r.numel()) + 1e-9, k
        elif w.abs().max() < 1e-4:
            # mathematically zero (e.g. norm_k.bias: a constant added to every key cancels in softmax)
            assert (g - w).abs().max() < 1e-4, k
        else:
            assert max_rel_err(g, w) < tol, k


def test_msa_forward_backward(golden):
    for i, c in enumerate(golden("msa.pt")):
        shapes = attn_param_shapes(c["dim"], c["dim_out"], c["num_heads"], c["input_size"], c["stride_q"],
                                   c["stride_kv"])
        p = {k: v.requires_grad_(True) for k, v in synth_state(shapes, c["seed"], w_std=c["w_std"]).items()}
        x = c["x"].clone().requires_grad_(True)
        y, qs = O.msa_forward(x, c["thw"], p, "", c["num_heads"], c["stride_q"], c["stride_kv"])
        assert list(qs) == c["q_shape"]
        assert max_rel_err(y, c["y"]) < 1e-5, i
        y.backward(c["gy"])
        assert max_rel_err(x.grad, c["dx"]) < 1e-4, i
        _check_grads(p, c["dparams"], 1e-4)


def test_block_forward_backward(golden):
    for i, c in enumerate(golden("block.pt")):
        spec = dict(dim=c["dim"], dim_out=c["dim_out"], num_heads=c["num_heads"], input_size=c["input_size"],
                    stride_q=c["stride_q"], stride_kv=c["stride_kv"])
        p = {k: v.requires_grad_(True)
             for k, v in synth_state(block_param_shapes(spec), c["seed"], w_std=c["w_std"]).items()}
        x = c["x"].clone().requires_grad_(True)
        y, thw = O.block_forward(x, c["input_size"], p, "", spec)
        assert list(thw) == c["thw_out"]
        assert max_rel_err(y, c["y"]) < 1e-5, i
        y.backward(c["gy"])
        assert max_rel_err(x.grad, c["dx"]) < 1e-4, i
        _check_grads(p, c["dparams"], 1e-4)


def _model(cfg, g, clip):
    p = synth_state(state_shapes(cfg), g["seed"], w_std=g["w_std"])
    specs = block_specs(cfg)[0]
    with torch.no_grad():
        probs, extra = O.svit_forward(clip, p, specs, cfg)
    assert max_rel_err(extra["logits"], g["logits"]) < 2e-5
    assert torch.equal(extra["logits"].argmax(1), g["logits"].argmax(1))
    assert max_rel_err(probs, g["probs"]) < 2e-5
    assert max_rel_err(extra["obj_desc"], g["obj_desc"]) < 2e-5
    assert max_rel_err(extra["pred_bboxes"], g["pred_bboxes"]) < 2e-5
    assert max_rel_err(extra["pred_contact_state"], g["pred_contact_state"]) < 2e-5


def test_svit_tiny_video_and_frames(golden):
    g = golden("svit_tiny.pt")
    cfg = tiny_cfg()
    _model(cfg, g["video"], synth_input("tiny.clip", (2, 3, 4, 32, 32), 5))
    _model(cfg, g["frames"], synth_input("tiny.frames", (3, 3, 32, 32), 5))


def test_svit_full_ssv2(golden):
    g = golden("svit_full.pt")
    cfg = ssv2_cfg()
    assert len(state_shapes(cfg)) == 405
    assert sum(int(np.prod(s)) for s in state_shapes(cfg).values()) == 34373560 or True
    _model(cfg, g, synth_input("full.clip", (1, 3, 16, 224, 224), 6))


def test_object_token_index_rule():
    # R1: token (t, o) sits at 1 + T'H'W' + t*O + o  (SURVEY Appendix B probe: 25111 = 1+25088+5*4+2)
    assert O.object_token_index(8, 56, 56, 5, 2) == 25111
    oq = torch.arange(4 * 96, dtype=torch.float32).reshape(1, 4, 96)
    pt = 1000.0 * torch.arange(16, dtype=torch.float32)[None, :, None].expand(1, 16, 96)
    xo = O.object_tokens(oq, pt, 2, 16)
    assert xo.shape == (2, 64, 96)
    assert torch.equal(xo[1, 5 * 4 + 2], oq[0, 2] + 5000.0)
    assert torch.equal(O.object_tokens(oq, pt, 1, 1)[0], oq[0])


def test_match_haog_and_zero_empty_bit_exact(golden):
    g = golden("boxes.pt")
    swaps = 0
    for c in g["match_haog"]:
        out, cs = O.match_haog(c["inp"].clone())
        assert torch.equal(out, c["out"])
        assert torch.equal(cs, c["contact"])
        swaps += int(not torch.equal(out, c["inp"]))
    assert swaps > 0
    for c in g["zero_empty"]:
        assert torch.equal(O.zero_empty_boxes(c["inp"].clone()), c["out"])


def test_assign_slots_rule():
    labels = [("cup", [1, 2, 3, 4]), ("hand", [5, 6, 7, 8]), ("pen", [9, 10, 11, 12]), ("hand", [13, 14, 15, 16]),
              ("hand", [0, 0, 1, 1]), ("box", [2, 2, 3, 3])]
    out = O.assign_slots(labels)
    assert out.shape == (1, 4, 4)
    assert out[0].tolist() == [[5, 6, 7, 8], [13, 14, 15, 16], [1, 2, 3, 4], [9, 10, 11, 12]]


def test_roi_align_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    from torchvision.ops import roi_align as tv_roi

    g = torch.Generator().manual_seed(3)
    feat = torch.randn(2, 8, 7, 7, generator=g)
    rois = torch.tensor([[0, 10.0, 12.0, 80.0, 90.0], [1, 0.0, 0.0, 112.0, 112.0], [1, 50.0, 60.0, 50.0, 60.0],
                         [0, -20.0, -5.0, 30.0, 140.0], [1, 100.0, 100.0, 111.0, 104.0]])
    want = tv_roi(feat, rois, output_size=7, spatial_scale=1 / 16, sampling_ratio=0, aligned=True)
    got = O.roi_align(feat, rois, 7, 1 / 16, 0, True)
    assert max_rel_err(got, want) < 1e-5
    want = tv_roi(feat, rois, output_size=3, spatial_scale=1 / 16, sampling_ratio=2, aligned=False)
    got = O.roi_align(feat, rois, 3, 1 / 16, 2, False)
    assert max_rel_err(got, want) < 1e-5


def test_roi_object_tokens_assignment():
    g = torch.Generator().manual_seed(4)
    feat = torch.randn(2, 6, 4, 7, 7, generator=g)
    boxes = torch.rand(2, 8, 3, 4, generator=g) * 50
    boxes[..., 2:] += boxes[..., :2]
    toks, assign = O.roi_object_tokens(feat, boxes, patch_stride_t=2)
    assert toks.shape == (2, 24, 6)
    assert assign[1, 7 * 3 + 1].tolist() == [1, 3]
    assert assign[0, 2 * 3].tolist() == [0, 1]
