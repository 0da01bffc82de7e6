"""GPU parity tests (run with -m gpu on the B200 box).  Every compute call goes through the C ABI
(libsvit_sm100.so) via svit_b200; results are compared with the CPU oracle and with the golden
fixtures generated from the unmodified reference.

Tolerances (north_star): fp32 mode max-rel-err <= 1e-4 per block (max|y-ref| / max|ref|);
bf16 mode <= 2e-2 on logits with top-1 agreement; integer outputs bit-exact.
"""
import numpy as np
import pytest
import torch
import torch.nn as nn

import svit_b200
from oracle import svit_oracle as O
from svit_b200 import _lib, msa, ops
from svit_b200.config import attn_param_shapes, block_param_shapes, block_specs, ssv2_cfg, state_shapes, tiny_cfg
from tests.conftest import max_rel_err
from tests.golden.recipe import synth_input, synth_state

pytestmark = pytest.mark.gpu
DEV = "cuda"
FP32_TOL = 1e-4
LN = lambda d: nn.LayerNorm(d, eps=1e-6)


@pytest.fixture(autouse=True, scope="module")
def _need_cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    assert _lib.lib().svit_abi_version() == 100
    yield
    torch.cuda.synchronize()


def cpu(t):
    return t.detach().float().cpu()


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(dtype):
    # the last two sizes take several grid-stride iterations of the two-rows-per-thread bf16 kernel
    for rows, C in ((37, 96), (130, 192), (65, 384), (9, 768), (20011, 384), (160001, 96)):
        x = synth_input(f"ln{C}", (rows, C), 1) * 2 + 0.5
        g = 1 + 0.3 * synth_input(f"lng{C}", (C,), 1)
        b = 0.3 * synth_input(f"lnb{C}", (C,), 1)
        gy = synth_input(f"lngy{C}", (rows, C), 1)
        xr = x.to(dtype).float().clone().requires_grad_(True)
        gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
        yr = torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-6)
        yr.backward(gy.to(dtype).float())
        xd = x.to(DEV, dtype).requires_grad_(True)
        gd, bd = g.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
        yd = ops.layer_norm(xd, gd, bd)
        yd.backward(gy.to(DEV, dtype))
        tol = 1e-5 if dtype == torch.float32 else 1e-2
        assert max_rel_err(cpu(yd), yr) < tol
        assert max_rel_err(cpu(xd.grad), xr.grad) < (1e-4 if dtype == torch.float32 else 2e-2)
        assert max_rel_err(cpu(gd.grad), gr.grad) < (1e-4 if dtype == torch.float32 else 2e-2)
        assert max_rel_err(cpu(bd.grad), br.grad) < (1e-4 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_scale_rows_exact(dtype):
    """y[m, :] = x[m, :] * scale[m // rows_per_sample] (DropPath backward, common.py:46-59): equal to the fp32 product
    rounded once to the activation dtype, for the 16-byte bf16 kernel (C % 8 == 0) and the scalar one (C = 100)."""
    g = torch.Generator().manual_seed(3)
    for B, N, C in ((3, 37, 96), (2, 1633, 384), (5, 11, 100)):
        x = torch.randn(B, N, C, generator=g).to(dtype).to(DEV)
        s = (torch.rand(B, generator=g) * 2).to(DEV)
        got = ops._scale_rows(x, s, N)
        want = (x.float() * s.view(B, 1, 1)).to(dtype)
        assert torch.equal(got, want), (B, N, C)


def _gemm_ref(A, B, tA, tB):
    a = A.t() if tA else A
    b = B.t() if tB else B
    return a.double() @ b.double()


@pytest.mark.parametrize("dtype,impl", [(torch.float32, ops.IMPL_SIMT), (torch.bfloat16, ops.IMPL_SIMT)])
def test_gemm_epilogues_and_transposes(dtype, impl):
    gen = torch.Generator().manual_seed(5)
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    for (M, N, K) in ((70, 96, 96), (257, 288, 192), (33, 40, 100)):
        for tA, tB in ((0, 1), (0, 0), (1, 0), (1, 1)):
            A = torch.randn((K, M) if tA else (M, K), generator=gen).to(dtype)
            B = torch.randn((N, K) if tB else (K, N), generator=gen).to(dtype)
            out = torch.empty(M, N, dtype=dtype, device=DEV)
            ops.gemm(A.to(DEV), B.to(DEV), out, M, N, K, A.shape[1], B.shape[1], N, tA, tB, impl=impl)
            assert max_rel_err(cpu(out), _gemm_ref(A.float(), B.float(), tA, tB)) < tol, (M, N, K, tA, tB)
    # fused epilogue: bias + gelu + pre_out, then gelu' * residual * sample scale + row remap
    M, N, K = 96, 192, 96
    A = torch.randn(M, K, generator=gen).to(dtype)
    W = (torch.randn(N, K, generator=gen) * 0.2).to(dtype)
    bias = torch.randn(N, generator=gen)
    res = torch.randn(2, 60, N, generator=gen).to(dtype)
    scale = torch.tensor([0.0, 1.25])
    out = res.clone().to(DEV)
    pre = torch.empty(M, N, dtype=dtype, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), out, M, N, K, K, K, N, 0, 1, bias=bias.to(DEV), residual=res.to(DEV), ldr=N,
             sample_scale=scale.to(DEV), rows_per_sample=48, act=1, pre_out=pre, ldp=N, remap=(48, 60, 5), impl=impl)
    z = A.float() @ W.float().t() + bias
    ref = res.float().clone()
    ref[:, 5:53] += (torch.nn.functional.gelu(z).reshape(2, 48, N) * scale[:, None, None])
    assert max_rel_err(cpu(pre), z) < tol
    assert max_rel_err(cpu(out), ref) < tol
    gout = torch.empty(M, N, dtype=dtype, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), gout, M, N, K, K, K, N, 0, 1, gelu_pre=pre, ldg=N, impl=impl)
    zz = cpu(pre).double().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    assert max_rel_err(cpu(gout), (A.float() @ W.float().t()).double() * zz.grad) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_pool_conv_golden(golden, dtype):
    for c in golden("attention_pool.pt")["conv"]:
        conv = nn.Conv3d(96, 96, (3, 3, 3), stride=tuple(c["stride"]), padding=(1, 1, 1), groups=96, bias=False).to(DEV)
        norm = LN(96).to(DEV)
        conv.weight.data.copy_(c["w"]); norm.weight.data.copy_(c["gamma"]); norm.bias.data.copy_(c["beta"])
        z = c["z"].to(DEV, dtype).requires_grad_(True)
        out, thw = msa.attention_pool(z, conv, c["thw"], has_cls_embed=True, norm=norm)
        assert list(thw) == c["thw_out"]
        assert out.shape == c["out"].shape
        if dtype == torch.float32:
            assert max_rel_err(cpu(out), c["out"]) < 1e-5
        else:
            assert max_rel_err(cpu(out), c["out"]) < 3e-2
        out.backward(c["gy"].to(DEV, dtype))
        tol = 1e-4 if dtype == torch.float32 else 5e-2
        # attention_pool() feeds the kernel a packed copy, so dz arrives through torch's cat/permute backward
        assert max_rel_err(cpu(z.grad), c["dz"]) < tol
        assert max_rel_err(cpu(conv.weight.grad), c["dw"]) < tol
        assert max_rel_err(cpu(norm.weight.grad), c["dgamma"]) < tol
        assert max_rel_err(cpu(norm.bias.grad), c["dbeta"]) < tol


@pytest.mark.parametrize("spec", [  # B, h, T, H, W, O, (sq, skv): multi-tile grids, ragged tile edges, T = 1
    (2, 2, 3, 20, 30, 8, (1, 2)), (1, 1, 8, 56, 56, 64, (1, 2)), (2, 4, 8, 14, 14, 64, (1, 2)), (1, 2, 1, 17, 23, 4, (2, 1)),
    (1, 8, 8, 7, 7, 64, (1, 1)), (1, 2, 5, 29, 15, 3, (2, 2)), (1, 1, 2, 28, 28, 64, (2, 4)), (1, 1, 4, 39, 39, 16, (2, 8)),
])
def test_pool_ln_tiled_bf16_vs_oracle(spec):
    """The smem-tiled bf16 pooling kernels (stride 1 and 2) and the direct kernel (stride >= 4) on the packed
    [B, N, 3, h, 96] layout, against the oracle's attention_pool restatement on the same bf16-rounded input."""
    B, h, T, H, W, Ot, (sq, skv) = spec
    g = torch.Generator().manual_seed(H * 100 + W)
    N = 1 + T * H * W + Ot
    qkv = torch.randn(B, N, 3 * h * 96, generator=g).bfloat16()
    ws = [torch.randn(96, 1, 3, 3, 3, generator=g) * 0.3 for _ in range(3)]
    gs = [1 + 0.2 * torch.randn(96, generator=g) for _ in range(3)]
    bs = [0.2 * torch.randn(96, generator=g) for _ in range(3)]
    dev = lambda t: t.to(DEV)
    got = ops.qkv_pool(dev(qkv), (T, H, W), Ot, sq, skv, dev(ws[0]), (dev(gs[0]), dev(bs[0])), dev(ws[1]),
                       (dev(gs[1]), dev(bs[1])), dev(ws[2]), (dev(gs[2]), dev(bs[2])))
    z = qkv.float().reshape(B, N, 3, h, 96).permute(2, 0, 3, 1, 4)
    for which, s in enumerate((sq, skv, skv)):
        want, thw2 = O.pool_tokens(z[which], ws[which], (1, s, s), gs[which], bs[which], [T, H, W])
        assert got[which].shape == want.shape
        err = (cpu(got[which]) - want).abs().max().item()
        assert err < 4e-2, f"which={which} s={s}: max abs err {err}"  # LN output is O(1); bf16 rounding of out ~ 2e-2 at |y|~4
        assert max_rel_err(cpu(got[which]), want) < 1e-2


@pytest.mark.parametrize("spec", [  # B, h, T, H, W, O, (sq, skv): strip form of the weight-gradient kernel (Wo % 7 == 0,
    # strides 1 / 2 / 4) and the token form (everything else)
    (2, 2, 4, 14, 14, 8, (1, 2)), (1, 1, 3, 28, 28, 4, (2, 4)), (1, 2, 2, 56, 56, 8, (1, 4)), (1, 4, 8, 7, 7, 64, (1, 1)),
    (1, 2, 3, 20, 30, 8, (1, 2)), (1, 1, 2, 28, 28, 4, (1, 8)),
])
def test_pool_ln_bf16_backward_vs_oracle(spec):
    """Backward of the bf16 conv-pool + LayerNorm (attention.py:13-65) on the packed qkv layout: gradients of the three
    depthwise conv weights, the LayerNorm parameters and the qkv input against the oracle's autograd in fp32 on the same
    bf16-rounded values; ||dg - dg_ref|| / ||dg_ref|| <= 2e-2 (dpre and dqkv are rounded to bf16 on the device)."""
    B, h, T, H, W, Ot, (sq, skv) = spec
    g = torch.Generator().manual_seed(H * 100 + W + sq)
    N = 1 + T * H * W + Ot
    qkv = torch.randn(B, N, 3 * h * 96, generator=g).bfloat16()
    ws = [torch.randn(96, 1, 3, 3, 3, generator=g) * 0.3 for _ in range(3)]
    gs = [1 + 0.2 * torch.randn(96, generator=g) for _ in range(3)]
    bs = [0.2 * torch.randn(96, generator=g) for _ in range(3)]
    leaf = lambda t: t.to(DEV).requires_grad_(True)
    qd = leaf(qkv)
    wd, gd, bd = [leaf(w) for w in ws], [leaf(x) for x in gs], [leaf(x) for x in bs]
    got = ops.qkv_pool(qd, (T, H, W), Ot, sq, skv, wd[0], (gd[0], bd[0]), wd[1], (gd[1], bd[1]), wd[2], (gd[2], bd[2]))
    qr = qkv.float().requires_grad_(True)
    wr, gr, br = [w.clone().requires_grad_(True) for w in ws], [x.clone().requires_grad_(True) for x in gs], \
        [x.clone().requires_grad_(True) for x in bs]
    z = qr.reshape(B, N, 3, h, 96).permute(2, 0, 3, 1, 4)
    loss_d, loss_r = 0, 0
    for which, s in enumerate((sq, skv, skv)):
        want, _ = O.pool_tokens(z[which], wr[which], (1, s, s), gr[which], br[which], [T, H, W])
        gy = torch.randn(want.shape, generator=g).bfloat16()
        loss_r = loss_r + (want * gy.float()).sum()
        loss_d = loss_d + (got[which].float() * gy.to(DEV).float()).sum()
    loss_r.backward()
    loss_d.backward()
    errs = {"dqkv": _norm_rel(cpu(qd.grad), qr.grad)}
    for i, nm in enumerate("qkv"):
        errs[f"dw_{nm}"] = _norm_rel(cpu(wd[i].grad), wr[i].grad)
        errs[f"dgamma_{nm}"] = _norm_rel(cpu(gd[i].grad), gr[i].grad)
        errs[f"dbeta_{nm}"] = _norm_rel(cpu(bd[i].grad), br[i].grad)
    worst = max(errs, key=errs.get)
    assert errs[worst] < 2e-2, errs


def test_attention_pool_skip_golden_exact(golden):
    for c in golden("attention_pool.pt")["skip"]:
        s = c["stride"]
        ks = [x + 1 if x > 1 else x for x in s]
        pool = nn.MaxPool3d(ks, s, [k // 2 for k in ks], ceil_mode=False)
        x = c["x"].to(DEV).requires_grad_(True)
        out, thw = msa.attention_pool(x, pool, c["thw"], has_cls_embed=True)
        assert list(thw) == c["thw_out"]
        assert torch.equal(cpu(out), c["out"])  # max / copy: bit exact
        out.backward(c["gy"].to(DEV))
        assert torch.equal(cpu(x.grad), c["dx"])


def _check_param_grads(mod, want, tol):
    named = dict(mod.named_parameters())
    for k, w in want.items():
        g = cpu(named[k].grad)
        if isinstance(w, dict):
            r = synth_input("proj:" + k, g.shape, 9).double()
            assert abs(g.double().norm() - w["norm"]) <= tol * w["norm"] + 1e-9, k
            assert abs((g.double() * r).sum() - w["proj"]) <= 30 * tol * w["norm"] * r.norm() / np.sqrt(r.numel()) + 1e-6, k
        elif w.abs().max() < 1e-4:
            assert (g - w).abs().max() < 1e-3, k
        else:
            assert max_rel_err(g, w) < tol, k


def _make_msa(c):
    m = svit_b200.MultiScaleAttention(c["dim"], c["dim_out"], input_size=c["input_size"], num_heads=c["num_heads"],
                                      qkv_bias=True, kernel_q=[3, 3, 3], kernel_kv=[3, 3, 3], stride_q=c["stride_q"],
                                      stride_kv=c["stride_kv"], norm_layer=LN, has_cls_embed=True, mode="conv",
                                      pool_first=False, rel_pos_spatial=True, rel_pos_temporal=True,
                                      rel_pos_zero_init=False, residual_pooling=True, separate_qkv=False)
    m.load_state_dict(synth_state({k: v.shape for k, v in m.state_dict().items()}, c["seed"], w_std=c["w_std"]))
    return m.to(DEV)


def test_msa_fp32_golden_forward_backward(golden):
    ops.set_impl(gemm=ops.IMPL_SIMT, attn=ops.IMPL_SIMT)
    for i, c in enumerate(golden("msa.pt")):
        m = _make_msa(c)
        x = c["x"].to(DEV).requires_grad_(True)
        y, qs = m(x, c["thw"])
        assert list(qs) == c["q_shape"]
        assert max_rel_err(cpu(y), c["y"]) < FP32_TOL, i
        y.backward(c["gy"].to(DEV))
        assert max_rel_err(cpu(x.grad), c["dx"]) < 5e-4, i
        _check_param_grads(m, c["dparams"], 5e-4)
    ops.set_impl(gemm=ops.IMPL_AUTO, attn=ops.IMPL_AUTO)


def test_msa_bf16_golden_forward(golden):
    for i, c in enumerate(golden("msa.pt")):
        m = _make_msa(c)
        with torch.no_grad():
            y, qs = m(c["x"].to(DEV, torch.bfloat16), c["thw"])
        assert list(qs) == c["q_shape"]
        err = max_rel_err(cpu(y), c["y"])
        print(f"msa golden case {i}: bf16 max-rel-err {err:.2e}")
        assert err < 2e-2, i  # north_star bf16 tolerance


def _make_block(c):
    m = svit_b200.MultiScaleBlock(dim=c["dim"], dim_out=c["dim_out"], num_heads=c["num_heads"], input_size=c["input_size"],
                                  mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0, drop_path=0.0, norm_layer=LN,
                                  kernel_q=[3, 3, 3], kernel_kv=[3, 3, 3], stride_q=c["stride_q"], stride_kv=c["stride_kv"],
                                  mode="conv", has_cls_embed=True, pool_first=False, rel_pos_spatial=True,
                                  rel_pos_temporal=True, rel_pos_zero_init=False, residual_pooling=True,
                                  dim_mul_in_att=True, separate_qkv=False)
    m.load_state_dict(synth_state({k: v.shape for k, v in m.state_dict().items()}, c["seed"], w_std=c["w_std"]))
    return m.to(DEV)


def test_block_fp32_golden_forward_backward(golden):
    ops.set_impl(gemm=ops.IMPL_SIMT, attn=ops.IMPL_SIMT)
    for i, c in enumerate(golden("block.pt")):
        m = _make_block(c)
        x = c["x"].to(DEV).requires_grad_(True)
        y, thw = m(x, c["input_size"])
        assert list(thw) == c["thw_out"]
        assert max_rel_err(cpu(y), c["y"]) < FP32_TOL, i
        y.backward(c["gy"].to(DEV))
        assert max_rel_err(cpu(x.grad), c["dx"]) < 5e-4, i
        _check_param_grads(m, c["dparams"], 5e-4)
    ops.set_impl(gemm=ops.IMPL_AUTO, attn=ops.IMPL_AUTO)


def test_block_bf16_golden_forward(golden):
    for i, c in enumerate(golden("block.pt")):
        m = _make_block(c)
        with torch.no_grad():
            y, _ = m(c["x"].to(DEV, torch.bfloat16), c["input_size"])
        err = max_rel_err(cpu(y), c["y"])
        print(f"block golden case {i}: bf16 max-rel-err {err:.2e}")
        assert err < 2e-2, i  # north_star bf16 tolerance


def test_block_folded_layernorm_matches_unfolded():
    """Inference blocks fold norm1 / norm2 into the consuming GEMMs (MultiScaleBlock._forward_folded); the result must
    agree with the LayerNorm-kernel path to bf16 rounding at the real stage widths, including a dim-changing block."""
    from svit_b200.msa import MultiScaleBlock
    torch.manual_seed(5)
    kw = dict(qkv_bias=True, kernel_q=(3, 3, 3), kernel_kv=(3, 3, 3), rel_pos_spatial=True, rel_pos_temporal=True,
              residual_pooling=True, dim_mul_in_att=True)
    for dim, dim_out, heads, size, sq, skv in ((96, 96, 1, (4, 16, 16), (1, 1, 1), (1, 4, 4)),
                                               (96, 192, 2, (4, 16, 16), (1, 2, 2), (1, 2, 2)),
                                               (384, 384, 4, (2, 14, 14), (1, 1, 1), (1, 2, 2))):
        m = MultiScaleBlock(dim, dim_out, heads, list(size), stride_q=sq, stride_kv=skv, **kw).to(DEV)
        for n, p in m.named_parameters():
            if n.endswith("norm1.weight") or n.endswith("norm2.weight"):
                p.data.add_(0.2 * torch.randn_like(p))
            if n.endswith("norm1.bias") or n.endswith("norm2.bias"):
                p.data.add_(0.2 * torch.randn_like(p))
        x = (torch.randn(2, 1 + size[0] * size[1] * size[2] + 8, dim, device=DEV) + 0.3).to(torch.bfloat16)
        was = ops._LN_FOLD["enabled"]
        try:
            with torch.no_grad():
                ops._LN_FOLD["enabled"] = True
                assert ops.ln_fold_applicable(x, m.attn.qkv.weight, m.mlp.fc1.weight)
                n0 = ops.launches()
                y1, thw1 = m(x, list(size))
                folded_calls = ops.launches() - n0
                ops._LN_FOLD["enabled"] = False
                n0 = ops.launches()
                y0, thw0 = m(x, list(size))
                plain_calls = ops.launches() - n0
        finally:
            ops._LN_FOLD["enabled"] = was
        assert thw0 == thw1 and folded_calls > 0 and plain_calls > 0
        err = max_rel_err(cpu(y1), cpu(y0))
        print(f"folded vs unfolded block dim {dim}->{dim_out}: {err:.2e}")
        assert err < 1.5e-2


# ------------------------------------------------------------------------------------------------ models
def _model(cfg, g, dtype):
    m = svit_b200.SViT(cfg, compute_dtype=dtype)
    m.load_state_dict(synth_state(state_shapes(cfg), g["seed"], w_std=g["w_std"]))
    return m.to(DEV)


def _check_model(m, g, clip, tol, top1=True):
    m.eval()
    with torch.no_grad():
        probs, extra = m([clip.to(DEV)])
    assert max_rel_err(cpu(extra["logits"]), g["logits"]) < tol
    if top1:
        assert torch.equal(cpu(extra["logits"]).argmax(1), g["logits"].argmax(1))
    assert max_rel_err(cpu(probs), g["probs"]) < 5 * tol
    assert max_rel_err(cpu(extra["obj_desc"]), g["obj_desc"]) < tol
    assert max_rel_err(cpu(extra["pred_bboxes"]), g["pred_bboxes"]) < tol
    assert max_rel_err(cpu(extra["pred_contact_state"]), g["pred_contact_state"]) < tol
    assert extra["obj_desc"].shape == g["obj_desc"].shape


def test_svit_tiny_fp32(golden):
    ops.set_impl(gemm=ops.IMPL_SIMT, attn=ops.IMPL_SIMT)
    g = golden("svit_tiny.pt")
    cfg = tiny_cfg()
    m = _model(cfg, g["video"], torch.float32)
    _check_model(m, g["video"], synth_input("tiny.clip", (2, 3, 4, 32, 32), 5), 2e-4)
    _check_model(m, g["frames"], synth_input("tiny.frames", (3, 3, 32, 32), 5), 2e-4)
    ops.set_impl(gemm=ops.IMPL_AUTO, attn=ops.IMPL_AUTO)


def test_svit_tiny_fp32_training_gradients(golden):
    ops.set_impl(gemm=ops.IMPL_SIMT, attn=ops.IMPL_SIMT)
    g = golden("svit_tiny.pt")["video"]
    cfg = tiny_cfg()
    m = _model(cfg, g, torch.float32)
    m.train()
    for mod in m.modules():
        if isinstance(mod, svit_b200.DropPath):
            mod.drop_prob = 0.0
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    clip = synth_input("tiny.clip", (2, 3, 4, 32, 32), 5).to(DEV)
    logits, extra = m([clip])
    assert max_rel_err(cpu(logits), g["train_logits"]) < 2e-4
    tgt = (torch.arange(2) % logits.shape[1]).to(DEV)
    loss = torch.nn.functional.cross_entropy(logits, tgt) + 0.1 * extra["pred_bboxes"].square().mean() \
        + 0.1 * extra["pred_contact_state"].square().mean()
    assert abs(loss.item() - g["loss"].item()) < 1e-4 * abs(g["loss"].item()) + 1e-6
    loss.backward()
    named = dict(m.named_parameters())
    bad = []
    for k, n in g["grad_norms"].items():
        got = named[k].grad.double().norm().item()
        if abs(got - n.item()) > 2e-3 * n.item() + 1e-7:
            bad.append((k, got, n.item()))
    assert not bad, bad[:10]
    for k, w in g["grads"].items():
        if w.abs().max() < 1e-6:  # mathematically zero (norm_k.bias: a constant key shift cancels in softmax)
            assert cpu(named[k].grad).abs().max() < 1e-5, k
        else:
            assert max_rel_err(cpu(named[k].grad), w) < 2e-3, k
    ops.set_impl(gemm=ops.IMPL_AUTO, attn=ops.IMPL_AUTO)


def test_svit_full_ssv2_fp32_and_bf16(golden):
    g = golden("svit_full.pt")
    cfg = ssv2_cfg()
    clip = synth_input("full.clip", (1, 3, 16, 224, 224), 6)
    ops.set_impl(gemm=ops.IMPL_SIMT, attn=ops.IMPL_SIMT)
    m = _model(cfg, g, torch.float32)
    _check_model(m, g, clip, 3e-4)
    ops.set_impl(gemm=ops.IMPL_AUTO, attn=ops.IMPL_AUTO)
    m.compute_dtype = torch.bfloat16
    m.eval()
    with torch.no_grad():
        probs, extra = m([clip.to(DEV)])
    err = max_rel_err(cpu(extra["logits"]), g["logits"])
    agree = torch.equal(cpu(extra["logits"]).argmax(1), g["logits"].argmax(1))
    print(f"bf16 logits max-rel-err {err:.4f}, top-1 agreement {agree}")
    assert err < 2e-2
    assert agree


def test_blocks_full_size_fp32_vs_oracle():
    """Per-block parity at the real ssv2 geometry (B=1): each of the 7 distinct block shapes against the
    CPU oracle on the same input, fp32 mode, <= 1e-4."""
    ops.set_impl(gemm=ops.IMPL_SIMT, attn=ops.IMPL_SIMT)
    cfg = ssv2_cfg()
    specs = block_specs(cfg)[0]
    for i in (0, 1, 2, 3, 4, 14, 15):
        sp = specs[i]
        p = synth_state(block_param_shapes(sp), 500 + i, w_std=0.06)
        T, H, W = sp["input_size"]
        x = synth_input(f"fullblk{i}", (1, 1 + T * H * W + 64, sp["dim"]), 7)
        with torch.no_grad():
            want, thw_w = O.block_forward(x, sp["input_size"], p, "", sp)
        m = svit_b200.MultiScaleBlock(dim=sp["dim"], dim_out=sp["dim_out"], num_heads=sp["num_heads"],
                                      input_size=sp["input_size"], qkv_bias=True, norm_layer=LN, kernel_q=[3, 3, 3],
                                      kernel_kv=[3, 3, 3], stride_q=sp["stride_q"], stride_kv=sp["stride_kv"],
                                      rel_pos_spatial=True, rel_pos_temporal=True, residual_pooling=True,
                                      dim_mul_in_att=True)
        m.load_state_dict(p)
        m = m.to(DEV)
        with torch.no_grad():
            got, thw_g = m(x.to(DEV), sp["input_size"])
        assert list(thw_g) == list(thw_w)
        err = max_rel_err(cpu(got), want)
        print(f"block {i}: fp32 max-rel-err {err:.2e}")
        assert err < FP32_TOL, i
    ops.set_impl(gemm=ops.IMPL_AUTO, attn=ops.IMPL_AUTO)


# ------------------------------------------------------------------------------------------------ object tokens
def test_object_token_indices_bit_exact():
    cfg = tiny_cfg()
    m = svit_b200.SViT(cfg, compute_dtype=torch.float32).to(DEV)
    with torch.no_grad():
        m.object_queries.copy_(torch.arange(4 * 96, dtype=torch.float32).reshape(1, 4, 96))
        m.pos_embed_temporal.copy_(1000.0 * torch.arange(4, dtype=torch.float32)[None, :, None].expand(1, 4, 96))
        m.cls_token.fill_(-7.0)
    clip = synth_input("idx.clip", (2, 3, 4, 32, 32), 1).to(DEV)
    pe = m.patch_embed.proj
    x = ops.patch_embed_tokens(clip, pe.weight, pe.bias, m.cls_token, m.object_queries, m.pos_embed_temporal,
                               pe.kernel_size, pe.stride, pe.padding, torch.float32)
    L = 2 * 8 * 8
    assert x.shape == (2, 1 + L + 16, 96)
    assert torch.equal(cpu(x[:, 0]), torch.full((2, 96), -7.0))
    for t in range(4):
        for o in range(4):
            idx = O.object_token_index(2, 8, 8, t, o)
            assert torch.equal(cpu(x[1, idx]), torch.arange(o * 96, (o + 1) * 96, dtype=torch.float32) + 1000.0 * t)
    want, thw = O.patch_embed(cpu(clip), cpu(pe.weight), cpu(pe.bias), pe.stride, pe.padding)
    assert thw == [2, 8, 8]
    assert max_rel_err(cpu(x[:, 1:1 + L]), want) < 1e-5
    co = ops.gather_cls_obj(x, 16)
    assert torch.equal(co[:, 0], x[:, 0]) and torch.equal(co[:, 1:], x[:, -16:])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("K", [1, 4, 16])
def test_roi_object_tokens_vs_oracle(K, dtype):
    """fp32: the channel-per-thread kernel; bf16: the 8-channels-per-thread kernel with per-box 1-D sample tables
    (roi_tokens_vec8_kernel) on bf16-rounded features, output rounded to bf16."""
    gen = torch.Generator().manual_seed(K)
    for (B, C, Tp, Hf, Tx, scale) in ((2, 96, 2, 7, 4, 1 / 16), (1, 192, 4, 14, 8, 1 / 8), (1, 96, 1, 7, 1, 1 / 16),
                                      (1, 384, 2, 14, 4, 1 / 16)):
        feat = torch.randn(B, C, Tp, Hf, Hf, generator=gen).to(dtype).float()
        size = Hf / scale
        boxes = torch.rand(B, Tx, K, 4, generator=gen) * size
        boxes[..., 2:] = boxes[..., :2] + torch.rand(B, Tx, K, 2, generator=gen) * size * 0.6
        boxes[0, 0, 0] = 0.0                                     # degenerate: zero area
        if K > 1:
            boxes[0, 0, 1] = torch.tensor([-30.0, -20.0, size + 40, size + 10])   # out of image
        want, assign_w = O.roi_object_tokens(feat, boxes, patch_stride_t=2, spatial_scale=scale)
        tokens = torch.zeros(B, 1 + Tp * Hf * Hf + 3, C)
        tokens[:, 1:1 + Tp * Hf * Hf] = feat.permute(0, 2, 3, 4, 1).reshape(B, -1, C)
        got, assign = ops.roi_tokens(tokens.to(DEV, dtype), [Tp, Hf, Hf], boxes, 2, scale, 7)
        assert torch.equal(cpu(assign).long(), assign_w.float().long())   # (batch, slice): bit exact
        assert max_rel_err(cpu(got), want) < (1e-5 if dtype == torch.float32 else 6e-3)


def test_roi_align_vs_torchvision():
    from torchvision.ops import roi_align as tv_roi

    gen = torch.Generator().manual_seed(11)
    feat = torch.randn(2, 32, 14, 14, generator=gen)
    rois = torch.tensor([[0, 10.0, 12.0, 180.0, 190.0], [1, 0.0, 0.0, 224.0, 224.0], [1, 50.0, 60.0, 50.0, 60.0],
                         [0, -20.0, -5.0, 30.0, 240.0], [1, 100.0, 100.0, 111.0, 104.0]])
    for P, sr, al in ((7, 0, True), (3, 2, False), (5, 0, False)):
        want = tv_roi(feat, rois, output_size=P, spatial_scale=1 / 16, sampling_ratio=sr, aligned=al)
        got = ops.roi_align_nhwc(feat.permute(0, 2, 3, 1).contiguous().to(DEV), rois, P, 1 / 16, sr, al)
        assert max_rel_err(cpu(got).permute(0, 3, 1, 2), want) < 1e-5


def test_match_haog_device_bit_exact(golden):
    g = golden("boxes.pt")
    inp = torch.cat([c["inp"] for c in g["match_haog"]]).contiguous().to(DEV)
    out, contact = ops.match_haog_device(inp)
    assert torch.equal(cpu(out), torch.cat([c["out"] for c in g["match_haog"]]))
    assert torch.equal(contact.cpu(), torch.stack([c["contact"] for c in g["match_haog"]]))
    z = torch.cat([c["inp"] for c in g["zero_empty"]]).contiguous().to(DEV)
    assert torch.equal(cpu(ops.zero_empty_boxes_device(z)), torch.cat([c["out"] for c in g["zero_empty"]]))


# ------------------------------------------------------------------------------------------------ properties at size
def test_full_size_properties_bf16():
    """Size-independent checks at BASELINE batch shapes (no CPU oracle at this size):
    batch independence (clip b's output does not depend on its neighbours) and determinism."""
    cfg = ssv2_cfg()
    m = svit_b200.SViT(cfg, compute_dtype=torch.bfloat16)
    m.load_state_dict(synth_state(state_shapes(cfg), 400, w_std=0.04))
    m = m.to(DEV).eval()
    clip = synth_input("prop.clip", (4, 3, 16, 224, 224), 8).to(DEV)
    with torch.no_grad():
        p4, e4 = m([clip])
        p4b, _ = m([clip])
        p1, e1 = m([clip[2:3]])
    assert torch.equal(p4, p4b)
    assert max_rel_err(cpu(e1["logits"]), cpu(e4["logits"][2:3])) < 1e-2
    assert torch.allclose(p4.float().sum(1).cpu(), torch.ones(4), atol=1e-3)
    assert e4["obj_desc"].shape == (4, 16, 4, 768) and e4["pred_bboxes"].shape == (4, 16, 4, 5)
    assert e4["pred_contact_state"].shape == (4, 16, 2, 5)


@pytest.mark.parametrize("spec", [(2, 192, 2, 9, 11, 5, 2), (1, 96, 3, 14, 14, 8, 2), (1, 384, 1, 7, 8, 4, 2)])
def test_skip_maxpool_bf16_backward_vectorised(spec):
    """bf16 skip-path max-pool (attention.py:549-555): forward bit-exact and the 8-channel-per-thread backward equal to
    ATen's max_pool3d backward (gradient routed to the first maximum in scan order) on the same bf16 values."""
    B, C, T, H, W, Ot, s = spec
    g = torch.Generator().manual_seed(C + H)
    N = 1 + T * H * W + Ot
    x = torch.randn(B, N, C, generator=g).bfloat16()
    # ties exercise the first-maximum rule
    x[:, 1:1 + T * H * W:3] = x[:, 2:2 + T * H * W:3][:, : x[:, 1:1 + T * H * W:3].shape[1]]
    xd = x.to(DEV).requires_grad_(True)
    out = ops.skip_pool(xd, (T, H, W), Ot, s)
    xr = x.float().requires_grad_(True)
    patch = xr[:, 1:1 + T * H * W].reshape(B, T, H, W, C).permute(0, 4, 1, 2, 3)
    pooled = torch.nn.functional.max_pool3d(patch, (1, 3, 3), (1, s, s), (0, 1, 1))
    ref = torch.cat([xr[:, :1], pooled.flatten(2).transpose(1, 2), xr[:, 1 + T * H * W:]], 1)
    assert torch.equal(cpu(out), ref.detach())
    gy = torch.randn(ref.shape, generator=g).bfloat16()
    out.backward(gy.to(DEV))
    ref.backward(gy.float())
    assert torch.equal(cpu(xd.grad), xr.grad.bfloat16().float())


def test_frames_pass_full_ssv2_vs_oracle():
    """the reference's no-grad pass over single frames (tools/train_net.py:105-110) at the real ssv2
    geometry -- T = 1 shapes of every kernel (N = 3141 / 789 / 201 / 54, rel_pos_t interpolated 15 -> 1) -- against
    the CPU oracle on the same two frames: fp32 mode <= 3e-4, bf16 mode <= 2e-2 on logits and object descriptors."""
    from svit_b200.distributed import consistency_loss, forward_video_frames
    cfg = ssv2_cfg()
    state = synth_state(state_shapes(cfg), 77, w_std=0.04)
    clip = synth_input("frames.clip", (1, 3, 2, 224, 224), 8)       # B = 1, T = 2 -> two frames
    frames = clip.transpose(1, 2).flatten(0, 1).unsqueeze(2)
    want_out, want = O.svit_forward(frames, state, block_specs(cfg)[0], cfg, training=False)
    m = svit_b200.SViT(cfg, compute_dtype=torch.float32)
    m.load_state_dict(state)
    m = m.to(DEV).eval()
    for dtype, tol, impl in ((torch.float32, 3e-4, ops.IMPL_SIMT), (torch.bfloat16, 2e-2, ops.IMPL_AUTO)):
        ops.set_impl(gemm=impl, attn=impl)
        try:
            m.compute_dtype = dtype
            probs, extra = forward_video_frames(m, clip.to(DEV))
        finally:
            ops.set_impl(gemm=ops.IMPL_AUTO, attn=ops.IMPL_AUTO)
        assert extra["obj_desc"].shape == want["obj_desc"].shape == (2, 1, 4, 768)
        assert max_rel_err(cpu(extra["logits"]), want["logits"]) < tol, dtype
        assert max_rel_err(cpu(extra["obj_desc"]), want["obj_desc"]) < tol, dtype
        assert max_rel_err(cpu(probs), want_out) < 5 * tol, dtype
    # losses.py:127-136 with the stock lambda keys: no consistency term (key mismatch kept on purpose)
    vid = {"obj_desc": torch.zeros(1, 2, 4, 768, device=DEV)}
    assert consistency_loss(m._lambda, vid, extra) == {}
    lam = dict(m._lambda, video_image_desc_l1_loss=1.5)
    got = consistency_loss(lam, vid, extra)["video_image_desc_l1_loss"]
    assert abs(float(got) - float(want["obj_desc"].abs().mean())) < 2e-2 * float(want["obj_desc"].abs().mean()) + 1e-6


def test_fused_adamw_matches_torch():
    """clip_grad_norm_(1.0) + torch.optim.AdamW over two weight-decay groups (optimizer.py:31-104,
    train_net.py:133-151) against the two-launch fused step, three steps, odd tensor sizes, fresh .grad tensors each step."""
    from svit_b200.optim import FusedAdamW
    gen = torch.Generator().manual_seed(3)
    shapes = [(96, 1, 3, 3, 3), (288, 96), (288,), (8193,), (1, 1, 96), (17, 33)]
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=gen).to(DEV)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    def groups(ps):
        return [{"params": [p for p in ps if p.dim() > 1], "weight_decay": 1e-2},
                {"params": [p for p in ps if p.dim() == 1], "weight_decay": 0.0}]
    ref = torch.optim.AdamW(groups(ref_p), lr=2e-3, eps=1e-8)
    ours = FusedAdamW(groups(our_p), lr=2e-3, eps=1e-8)
    for step in range(3):
        grads = [torch.randn(s, generator=gen).to(DEV) * (3.0 if step == 1 else 0.01) for s in shapes]  # clipped / not
        for p, q, g in zip(ref_p, our_p, grads):
            p.grad, q.grad = g.clone(), g.clone()
        norm = torch.nn.utils.clip_grad_norm_(ref_p, 1.0)
        ref.step()
        ours.step(max_norm=1.0)
        assert abs(float(ours.grad_norm()) - float(norm)) <= 1e-5 * float(norm)
        for p, q in zip(ref_p, our_p):
            assert max_rel_err(cpu(q), cpu(p)) < 2e-6, (step, tuple(p.shape))
    for p, q in zip(ref_p, our_p):
        # second moments: (1 - beta2) g * g is rounded in a different order than addcmul_ and g carries the clip factor
        assert max_rel_err(cpu(ours.state[q]["exp_avg_sq"]), cpu(ref.state[p]["exp_avg_sq"])) < 5e-5
        assert max_rel_err(cpu(ours.state[q]["exp_avg"]), cpu(ref.state[p]["exp_avg"])) < 5e-6


def test_normalize_u8_bit_exact_and_model_accepts_frames():
    """uint8 frames [B,T,H,W,3] -> (x/255 - mean)/std, CTHW layout (datasets/utils.py:287-303): fp32 output
    bit-identical to the reference expression, bf16 output = its rounding; SViT.forward takes the uint8 tensor directly."""
    gen = torch.Generator().manual_seed(11)
    frames = torch.randint(0, 256, (2, 4, 32, 32, 3), generator=gen, dtype=torch.uint8)
    mean, std = [0.45, 0.40, 0.35], [0.225, 0.25, 0.2]
    ref = ((frames.float() / 255.0 - torch.tensor(mean)) / torch.tensor(std)).permute(0, 4, 1, 2, 3).contiguous()
    got32 = ops.normalize_u8(frames.to(DEV), mean, std, torch.float32)
    assert torch.equal(cpu(got32), ref)
    got16 = ops.normalize_u8(frames.to(DEV), mean, std, torch.bfloat16)
    assert torch.equal(cpu(got16), ref.bfloat16().float())
    cfg = tiny_cfg()
    m = svit_b200.SViT(cfg, compute_dtype=torch.float32).to(DEV).eval()
    ops.set_impl(gemm=ops.IMPL_SIMT, attn=ops.IMPL_SIMT)
    try:
        with torch.no_grad():
            a, _ = m([frames.to(DEV)])
            clip = ((frames.float() / 255.0 - torch.tensor(cfg.DATA.MEAN)) / torch.tensor(cfg.DATA.STD)).permute(0, 4, 1, 2, 3)
            b, _ = m([clip.contiguous().to(DEV)])
    finally:
        ops.set_impl(gemm=ops.IMPL_AUTO, attn=ops.IMPL_AUTO)
    assert torch.equal(a, b)


def test_fused_adamw_invalidates_weight_caches():
    """The fused step writes parameters through raw pointers; the bf16 GEMM-weight cache (keyed on param._version)
    must see the update: a linear layer evaluated after the step uses the new weights."""
    from svit_b200.optim import FusedAdamW
    gen = torch.Generator().manual_seed(9)
    w = torch.nn.Parameter(torch.randn(96, 96, generator=gen).to(DEV))
    x = torch.randn(128, 96, generator=gen).to(torch.bfloat16).to(DEV)
    y0 = ops.linear(x, w)
    opt = FusedAdamW([w], lr=0.5, weight_decay=0.0)
    w.grad = torch.ones_like(w)
    opt.step()
    y1 = ops.linear(x, w)
    ref = x.float() @ w.detach().float().bfloat16().float().t()
    assert max_rel_err(cpu(y1), cpu(ref)) < 1e-2
    assert max_rel_err(cpu(y0), cpu(ref)) > 1e-1  # the step really moved the weights


def _roi_reference_tokens(feat, boxes, pst, scale, P=7):
    """Differentiable CPU reference of roi_object_tokens: torchvision roi_align per (b, t) and torch.max over the bins
    flattened row-major (gradient to the FIRST maximal bin, the kernel's rule)."""
    from torchvision.ops import roi_align as tv_roi
    B, C, Tp, H, W = feat.shape
    _, Tx, K, _ = boxes.shape
    rows = []
    for b in range(B):
        for t in range(Tx):
            s = 0 if Tp == 1 else (t if Tx == 1 else t // pst)
            rois = torch.cat([torch.zeros(K, 1), boxes[b, t]], dim=1)
            r = tv_roi(feat[b:b + 1, :, s], rois, output_size=P, spatial_scale=scale, sampling_ratio=0, aligned=True)
            rows.append(r.flatten(-2).max(dim=-1).values)
    return torch.stack(rows).reshape(B, Tx * K, C)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_roi_tokens_backward_vs_torchvision(dtype):
    """SURVEY R3 backward (VERDICT r1 missing #1): d tokens / d features through the arg-max bin's bilinear taps."""
    gen = torch.Generator().manual_seed(5)
    for (B, C, Tp, Hf, Tx, K, scale) in ((2, 96, 2, 7, 4, 4, 1 / 16), (1, 192, 4, 14, 8, 3, 1 / 8), (1, 96, 1, 7, 1, 2, 1 / 16)):
        feat = torch.randn(B, C, Tp, Hf, Hf, generator=gen).to(dtype).float()
        size = Hf / scale
        boxes = torch.rand(B, Tx, K, 4, generator=gen) * size * 0.7
        boxes[..., 2:] = boxes[..., :2] + 8 + torch.rand(B, Tx, K, 2, generator=gen) * size * 0.5
        gy = torch.randn(B, Tx * K, C, generator=gen).to(dtype).float()
        fr = feat.clone().requires_grad_(True)
        want = _roi_reference_tokens(fr, boxes, 2, scale)
        want.backward(gy)
        L = Tp * Hf * Hf
        seq = torch.zeros(B, 1 + L + 5, C)
        seq[:, 1:1 + L] = feat.permute(0, 2, 3, 4, 1).reshape(B, L, C)
        x = seq.to(DEV, dtype).requires_grad_(True)
        got, _ = ops.roi_tokens(x, [Tp, Hf, Hf], boxes, 2, scale, 7)
        assert max_rel_err(cpu(got), want.detach()) < (1e-5 if dtype == torch.float32 else 6e-3)
        got.backward(gy.to(DEV, dtype))
        dx = cpu(x.grad)
        assert dx[:, 0].abs().max() == 0 and dx[:, 1 + L:].abs().max() == 0
        dwant = fr.grad.permute(0, 2, 3, 4, 1).reshape(B, L, C)
        # bf16: an arg-max decided on bf16-rounded bin values can differ from the fp32 reference's on near ties
        tol = 1e-5 if dtype == torch.float32 else 3e-2
        err = max_rel_err(dx[:, 1:1 + L], dwant)
        if dtype == torch.float32:
            assert err < tol, err
        else:
            bad = ((dx[:, 1:1 + L] - dwant).abs() > tol * dwant.abs().max()).float().mean().item()
            assert bad < 5e-3, (err, bad)


def test_roi_scatter_into_sequence_and_model_option():
    """RoI tokens land at sequence rows 1 + T'H'W' + t*K + k (bit-exact index rule), 'add' keeps the learned object token,
    and SViT.forward(bboxes=...) with cfg.SVIT.BOX_TOKENS routes them into the head with gradients to the backbone."""
    gen = torch.Generator().manual_seed(9)
    B, C, Tp, Hf, Tx, K = 2, 96, 2, 7, 4, 4
    L = Tp * Hf * Hf
    seq = torch.randn(B, 1 + L + Tx * K, C, generator=gen).to(DEV)
    boxes = torch.rand(B, Tx, K, 4, generator=gen) * 60
    boxes[..., 2:] = boxes[..., :2] + 10 + torch.rand(B, Tx, K, 2, generator=gen) * 40
    toks, _ = ops.roi_tokens(seq, [Tp, Hf, Hf], boxes, 2, 1 / 16, 7)
    rep, _ = ops.roi_scatter_tokens(seq, [Tp, Hf, Hf], boxes, 2, 1 / 16, 7, "replace")
    add, _ = ops.roi_scatter_tokens(seq, [Tp, Hf, Hf], boxes, 2, 1 / 16, 7, "add")
    assert torch.equal(rep[:, :1 + L], seq[:, :1 + L]) and torch.equal(add[:, :1 + L], seq[:, :1 + L])
    for t in range(Tx):
        for k in range(K):
            row = 1 + L + t * K + k
            assert torch.equal(rep[:, row], toks[:, t * K + k])
    assert torch.allclose(add[:, 1 + L:], seq[:, 1 + L:] + toks, rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        ops.roi_scatter_tokens(seq[:, :-1], [Tp, Hf, Hf], boxes, 2, 1 / 16, 7, "replace")
    # model option
    cfg = tiny_cfg()
    cfg.SVIT.BOX_TOKENS = "add"
    ops.set_impl(gemm=ops.IMPL_SIMT, attn=ops.IMPL_SIMT)
    try:
        m = svit_b200.SViT(cfg, compute_dtype=torch.float32).to(DEV).eval()
        clip = synth_input("tiny.clip", (2, 3, 4, 32, 32), 5).to(DEV)
        bx = torch.tensor([[4.0, 4.0, 20.0, 24.0], [0.0, 0.0, 32.0, 32.0], [8.0, 2.0, 30.0, 12.0], [1.0, 16.0, 14.0, 31.0]]).repeat(2, 4, 1, 1)
        _, e0 = m([clip])
        _, e1 = m([clip], bboxes=bx)
        assert e1["obj_desc"].shape == e0["obj_desc"].shape
        assert torch.allclose(e1["obj_desc"].reshape(2, 16, -1), e0["obj_desc"].reshape(2, 16, -1) + e1["roi_tokens"].float(),
                              rtol=1e-5, atol=1e-5)
        m.zero_grad()
        e1["obj_desc"].square().sum().backward()
        assert m.patch_embed.proj.weight.grad is not None and m.patch_embed.proj.weight.grad.abs().max() > 0
    finally:
        ops.set_impl(gemm=ops.IMPL_AUTO, attn=ops.IMPL_AUTO)


# ------------------------------------------------------------------------------------------------ round-2 parity holes
def _norm_rel(got, want):
    """||got - want|| / ||want||: the gradient metric for bf16 (element-wise max ratios are dominated by near-zero entries)."""
    return float((got.double() - want.double()).norm() / want.double().norm().clamp_min(1e-30))


def test_svit_tiny_bf16_training_gradients(golden):
    """the bf16 training path (tcgen05 GEMMs, attention forward / backward, vectorised pooling backward)
    against the gradients of the UNMODIFIED reference (fp32, tests/golden/svit_tiny.pt).  Stated tolerance: 3e-2 on the
    loss, on every parameter's gradient norm and on ||g - g_ref|| / ||g_ref|| of the stored gradients (5e-2 for the rel-pos
    tables)."""
    g = golden("svit_tiny.pt")["video"]
    cfg = tiny_cfg()
    m = _model(cfg, g, torch.bfloat16)
    m.train()
    for mod in m.modules():
        if isinstance(mod, svit_b200.DropPath):
            mod.drop_prob = 0.0
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    clip = synth_input("tiny.clip", (2, 3, 4, 32, 32), 5).to(DEV)
    logits, extra = m([clip])
    assert max_rel_err(cpu(logits), g["train_logits"]) < 2e-2
    tgt = (torch.arange(2) % logits.shape[1]).to(DEV)
    loss = torch.nn.functional.cross_entropy(logits.float(), tgt) + 0.1 * extra["pred_bboxes"].float().square().mean() \
        + 0.1 * extra["pred_contact_state"].float().square().mean()
    assert abs(loss.item() - g["loss"].item()) < 3e-2 * abs(g["loss"].item())
    loss.backward()
    named = dict(m.named_parameters())
    worst_norm, worst = 0.0, 0.0
    total = sum(float(n) ** 2 for n in g["grad_norms"].values()) ** 0.5
    for k, n in g["grad_norms"].items():
        if n.item() < 1e-3 * total:  # gradients at noise level relative to the model's total: covered by the total below
            continue
        e = abs(named[k].grad.double().norm().item() - n.item()) / n.item()
        worst_norm = max(worst_norm, e)
        assert e < 3e-2, (k, e)
    got_total = sum(float(p.grad.double().norm()) ** 2 for p in named.values()) ** 0.5
    assert abs(got_total - total) < 3e-2 * total
    errs = {}
    for k, w in g["grads"].items():
        if w.norm() < 1e-3 * total:
            continue
        errs[k] = _norm_rel(cpu(named[k].grad), w)
    print(f"bf16 tiny training gradients vs reference: worst norm error {worst_norm:.2e}; ||dg||/||g||: "
          + ", ".join(f"{k} {v:.2e}" for k, v in sorted(errs.items(), key=lambda kv: -kv[1])))
    for k, e in errs.items():
        # the rel-pos tables collect a bf16-rounded bias gradient dE over every (query, key) pair of a head: 5e-2 there
        assert e < (5e-2 if "rel_pos" in k else 3e-2), (k, e)


def test_svit_tiny_droppath_training_parity(golden):
    """VERDICT r1 weak #1 / SURVEY App. C: training-mode parity with DropPath ON.  The reference draws two
    torch.rand((B,1,1)) masks per block (attention branch, MLP branch) in block order from the global CPU generator;
    the same draws (same seed, same order) are injected into svit_b200.DropPath, fp32 mode."""
    g = golden("svit_tiny_droppath.pt")
    cfg = tiny_cfg()
    ops.set_impl(gemm=ops.IMPL_SIMT, attn=ops.IMPL_SIMT)
    try:
        m = _model(cfg, g, torch.float32)
        m.train()
        n_dp = 0
        for mod in m.modules():
            if isinstance(mod, svit_b200.DropPath):
                mod.drop_prob = g["drop_prob"]
                n_dp += 1
            if isinstance(mod, nn.Dropout):
                mod.p = 0.0
        assert n_dp == g["n_droppath_modules"]
        B = g["logits"].shape[0]
        torch.manual_seed(g["rng_seed"])
        draws = []
        real_rand = torch.rand

        def cpu_order_rand(shape, *a, dtype=None, device=None, **kw):
            r = real_rand(tuple(shape), dtype=torch.float32)  # global CPU generator, the reference's stream
            draws.append(r.reshape(-1))
            return r.to(device=device, dtype=dtype or torch.float32)
        msa.torch.rand = cpu_order_rand
        try:
            clip = synth_input("tiny.clip.dp", (B, 3, 4, 32, 32), 15).to(DEV)
            logits, extra = m([clip])
        finally:
            msa.torch.rand = real_rand
        assert torch.equal(torch.stack(draws), g["rand"])  # same number of draws, same order, same values
        assert max_rel_err(cpu(logits), g["logits"]) < 2e-4
        tgt = (torch.arange(B) % logits.shape[1]).to(DEV)
        loss = torch.nn.functional.cross_entropy(logits, tgt) + 0.1 * extra["pred_bboxes"].square().mean()
        assert abs(loss.item() - g["loss"].item()) < 1e-4 * abs(g["loss"].item())
        loss.backward()
        named = dict(m.named_parameters())
        for k, n in g["grad_norms"].items():
            # the reference keeps unused heads in the graph (+ sum(p) * 0): their gradient is zero there, None here
            got = named[k].grad.double().norm().item() if named[k].grad is not None else 0.0
            assert abs(got - n.item()) <= 2e-3 * n.item() + 1e-7, (k, got, n.item())
        for k, w in g["grads"].items():
            assert max_rel_err(cpu(named[k].grad), w) < 2e-3, k
    finally:
        ops.set_impl(gemm=ops.IMPL_AUTO, attn=ops.IMPL_AUTO)


def test_blocks_full_size_bf16_forward_backward_vs_oracle():
    """One real-geometry block per stage (blocks 0, 2, 4, 15 of configs/ssv2.yaml, B = 1), bf16 production kernels,
    forward AND backward against the CPU oracle's autograd in fp32 on the same bf16-rounded input: forward <= 2e-2
    (max|dy| / max|y|), gradients ||dg|| / ||g|| <= 3e-2 (input gradient, GEMM / pooling / rel-pos / norm parameters)."""
    cfg = ssv2_cfg()
    specs = block_specs(cfg)[0]
    for i in (0, 2, 4, 15):
        sp = specs[i]
        p = synth_state(block_param_shapes(sp), 600 + i, w_std=0.06)
        T, H, W = sp["input_size"]
        x = synth_input(f"fullblk.bf16.{i}", (1, 1 + T * H * W + 64, sp["dim"]), 8).bfloat16().float()
        gy_shape = None
        po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        xo = x.clone().requires_grad_(True)
        want, thw_w = O.block_forward(xo, sp["input_size"], po, "", sp)
        gy = synth_input(f"fullblk.bf16.gy{i}", tuple(want.shape), 9).bfloat16().float()
        want.backward(gy)
        m = svit_b200.MultiScaleBlock(dim=sp["dim"], dim_out=sp["dim_out"], num_heads=sp["num_heads"],
                                      input_size=sp["input_size"], qkv_bias=True, norm_layer=LN, kernel_q=[3, 3, 3],
                                      kernel_kv=[3, 3, 3], stride_q=sp["stride_q"], stride_kv=sp["stride_kv"],
                                      rel_pos_spatial=True, rel_pos_temporal=True, residual_pooling=True,
                                      dim_mul_in_att=True)
        m.load_state_dict(p)
        m = m.to(DEV)
        xg = x.to(DEV, torch.bfloat16).requires_grad_(True)
        got, thw_g = m(xg, sp["input_size"])
        assert list(thw_g) == list(thw_w)
        ferr = max_rel_err(cpu(got), want.detach())
        got.backward(gy.to(DEV, torch.bfloat16))
        named = dict(m.named_parameters())
        errs = {"dx": _norm_rel(cpu(xg.grad), xo.grad)}
        for k in ("attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.bias", "attn.pool_q.weight",
                  "attn.pool_k.weight", "attn.norm_q.weight", "attn.rel_pos_h", "attn.rel_pos_t", "norm1.weight", "norm2.bias"):
            errs[k] = _norm_rel(cpu(named[k].grad), po[k].grad)
        worst = max(errs, key=errs.get)
        print(f"block {i}: bf16 forward max-rel-err {ferr:.2e}; worst gradient {worst} {errs[worst]:.2e}")
        assert ferr < 2e-2, (i, ferr)
        assert errs[worst] < 3e-2, (i, errs)


def test_msa_config5_shapes_vs_oracle():
    """BASELINE configs[4] shapes: a full MultiScaleAttention at each of the 7 distinct stage shapes of a 32x312^2 clip
    (patch grid 16x78x78, 128 object tokens; run-time rel-pos table interpolation, non-integer q/k ratios), B = 1, bf16,
    against the CPU oracle in fp32.  Blocks 1, 3, 14 have kh + kw + kt = 56 > 47 bias entries and take the older
    attn_tc kernel, the others attn_tc3: both paths are covered.  Tolerance 2e-2 (max|dy| / max|y|)."""
    cfg = ssv2_cfg()
    cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE, cfg.DATA.TEST_CROP_SIZE = 32, 312, 312
    specs = block_specs(cfg)[0]
    ps = cfg.MVIT.PATCH_STRIDE
    thw = [32 // ps[0], 312 // ps[1], 312 // ps[2]]
    Otot = 32 * cfg.SVIT.O
    seen = set()
    for i, sp in enumerate(specs):
        sq = sp["stride_q"][1] if sp["stride_q"] else 1
        key = (sp["dim"], sp["dim_out"], tuple(thw), sq, tuple(sp["stride_kv"]))
        nthw = [thw[0], (thw[1] - 1) // sq + 1, (thw[2] - 1) // sq + 1]
        if key not in seen:
            seen.add(key)
            p = synth_state(attn_param_shapes(sp["dim"], sp["dim_out"], sp["num_heads"], sp["input_size"], sp["stride_q"],
                                                sp["stride_kv"]), 700 + i, w_std=0.06)
            N = 1 + thw[0] * thw[1] * thw[2] + Otot
            x = synth_input(f"cfg5.msa{i}", (1, N, sp["dim"]), 10).bfloat16().float()
            with torch.no_grad():
                want, q_thw = O.msa_forward(x, thw, p, "", sp["num_heads"], sp["stride_q"], sp["stride_kv"])
            m = svit_b200.MultiScaleAttention(sp["dim"], sp["dim_out"], sp["input_size"], num_heads=sp["num_heads"],
                                              qkv_bias=True, kernel_q=sp["kernel_q"], kernel_kv=sp["kernel_kv"],
                                              stride_q=sp["stride_q"], stride_kv=sp["stride_kv"], norm_layer=LN,
                                              rel_pos_spatial=True, rel_pos_temporal=True, residual_pooling=True)
            m.load_state_dict(p)
            m = m.to(DEV)
            with torch.no_grad():
                got, q_got = m(x.to(DEV, torch.bfloat16), thw)
            assert list(q_got) == list(q_thw) == nthw
            err = max_rel_err(cpu(got), want)
            skv = sp["stride_kv"][1]
            ne = thw[0] + 2 * ((thw[1] - 1) // skv + 1)
            print(f"config-5 block {i}: Nq {got.shape[1]} heads {sp['num_heads']} bias entries {ne}: bf16 max-rel-err {err:.2e}")
            assert err < 2e-2, (i, err)
        thw = nthw
    assert len(seen) == 7


@pytest.mark.parametrize("geom", [  # B, T, H, W (kernel (3,7,7), stride (2,4,4), padding (1,3,3), 3 -> 96 channels)
    (2, 16, 224, 224), (3, 4, 32, 32), (5, 1, 224, 224), (1, 8, 312, 312), (2, 3, 30, 50)])
def test_patch_embed_implicit_gemm(geom):
    """VERDICT r1 missing #4: PatchEmbed as an implicit GEMM over space-to-depth cells (csrc/patch_embed_tc.cu) against
    (a) the reference conv3d (stem_helper.py:309-320) on the same bf16-rounded clip and bf16-rounded weights, fp32 math, and
    (b) the im2col + GEMM path it replaces; also from uint8 frames (normalisation fused into the cell layout kernel)."""
    B, T, H, W = geom
    gen = torch.Generator().manual_seed(T * 1000 + H)
    E = 96
    w = (torch.randn(E, 3, 3, 7, 7, generator=gen) * 0.1)
    b = torch.randn(E, generator=gen) * 0.1
    cls, qs, pt = torch.randn(1, 1, E, generator=gen), torch.randn(1, 4, E, generator=gen), torch.randn(1, T, E, generator=gen)
    clip = torch.randn(B, 3, T, H, W, generator=gen).bfloat16()
    kernel, stride, padding = (3, 7, 7), (2, 4, 4), (1, 3, 3)
    if not _lib.lib().svit_patch_embed_s2d_supported(3, *kernel, *stride, *padding, E):
        pytest.fail("the ssv2.yaml stem geometry must be supported by the implicit GEMM")
    dev = lambda t: t.to(DEV)
    with torch.no_grad():
        got = ops.patch_embed_tokens(dev(clip), dev(w), dev(b), dev(cls), dev(qs), dev(pt), kernel, stride, padding, torch.bfloat16)
        ops._state["implicit_patch_embed"] = False
        try:
            old = ops.patch_embed_tokens(dev(clip), dev(w), dev(b), dev(cls), dev(qs), dev(pt), kernel, stride, padding, torch.bfloat16)
        finally:
            ops._state["implicit_patch_embed"] = True
    want = torch.nn.functional.conv3d(clip.float(), w.bfloat16().float(), b, stride=stride, padding=padding)
    L = want.shape[2] * want.shape[3] * want.shape[4]
    want = want.flatten(2).transpose(1, 2)
    assert got.shape == old.shape == (B, 1 + L + T * 4, E)
    assert max_rel_err(cpu(got[:, 1:1 + L]), want) < 6e-3          # bf16 rounding of the output only
    assert max_rel_err(cpu(got), cpu(old)) < 6e-3
    assert torch.equal(got[:, 0], old[:, 0]) and torch.equal(got[:, 1 + L:], old[:, 1 + L:])   # cls / object rows
    # uint8 frames: (x / 255 - mean) / std fused into the cell layout
    frames = torch.randint(0, 256, (B, T, H, W, 3), generator=gen, dtype=torch.uint8)
    mean, std = [0.45, 0.40, 0.35], [0.225, 0.25, 0.2]
    with torch.no_grad():
        g8 = ops.patch_embed_tokens(dev(frames), dev(w), dev(b), dev(cls), dev(qs), dev(pt), kernel, stride, padding,
                                    torch.bfloat16, mean=mean, std=std)
        ref_clip = ((frames.float() / 255.0 - torch.tensor(mean)) / torch.tensor(std)).permute(0, 4, 1, 2, 3).contiguous()
        g16 = ops.patch_embed_tokens(dev(ref_clip.bfloat16()), dev(w), dev(b), dev(cls), dev(qs), dev(pt), kernel, stride, padding, torch.bfloat16)
    assert torch.equal(g8, g16)
