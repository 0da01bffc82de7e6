"""GPU tests of the tcgen05 (tensor-core) kernels, forced with impl=IMPL_TC so a silent fall back to the
CUDA-core path cannot make them pass.  Reference = fp64 matmul of the same bf16-rounded operands on the CPU."""
import pytest
import torch

from svit_b200 import ops
from tests.conftest import max_rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TC = ops.IMPL_TC


def cpu(t):
    return t.detach().float().cpu()


def _ref(A, B, tA, tB):
    a = A.float().t() if tA else A.float()
    b = B.float().t() if tB else B.float()
    return a.double() @ b.double()


@pytest.mark.parametrize("tA,tB", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_gemm_tc_layouts(tA, tB):
    gen = torch.Generator().manual_seed(17 + 2 * tA + tB)
    for (M, N, K) in ((200, 96, 96), (1000, 288, 192), (264, 192, 448), (136, 384, 1536), (128, 64, 64),
                      (520, 128, 320), (4104, 768, 384), (96, 2304, 768), (8, 96, 8)):
        A = torch.randn((K, M) if tA else (M, K), generator=gen).to(torch.bfloat16)
        B = torch.randn((N, K) if tB else (K, N), generator=gen).to(torch.bfloat16)
        for odt in (torch.bfloat16, torch.float32):
            out = torch.full((M, N), float("nan"), dtype=odt, device=DEV)
            ops.gemm(A.to(DEV), B.to(DEV), out, M, N, K, A.shape[1], B.shape[1], N, tA, tB, impl=TC)
            err = max_rel_err(cpu(out), _ref(A, B, tA, tB))
            assert err < (6e-3 if odt == torch.bfloat16 else 1e-5), (M, N, K, tA, tB, odt, err)


def test_gemm_tc_fused_epilogue():
    gen = torch.Generator().manual_seed(23)
    M, N, K = 2 * 300, 384, 192
    A = torch.randn(M, K, generator=gen).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=gen) * 0.1).to(torch.bfloat16)
    bias = torch.randn(N, generator=gen)
    res = torch.randn(2, 330, N, generator=gen).to(torch.bfloat16)
    scale = torch.tensor([0.0, 1.25])
    out = res.clone().to(DEV)
    pre = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), out, M, N, K, K, K, N, 0, 1, bias=bias.to(DEV), residual=res.to(DEV), ldr=N,
             sample_scale=scale.to(DEV), rows_per_sample=300, act=1, pre_out=pre, ldp=N, remap=(300, 330, 7), impl=TC)
    z = A.float() @ W.float().t() + bias
    ref = res.float().clone()
    ref[:, 7:307] += torch.nn.functional.gelu(z).reshape(2, 300, N) * scale[:, None, None]
    assert max_rel_err(cpu(pre), z) < 6e-3
    assert max_rel_err(cpu(out), ref) < 6e-3
    assert torch.equal(cpu(out[:, :7]), res[:, :7].float()) and torch.equal(cpu(out[:, 307:]), res[:, 307:].float())
    gout = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), gout, M, N, K, K, K, N, 0, 1, gelu_pre=pre, ldg=N, impl=TC)
    zz = cpu(pre).double().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    assert max_rel_err(cpu(gout), (A.float() @ W.float().t()).double() * zz.grad) < 6e-3


@pytest.mark.parametrize("M,N,K,mode", [
    (768, 288, 96, "gelu"), (768, 384, 192, "res_remap"), (1000, 1536, 384, "gelu_pre"), (50000, 1152, 384, "res"),
    (33000, 96, 384, "res_scale"), (4100, 200, 72, "bias"), (20000, 3072, 768, "gelu"), (3000, 768, 3072, "res_scale"),
    (130, 40, 40, "res"),
])
def test_gemm_tc_tma_epilogues(M, N, K, mode):
    """The TMA-store epilogue kernel (bf16 out): every fused epilogue, ragged M / N / K, many tiles per CTA so the
    staging-slot ring, the aux prefetch and both TMEM accumulator phases wrap several times."""
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=gen).to(torch.bfloat16).to(DEV)
    W = (torch.randn(N, K, generator=gen) * K ** -0.5).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=gen).to(DEV)
    z = A.float() @ W.float().t() + bias
    if mode == "gelu":
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.gemm(A, W, out, M, N, K, K, K, N, 0, 1, bias=bias, act=1, impl=TC)
        ref = torch.nn.functional.gelu(z)
    elif mode == "bias":
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.gemm(A, W, out, M, N, K, K, K, N, 0, 1, bias=bias, impl=TC)
        ref = z
    elif mode == "res":
        res = torch.randn(M, N, generator=gen).to(torch.bfloat16).to(DEV)
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.gemm(A, W, out, M, N, K, K, K, N, 0, 1, bias=bias, residual=res, ldr=N, impl=TC)
        ref = z + res.float()
    elif mode == "res_scale":
        nb = 4 if M % 4 == 0 else 1
        res = torch.randn(M, N, generator=gen).to(torch.bfloat16).to(DEV)
        scale = torch.tensor([0.0, 1.25, 1.0, 2.0][:nb]).to(DEV)
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.gemm(A, W, out, M, N, K, K, K, N, 0, 1, bias=bias, residual=res, ldr=N, sample_scale=scale,
                 rows_per_sample=M // nb, impl=TC)
        ref = z * scale.repeat_interleave(M // nb)[:, None] + res.float()
    elif mode == "res_remap":
        nb, rin, rout, off = M // 384, 384, 400, 9
        res = torch.randn(nb, rout, N, generator=gen).to(torch.bfloat16).to(DEV)
        out = res.clone()
        ops.gemm(A, W, out, M, N, K, K, K, N, 0, 1, bias=bias, residual=res, ldr=N, remap=(rin, rout, off), impl=TC)
        ref = res.float().clone()
        ref[:, off:off + rin] += z.reshape(nb, rin, N)
        assert torch.equal(out[:, :off], res[:, :off]) and torch.equal(out[:, off + rin:], res[:, off + rin:])
    elif mode == "gelu_pre":
        pre = torch.randn(M, N, generator=gen).to(torch.bfloat16).to(DEV)
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.gemm(A, W, out, M, N, K, K, K, N, 0, 1, gelu_pre=pre, ldg=N, impl=TC)
        zz = pre.float().double().requires_grad_(True)
        torch.nn.functional.gelu(zz).sum().backward()
        ref = ((z - bias).double() * zz.grad).float()
    assert not torch.isnan(out.float()).any()
    assert max_rel_err(cpu(out), cpu(ref)) < 6e-3, (M, N, K, mode)


@pytest.mark.parametrize("M,N", [(5000, 384), (201224, 96), (13064, 1536), (300, 2304), (77, 8)])
def test_colsum_bf16(M, N):
    gen = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, N, generator=gen).to(torch.bfloat16).to(DEV)
    got = ops._colsum(x, M, N)
    ref = x.double().sum(0)
    assert (cpu(got).double() - ref.cpu()).abs().max() / ref.abs().max().clamp_min(1.0) < 1e-4


@pytest.mark.parametrize("Mo,No,Kr", [(96, 384, 201224), (384, 1536, 13064), (1536, 384, 13064), (96, 448, 200704), (192, 768, 50696)])
def test_gemm_tc_wgrad_split_k(Mo, No, Kr):
    """Weight-gradient shape dW[Mo, No] = G[Kr, Mo]^T X[Kr, No] (fp32 out, both operands MN-major): few output tiles and a
    very long reduction -> split-K with fp32 atomics."""
    gen = torch.Generator().manual_seed(Mo + No)
    G = (torch.randn(Kr, Mo, generator=gen) * 0.1).to(torch.bfloat16).to(DEV)
    X = torch.randn(Kr, No, generator=gen).to(torch.bfloat16).to(DEV)
    out = torch.full((Mo, No), float("nan"), dtype=torch.float32, device=DEV)
    ops.gemm(G, X, out, Mo, No, Kr, Mo, No, No, 1, 0, impl=TC)
    ref = G.float().t().double() @ X.float().double()
    assert max_rel_err(cpu(out), ref.float().cpu()) < 2e-4


@pytest.mark.parametrize("Bn,h,Nq,Nk", [(2, 4, 1633, 457), (1, 2, 700, 1633), (3, 1, 130, 65)])
def test_gemm_tc_batched_attention_backward_shapes(Bn, h, Nq, Nk):
    """Batched tcgen05 GEMM as the attention backward uses it: dV[b,h] = P[b,h]^T dO[b,:,h,:] -- A stored [Nq, Nkp]
    (MN-major, padded key count), B a (sample, head) slice of a [B, Nq, h, 96] tensor (two-level batch index)."""
    gen = torch.Generator().manual_seed(Nq + Nk)
    Nkp = (Nk + 63) // 64 * 64
    P = (torch.rand(Bn * h, Nq, Nkp, generator=gen) / Nk).to(torch.bfloat16).to(DEV)
    dO = torch.randn(Bn, Nq, h, 96, generator=gen).to(torch.bfloat16).to(DEV)
    dV = torch.full((Bn * h, Nk, 96), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.gemm(P, dO, dV, Nk, 96, Nq, Nkp, h * 96, 96, 1, 0, impl=TC, batch=Bn * h, strideA=Nq * Nkp, strideB=Nq * h * 96,
             strideC=Nk * 96, b_inner=h, strideB_inner=96)
    ref = torch.einsum("bqk,bqd->bkd", P.float()[:, :, :Nk], dO.float().permute(0, 2, 1, 3).reshape(Bn * h, Nq, 96))
    assert not torch.isnan(dV.float()).any()
    assert max_rel_err(cpu(dV), cpu(ref)) < 6e-3


def test_gemm_tc_matches_simt_on_model_shapes():
    """Every forward GEMM shape of configs/ssv2.yaml (M = tokens of one clip)."""
    gen = torch.Generator().manual_seed(29)
    shapes = [(25153, 288, 96), (25153, 96, 96), (25153, 384, 96), (25153, 96, 384), (25153, 576, 96),
              (6337, 192, 192), (6337, 768, 192), (6337, 192, 768), (6337, 1152, 192), (1633, 384, 384),
              (1633, 1536, 384), (1633, 384, 1536), (1633, 2304, 384), (457, 768, 768), (457, 3072, 768),
              (457, 768, 3072), (25088, 96, 448)]
    for (M, N, K) in shapes:
        A = torch.randn(M, K, generator=gen).to(torch.bfloat16).to(DEV)
        W = (torch.randn(N, K, generator=gen) * K ** -0.5).to(torch.bfloat16).to(DEV)
        bias = torch.randn(N, generator=gen).to(DEV)
        o1 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        o2 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        ops.gemm(A, W, o1, M, N, K, K, K, N, 0, 1, bias=bias, impl=TC)
        ops.gemm(A, W, o2, M, N, K, K, K, N, 0, 1, bias=bias, impl=ops.IMPL_SIMT)
        assert max_rel_err(cpu(o1), cpu(o2)) < 1e-2, (M, N, K)


def _attn_inputs(B, h, q_thw, k_thw, O, seed, rel_std=0.2):
    from svit_b200 import msa
    gen = torch.Generator().manual_seed(seed)
    Nq = 1 + q_thw[0] * q_thw[1] * q_thw[2] + O
    Nk = 1 + k_thw[0] * k_thw[1] * k_thw[2] + O
    q = torch.randn(B, h, Nq, 96, generator=gen).to(torch.bfloat16).to(DEV)
    k = torch.randn(B, h, Nk, 96, generator=gen).to(torch.bfloat16).to(DEV)
    v = torch.randn(B, h, Nk, 96, generator=gen).to(torch.bfloat16).to(DEV)
    rels = [(rel_std * torch.randn(2 * max(a, b) - 1, 96, generator=gen)).to(DEV)
            for a, b in ((q_thw[1], k_thw[1]), (q_thw[2], k_thw[2]), (q_thw[0], k_thw[0]))]
    R = [msa.gathered_rel_pos(r, a, b) for r, (a, b) in
         zip(rels, ((q_thw[1], k_thw[1]), (q_thw[2], k_thw[2]), (q_thw[0], k_thw[0])))]
    tabs = [r.to(torch.bfloat16) for r in rels]
    tc_tables = (torch.cat(tabs).contiguous(), [t.shape[0] for t in tabs],
                 msa._index32_on(q.device, q_thw[1], k_thw[1]), msa._index32_on(q.device, q_thw[2], k_thw[2]),
                 msa._index32_on(q.device, q_thw[0], k_thw[0]), msa.key_column_codes(k_thw, O, q.device),
                 msa.key_select_table(k_thw, O, q.device))
    return q, k, v, R, tc_tables


ATTN_SHAPES = [
    # B, h, q_thw, k_thw, O      (tiny / ragged, then every ssv2.yaml stage shape at B=1)
    (2, 2, (2, 8, 8), (2, 2, 2), 8), (1, 1, (2, 4, 4), (2, 4, 4), 8), (1, 3, (3, 7, 7), (3, 4, 4), 12),
    (1, 1, (1, 9, 9), (1, 3, 3), 4), (2, 1, (2, 9, 9), (2, 5, 5), 8),
    (1, 1, (8, 56, 56), (8, 7, 7), 64), (1, 2, (8, 28, 28), (8, 14, 14), 64), (1, 2, (8, 28, 28), (8, 7, 7), 64),
    (1, 4, (8, 14, 14), (8, 14, 14), 64), (2, 4, (8, 14, 14), (8, 7, 7), 64), (1, 8, (8, 7, 7), (8, 14, 14), 64),
    (2, 8, (8, 7, 7), (8, 7, 7), 64),
    # 312^2-style key grids (kw 10 / 20 fast paths), 32-frame object tail (two object-key tiles), frame mode (kt 1)
    (1, 2, (4, 20, 20), (4, 10, 10), 128), (1, 1, (3, 10, 10), (3, 20, 20), 12), (2, 2, (1, 14, 14), (1, 7, 7), 4),
    (1, 1, (16, 14, 14), (16, 14, 14), 128),
]


@pytest.mark.parametrize("bias_in_mma", [True, False])
@pytest.mark.parametrize("shape", ATTN_SHAPES)
def test_attention_tc_matches_simt(shape, bias_in_mma):
    """bias_in_mma: the kernel of attn_tc3.cu (selection table present and kh + kw + kt <= 31), otherwise the
    kernel of attn_tc.cu (bias added by the softmax warps)."""
    B, h, q_thw, k_thw, O = shape
    q, k, v, R, tc_tables = _attn_inputs(B, h, q_thw, k_thw, O, seed=31)
    if bias_in_mma and tc_tables[6] is None:
        pytest.skip("bias columns do not fit the 32-column selection table")
    if not bias_in_mma:
        tc_tables = tc_tables[:6]
    scale = 96 ** -0.5
    ops.set_impl(attn=ops.IMPL_SIMT)
    with torch.no_grad():
        ref = ops.attention(q.float(), k.float(), v.float(), R[0], R[1], R[2], q_thw, k_thw, O, scale)
    ops.set_impl(attn=TC)
    try:
        with torch.no_grad():
            got = ops.attention(q, k, v, R[0], R[1], R[2], q_thw, k_thw, O, scale, tc_tables)
        torch.cuda.synchronize()
    finally:
        ops.set_impl(attn=ops.IMPL_AUTO)
    err = max_rel_err(cpu(got), cpu(ref))
    assert err < 1.5e-2, (shape, err)


@pytest.mark.parametrize("Bn,h,Nq,Nk", [(2, 4, 1633, 457), (1, 1, 300, 1633), (3, 2, 130, 73)])
def test_gemm_tc_batched_ragged_scores(Bn, h, Nq, Nk):
    """Batched score-shaped GEMMs of the attention backward: N (= key count) not a multiple of 8 (pad columns of the
    fp32 output are written as zeros), A a (sample, head) slice of the head-merged dO, alpha scaling, and a K-major A
    whose reduction extent (= key count) is ragged."""
    gen = torch.Generator().manual_seed(Nq * 3 + Nk)
    Nkp = (Nk + 7) // 8 * 8
    dO = torch.randn(Bn, Nq, h, 96, generator=gen).to(torch.bfloat16).to(DEV)
    V = torch.randn(Bn * h, Nk, 96, generator=gen).to(torch.bfloat16).to(DEV)
    dP = torch.full((Bn * h, Nq, Nkp), float("nan"), dtype=torch.float32, device=DEV)
    ops.gemm(dO, V, dP, Nq, Nk, 96, h * 96, 96, Nkp, 0, 1, impl=TC, batch=Bn * h, strideA=Nq * h * 96, strideB=Nk * 96,
             strideC=Nq * Nkp, a_inner=h, strideA_inner=96, alpha=0.5)
    ref = 0.5 * torch.einsum("bqd,bkd->bqk", dO.float().permute(0, 2, 1, 3).reshape(Bn * h, Nq, 96), V.float())
    assert not torch.isnan(dP).any()
    assert max_rel_err(cpu(dP[:, :, :Nk]), cpu(ref)) < 1e-5
    assert float(dP[:, :, Nk:].abs().max()) == 0.0 if Nkp > Nk else True
    # dQ-shaped: A = dS [Nq, Nk] with ragged K, B = k stored [Nk, 96] (MN-major)
    dS = (torch.randn(Bn * h, Nq, Nkp, generator=gen) / Nk ** 0.5).to(torch.bfloat16).to(DEV)
    dS[:, :, Nk:] = 0
    dQ = torch.full((Bn * h, Nq, 96), float("nan"), dtype=torch.float32, device=DEV)
    ops.gemm(dS, V, dQ, Nq, 96, Nk, Nkp, 96, 96, 0, 0, impl=TC, batch=Bn * h, strideA=Nq * Nkp, strideB=Nk * 96,
             strideC=Nq * 96, alpha=2.0)
    ref = 2.0 * torch.einsum("bqk,bkd->bqd", dS.float()[:, :, :Nk], V.float())
    assert max_rel_err(cpu(dQ), cpu(ref)) < 1e-5


BWD_SHAPES = [ATTN_SHAPES[i] for i in (0, 2, 3, 4, 6, 9, 10, 12, 14)]


@pytest.mark.parametrize("unfused", [False, True])
@pytest.mark.parametrize("shape", BWD_SHAPES)
def test_attention_backward_tc_matches_simt(shape, unfused):
    """attn_bwd_tc.cu (batched tcgen05 GEMMs + streaming softmax/dS kernels, bf16) against the CUDA-core fp32
    backward on the same bf16-rounded inputs: dq, dk, dv and the three rel-pos table gradients."""
    B, h, q_thw, k_thw, O = shape
    q, k, v, R, _ = _attn_inputs(B, h, q_thw, k_thw, O, seed=47)
    scale = 96 ** -0.5
    gen = torch.Generator().manual_seed(5)
    Nq = q.shape[2]
    dout = torch.randn(B, Nq, h * 96, generator=gen).to(torch.bfloat16).to(DEV)

    def run(dtype, impl):
        ops.set_impl(attn=impl)
        ops._state["attn_bwd_unfused"] = unfused  # fp32 S / dP scratch + streaming softmax instead of attn_bwd_sdp.cu
        try:
            ins = [t.detach().to(dtype).requires_grad_(True) for t in (q, k, v)]
            Rs = [r.detach().to(torch.bfloat16).float().requires_grad_(True) for r in R]
            out = ops.attention(ins[0], ins[1], ins[2], Rs[0], Rs[1], Rs[2], q_thw, k_thw, O, scale)
            out.backward(dout.to(dtype))
            torch.cuda.synchronize()
            return [cpu(t.grad) for t in ins + Rs]
        finally:
            ops.set_impl(attn=ops.IMPL_AUTO)
            ops._state["attn_bwd_unfused"] = False

    ref = run(torch.float32, ops.IMPL_SIMT)
    got = run(torch.bfloat16, ops.IMPL_AUTO)
    for name, g, r in zip(("dq", "dk", "dv", "dRh", "dRw", "dRt"), got, ref):
        assert torch.isfinite(g).all(), name
        err = max_rel_err(g, r)
        # table gradients sum dE = -(dS over the few cls / object keys) over thousands of rows: heavy cancellation on
        # top of the bf16 rounding of dS, hence the wider bound
        assert err < (5e-2 if name.startswith("dR") else 2e-2), (shape, name, err)


TAB_SHAPES = [ATTN_SHAPES[i] for i in (5, 7, 8, 9, 10, 11, 12, 14)]  # 237 / 125 / 69 / 41 / 85 table rows; the last one (Nk = 54) falls back to the gathered path


@pytest.mark.parametrize("shape", TAB_SHAPES)
def test_attention_backward_table_space_gradient(shape):
    """Rel-pos gradient in table-row space (attn_bwd_tc.cu: G scatter, dq += G . T and dT = G^T . q as tcgen05 GEMMs;
    attention.py:116-119 is the gather it differentiates through) against the CUDA-core fp32 backward that goes through
    the gathered tables, with the gradients compared ON THE PARAMETERS (the three un-gathered tables)."""
    from svit_b200 import msa
    B, h, q_thw, k_thw, O = shape
    q, k, v, _, tc_tables = _attn_inputs(B, h, q_thw, k_thw, O, seed=53)
    scale = 96 ** -0.5
    pairs = ((q_thw[1], k_thw[1]), (q_thw[2], k_thw[2]), (q_thw[0], k_thw[0]))
    gen = torch.Generator().manual_seed(7)
    Nq, Nk = q.shape[2], k.shape[2]
    dout = torch.randn(B, Nq, h * 96, generator=gen).to(torch.bfloat16).to(DEV)
    base = [t.float() for t in torch.split(tc_tables[0], tc_tables[1])]  # bf16-representable table values

    def run(dtype, impl, table_space):
        ops.set_impl(attn=impl)
        try:
            ins = [t.detach().to(dtype).requires_grad_(True) for t in (q, k, v)]
            rels = [t.detach().clone().requires_grad_(True) for t in base]
            Rs = [msa.gathered_rel_pos(r, a, b) for r, (a, b) in zip(rels, pairs)]
            if table_space and ops.attention_tables_only(dtype, q_thw, k_thw, O, sum(tc_tables[1])):
                Rs = [None, None, None]  # what MultiScaleAttention passes in training: no gathered tables at all
            tabs = tc_tables if dtype == torch.bfloat16 else None
            out = ops.attention(ins[0], ins[1], ins[2], Rs[0], Rs[1], Rs[2], q_thw, k_thw, O, scale, tabs,
                                torch.cat(rels) if table_space else None)
            out.backward(dout.to(dtype))
            torch.cuda.synchronize()
            return [cpu(t.grad) for t in ins + rels]
        finally:
            ops.set_impl(attn=ops.IMPL_AUTO)

    ref = run(torch.float32, ops.IMPL_SIMT, False)
    gathered = run(torch.bfloat16, ops.IMPL_AUTO, False)
    got = run(torch.bfloat16, ops.IMPL_AUTO, True)
    for name, g, r, o in zip(("dq", "dk", "dv", "d_rel_h", "d_rel_w", "d_rel_t"), got, ref, gathered):
        assert torch.isfinite(g).all(), name
        err, err_g = max_rel_err(g, r), max_rel_err(o, r)
        print(f"{shape} {name}: table-space {err:.2e}, gathered {err_g:.2e}")
        # dq / dk / dv: 2e-2, or -- where the bf16 rounding of dS alone already costs the gathered path more than that
        # (measured: 2.5e-2 on the 4 x 10 x 10 key grid with 128 object keys) -- no worse than the gathered path, 3e-2 at most
        # table gradients: 5e-2 as in the gathered-path test, or no worse than the gathered path where that one is beyond
        # it already (frame mode, kt = 1: ONE table row collects the heavily cancelling sum over every query row, 7.1e-2)
        if name.startswith("d_rel"):
            bound = max(5e-2, min(1e-1, 1.05 * err_g))
        else:
            bound = max(2e-2, min(3e-2, 1.05 * err_g))
        assert err < bound, (shape, name, err, err_g)


# ------------------------------------------------------------------------------------------------ folded LayerNorm
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_row_stats(dtype):
    gen = torch.Generator().manual_seed(41)
    for (M, Cn) in ((1, 96), (777, 96), (1000, 192), (333, 384), (130, 768), (257, 128), (64, 512)):
        if dtype == torch.float32 and Cn > 512:
            continue  # at most 128 16-byte vectors per row
        x = (torch.randn(M, Cn, generator=gen) * 2.0 + 0.7 * torch.randn(M, 1, generator=gen)).to(dtype)
        st = cpu(ops.row_stats(x.to(DEV), 1e-6))
        xd = x.double()
        mean = xd.mean(1)
        rstd = 1.0 / torch.sqrt(xd.var(1, unbiased=False) + 1e-6)
        assert max_rel_err(st[:, 0], mean) < 1e-5, (M, Cn)
        assert max_rel_err(st[:, 1], rstd) < 1e-5, (M, Cn)


@pytest.mark.parametrize("M,N,K,act", [(300, 288, 96, 0), (1000, 384, 96, 1), (640, 576, 192, 0), (520, 1536, 384, 1),
                                        (4100, 1152, 384, 0), (2304, 3072, 768, 1), (2100, 2304, 768, 0)])
def test_gemm_folded_layernorm(M, N, K, act):
    """LayerNorm(x) W^T + b through the folded epilogue (rstd * (x W'^T - mean * colsum(W')) + b') against an fp64
    LayerNorm followed by the product with the fp32 weight (attention.py:558-561, 566-567)."""
    gen = torch.Generator().manual_seed(43 + N)
    x = (torch.randn(M, K, generator=gen) * 1.5 + 0.5 * torch.randn(M, 1, generator=gen)).to(torch.bfloat16)
    W = torch.randn(N, K, generator=gen) * 0.05
    b = torch.randn(N, generator=gen) * 0.1
    gamma = 1.0 + 0.2 * torch.randn(K, generator=gen)
    beta = 0.1 * torch.randn(K, generator=gen)
    Wp, bp, gp, btp = (torch.nn.Parameter(t.to(DEV)) for t in (W, b, gamma, beta))
    with torch.no_grad():
        xs = x.to(DEV)
        st = ops.row_stats(xs, 1e-6)
        if act:
            w2 = torch.nn.Parameter(torch.eye(N, device=DEV)[:96].contiguous())
            b2 = torch.nn.Parameter(torch.zeros(96, device=DEV))
            y = ops.mlp_ln(xs, st, gp, btp, Wp, bp, w2, b2)
        else:
            y = ops.linear_ln(xs, st, gp, btp, Wp, bp)
    xn = torch.nn.functional.layer_norm(x.double(), (K,), gamma.double(), beta.double(), 1e-6)
    ref = xn @ W.double().t() + b.double()
    if act:
        ref = torch.nn.functional.gelu(ref).to(torch.bfloat16).double()[:, :96]
    err = max_rel_err(cpu(y), ref)
    assert err < 8e-3, (M, N, K, act, err)


def test_folded_weight_cache_follows_parameter_updates():
    W = torch.nn.Parameter(torch.randn(96, 96, device=DEV) * 0.05)
    g = torch.nn.Parameter(torch.ones(96, device=DEV))
    bt = torch.nn.Parameter(torch.zeros(96, device=DEV))
    a = ops.folded_ln_weight(W, None, g, bt, torch.bfloat16)
    assert ops.folded_ln_weight(W, None, g, bt, torch.bfloat16) is a
    with torch.no_grad():
        g.mul_(2.0)
    b = ops.folded_ln_weight(W, None, g, bt, torch.bfloat16)
    assert b is not a and max_rel_err(cpu(b[0]), 2.0 * cpu(a[0])) < 1e-6


# ------------------------------------------------------------------------------------------------ fused MLP
@pytest.mark.parametrize("Cn,M,mode", [(96, 128, "res"), (96, 1000, "res"), (96, 40000, "none"), (96, 257, "ln"),
                                       (96, 30000, "ln"), (96, 5000, "self"), (96, 129, "ln_nores"), (192, 128, "res"),
                                       (192, 777, "none"), (192, 50000, "self"), (192, 19000, "res")])
def test_mlp_fused_matches_reference(Cn, M, mode):
    """svit_mlp_fused (hidden activation kept on chip, optional LayerNorm prologue) against fp64 LayerNorm -> fc1 ->
    exact-erf GELU -> fc2 (+ residual) of the same bf16 operands (common.py:27-34, attention.py:566-570), and against the
    LayerNorm + two-GEMM path it replaces."""
    gen = torch.Generator().manual_seed(51 + M + Cn)
    Hd = 4 * Cn
    x = (torch.randn(M, Cn, generator=gen) * 1.3 + 0.4 * torch.randn(M, 1, generator=gen)).to(torch.bfloat16)
    w1 = (torch.randn(Hd, Cn, generator=gen) * Cn ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn(Cn, Hd, generator=gen) * Hd ** -0.5).to(torch.bfloat16)
    b1 = torch.randn(Hd, generator=gen) * 0.2
    b2 = torch.randn(Cn, generator=gen) * 0.2
    gamma = 1.0 + 0.2 * torch.randn(Cn, generator=gen)
    beta = 0.1 * torch.randn(Cn, generator=gen)
    res = torch.randn(M, Cn, generator=gen).to(torch.bfloat16) if mode == "res" else None
    dev = lambda t: None if t is None else t.to(DEV)
    xd, w1d, w2d, b1d, b2d = dev(x), dev(w1), dev(w2), dev(b1), dev(b2)
    use_ln = mode in ("ln", "ln_nores")
    with torch.no_grad():
        assert ops.mlp_fused_applicable(xd, w1d, w2d)
        rd = dev(res) if mode == "res" else (xd if mode in ("ln", "self") else None)
        got = ops.mlp_fused(xd, w1d, b1d, w2d, b2d, rd, ln=(dev(gamma), dev(beta), 1e-6) if use_ln else None)
        ops._MLP_FUSED["enabled"] = False
        try:
            xin = ops.layer_norm(xd, dev(gamma), dev(beta), 1e-6) if use_ln else xd
            two = ops.mlp(xin, w1d, b1d, w2d, b2d, rd)
        finally:
            ops._MLP_FUSED["enabled"] = True
    torch.cuda.synchronize()
    xin = x.double()
    if use_ln:
        xin = torch.nn.functional.layer_norm(xin, (Cn,), gamma.double(), beta.double(), 1e-6).to(torch.bfloat16).double()
    hid = torch.nn.functional.gelu(xin @ w1.double().t() + b1.double()).to(torch.bfloat16).double()
    ref = hid @ w2.double().t() + b2.double()
    if mode == "res":
        ref = ref + res.double()
    elif mode in ("ln", "self"):
        ref = ref + x.double()
    err, err2 = max_rel_err(cpu(got), ref), max_rel_err(cpu(got), cpu(two))
    assert err < 8e-3 and err2 < 8e-3, (M, mode, err, err2)
