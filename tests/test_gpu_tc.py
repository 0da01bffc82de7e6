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
