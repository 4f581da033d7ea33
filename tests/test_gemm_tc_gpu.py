"""GPU: the tcgen05 / TMEM / TMA projection GEMMs (csrc/gemm_tc.cu) through the C ABI against fp64 matmuls of the same
bf16 operands.  bf16 inputs, fp32 accumulation: an fp32-output result must agree to fp32 summation noise (2e-5 of the
largest element); a bf16 output to one bf16 rounding (2^-8 relative to the largest element).  Shapes: every projection of
the shipped network (tokens x C with C in 48..1536, the 144-wide x_proj) plus ragged M / N / K tails and strided views."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K)
    (300, 96, 96), (25600, 96, 96), (1000, 48, 48), (6400, 192, 96), (4000, 1536, 768), (4000, 768, 1536),
    (3400, 144, 96), (777, 96, 48), (129, 256, 128), (128, 8, 8), (20480, 16, 48), (1600, 16, 384), (5000, 384, 192), (1111, 200, 72), (4100, 768, 384),
]


def _mk(M, N, K, seed=0):
    g = torch.Generator().manual_seed(seed + M + N + K)
    x = (torch.randn(M, K, generator=g)).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g)
    return x, w, b


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_fwd_matches_fp64(M, N, K):
    from mlagg_unet_b200 import gemm
    x, w, b = _mk(M, N, K)
    ref = x.double() @ w.double().t() + b.double()
    y32, _ = gemm.linear_fwd(x.cuda(), w.cuda(), b.cuda(), out_dtype=torch.float32)
    assert rel_err(y32.cpu(), ref) < 2e-5
    y16, _ = gemm.linear_fwd(x.cuda(), w.cuda(), b.cuda())
    assert y16.dtype == torch.bfloat16 and rel_err(y16.float().cpu(), ref) < 2 ** -8
    ynb, _ = gemm.linear_fwd(x.cuda(), w.cuda(), None, out_dtype=torch.float32)
    assert rel_err(ynb.cpu(), ref - b.double()) < 2e-5


@pytest.mark.parametrize("act", ["gelu", "silu"])
def test_linear_fwd_activation_and_preactivation(act):
    from mlagg_unet_b200 import gemm
    M, N, K = 2500, 192, 96
    x, w, b = _mk(M, N, K)
    pre_ref = x.double() @ w.double().t() + b.double()
    f = torch.nn.functional.gelu if act == "gelu" else torch.nn.functional.silu
    y, pre = gemm.linear_fwd(x.cuda(), w.cuda(), b.cuda(), act=act, out_dtype=torch.float32, want_pre=True)
    assert rel_err(y.cpu(), f(pre_ref)) < 2e-5
    assert pre.dtype == torch.bfloat16 and rel_err(pre.float().cpu(), pre_ref) < 2 ** -8
    # bf16 output: the persistent kernel (two TMA-store passes over the accumulator: pre-activation, then activation)
    for (M2, N2, K2) in ((M, N, K), (4100, 768, 384), (900, 1536, 768)):
        x, w, b = _mk(M2, N2, K2)
        pre_ref = x.double() @ w.double().t() + b.double()
        y16, pre16 = gemm.linear_fwd(x.cuda(), w.cuda(), b.cuda(), act=act, want_pre=True)
        assert rel_err(y16.float().cpu(), f(pre_ref)) < 2 ** -8
        assert rel_err(pre16.float().cpu(), pre_ref) < 2 ** -8


def test_linear_on_strided_views_in_place():
    """a channel slice of a wider activation as input, a channel slice of a wider buffer as output"""
    from mlagg_unet_b200 import gemm
    M, N, K = 1500, 96, 48
    x, w, b = _mk(M, N, K)
    wide = torch.randn(M, 2 * K).bfloat16()
    wide[:, K:] = x
    xc = wide.cuda()[:, K:]
    out = torch.full((M, 3 * N), 7.0, device="cuda", dtype=torch.bfloat16)
    gemm.linear_fwd(xc, w.cuda(), b.cuda(), out=out[:, N:2 * N])
    ref = x.double() @ w.double().t() + b.double()
    assert rel_err(out[:, N:2 * N].float().cpu(), ref) < 2 ** -8
    assert bool((out[:, :N] == 7).all()) and bool((out[:, 2 * N:] == 7).all())


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_bwd_data_and_weight_match_fp64(M, N, K):
    from mlagg_unet_b200 import gemm
    x, w, _ = _mk(M, N, K)
    g = torch.Generator().manual_seed(5)
    dy = torch.randn(M, N, generator=g).bfloat16()
    dx = gemm.linear_bwd_data(dy.cuda(), w.cuda(), out_dtype=torch.float32)
    assert rel_err(dx.cpu(), dy.double() @ w.double()) < 2e-5
    dx16 = gemm.linear_bwd_data(dy.cuda(), w.cuda())
    assert rel_err(dx16.float().cpu(), dy.double() @ w.double()) < 2 ** -8
    dw = gemm.linear_bwd_weight(dy.cuda(), x.cuda())
    assert dw.dtype == torch.float32 and rel_err(dw.cpu(), dy.double().t() @ x.double()) < 2e-5
    dw2, db = gemm.linear_bwd_weight(dy.cuda(), x.cuda(), want_db=True)      # bias gradient from the ones-tile MMA
    assert rel_err(dw2.cpu(), dy.double().t() @ x.double()) < 2e-5
    assert db.shape == (N,) and rel_err(db.cpu(), dy.double().sum(0)) < 2e-5


@pytest.mark.parametrize("act", ["gelu", "silu"])
def test_linear_bwd_data_applies_the_activation_gradient(act):
    from mlagg_unet_b200 import gemm
    M, N, K = 1300, 96, 192
    _, w, _ = _mk(M, N, K)
    g = torch.Generator().manual_seed(6)
    dy = torch.randn(M, N, generator=g).bfloat16()
    pre = torch.randn(M, K, generator=g).bfloat16()
    p = pre.double().requires_grad_()
    f = torch.nn.functional.gelu if act == "gelu" else torch.nn.functional.silu
    (dact,) = torch.autograd.grad(f(p).sum(), p)
    ref = (dy.double() @ w.double()) * dact
    dx = gemm.linear_bwd_data(dy.cuda(), w.cuda(), aux=pre.cuda(), act=act, out_dtype=torch.float32)
    assert rel_err(dx.cpu(), ref) < 2e-5
    dx16 = gemm.linear_bwd_data(dy.cuda(), w.cuda(), aux=pre.cuda(), act=act)
    assert rel_err(dx16.float().cpu(), ref) < 2 ** -8


def test_linear_autograd_function_uses_the_tensor_core_path_under_autocast():
    """ops.linear_tokens under bf16 autocast == F.linear on the bf16-rounded operands (forward and all three gradients)"""
    from mlagg_unet_b200 import _lib
    from mlagg_unet_b200.ops import linear_tokens
    torch.manual_seed(0)
    lin = torch.nn.Linear(96, 192).cuda()
    x = torch.randn(4, 333, 96, device="cuda", requires_grad=True)
    _lib.STATS["launches"] = 0
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = linear_tokens(x, lin)
    dy = torch.randn_like(y)
    y.backward(dy)
    assert _lib.STATS["launches"] >= 3            # fwd, bwd_data, bwd_weight (+ bias gradient) went through the C ABI
    xr = x.detach().bfloat16().double().requires_grad_()
    wr = lin.weight.detach().bfloat16().double().requires_grad_()
    br = lin.bias.detach().double().requires_grad_()
    yr = torch.nn.functional.linear(xr, wr, br)
    yr.backward(dy.double())
    assert rel_err(y.float(), yr) < 2 ** -8
    assert rel_err(x.grad, xr.grad) < 2 ** -7
    assert rel_err(lin.weight.grad, wr.grad) < 2 ** -8
    assert rel_err(lin.bias.grad, br.grad) < 2 ** -8
