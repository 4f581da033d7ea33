"""GPU parity: sm_100a selective scan (through the C ABI via selective_scan_fn) vs the fp64 CPU oracle.
Tolerance: 1e-4 relative (max|diff| / max|ref|), the fp32 bar BASELINE.json:north_star states."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _inputs(Bn, D, L, G, N=16, seed=0, realistic=True):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    if realistic:
        A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(D, 1) * (1 + 0.1 * r(D, N))
        dt = torch.exp(torch.rand(D, generator=g) * (torch.log(torch.tensor(0.1)) - torch.log(torch.tensor(1e-3)))
                       + torch.log(torch.tensor(1e-3)))
        bias = dt + torch.log(-torch.expm1(-dt))
    else:
        A, bias = -torch.rand(D, N, generator=g) * 4 - 0.1, r(D)
    return dict(u=r(Bn, D, L), delta=r(Bn, D, L), A=A, B=r(Bn, G, N, L), C=r(Bn, G, N, L), D=r(D), delta_bias=bias)


def _run(t, softplus=True, use_D=True, use_bias=True, check_bwd=True):
    from mlagg_unet_b200.selective_scan_interface import selective_scan_fn
    from oracle.scan import scan_bwd_c, scan_fwd_c
    D = t["D"] if use_D else None
    bias = t["delta_bias"] if use_bias else None
    ref = scan_fwd_c(t["u"], t["delta"], t["A"], t["B"], t["C"], D, bias, softplus, fp64=True)
    names = ["u", "delta", "A", "B", "C"] + (["D"] if use_D else []) + (["delta_bias"] if use_bias else [])
    cu = {k: t[k].cuda().requires_grad_() for k in names}
    out, last = selective_scan_fn(cu["u"], cu["delta"], cu["A"], cu["B"], cu["C"], cu.get("D"), None,
                                  cu.get("delta_bias"), softplus, return_last_state=True)
    assert out.dtype == torch.float32 and out.shape == t["u"].shape
    assert rel_err(out.cpu(), ref) < TOL
    _, last_ref = scan_fwd_c(t["u"], t["delta"], t["A"], t["B"], t["C"], D, bias, softplus, fp64=True,
                             return_last_state=True)
    assert rel_err(last.cpu(), last_ref) < TOL
    if not check_bwd:
        return
    dy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(7))
    gref = scan_bwd_c(t["u"], t["delta"], t["A"], t["B"], t["C"], D, bias, softplus, dy, fp64=True)
    gref = dict(zip(["u", "delta", "A", "B", "C", "D", "delta_bias"], gref))
    out.backward(dy.cuda())
    for k in names:
        assert rel_err(cu[k].grad.cpu(), gref[k]) < TOL, k


@pytest.mark.parametrize("shape", [(2, 64, 256, 2), (1, 32, 64, 1), (2, 96, 1000, 3), (3, 40, 4096, 1),
                                   (2, 20, 128, 2), (1, 8, 16, 1), (1, 384, 1360, 4)])
def test_scan_matches_oracle_aligned(shape):
    _run(_inputs(*shape))


@pytest.mark.parametrize("L", [1, 3, 67, 130, 1001])
def test_scan_matches_oracle_ragged_lengths(L):
    """L % 4 != 0 takes the non-bulk loader; tails inside a 16-step chunk and a 64-step tile."""
    _run(_inputs(2, 24, L, 2, seed=L))


@pytest.mark.parametrize("softplus,use_D,use_bias", [(False, True, True), (True, False, True), (True, True, False),
                                                     (False, False, False)])
def test_scan_optional_arguments(softplus, use_D, use_bias):
    t = _inputs(2, 32, 200, 2, seed=3, realistic=False)
    if not softplus:
        t["delta"] = t["delta"].abs() * 0.05
        t["delta_bias"] = t["delta_bias"].abs() * 0.01
    _run(t, softplus, use_D, use_bias)


def test_scan_rank3_BC_and_z_gate():
    from mlagg_unet_b200.selective_scan_interface import selective_scan_fn
    from oracle.scan import selective_scan_loop
    t = _inputs(2, 16, 96, 1, seed=5)
    z = torch.randn(2, 16, 96)
    ref = selective_scan_loop(t["u"].double(), t["delta"].double(), t["A"].double(), t["B"][:, 0].double(),
                              t["C"][:, 0].double(), t["D"].double(), z.double(), t["delta_bias"].double(), True)
    out = selective_scan_fn(t["u"].cuda(), t["delta"].cuda(), t["A"].cuda(), t["B"][:, 0].cuda(), t["C"][:, 0].cuda(),
                            t["D"].cuda(), z.cuda(), t["delta_bias"].cuda(), True)
    assert rel_err(out.cpu(), ref) < TOL


def test_scan_large_softplus_inputs_and_linearity():
    """softplus threshold branch (x > 20) and a size-independent property: the scan is linear in u."""
    from mlagg_unet_b200.selective_scan_interface import selective_scan_fn
    t = _inputs(2, 32, 512, 2, seed=9)
    t["delta"][:, :, ::7] = 25.0
    _run(t, check_bwd=True)
    cu = {k: v.cuda() for k, v in t.items()}
    f = lambda u: selective_scan_fn(u, cu["delta"], cu["A"], cu["B"], cu["C"], cu["D"], None, cu["delta_bias"], True)
    u2 = torch.randn_like(cu["u"])
    assert rel_err(f(cu["u"] + 2 * u2), f(cu["u"]) + 2 * f(u2)) < 1e-5


def test_scan_cpu_tensor_fails_loudly():
    from mlagg_unet_b200._lib import MlaggError
    from mlagg_unet_b200.selective_scan_interface import selective_scan_fn
    t = _inputs(1, 8, 16, 1)
    with pytest.raises(MlaggError):
        selective_scan_fn(t["u"], t["delta"], t["A"], t["B"], t["C"])


def test_scan_config5_length_properties():
    """BASELINE config 5 (Endovis17-shaped 3x512x512): L_cat = 256^2 + 128^2 + 64^2 + 32^2 = 87040 per image.  Too long
    for the CPU oracle in test time, so size-independent properties: linearity in u, and agreement of the first 4096
    steps with the fp64 oracle run on that prefix (the scan is causal)."""
    import torch
    from mlagg_unet_b200.selective_scan_interface import selective_scan_fn
    from oracle.scan import scan_fwd_c
    Bn, D, G, N, L = 2, 64, 4, 16, 87040
    g = torch.Generator().manual_seed(5)
    u1, u2 = torch.randn(Bn, D, L, generator=g), torch.randn(Bn, D, L, generator=g)
    dl = torch.randn(Bn, D, L, generator=g)
    A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(D, 1)
    Bm, Cm = torch.randn(Bn, G, N, L, generator=g), torch.randn(Bn, G, N, L, generator=g)
    Dk, bias = torch.ones(D), torch.randn(D, generator=g) - 3.0
    run = lambda u: selective_scan_fn(u.cuda(), dl.cuda(), A.cuda(), Bm.cuda(), Cm.cuda(), Dk.cuda(), None, bias.cuda(), True)
    y1, y2, y12 = run(u1), run(u2), run(1.5 * u1 - 0.25 * u2)
    assert torch.isfinite(y12).all()
    assert float((y12 - (1.5 * y1 - 0.25 * y2)).abs().max() / y12.abs().max()) < 1e-4
    P = 4096
    ref = scan_fwd_c(u1[..., :P].contiguous(), dl[..., :P].contiguous(), A, Bm[..., :P].contiguous(),
                     Cm[..., :P].contiguous(), Dk, bias, True, fp64=True)
    assert float((y1[..., :P].cpu().double() - ref.double()).abs().max() / ref.abs().max()) < 1e-4


def test_forward_matches_mamba_ssm_derived_cuda_kernel():
    """Our forward against vLLM's `selective_scan_fwd` -- mamba-ssm's CUDA forward kernel carried into vLLM (the image
    has vLLM 0.22, not mamba-ssm) -- on the MSMM call's argument pattern (grouped B / C, D skip, delta_bias, softplus),
    output and final state, 1e-4 relative.  Runs `tools/mamba_kernel_compare.py --check` in a subprocess (a foreign
    extension stays out of this process); skipped when vLLM's op cannot be imported or rejects the call."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "mamba_kernel_compare.py"), "--check"],
                           capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired:
        pytest.skip("vllm import / call timed out")
    lines = [l for l in r.stdout.splitlines() if l.startswith("PARITY")]
    if not lines:
        pytest.skip("vllm selective scan unavailable: " + (r.stdout + r.stderr)[-300:])
    _, rel_out, rel_state = lines[-1].split()
    assert float(rel_out) < 1e-4 and float(rel_state) < 1e-4


@pytest.mark.parametrize("warps", ["1", "4"])
def test_forward_cta_shapes_agree_with_the_oracle(warps, monkeypatch):
    """The small-batch CTA shape (one scan warp + one helper warp, chosen automatically when the 4-warp grid would leave
    most SMs empty -- B = 1 sliding-window tiles) and the training shape compute the same thing: both against the fp64
    oracle on a ragged multi-tile shape, output, final state and checkpoints (through a backward pass)."""
    from mlagg_unet_b200.selective_scan_interface import selective_scan_fn
    from oracle.scan import scan_bwd_c, scan_fwd_c
    monkeypatch.setenv("MLAGG_SCAN_WARPS", warps)
    g = torch.Generator().manual_seed(17)
    Bn, D, G, N, L = 1, 96, 4, 16, 333
    u, dl = torch.randn(Bn, D, L, generator=g), 0.5 * torch.randn(Bn, D, L, generator=g)
    A = -torch.rand(D, N, generator=g) * 4 - 0.1
    Bm, Cm = torch.randn(Bn, G, N, L, generator=g), torch.randn(Bn, G, N, L, generator=g)
    Dk, bias = torch.randn(D, generator=g), 0.3 * torch.randn(D, generator=g)
    ref, last = scan_fwd_c(u, dl, A, Bm, Cm, Dk, bias, True, fp64=True, return_last_state=True)
    dout = torch.randn(Bn, D, L, generator=g)
    gref = scan_bwd_c(u, dl, A, Bm, Cm, Dk, bias, True, dout, fp64=True)
    ts = [t.cuda().requires_grad_() for t in (u, dl, A, Bm, Cm, Dk, bias)]
    out, st = selective_scan_fn(ts[0], ts[1], ts[2], ts[3], ts[4], ts[5], None, ts[6], True, return_last_state=True)
    assert rel_err(out.cpu(), ref) < 1e-4 and rel_err(st.cpu(), last) < 1e-4
    out.backward(dout.cuda())
    for t, r in zip(ts, gref):
        assert rel_err(t.grad.cpu(), r) < 1e-4
