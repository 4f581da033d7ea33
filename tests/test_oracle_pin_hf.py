"""CPU: pins the scan / causal-conv1d oracles to an INDEPENDENT third-party implementation that ships in this image:
Hugging Face transformers' `MambaMixer.slow_forward` (transformers 5.x, models/mamba/modeling_mamba.py) -- the pure-torch
path transformers falls back to when mamba-ssm's CUDA kernels are absent, written from the same published S6 definition
as mamba-ssm's `selective_scan_ref` (softplus(dt_proj(.)) discretisation, exp(delta A) state decay, delta B u input,
D skip, SiLU(z) gate) behind a causal depthwise conv1d + SiLU.  The reference repo calls the mamba-ssm / causal-conv1d
packages for this arithmetic and vendors neither; this is the closest executable restatement by a
party other than this repo.  The whole mixer is rebuilt from oracle pieces on the same weights and compared forward and
backward (input and every parameter gradient), fp32 arithmetic against the fp64 C oracle, tolerance 2e-5.
Skipped when transformers (or its Mamba model) cannot be imported."""
import pytest
import torch

from conftest import rel_err
from oracle.convs import causal_conv1d
from oracle.scan import selective_scan_oracle

TOL = 2e-5


def _mixer(hidden=12, state=16, expand=2, conv_kernel=4, rank=3, seed=0):
    try:
        from transformers import MambaConfig
        from transformers.models.mamba.modeling_mamba import MambaMixer
    except Exception as e:  # pragma: no cover
        pytest.skip(f"transformers Mamba unavailable: {e}")
    torch.manual_seed(seed)
    cfg = MambaConfig(hidden_size=hidden, state_size=state, expand=expand, conv_kernel=conv_kernel, time_step_rank=rank,
                      num_hidden_layers=1, vocab_size=16, use_mambapy=False)
    m = MambaMixer(cfg, layer_idx=0).float()
    with torch.no_grad():   # move the parameters off their structured initial values
        m.A_log.add_(0.3 * torch.randn_like(m.A_log))
        m.D.add_(0.3 * torch.randn_like(m.D))
        m.dt_proj.bias.add_(0.5 * torch.randn_like(m.dt_proj.bias))
    return m


def _oracle_mixer(m, x):
    """MambaMixer rebuilt from oracle pieces: in_proj -> causal_conv1d + SiLU -> x_proj -> selective scan -> out_proj"""
    R, N = m.time_step_rank, m.ssm_state_size
    xz = torch.nn.functional.linear(x, m.in_proj.weight, m.in_proj.bias).transpose(1, 2)
    u, z = xz.chunk(2, dim=1)
    u = causal_conv1d(u, m.conv1d.weight[:, 0, :], m.conv1d.bias, silu=True)
    dbl = torch.nn.functional.linear(u.transpose(1, 2), m.x_proj.weight)
    dt, Bm, Cm = torch.split(dbl, [R, N, N], dim=-1)
    delta = torch.nn.functional.linear(dt, m.dt_proj.weight).transpose(1, 2)        # bias enters through delta_bias
    y = selective_scan_oracle(u, delta, -torch.exp(m.A_log.float()), Bm.transpose(1, 2).contiguous(),
                              Cm.transpose(1, 2).contiguous(), m.D.float(), z=z, delta_bias=m.dt_proj.bias.float(),
                              delta_softplus=True, fp64=True)
    return torch.nn.functional.linear(y.transpose(1, 2), m.out_proj.weight, m.out_proj.bias)


@pytest.mark.parametrize("L", [1, 5, 64, 203])
def test_oracle_mixer_matches_transformers_slow_path(L):
    m = _mixer(seed=L)
    x = torch.randn(2, L, m.hidden_size if hasattr(m, "hidden_size") else m.in_proj.in_features,
                    generator=torch.Generator().manual_seed(L))
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    ya = m.slow_forward(xa)
    g = torch.randn(ya.shape, generator=torch.Generator().manual_seed(99))
    params = [p for p in m.parameters()]
    ga = torch.autograd.grad(ya, [xa] + params, g, allow_unused=True)
    yb = _oracle_mixer(m, xb)
    gb = torch.autograd.grad(yb, [xb] + params, g, allow_unused=True)
    assert rel_err(yb, ya) < TOL
    names = ["input"] + [n for n, _ in m.named_parameters()]
    for n, a, b in zip(names, ga, gb):
        assert (a is None) == (b is None), n
        if a is not None:
            assert rel_err(b, a) < TOL, n
