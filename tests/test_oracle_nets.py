"""CPU: sanity of the (parity-unpinned) TF-graph restatement: fp32 vs fp64 agreement bounds the
round-off budget, gradients agree with finite differences, TF-RMSProp semantics (ms0=1, eps in sqrt)."""
import numpy as np
import pytest
import torch

from helpers import random_params
from oracle import nets
from oracle.cacc import chain_masks


def _rollout_inputs(variant, T=5, B=2, seed=0):
    rs = np.random.RandomState(seed)
    mask, _ = chain_masks(8)
    n_s_ls = [5] * 8 if variant != 'ia2c' else [10, 15, 15, 15, 15, 15, 15, 10]
    obs = [[rs.randn(B, n).astype(np.float32) for n in n_s_ls] for _ in range(T)]
    ps = rs.dirichlet(np.ones(4), size=(T, B, 8)).astype(np.float32)
    acts = rs.randint(0, 4, size=(T, B, 8))
    dones = np.zeros((T, B)); dones[0] = 1; dones[3, 0] = 1
    Rs, Advs = rs.randn(T, B, 8).astype(np.float32), rs.randn(T, B, 8).astype(np.float32)
    return mask, n_s_ls, obs, ps, acts, dones, Rs, Advs


@pytest.mark.parametrize('variant', nets.VARIANTS)
def test_fp32_close_to_fp64_and_grad_check(variant):
    mask, n_s_ls, obs, ps, acts, dones, Rs, Advs = _rollout_inputs(variant)
    params = random_params(nets.param_shapes(variant, n_s_ls, 4, mask), seed=3)
    p32 = nets.OraclePolicy(variant, n_s_ls, 4, mask, params=params, dtype=torch.float32, n_env=2)
    p64 = nets.OraclePolicy(variant, n_s_ls, 4, mask, params=params, dtype=torch.float64, n_env=2)
    s32 = p32.backward(obs, ps, acts, dones, Rs, Advs, 5e-4, apply=False)
    s64 = p64.backward(obs, ps, acts, dones, Rs, Advs, 5e-4, apply=False)
    assert abs(s32['total_loss'] - s64['total_loss']) < 1e-5
    np.testing.assert_allclose(p32.last_pi.numpy(), p64.last_pi.numpy(), atol=2e-6)
    np.testing.assert_allclose(p32.last_v.numpy(), p64.last_v.numpy(), atol=5e-6)
    # finite differences on a few fp64 weights
    name = [n for n in p64.names if n.endswith('w_ob') or n.endswith('fc/w')][2]
    g = p64.grads[name]
    for idx in [(0, 0), (3, 17)]:
        eps = 1e-6
        vals = []
        for sgn in (+1, -1):
            with torch.no_grad():
                p64.p[name][idx] += sgn * eps
            pi, v = p64.unroll(obs, ps, acts, dones, torch.zeros(2, 8, 128, dtype=torch.float64))
            pl, vl, el = p64.loss_terms(pi, v, acts, Rs, Advs, 0.5, 0.05)
            vals.append(float(pl.sum() + vl.sum() + el.sum()))
            with torch.no_grad():
                p64.p[name][idx] -= sgn * eps
        assert abs((vals[0] - vals[1]) / (2 * eps) - float(g[idx])) < 1e-6


def test_tf_rmsprop_semantics():
    mask, n_s_ls, obs, ps, acts, dones, Rs, Advs = _rollout_inputs('ma2c_ic3')
    params = random_params(nets.param_shapes('ma2c_ic3', n_s_ls, 4, mask), seed=3)
    p = nets.OraclePolicy('ma2c_ic3', n_s_ls, 4, mask, params=params, n_env=2)
    name = 'ic3/pi_0/w'
    w0 = p.p[name].detach().clone()
    s = p.backward(obs, ps, acts, dones, Rs, Advs, 1e-2, max_grad_norm=0.5)
    gn = s['grad_norm'][0]
    g = p.grads[name] * (0.5 / max(gn, 0.5))
    ms = 0.99 * 1.0 + 0.01 * g * g                      # slot starts at ONE
    torch.testing.assert_close(p.p[name].detach(), w0 - 1e-2 * g / torch.sqrt(ms + 1e-5))
    assert gn > 0.5                                       # the clip was active


def test_ia2c_clips_per_agent():
    mask, n_s_ls, obs, ps, acts, dones, Rs, Advs = _rollout_inputs('ia2c')
    p = nets.OraclePolicy('ia2c', n_s_ls, 4, mask, params=random_params(nets.param_shapes('ia2c', n_s_ls, 4, mask)), n_env=2)
    s = p.backward(obs, None, acts, dones, Rs, Advs, 5e-4, apply=False)
    assert len(s['grad_norm']) == 8 and len(set(np.round(s['grad_norm'], 6))) == 8
