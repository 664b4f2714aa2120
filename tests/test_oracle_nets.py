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
    n_s_ls = {'ia2c': [10, 15, 15, 15, 15, 15, 15, 10], 'ia2c_fp': [14, 23, 23, 23, 23, 23, 23, 14]}.get(variant, [5] * 8)
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
    name = [n for n in p64.names if n.endswith('w_ob') or n.endswith('fc/w') or n.endswith('fcs/w') or n.startswith('cu/fc_')][2]
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


def test_fp_policy_is_a_neurcomm_cell_with_null_message_encoder():
    """The identity the CUDA path relies on for ia2c_fp (deeprl_network_b200/layout.py): FPPolicy == NeurComm cell
    with w_msg = b_msg = 0 and wx_hid rows 128..191 = 0, fed with own+neighbour observations and fingerprints."""
    mask, n_s_ls, obs, ps, acts, dones, Rs, Advs = _rollout_inputs('ia2c_fp')
    fp_params = random_params(nets.param_shapes('ia2c_fp', n_s_ls, 4, mask), seed=5)
    nbr = [list(np.where(mask[i] == 1)[0]) for i in range(8)]
    nc_params = {n: np.zeros(s, dtype=np.float32) for n, s in nets.param_shapes('ma2c_nc', [5] * 8, 4, mask)}
    for i in range(8):
        s, d = 'lstm_%d/' % i, 'nc/lstm_comm_%d/' % i
        nc_params[d + 'w_ob'], nc_params[d + 'b_ob'] = fp_params[s + 'fcs/w'], fp_params[s + 'fcs/b']
        nc_params[d + 'w_fp'], nc_params[d + 'b_fp'] = fp_params[s + 'fcp/w'], fp_params[s + 'fcp/b']
        nc_params[d + 'wx_hid'][:128] = fp_params[s + 'lstm/wx']
        nc_params[d + 'wh_hid'], nc_params[d + 'b_hid'] = fp_params[s + 'lstm/wh'], fp_params[s + 'lstm/b']
        for h in ('pi', 'v'):
            for k in ('w', 'b'):
                nc_params['nc/%s_%d/%s' % (h, i, k)] = fp_params[s + '%s/%s' % (h, k)]
    # the FP observation is [own, neighbours' obs, neighbours' fingerprints]; NeurComm gets own obs + ps separately
    own = [[o[i][:, :5] for i in range(8)] for o in obs]
    T, B = ps.shape[:2]
    for t in range(T):
        for i in range(8):
            for k, j in enumerate(nbr[i]):
                obs[t][i][:, 5 * (k + 1):5 * (k + 2)] = own[t][j]
                obs[t][i][:, 5 * (1 + len(nbr[i])) + 4 * k:5 * (1 + len(nbr[i])) + 4 * (k + 1)] = ps[t][:, j]
    fp = nets.OraclePolicy('ia2c_fp', n_s_ls, 4, mask, params=fp_params, n_env=2)
    nc = nets.OraclePolicy('ma2c_nc', [5] * 8, 4, mask, params=nc_params, n_env=2)
    fp.backward(obs, None, acts, dones, Rs, Advs, 5e-4, apply=False)
    nc.backward(own, ps, acts, dones, Rs, Advs, 5e-4, apply=False)
    torch.testing.assert_close(fp.last_pi, nc.last_pi, rtol=0, atol=1e-6)
    torch.testing.assert_close(fp.last_v, nc.last_v, rtol=0, atol=1e-6)
    for i in range(8):
        torch.testing.assert_close(fp.grads['lstm_%d/lstm/wx' % i], nc.grads['nc/lstm_comm_%d/wx_hid' % i][:128], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(fp.grads['lstm_%d/fcp/w' % i], nc.grads['nc/lstm_comm_%d/w_fp' % i], rtol=1e-4, atol=1e-6)
        assert not nc.grads['nc/lstm_comm_%d/w_msg' % i].any() and not nc.grads['nc/lstm_comm_%d/b_msg' % i].any()
        assert not nc.grads['nc/lstm_comm_%d/wx_hid' % i][128:].any()         # the padding receives exact zeros


def test_consensus_update_is_a_neighbourhood_mean_of_lstm_weights():
    mask, n_s_ls, obs, ps, acts, dones, Rs, Advs = _rollout_inputs('ma2c_cu')
    p = nets.OraclePolicy('ma2c_cu', n_s_ls, 4, mask, params=random_params(nets.param_shapes('ma2c_cu', n_s_ls, 4, mask)), n_env=2)
    before = {n: v.detach().clone() for n, v in p.p.items()}
    with torch.no_grad():
        p.consensus_update()
    for key in ('wx', 'wh', 'b'):
        torch.testing.assert_close(p.p['cu/lstm_0a/' + key], (before['cu/lstm_0a/' + key] + before['cu/lstm_1a/' + key]) / 2)
        torch.testing.assert_close(p.p['cu/lstm_4a/' + key], (before['cu/lstm_4a/' + key] + before['cu/lstm_3a/' + key] + before['cu/lstm_5a/' + key]) / 3)
    for n in p.names:
        if '/lstm_' not in n:
            assert torch.equal(p.p[n], before[n]), n                          # fc, pi, v stay local
    s = p.backward(obs, ps, acts, dones, Rs, Advs, 5e-4)                       # one global clip group
    assert len(s['grad_norm']) == 1
