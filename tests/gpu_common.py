"""Helpers shared by the GPU parity tests: build a CUDA engine and an oracle policy holding the
same weights, and convert between the kernel layout [agent][env][..] and the oracle's [env][agent][..]."""
import numpy as np
import torch

from helpers import random_params
from oracle import nets
from oracle.cacc import chain_masks

N_S = {'ia2c': [10, 15, 15, 15, 15, 15, 15, 10]}
HP = dict(v_coef=0.5, e_coef=0.05, max_grad_norm=40.0, alpha=0.99, epsilon=1e-5, gamma=0.99,
          reward_norm=5000.0, reward_clip=-1.0)


def make_pair(variant, B, T=4, seed=0, dtype=torch.float32, mask=None, n_a=4, hp=None, scale=0.3):
    from deeprl_network_b200.agents.engine import PolicyEngine
    from deeprl_network_b200.layout import ModelLayout
    if mask is None:
        mask, _ = chain_masks(8)
    N = len(mask)
    nm = [int(mask[i].sum()) for i in range(N)]
    n_s_ls = {'ia2c': [5 * (1 + k) for k in nm], 'ia2c_fp': [5 * (1 + k) + n_a * k for k in nm]}.get(variant, [5] * N)
    lay = ModelLayout(variant, n_s_ls, n_a, mask, obs_mode='gather')
    params = random_params(lay.creation_order(), seed=seed, scale=scale)
    eng = PolicyEngine(lay, B, T, dict(HP if hp is None else hp), flat_params=lay.pack(params))
    orc = nets.OraclePolicy(variant, n_s_ls, n_a, mask, params=params, dtype=dtype, n_env=B)
    return eng, orc, lay, params


def oracle_obs(lay, base):
    """base [B, N, 5] own features -> per-agent oracle inputs (IA2C: own + neighbours concatenated)."""
    if lay.variant not in ('ia2c', 'ia2c_fp'):      # ia2c_fp: the fingerprints travel separately (ps)
        return [base[:, i] for i in range(lay.N)]
    return [np.concatenate([base[:, i]] + [base[:, j] for j in lay.nbr[i]], axis=1) for i in range(lay.N)]


def to_dev(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).cuda()


def nb(x):
    """[B, N, ...] -> contiguous [N, B, ...] device tensor."""
    return to_dev(np.swapaxes(x, 0, 1))


def bn(t):
    """device [N, B, ...] -> numpy [B, N, ...]."""
    return np.swapaxes(t.detach().cpu().numpy(), 0, 1)


def obs_dev(lay, base):
    B, N, _ = base.shape
    o = np.zeros((N, B, lay.obs_stride), dtype=np.float32)
    o[:, :, :5] = np.swapaxes(base, 0, 1)
    return to_dev(o)
