"""CPU, world_size 2 (gloo): the data-parallel contract of the N>1 path.  Envs shard across
ranks, the loss is scaled by 1/(T * B_total), ONE sum all-reduce of the flat gradient follows,
and clipping happens after the reduce -- so every rank ends with the same weights as a single
process holding all envs.  (Uses the CPU oracle for the math; the CUDA kernels honour the same
contract via nmarl_bwd_args.B_total, checked on the GPU in test_gpu_backward.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from test_oracle_nets import _rollout_inputs
    from helpers import random_params
    from oracle import nets
    torch.set_num_threads(1)
    B = 4
    mask, n_s_ls, obs, ps, acts, dones, Rs, Advs = _rollout_inputs('ma2c_nc', T=4, B=B, seed=1)
    params = random_params(nets.param_shapes('ma2c_nc', n_s_ls, 4, mask), seed=2)
    lo, hi = rank * B // world, (rank + 1) * B // world
    pol = nets.OraclePolicy('ma2c_nc', n_s_ls, 4, mask, params=params, n_env=hi - lo)
    obs_l = [[o[lo:hi] for o in ob] for ob in obs]
    pol.backward(obs_l, ps[:, lo:hi], acts[:, lo:hi], dones[:, lo:hi], Rs[:, lo:hi], Advs[:, lo:hi], 5e-4, apply=False)
    # local mean over (t, b_local) -> rescale to 1/(T*B_total) and sum-reduce one flat buffer
    flat = torch.cat([pol.grads[n].reshape(-1) for n in pol.names]) * ((hi - lo) / B)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    gn = float(torch.sqrt((flat ** 2).sum()))
    if rank == 0:
        full = nets.OraclePolicy('ma2c_nc', n_s_ls, 4, mask, params=params, n_env=B)
        s = full.backward(obs, ps, acts, dones, Rs, Advs, 5e-4, apply=False)
        ref = torch.cat([full.grads[n].reshape(-1) for n in full.names])
        out.put((float((flat - ref).abs().max()), float(ref.abs().max()), gn, s['grad_norm'][0]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradient_allreduce_equals_single_process():
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    err, scale, gn, gn_ref = q.get()
    assert err <= 1e-5 * scale + 1e-8
    assert abs(gn - gn_ref) <= 1e-4 * gn_ref
