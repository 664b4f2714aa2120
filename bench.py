"""bench.py -- agent-env-steps/sec of the CACC + A2C + NeurComm hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full update of the hot path over one batch: n_step (60) env steps of B parallel
CACC Catch-up episodes x 8 agents [p-call, sampling, v-call, env step], bootstrap, n-step returns,
training forward + BPTT + weight gradients, (NCCL all-reduce), clip + RMSProp.  Workload =
BASELINE.json configs[1]: config_ma2c_nc_catchup.ini, 4096 parallel envs per GPU (weak scaling).

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` runs
the same update with HOST buffers: the action uniforms (the reference draws them with the host
NumPy RNG) are copied from pinned memory every step and the per-step rewards + loss terms are
read back, copies inside the timed region.  `roofline` is for the fused step+message+cell
forward kernel (cell_fwd p-call), `cpu_baseline` times the restated reference (TF unavailable)
on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = 'agent-env-steps/sec CACC Catch-up NeurComm A2C'
UNIT = 'agent-env-steps/s'
CONFIG = 'config_ma2c_nc_catchup.ini'
N_ENV = 4096


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--n-env', type=int, default=N_ENV, help='parallel envs per GPU')
    ap.add_argument('--config', default=CONFIG)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


def load_cfg(name, **env_over):
    import configparser
    cp = configparser.ConfigParser()
    assert cp.read(os.path.join(ROOT, 'config', name)), name
    for k, v in env_over.items():
        cp['ENV_CONFIG'][k] = str(v)
    return cp


# ---- CPU arm: the restated reference trainer (oracle/) on the host cores -----------------------
def cpu_reference_best(cfg_name, updates):
    """The reference is single-process with TF's default thread pools; M=1 GEMVs gain nothing from
    threads, so time it with 1 thread and with all host threads and keep the faster."""
    import torch
    n_all = torch.get_num_threads()
    best = None
    for nt in sorted({1, n_all}):
        torch.set_num_threads(nt)
        v, cores, sample, dt = cpu_reference(cfg_name, updates=max(2, updates // 2), warm_updates=1)
        r = (v, nt, sample + ', torch threads=%d of %d host cores' % (nt, os.cpu_count()), dt / max(2, updates // 2))
        if best is None or v > best[0]:
            best = r
    torch.set_num_threads(n_all)
    return best


def cpu_reference(cfg_name, updates, warm_updates=1):
    """Times `updates` update cycles (n_step env steps each, B=1, per-agent Python loops, one
    forward per call -- the reference's structure) of the restated reference.  Returns
    (agent-env-steps/s, cores, sample description, seconds)."""
    import numpy as np
    import torch
    from oracle.cacc import OracleCACC
    from oracle.trainer import Counter, OracleAgent, OracleTrainer
    cp = load_cfg(cfg_name)
    env = OracleCACC(cp['ENV_CONFIG'])
    variant = cp['ENV_CONFIG']['agent']
    ag = OracleAgent(variant, env.n_s_ls, env.n_a_ls, env.neighbor_mask, env.distance_mask, env.coop_gamma, 10 ** 6,
                     cp['MODEL_CONFIG'], seed=12)
    tr = OracleTrainer(env, ag, Counter(10 ** 9, 10 ** 9, 10 ** 9))
    cores = torch.get_num_threads()
    done_updates, steps, t0 = 0, 0, None
    while done_updates < updates + warm_updates:
        ob = env.reset(); done = True; ag.reset(); tr.cur_step = 0; tr.episode_rewards = []
        while True:
            if done_updates == warm_updates and t0 is None:
                t0 = time.perf_counter(); steps = 0
            c0 = tr.global_counter.cur_step
            ob, done, R = tr.explore(ob, done)
            ag.backward(R)
            steps += tr.global_counter.cur_step - c0
            done_updates += 1
            if done or done_updates >= updates + warm_updates:
                break
    dt = time.perf_counter() - t0
    n = env.n_agent
    return steps * n / dt, cores, '%d update cycles of %d env steps, B=1, %s' % (updates, ag.n_step, cfg_name), dt


class ClockSampler:
    """Streams `nvidia-smi -lms 100` while the timed region runs (B200_PROFILING.md clocks line)."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.5)
        except Exception:
            self.proc = None

    def summary(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        time.sleep(0.2)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ''
        samples = [[x.strip() for x in l.split(',')] for l in out.strip().splitlines() if l.count(',') >= 5]
        if not samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith('active') for s in samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(samples[0][1]), 'reasons': reasons, 'samples': len(sm)}


def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        if rank != 0:
            return
        import torch
        val, cores, sample, dt = cpu_reference_best(args.config, updates=max(2, 2 * args.steps))
        cp = load_cfg(args.config)
        T, N = int(cp['MODEL_CONFIG']['batch_size']), int(cp['ENV_CONFIG']['n_vehicle'])
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': args.config + ', restated reference (TF unavailable), 1 env x %d agents, CPU' % N},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    # stdout carries exactly one JSON line: NCCL (version banner, NCCL_DEBUG output) and any library chatter write to
    # file descriptor 1 as well, so fd 1 is pointed at stderr for the run and the line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from deeprl_network_b200 import _lib as L
    from deeprl_network_b200.agents.models import MA2C_NC, MA2C_IC3, MA2C_DIAL, IA2C, IA2C_FP, IA2C_CU
    from deeprl_network_b200.envs.cacc_env import CACCEnv
    from deeprl_network_b200.utils import VecTrainer
    B = args.n_env
    cp = load_cfg(args.config, n_env=B, seed=12 + 1000 * rank)
    env = CACCEnv(cp['ENV_CONFIG'])
    cls = {'ma2c_nc': MA2C_NC, 'ma2c_ic3': MA2C_IC3, 'ma2c_dial': MA2C_DIAL, 'ia2c': IA2C, 'ia2c_fp': IA2C_FP,
           'ma2c_cu': IA2C_CU}[env.agent]
    np.random.seed(12)                                   # identical initial weights on every rank
    kw = dict(obs_mode='gather') if env.agent == 'ia2c' else {}
    model = cls(env.n_s_ls, env.n_a_ls, env.neighbor_mask, env.distance_mask, env.coop_gamma, 10 ** 9,
                cp['MODEL_CONFIG'], seed=12 + rank, n_env=B, **kw)
    e = model.engine
    T, N = e.T, e.N
    steps_per_update = T * B * N

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(K):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm (Philox sampling inside the kernels) --------------------------------------
    vt = VecTrainer(env, model, graph=True, sample='philox')
    vt.start()
    l0 = e.launches
    vt.update()                                          # eager warm-up + capture
    launches_per_update = (e.launches - l0) // 2         # eager pass + capture pass issue the same calls
    for _ in range(max(0, args.warmup - 1)):
        vt.update()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(vt.update, args.steps)
    clocks = sampler.summary()
    value = args.steps * steps_per_update * world / (ms * 1e-3)

    # ---- e2e arm: host-supplied uniforms in, rewards + losses out, copies inside the timed region ------
    e2e = None
    if not args.no_e2e:
        uni_host = torch.rand(T + 1, N, B, dtype=torch.float64).pin_memory()
        uni_dev = torch.zeros(T + 1, N, B, dtype=torch.float64, device='cuda')
        rew_host = torch.zeros(T, B, dtype=torch.float64).pin_memory()
        loss_host = torch.zeros(N, 4, dtype=torch.float32).pin_memory()
        vt2 = VecTrainer(env, model, graph=True, sample='uniform')
        vt2._seed = vt._seed

        def e2e_step():
            uni_dev.copy_(uni_host, non_blocking=True)                     # H2D: this step's action uniforms
            vt2.update(uniforms=uni_dev)
            rew_host.copy_(e.grew_buf, non_blocking=True)                  # D2H: per-step global rewards
            loss_host.copy_(e.loss_part.sum(dim=(0, 2)), non_blocking=True)  # D2H: loss terms
            torch.cuda.current_stream().synchronize()                      # the caller reads the results
        for _ in range(max(2, args.warmup)):
            e2e_step()
        ms2 = timed(e2e_step, args.steps)
        e2e = {'value': args.steps * steps_per_update * world / (ms2 * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': uni_host.numel() * 8,
               'd2h_bytes_per_step': rew_host.numel() * 8 + loss_host.numel() * 4,
               'ms_per_step': ms2 / args.steps,
               'api': 'VecTrainer.update(uniforms=<host RNG stream>) -> rewards, loss terms'}

    # ---- roofline of the dominant kernel: fused gather + encoders + LSTM cell + heads (rollout p-call) ----
    # Timed IN SITU: one more rollout, launched eagerly, with CUDA events on the launching stream around each of its
    # T p-calls (every call works on its own state / activation slots, so the caches are in their real state); the
    # v-calls stay on the main stream for this pass so nothing shares the SMs with the kernel being timed.
    pi = torch.zeros(N, B, e.n_a, device='cuda'); act = torch.zeros(N, B, dtype=torch.int32, device='cuda')
    ms_warm = None
    if getattr(e, 'fuse_save', False) and e.use_tc:
        ov = e.overlap_v
        e.overlap_v, e.kernel_events = False, []
        e.rollout(env, sample=vt.sample)
        torch.cuda.synchronize()
        e.overlap_v = ov
        durs = [a.elapsed_time(b) for a, b in e.kernel_events[:T]]          # the T saving p-calls (not the bootstrap)
        e.kernel_events = None
        e.saved_rollout = False
        ms_k = float(np.mean(durs))
        saves = True
    else:
        saves = False

    def pcall():
        e.step_p(e.obs_buf[0], e.fp_buf[0], e.done_buf[0], pi, act, L.SAMPLE_PHILOX, rng_offset=0)
    for _ in range(5):
        pcall()
    ms_warm = timed(pcall, 50) / 50                       # same kernel without the saves, back to back on one input
    if not saves:
        ms_k = ms_warm
    bytes_per_agent_step = 1640 if env.agent == 'ma2c_nc' else {'ma2c_ic3': 1612, 'ma2c_dial': 1628}.get(env.agent, 1100)
    # K1+K2+state+outputs per agent-env-step; the rollout p-call also writes the activations BPTT needs
    # (8(d) "Backward": s, gates, c, h, encoder pre-activations ~ 768 floats, booked there under the training forward)
    save_bytes = 3072 if saves else 0
    alg_bytes = (bytes_per_agent_step + save_bytes) * N * B
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(
            ('tc_cell_fwd_ps_bytes_per_launch' if saves else 'tc_cell_fwd_p_bytes_per_launch') if e.use_tc
            else 'cell_fwd_p_bytes_per_launch')
    except Exception:
        pass
    achieved = alg_bytes / (ms_k * 1e-3) / 1e9
    roofline = {'kernel': (('tc_cell_fwd_kernel<PS> (tcgen05 3xTF32: fused gather+encoders+LSTM cell+heads+sampling+activation save)'
                            if saves else 'tc_cell_fwd_kernel<P> (tcgen05 3xTF32: fused gather+encoders+LSTM cell+heads+sampling)')
                           if e.use_tc else 'cell_fwd_kernel<P> (FP32 FFMA)'), 'bound': 'hbm',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                'us_per_launch': ms_k * 1e3, 'us_per_launch_warm_no_saves': ms_warm * 1e3,
                'timing': 'CUDA events around each p-call of one eagerly launched rollout (in situ)' if saves else 'back-to-back launches',
                'algorithmic_bytes_per_launch': alg_bytes, 'algorithmic_bytes_per_agent_step': bytes_per_agent_step + save_bytes,
                'peak_source': 'MEASURED_PEAKS.json (burst)' if peaks else 'fallback 6.65 TB/s',
                'tensor_tflops_3xtf32': 3 * 2 * 74359 * N * B / (ms_k * 1e-3) / 1e12,
                'note': 'algorithmic bytes per SURVEY 8(d): %d B forward + %d B saved activations per agent-env-step; the kernel '
                        'issues 3 TF32 MMAs per fp32 product (148.7 kFLOP fp32-equivalent per agent-step), see DESIGN.md'
                        % (bytes_per_agent_step, save_bytes)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, dt = cpu_reference_best(args.config, updates=16)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
               'label': 'restated reference (TF unavailable)', 'seconds': dt}

    if rank == 0:
        emit = lambda line: os.write(json_fd, (line + '\n').encode())
        emit(json.dumps({
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': '%s, %d parallel envs per GPU x %d agents, n_step %d (BASELINE configs[1])' %
                                   (args.config, B, N, T), 'global_envs': B * world, 'parallelism': 'dp%d' % world,
                       'l2_policy': 'per-step working set (activations %.1f GB) exceeds L2' %
                                    (T * N * B * 800 * 4 / 1e9)},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches_per_update * args.steps,
            'roofline': roofline, 'cpu_baseline': cpu}))
    if world > 1:
        # captured CUDA graphs hold NCCL kernels: tearing the communicator down under them can block, so leave
        # together after a final barrier instead of destroy_process_group()
        sys.stdout.flush()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == '__main__':
    main()
