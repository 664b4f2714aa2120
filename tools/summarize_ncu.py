"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
    python tools/summarize_ncu.py <tag>
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1]
os.makedirs('profiles', exist_ok=True)

# ---- launch lists (headline + one per BASELINE configuration: launches_<tag>_<cfg>.csv) --------------------
import glob
for path in sorted(glob.glob('gpurun_out/launches_%s*.csv' % tag)):
    suffix = os.path.basename(path)[len('launches_%s' % tag):-4]          # '' or '_cfg3' ...
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    n = 0
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}[r['Metric Unit']]
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; n += 1
    if not n:
        continue
    tot = sum(a[1] for a in agg.values())
    out = 'profiles/%s%s_launches.txt' % (tag, suffix)
    with open(out, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n')
        f.write('# command: python bench.py [--config ... --n-env ...] --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra ; %d launches, %.0f us total\n' % (n, tot))
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-75s n=%5d total=%10.0f us avg=%9.1f us share=%5.1f%%\n' % (k[:75], c, t, t / c, 100 * t / tot))
    if not suffix:
        print(open(out).read())

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_mio_throttle']
traffic = {}
for which in ('fwd', 'bwd', 'tc', 'wg'):
    rep = 'gpurun_out/prof_%s_%s.ncu-rep' % (which, tag)
    if not os.path.exists(rep):
        continue
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open('profiles/%s_%s_full.txt' % (tag, which), 'w') as f:
        f.write('# ncu --set full --clock-control none --import-source on ; selected metrics of %s\n' % rep)
        for r in rows[2:]:
            name = re.sub(r'\(.*', '', r[idx['Kernel Name']])
            f.write('---- %s\n' % name)
            for w in WANT:
                if w in idx:
                    f.write('%-72s %s %s\n' % (w, r[idx[w]], units[idx[w]]))
            def val(m):
                v = float(r[idx[m]].replace(',', ''))
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[idx[m]]]
            tr = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
            f.write('%-72s %.0f byte\n' % ('DRAM traffic per launch (read+write)', tr))
            traffic[name.replace('void ', '').replace('<unnamed>::', '')] = tr
    print(open('profiles/%s_%s_full.txt' % (tag, which)).read()[:3000])
if traffic:
    old = {}
    try:
        old = json.load(open('profiles/traffic.json'))
    except Exception:
        pass
    p = [v for k, v in traffic.items() if k.startswith('cell_fwd_kernel<1, 0')]
    ptc = [v for k, v in traffic.items() if k.startswith('tc_cell_fwd_kernel<1, 0')]
    old.setdefault('per_kernel_bytes_per_launch', {}).update(traffic)
    old['tag'] = tag
    if p:
        old['cell_fwd_p_bytes_per_launch'] = p[0]
    if ptc:
        old['tc_cell_fwd_p_bytes_per_launch'] = ptc[0]
    pps = [v for k, v in traffic.items() if k.startswith('tc_cell_fwd_kernel<1, 3')]      # rollout p-call that also saves
    if pps:
        old['tc_cell_fwd_ps_bytes_per_launch'] = pps[0]
    json.dump(old, open('profiles/traffic.json', 'w'), indent=1)
