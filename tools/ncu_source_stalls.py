"""Warp-stall samples by SASS region from an ncu report with source info:
    python tools/ncu_source_stalls.py gpurun_out/prof_tc_<tag>.ncu-rep tc_cell_fwd_kernel [launch_index]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern],
                         capture_output=True, text=True).stdout.splitlines()
    hdr = [i for i, l in enumerate(out) if l.startswith('"Address"')]
    rows = list(csv.reader(out[hdr[which]:(hdr[which + 1] - 1 if len(hdr) > which + 1 else len(out))]))
    h, rows = rows[0], rows[1:]
    ia, isamp = h.index('Source'), h.index('# Samples')
    samp = [int(r[isamp] or 0) for r in rows]
    src = [r[ia].strip() for r in rows]
    tot = sum(samp)
    print('# %s, launch %d of %s: %d stall samples over %d SASS instructions' % (kern, which, rep, tot, len(rows)))
    B = 400
    for b0 in range(0, len(rows), B):
        ops = {}
        for k in range(b0, min(len(rows), b0 + B)):
            w = src[k].split()
            op = (w[1] if w[0].startswith('@') else w[0]).split('.')[0]
            ops[op] = ops.get(op, 0) + samp[k]
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
        print('instr %5d-%5d: %5.1f%%  %s' % (b0, b0 + B, 100 * sum(samp[b0:b0 + B]) / tot, ' '.join('%s:%d' % kv for kv in top)))
    kinds = {}
    for k, s in enumerate(src):
        for w in s.split():
            key = w.split('.')[0]
            if key in ('LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL', 'UTCHMMA', 'LDTM', 'STTM', 'MUFU', 'SYNCS', 'UBLKCP', 'BRA', 'EXIT'):
                kinds.setdefault(key, [0, 0]); kinds[key][0] += 1; kinds[key][1] += samp[k]
                break
    print('by opcode (instructions, samples):', ' '.join('%s=(%d,%d)' % (k, v[0], v[1]) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1][1])))


if __name__ == '__main__':
    main()
