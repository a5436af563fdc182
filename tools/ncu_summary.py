#!/usr/bin/env python
"""Markdown summary of one `ncu --set full` report (run in the build container, ncu is installed):

    python tools/ncu_summary.py gpurun_out/X.ncu-rep --evals 115840000 [--title "..."] [--traffic-json out.json --md5 <lib md5>]

Prints the counters this repository's profile summaries quote (duration, FP64 pipe, issue, stalls, DRAM bytes, registers,
local-memory instructions) and the dynamic SASS instruction mix per evaluation from the source page."""
import argparse
import collections
import csv
import json
import re
import subprocess


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def to_num(v):
    try:
        return float(v.replace(',', ''))
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--evals', type=float, required=True, help='evaluations of one launch')
    ap.add_argument('--title', default='')
    ap.add_argument('--traffic-json', default='')
    ap.add_argument('--md5', default='')
    a = ap.parse_args()
    raw = ncu_csv(a.report, 'raw')
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (to_num(v), u) for h, u, v in zip(hdr, units, vals)}
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'

    def g(k, scale=1.0):
        v = m.get(k, (None, ''))[0]
        return None if v is None else v * scale

    def to_bytes(k):
        v, u = m.get(k, (None, ''))
        if v is None:
            return None
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(u, 1)

    def to_ms(k):
        v, u = m.get(k, (None, ''))
        if v is None:
            return None
        return v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(u, 1)

    dur = to_ms('gpu__time_duration.sum')
    rd, wr = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum')
    print('## %s' % (a.title or name))
    print()
    print('`%s`, grid %d x block %d, %d registers/thread, dynamic shared memory %s %s/block' % (
        name[:140], g('launch__grid_size') or 0, g('launch__block_size') or 0, g('launch__registers_per_thread') or 0,
        vals[hdr.index('launch__shared_mem_per_block_dynamic')] if 'launch__shared_mem_per_block_dynamic' in hdr else '?',
        units[hdr.index('launch__shared_mem_per_block_dynamic')] if 'launch__shared_mem_per_block_dynamic' in hdr else ''))
    print()
    print('| metric | value |')
    print('|---|---|')
    print('| duration (ncu, cold cache, serialised) | %.3f ms -> %.3g evals/s |' % (dur, a.evals / (dur * 1e-3)))
    for k, label in (('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', '`sm__pipe_fp64_cycles_active` (% of peak)'),
                     ('smsp__issue_active.avg.pct_of_peak_sustained_active', '`smsp__issue_active` (%)'),
                     ('sm__warps_active.avg.pct_of_peak_sustained_active', '`sm__warps_active` (% of peak)'),
                     ('sm__throughput.avg.pct_of_peak_sustained_elapsed', '`sm__throughput` (% of peak)'),
                     ('gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', 'memory throughput (% of peak)'),
                     ('dram__throughput.avg.pct_of_peak_sustained_elapsed', '`dram__throughput` (% of peak)')):
        if g(k) is not None:
            print('| %s | %.1f |' % (label, g(k)))
    if rd is not None and wr is not None:
        print('| `dram__bytes_read.sum` + `dram__bytes_write.sum` | %.3f + %.3f GB = %.3f GB per launch = %.1f B/eval -> %.0f GB/s |' % (
            rd / 1e9, wr / 1e9, (rd + wr) / 1e9, (rd + wr) / a.evals, (rd + wr) / (dur * 1e-3) / 1e9))
    ll, ls = g('sass__inst_executed_local_loads'), g('sass__inst_executed_local_stores')
    if ll is not None:
        print('| local-memory instructions (warp level) | %.3g loads, %s stores per launch = %.2f + %.2f per eval (thread level) |' % (
            ll, '%.3g' % ls if ls is not None else '?', ll * 32 / a.evals, (ls or 0) * 32 / a.evals))
    stalls = sorted(((v[0], k.split('issue_stalled_')[1].split('_per_issue')[0]) for k, v in m.items()
                     if 'issue_stalled_' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k and v[0]), reverse=True)
    print('| stalls per issue | %s |' % ', '.join('%s %.2f' % (n, v) for v, n in stalls[:7] if n != 'selected'))
    # instruction mix from the source page
    src = ncu_csv(a.report, 'source', ['--print-source', 'sass'])
    try:
        h = src[1]
        iS, iE = h.index('Source'), h.index('Instructions Executed')
        ops = collections.Counter()
        for r in src[2:]:
            if len(r) <= iE or not r[iE].isdigit():
                continue
            mm = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS])
            if mm:
                ops[mm.group(2).split('.')[0]] += int(r[iE])
        tot = sum(ops.values())
        per = lambda n: n * 32.0 / a.evals
        fp64 = sum(ops[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
        print('| dynamic instructions per eval (SASS, thread level) | total %.0f: FP64 %.0f (DFMA %.0f, DMUL %.0f, DADD %.0f, DSETP %.0f); other %.0f (%s) |' % (
            per(tot), per(fp64), per(ops['DFMA']), per(ops['DMUL']), per(ops['DADD']), per(ops['DSETP']), per(tot - fp64),
            ', '.join('%s %.0f' % (k, per(v)) for k, v in ops.most_common(14) if k not in ('DFMA', 'DMUL', 'DADD', 'DSETP'))))
        if fp64:
            lb = (2.18 * per(fp64) + per(tot - fp64)) / 128.0
            sm_clk = g('sm__cycles_elapsed.max') or 0
            n_sm = 148
            meas = sm_clk * n_sm / a.evals if sm_clk else None
            print('| issue model `(2.18 N_fp64 + N_other)/128` | %.1f cycles/eval/SM lower bound%s |' % (
                lb, (', %.1f measured (%.0f %%)' % (meas, 100 * lb / meas)) if meas else ''))
    except Exception as e:
        print('| instruction mix | unavailable: %r |' % (e,))
    print()
    if a.traffic_json and rd is not None:
        json.dump(dict(kernel=name[:120], source='ncu --set full, %s' % a.report, evals_per_launch=a.evals, dram_bytes_read=rd,
                       dram_bytes_write=wr, dram_bytes_per_eval=(rd + wr) / a.evals, stack_sass_md5=a.md5), open(a.traffic_json, 'w'), indent=1)


if __name__ == '__main__':
    main()
