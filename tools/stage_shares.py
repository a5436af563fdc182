#!/usr/bin/env python
"""Issue-slot shares of the stack kernel by algorithmic stage: joins the per-instruction execution counts of an
`ncu --set full --import-source on` report (SASS page) with the inline chains `nvdisasm -gi` prints for the same device code,
and attributes every instruction to the outermost stage-defining function / source region of its chain.

    python tools/stage_shares.py gpurun_out/r02y_stack.ncu-rep --evals 115840000

FP64 instructions are weighted 2.18 issue slots, everything else 1 (tools/mix_probe.cu).  The library must be the build the
report was captured on (tools/sass_md5.py)."""
import argparse
import collections
import csv
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL = 'stack_kernelILi0ELi384ELi1ELb1'
FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')


def function_ranges(path):
    """(start line, name) of every function definition in tamm_core.cuh (top-level `TAMM_HD` / template / noinline forms)."""
    out = []
    lines = open(path).read().split('\n')
    for i, l in enumerate(lines, 1):
        m = re.match(r'^(?:TAMM_HD|inline|static)?\s*(?:[\w:<>]+\s+)+?(\w+)\s*\(', l)
        if m and not l.startswith(' ') and not l.strip().endswith(';') and m.group(1) not in ('if', 'for', 'while', 'return', 'defined'):
            out.append((i, m.group(1)))
    return out, lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--evals', type=float, required=True)
    ap.add_argument('--lib', default=os.path.join(ROOT, 'tamm_b200', 'libtamm_b200.so'))
    a = ap.parse_args()
    core = os.path.join(ROOT, 'tamm_b200', 'csrc', 'tamm_core.cuh')
    franges, core_lines = function_ranges(core)

    def func_of(line):
        name = None
        for s, n in franges:
            if s <= line:
                name = n
            else:
                break
        return name

    def text_of(line):
        return core_lines[line - 1] if 0 < line <= len(core_lines) else ''

    # --- inline chains per instruction offset
    tmp = tempfile.mkdtemp()
    subprocess.run('cd %s && cuobjdump -xelf all %s > /dev/null 2>&1' % (tmp, a.lib), shell=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    gi = subprocess.run(['nvdisasm', '-gi', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
    chains, cur, insec = {}, [], False
    for l in gi:
        if l.startswith('\t.section'):
            insec = ('.text.' in l and KERNEL in l)
            cur = []
            continue
        if not insec:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            cur.append((os.path.basename(m.group(1)), int(m.group(2))))
            if m.group(3):
                cur.append((os.path.basename(m.group(3)), int(m.group(4))))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if m:
            off = int(m.group(1), 16)
            if cur:
                last = cur
            chains[off] = list(last) if 'last' in dir() else []
            cur = []

    def stage(chain, op):
        files_lines = [(f, ln) for f, ln in chain if f == 'tamm_core.cuh']
        funcs = [func_of(ln) for f, ln in files_lines]
        texts = ' '.join(text_of(ln) for f, ln in files_lines)
        fs = set(funcs)
        if 'refine_root' in fs:
            return 'eigen-solver: entry-wise Newton / MSLP refinement'
        if fs & {'mslp_eigs_f', 'mslp_core', 'mslp_step_cold'}:
            return 'eigen-solver: pair step'
        if fs & {'eig4_fast', 'eig4_zeta0', 'eig4_fallback', 'quartic_coef'}:
            return 'eigen-solver: seeds, Vieta start, null vectors, checks'
        if fs & {'layer_phases', 'sincos_f', 'exp_small', 'exp_small_re', 'exp_f', 'phase_f', 'layer_phases_full'}:
            return 'phases'
        if 'interface_fb' in fs or 'n2n_blocks' in fs or 'nonfinite_key' in fs:
            return 'interface products (7 two-by-two products, M = (Yt\' - Yr)^-1)'
        if 'layer_solve_field' in fs and 'layer_modes' not in fs:
            return 'field blocks Yt, Yr, P, Q'
        if 'layer_modes' in fs:
            lm_lines = [ln for (f, ln), fn in zip(files_lines, funcs) if fn == 'layer_modes']
            t = ' '.join(text_of(ln) for ln in lm_lines)
            ln = max(lm_lines) if lm_lines else 0
            # regions of layer_modes by source position
            src = core_lines
            start = next(s for s, n in franges if n == 'layer_modes')
            pos = {k: next((i for i in range(start, len(src)) if key in src[i]), 10 ** 9) + 1 for k, key in
                   (('perm', 'const Sym3 e = permittivity'), ('clean', 'clean-up (layer_PPJ.py:199-203)'), ('sort', 'const int tmask'),
                    ('gamma', '---- gamma (layer_PPJ.py:309-389)'))}
            if ln >= pos['gamma']:
                return 'gamma (unnormalised field vectors)'
            if ln >= pos['sort']:
                return 'p/s sort: partition and comparisons'
            if ln >= pos['clean']:
                return 'p/s sort: Poynting measures of the four modes'
            return 'permittivity, wave coefficients, Berreman a-entries'
        if fs & {'permittivity', 'wave_coef'}:
            return 'permittivity, wave coefficients, Berreman a-entries'
        return 'layer loop, TMA staging, carry slots, stores'

    # --- executed counts per address
    out = subprocess.run(['ncu', '-i', a.report, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    iA, iS, iE = h.index('Address'), h.index('Source'), h.index('Instructions Executed')
    base = None
    slots, fp, other = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        addr = int(r[iA], 16)
        if base is None:
            base = addr
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS])
        if not m:
            continue
        op = m.group(2).split('.')[0]
        n = int(r[iE]) * 32.0 / a.evals
        st = stage(chains.get(addr - base, []), op)
        if op in FP64:
            fp[st] += n
            slots[st] += 2.18 * n
        else:
            other[st] += n
            slots[st] += n
    tot = sum(slots.values())
    print('| stage | FP64 instr / eval | other instr / eval | issue slots / eval | share |')
    print('|---|---|---|---|---|')
    for st, v in sorted(slots.items(), key=lambda x: -x[1]):
        print('| %s | %.0f | %.0f | %.0f | %.1f %% |' % (st, fp[st], other[st], v, 100 * v / tot))
    print('| **total** | %.0f | %.0f | %.0f | 100 %% |' % (sum(fp.values()), sum(other.values()), tot))


if __name__ == '__main__':
    main()
