#!/usr/bin/env python
"""Static cost of the returns tail inside stack_returns_kernel (tamm_stack_returns): the instructions of the mu = 1 instantiation
whose inline chain (`nvdisasm -gi`, the library is built with -lineinfo) passes through tamm::returns_point or through the
MODE_RET block of stack_body -- straight-line code executed once per (lane, interface), i.e. per eval.  The cold routine
(returns_point_store_cold) is a separate function and not counted.  No GPU needed.

    python tools/direct_static.py [library]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL = 'stack_returns_kernelILi384ELi1ELb1'
FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU')


def source_range(path, start_pat, end_pat):
    lines = open(path).read().split('\n')
    a = next(i for i, l in enumerate(lines, 1) if start_pat in l)
    b = next(i for i, l in enumerate(lines, 1) if i > a and end_pat in l)
    return a, b


def function_range(path, name):
    lines = open(path).read().split('\n')
    a = next(i for i, l in enumerate(lines, 1) if re.match(r'^TAMM_HD .*\b%s\s*\(' % name, l))
    b = next(i for i, l in enumerate(lines, 1) if i > a and l.startswith('}'))
    return a, b


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'tamm_b200', 'libtamm_b200.so')
    ker = os.path.join(ROOT, 'tamm_b200', 'csrc', 'tamm_kernels.cu')
    core = os.path.join(ROOT, 'tamm_b200', 'csrc', 'tamm_core.cuh')
    # the block inside the layer loop: from its comment to the first line of the plane-writing branch
    blk = source_range(ker, '// one chunk: o = URD_n is final.', 'cd* op = out + (size_t)n * 4 * QF;')
    rp = function_range(core, 'returns_point')
    tmp = tempfile.mkdtemp()
    subprocess.run('cd %s && cuobjdump -xelf all %s > /dev/null 2>&1' % (tmp, lib), shell=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    gi = subprocess.run(['nvdisasm', '-gi', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
    insec, cur, last = False, [], []
    ops = collections.Counter()
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
            if cur:
                last = cur
            cur = []
            hit = any((f == 'tamm_kernels.cu' and blk[0] <= ln <= blk[1]) or (f == 'tamm_core.cuh' and rp[0] <= ln <= rp[1])
                      for f, ln in last)
            if hit:
                ops[m.group(3).split('.')[0]] += 1
    fp = sum(v for k, v in ops.items() if k in FP64)
    tot = sum(ops.values())
    print('returns tail inside %s: %d instructions per eval, %d on the FP64 pipe (%s), %d others (%s)' % (
        KERNEL, tot, fp, ', '.join('%s %d' % (k, ops[k]) for k in FP64 if ops[k]), tot - fp,
        ', '.join('%s %d' % (k, v) for k, v in ops.most_common() if k not in FP64)))
    print('issue slots per eval at 2.18 per FP64 instruction: %.0f (stack_kernel: 4919)' % (2.18 * fp + (tot - fp)))


if __name__ == '__main__':
    main()
