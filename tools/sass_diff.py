#!/usr/bin/env python
"""Per-kernel comparison of the device code of two builds of libtamm_b200.so (cuobjdump, no GPU needed): which kernels are
byte-identical (instruction text, addresses stripped), which changed, which are new.  Used to show that a source change (new
entry points, spelling changes for the host emulation) left the measured kernels alone:

    python tools/sass_diff.py old.so [new.so]

Round 2, build of the last GPU measurement against the final one: 32 kernels identical, 4 changed (the opt-in
returns_tma_kernel instantiations: a bug fix), 6 new (stack_returns_kernel x 2, power_inplace_kernel x 4)."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(lib):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    d, name, body = {}, None, []
    for l in out.splitlines():
        m = re.search(r'Function : (\S+)', l)
        if m:
            if name:
                d[name] = hashlib.md5('\n'.join(body).encode()).hexdigest()
            name, body = m.group(1), []
        elif name and '/*' in l:
            body.append(re.sub(r'^\s*/\*[0-9a-f]+\*/', '', l))
    if name:
        d[name] = hashlib.md5('\n'.join(body).encode()).hexdigest()
    return d


def short(k):
    return re.sub(r'_ZN\d+_GLOBAL__N__\w+?_cu_\w{8}\d*', '', k)[:110]


def main():
    old = sys.argv[1]
    new = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'tamm_b200', 'libtamm_b200.so')
    a, b = kernels(old), kernels(new)
    same = [k for k in a if k in b and a[k] == b[k]]
    diff = [k for k in a if k in b and a[k] != b[k]]
    print('%d kernels identical, %d changed, %d new, %d removed' % (len(same), len(diff), len([k for k in b if k not in a]),
                                                                   len([k for k in a if k not in b])))
    for tag, ks in (('CHANGED', diff), ('NEW', [k for k in b if k not in a]), ('REMOVED', [k for k in a if k not in b])):
        for k in ks:
            print(tag, short(k))
    return 1 if diff else 0


if __name__ == '__main__':
    sys.exit(main())
