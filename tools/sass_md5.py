#!/usr/bin/env python
"""md5 of the stack kernel's SASS in a built library (default: the in-tree one) -- the key under which
profiles/r02_stack_traffic.json records the ncu capture of that kernel (bench.py: stack_traffic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
argv, sys.argv = sys.argv, sys.argv[:1]
import bench  # noqa: E402

lib = argv[1] if len(argv) > 1 else os.path.join(ROOT, 'tamm_b200', 'libtamm_b200.so')
print(bench.stack_sass_md5(lib), lib)
