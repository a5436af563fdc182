"""TEST / BENCH INFRASTRUCTURE -- stages the UNMODIFIED reference implementation of the hot path as `oracle/_ref/`.

The reference (nicholasmr/tamm) is pure Python: "building" it is copying the three modules of the path,
lib/layerstack.py, lib/layer.py and lib/layer_PPJ.py, byte for byte from the read-only reference tree into
oracle/_ref/ (git-ignored, so no reference source enters this repository's history; not gpurun-ignored, so the copy
travels to the GPU box like the built .so files).  A MANIFEST with the sha256 of every file is written next to them.

Users: bench.py's reference arm (`--impl reference`, cpu_baseline.kind = "reference") times these files on the host
cores; tests/test_oracle_vs_reference.py pins the oracle against them.  Nothing in tamm_b200/ or lib/ imports them.

    python oracle/stage_ref.py            # stage if the reference tree is present, report otherwise
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(os.sep, 'root', 'reference', 'lib')
DEST = os.path.join(HERE, '_ref')
FILES = ('layerstack.py', 'layer.py', 'layer_PPJ.py')


def staged():
    """True when oracle/_ref holds all three modules."""
    return all(os.path.exists(os.path.join(DEST, f)) for f in FILES)


def stage(verbose=False):
    """Copy the reference modules if the reference tree exists here (build container); never modifies them."""
    if not os.path.isdir(REF_LIB):
        if verbose:
            print('reference tree not present; oracle/_ref %s' % ('already staged' if staged() else 'NOT staged'))
        return staged()
    os.makedirs(DEST, exist_ok=True)
    lines = []
    for f in FILES:
        src, dst = os.path.join(REF_LIB, f), os.path.join(DEST, f)
        data = open(src, 'rb').read()
        if not os.path.exists(dst) or open(dst, 'rb').read() != data:
            shutil.copyfile(src, dst)
        lines.append('%s  %s' % (hashlib.sha256(data).hexdigest(), f))
    with open(os.path.join(DEST, 'MANIFEST'), 'w') as fh:
        fh.write('unmodified copies of nicholasmr/tamm lib/ (GPL-3.0), staged by oracle/stage_ref.py\n' + '\n'.join(lines) + '\n')
    if verbose:
        print('staged %d reference modules into %s' % (len(FILES), DEST))
    return True


def load():
    """Import the staged reference as a module namespace: returns the `layerstack` module of oracle/_ref.  The reference
    modules import each other by bare name (layerstack.py:9-10), so oracle/_ref is put first on sys.path for the import
    and the three names are registered under a private prefix afterwards (lib/ of this repository uses the same names)."""
    import importlib
    import sys
    if not staged():
        raise FileNotFoundError('oracle/_ref is not staged (run oracle/stage_ref.py in the build container)')
    saved = {k: sys.modules.pop(k) for k in ('layerstack', 'layer', 'layer_PPJ') if k in sys.modules}
    sys.path.insert(0, DEST)
    try:
        mod = importlib.import_module('layerstack')
        assert os.path.dirname(os.path.abspath(mod.__file__)) == DEST, mod.__file__
        for k in ('layerstack', 'layer', 'layer_PPJ'):
            sys.modules['_tamm_ref_' + k] = sys.modules.pop(k)
    finally:
        sys.path.remove(DEST)
        sys.modules.update(saved)
    return mod


if __name__ == '__main__':
    stage(verbose=True)
