"""TEST / BENCH INFRASTRUCTURE -- stages the UNMODIFIED reference implementation of the hot path as `oracle/_ref/`.

The reference (nicholasmr/tamm) is pure Python: "building" it is packing the three modules of the path,
lib/layerstack.py, lib/layer.py and lib/layer_PPJ.py, byte for byte from the read-only reference tree into ONE archive,
oracle/_ref/tamm_reference_lib.zip (git-ignored, so no reference source enters this repository's history or its source
tree; not gpurun-ignored, so the archive travels to the GPU box like the built .so files).  Python imports the modules
straight from the archive (zipimport).  A MANIFEST with the sha256 of every member is written next to it.

Users: bench.py's reference arm (`--impl reference`, cpu_baseline.kind = "reference") times these files on the host
cores; tests/test_oracle_vs_reference.py pins the oracle against them.  Nothing in tamm_b200/ or lib/ imports them.

    python oracle/stage_ref.py            # stage if the reference tree is present, report otherwise
"""
import hashlib
import os
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(os.sep, 'root', 'reference', 'lib')
DEST = os.path.join(HERE, '_ref')
ARCHIVE = os.path.join(DEST, 'tamm_reference_lib.zip')
FILES = ('layerstack.py', 'layer.py', 'layer_PPJ.py')


def staged():
    """True when oracle/_ref holds the archive with all three modules."""
    if not os.path.exists(ARCHIVE):
        return False
    try:
        with zipfile.ZipFile(ARCHIVE) as z:
            return all(f in z.namelist() for f in FILES)
    except Exception:
        return False


def stage(verbose=False):
    """Pack the reference modules if the reference tree exists here (build container); never modifies them."""
    if not os.path.isdir(REF_LIB):
        if verbose:
            print('reference tree not present; oracle/_ref %s' % ('already staged' if staged() else 'NOT staged'))
        return staged()
    os.makedirs(DEST, exist_ok=True)
    data = {f: open(os.path.join(REF_LIB, f), 'rb').read() for f in FILES}
    lines = ['%s  %s' % (hashlib.sha256(data[f]).hexdigest(), f) for f in FILES]
    current = None
    if staged():
        with zipfile.ZipFile(ARCHIVE) as z:
            current = {f: z.read(f) for f in FILES}
    if current != data:
        with zipfile.ZipFile(ARCHIVE, 'w', compression=zipfile.ZIP_DEFLATED) as z:
            for f in FILES:
                z.writestr(zipfile.ZipInfo(f, date_time=(2020, 1, 1, 0, 0, 0)), data[f])
    for f in FILES:                       # plain copies of an earlier staging scheme
        old = os.path.join(DEST, f)
        if os.path.exists(old):
            os.remove(old)
    with open(os.path.join(DEST, 'MANIFEST'), 'w') as fh:
        fh.write('members of tamm_reference_lib.zip: unmodified nicholasmr/tamm lib/ modules (GPL-3.0), packed by oracle/stage_ref.py\n'
                 + '\n'.join(lines) + '\n')
    if verbose:
        print('packed %d reference modules into %s' % (len(FILES), ARCHIVE))
    return True


def load():
    """Import the staged reference: returns the `layerstack` module of the archive.  The reference modules import each other
    by bare name (layerstack.py:9-10), so the archive is put first on sys.path for the import and the three names are
    registered under a private prefix afterwards (lib/ of this repository uses the same names)."""
    import importlib
    import sys
    if not staged():
        raise FileNotFoundError('oracle/_ref is not staged (run oracle/stage_ref.py in the build container)')
    saved = {k: sys.modules.pop(k) for k in ('layerstack', 'layer', 'layer_PPJ') if k in sys.modules}
    sys.path.insert(0, ARCHIVE)
    try:
        mod = importlib.import_module('layerstack')
        assert os.path.abspath(mod.__file__).startswith(ARCHIVE), mod.__file__
        for k in ('layerstack', 'layer', 'layer_PPJ'):
            sys.modules['_tamm_ref_' + k] = sys.modules.pop(k)
    finally:
        sys.path.remove(ARCHIVE)
        sys.modules.update(saved)
    return mod


if __name__ == '__main__':
    stage(verbose=True)
