"""Live pin: the oracle against the reference itself, when /root/reference is mounted (build container only)."""
import os
import sys

import numpy as np
import pytest

import helpers as h
from tamm_b200 import profiles
import tamm_oracle as orc

REF = '/root/reference/lib'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference not mounted (GPU box)')


def _ref():
    saved = {k: sys.modules.pop(k) for k in ('layerstack', 'layer', 'layer_PPJ') if k in sys.modules}
    sys.path.insert(0, REF)
    try:
        import layerstack as ref
        return ref
    finally:
        sys.path.remove(REF)
        for k in ('layerstack', 'layer', 'layer_PPJ'):
            sys.modules.pop(k, None)
        sys.modules.update(saved)


@pytest.mark.parametrize('case', ['rathmann1', 'rathmann3_oblique', 'ensemble_member'])
def test_oracle_equals_reference_bitwise(case):
    ref = _ref()
    if case == 'rathmann1':
        (nlm, z), F, alpha = profiles.rathmann_transition(41, 800, 1), 10, 0.0
    elif case == 'rathmann3_oblique':
        (nlm, z), F, alpha = profiles.rathmann_transition(41, 800, 3), 10, np.deg2rad(10)
    else:
        nl, z = profiles.ensemble(1, 50, 250, seed0=1000)
        (nlm, z), F, alpha = (nl[0], z), 8, np.deg2rad(10)
    ls = ref.LayerStack(nlm, z, N_frames=F, VERBOSE=0)
    want = ls.get_returns(1e3, f=179e6, alpha=alpha)
    got = orc.OracleStack(nlm, z, N_frames=F).get_returns(1e3, f=179e6, alpha=alpha, faithful=True)
    for n, a, b in zip(h.RET_NAMES, got, want):
        assert np.array_equal(a, b, equal_nan=True), n
