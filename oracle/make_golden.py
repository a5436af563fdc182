"""
Generates tests/golden/*.npz by running the UNMODIFIED reference (nicholasmr/tamm, /root/reference/lib) in the build
container.  The reference has no tests or stored outputs, so these fixtures are what pins the oracle.
Run:  python oracle/make_golden.py      (needs /root/reference; the fixtures are committed)
"""
import os
import sys

import numpy as np

REF = os.environ.get('TAMM_REFERENCE', '/root/reference/lib')
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
sys.path.insert(0, os.path.dirname(HERE))
from tamm_b200 import profiles  # noqa: E402


def run_case(name, nlm, z, F, alpha, f=179e6, E0=1e3, keep_layers=((0, 0), (1, 3)), **kw):
    sys.path.insert(0, REF)
    import layerstack as ref
    stack_kw = {k: kw[k] for k in ('epsa', 'epsc', 'sigma', 'mu', 'modeltype') if k in kw}
    ret_kw = {k: kw[k] for k in ('Tx_pol', 'URD', 'nn') if k in kw}
    ls = ref.LayerStack(nlm, z, N_frames=F, VERBOSE=0, **stack_kw)
    r = ls.get_returns(E0, f=f, alpha=alpha, **ret_kw)
    d = dict(nlm=nlm, z=z, F=F, alpha=alpha, f=f, E0=E0)
    for k, v in stack_kw.items():
        d['kw_' + k] = np.asarray(v)
    for k, v in ret_kw.items():
        d['kw_' + k] = np.asarray(v)
    for n, a in zip(['Pm_HH', 'Pm_HV', 'dP_HH', 'dP_HV', 'c_HHVV', 'E_HH', 'E_HV'], r):
        d['ret_' + n] = a
    if ret_kw.get('URD', 'URD') == 'URD':
        d['URD'] = ls.get_URD(ls.frm_layers)
    else:
        d['URD'] = ls.get_U(ls.frm_layers)
    d['propconst'] = ls.get_propconst()
    fp = stack_kw.get('modeltype', 'GTM') == 'FP'
    for (ff, nn) in keep_layers:
        L = ls.frm_layers[ff][nn]
        if fp:
            for attr in ['ai_horiz', 'eigang', 'ks', 'qs', 'Mfwd', 'Mrev']:
                d['layer_%d_%d_%s' % (ff, nn, attr)] = np.asarray(getattr(L, attr))
            continue
        for attr in ['epsilon', 'Delta', 'qs', 'gamma', 'Ai', 'Ai_inv', 'Ki', 'Mfwd', 'Mrev', 'Py', 'ccavg', 'nlm']:
            d['layer_%d_%d_%s' % (ff, nn, attr)] = np.asarray(getattr(L, attr))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **d)
    print('wrote', name, {k: np.shape(v) for k, v in d.items() if k.startswith('ret_')})


def main():
    os.makedirs(OUT, exist_ok=True)
    nlm, z = profiles.demo(100, 1500)
    run_case('demo_a10', nlm, z, 20, np.deg2rad(10), keep_layers=((0, 0), (3, 50), (19, 99)))
    run_case('demo_a0', nlm, z, 20, 0.0, keep_layers=((0, 0), (3, 50)))
    nlm, z = profiles.tilted(60, 600, seed=4)
    run_case('tilted_a10', nlm, z, 16, np.deg2rad(10), keep_layers=((0, 0), (5, 17), (15, 59)))
    run_case('tilted_a0_U', nlm, z, 16, 0.0, URD='U', keep_layers=((2, 9),))
    run_case('tilted_a25_nn', nlm, z, 9, np.deg2rad(25), nn=7, Tx_pol=([1, 0], [0, 1]), keep_layers=((4, 30),))
    N = 40
    nlm, z = profiles.tilted(N, 400, seed=11)
    rng = np.random.default_rng(5)
    run_case('perlayer_eps', nlm, z, 12, np.deg2rad(10), f=300e6,
             epsa=3.10 + 0.05 * rng.random(N), epsc=3.17 + 0.02 * rng.random(N), sigma=1e-5 * (1 + rng.random(N)),
             Tx_pol=([0.6, 0.8j],), keep_layers=((7, 21),))
    nlm, z = profiles.rathmann_transition(61, 2000, fabric_type=2)
    run_case('rathmann_t2_a0', nlm, z, 24, 0.0, keep_layers=((6, 30),))
    nlm, z = profiles.rathmann_transition(61, 2000, fabric_type=4)
    run_case('rathmann_t4_a10', nlm, z, 24, np.deg2rad(10), keep_layers=((6, 30),))
    main_ensemble()


def main_ensemble():
    """configs[4] recipe (profile 0 of the bench ensemble, seed 1000) at its layer spacing (0.5 m), 60 layers."""
    nlm, z = profiles.ensemble(1, 60, H=29.5)
    run_case('ensemble_a10', nlm[0], z, 16, np.deg2rad(10), keep_layers=((0, 0), (5, 17), (15, 59)))


def main_fp():
    nlm, z = profiles.demo(60, 900)
    run_case('fp_demo', nlm, z, 12, 0.0, modeltype='FP', keep_layers=((0, 0), (5, 30)))
    nlm, z = profiles.rathmann_transition(41, 800, fabric_type=4)
    run_case('fp_rathmann_t4', nlm, z, 10, 0.0, modeltype='FP', Tx_pol=([0.6, 0.8],), keep_layers=((3, 20),))
    N = 30
    nlm, z = profiles.tilted(N, 300, seed=8)
    rng = np.random.default_rng(6)
    run_case('fp_perlayer', nlm, z, 7, 0.0, f=300e6, modeltype='FP', epsa=3.10 + 0.05 * rng.random(N),
             epsc=3.17 + 0.02 * rng.random(N), sigma=1e-5 * (1 + rng.random(N)), nn=11, keep_layers=((2, 9),))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'fp':
        main_fp()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'ensemble':
        main_ensemble()
        sys.exit(0)
    main()
