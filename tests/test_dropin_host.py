"""The drop-in lib/ API (LayerStack.get_returns / get_URD / get_U / frm_layers / get_propconst, Layer attributes, FP
model) against the reference fixtures WITHOUT a GPU: the very test bodies of tests/test_gpu_parity.py run with the
engine replaced by the host-simulator double (tests/hostsim_engine.py).  This covers the host glue of lib/ and the kernel
arithmetic together; the GPU tier repeats them on the real kernels."""
import numpy as np
import pytest

import helpers as h
from hostsim_engine import HostSimEngine
import test_gpu_parity as tg


@pytest.fixture()
def eng(built, monkeypatch):
    import layer
    import layerstack
    e = HostSimEngine()
    monkeypatch.setattr(layer, '_get_engine', lambda: e)
    monkeypatch.setattr(layerstack, '_get_engine', lambda: e)
    return e


@pytest.mark.parametrize('name', h.golden_names())
def test_get_returns_fixture(eng, name):
    tg.test_get_returns_matches_reference_fixture(eng, name)


@pytest.mark.parametrize('name', h.golden_names())
def test_get_URD_get_U_fixture(eng, name):
    tg.test_get_URD_get_U_match_reference_fixture(eng, name)


@pytest.mark.parametrize('name', h.golden_names())
def test_layer_attributes_fixture(eng, name):
    tg.test_layer_attributes_match_reference_fixture(eng, name)


@pytest.mark.parametrize('name', h.golden_names('FP'))
def test_fujita_paren_fixture(eng, name):
    tg.test_fujita_paren_model_matches_reference_fixture(eng, name)


def test_frequency_axis(eng):
    tg.test_frequency_axis_equals_scalar_calls(eng)


def test_polarimetric(eng):
    tg.test_polarimetric_extension_matches_oracle(eng)


def test_isotropic_and_degenerate(eng):
    tg.test_isotropic_and_degenerate_layers(eng)


def test_nonunit_permeability(eng):
    tg.test_nonunit_permeability(eng)


@pytest.mark.parametrize('name,kw,alpha_deg', [('lossless_tilted', dict(sigma=[0.0]), 10.0), ('grazing_tilted', {}, 89.0),
                                               ('strong_loss', dict(sigma=[1e-2]), 30.0)])
def test_unusual_regimes(eng, name, kw, alpha_deg):
    tg.test_unusual_regimes(eng, name, kw, alpha_deg)


def test_ensemble_axis(eng):
    tg.test_ensemble_axis_equals_single_profiles(eng)
