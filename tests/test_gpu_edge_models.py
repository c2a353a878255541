"""Edge cases of the SVR path on the device: support-vector counts around the kernels' chunk sizes (16 per chunk in
the factored kernel, 64 per slab in the dense one) and a configuration the factored kernel cannot take (an arm
longer than its tables allow), where auto mode must fall back to the dense contraction and stay exact."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mipgen_b200 as mg  # noqa: E402
from mipgen_b200 import panel  # noqa: E402
from helpers import random_model, rel_err, small_config, synthetic_regions, tmpdir  # noqa: E402

pytestmark = pytest.mark.gpu
SVR_RTOL = 1e-9


@pytest.fixture(scope="module")
def oracle():
    from oracle_api import Oracle
    return Oracle()


@pytest.fixture(scope="module")
def ctx():
    c = mg.Context(0)
    yield c
    c.close()


def oracle_svr(oracle, cfg, regions, model):
    h = oracle.svm_load_model(model)
    want = np.concatenate([oracle.grid_region(r, cfg, h, want_logistic=False, want_svr=True)[2] for r in regions])
    oracle.svm_free(h)
    return want


@pytest.mark.parametrize("n_sv", [1, 15, 16, 17, 33, 65])
def test_support_vector_counts_around_chunk_sizes(ctx, oracle, n_sv):
    cfg = small_config((40, 45))
    _genome, regions = synthetic_regions(oracle, cfg, 2, 20, 45, 400 + n_sv)
    model = random_model(oracle, cfg, n_sv, 50 + n_sv, os.path.join(tmpdir(), "m.model"))
    ctx.set_config(cfg)
    ctx.load_svr_model(model)
    want = oracle_svr(oracle, cfg, regions, model)
    assert ctx.svr_factored_available() > 0
    for mode in (2, 1):  # factored, dense
        ctx.set_svr_mode(mode)
        _o, valid, _l, got, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
        assert rel_err(got, want) <= SVR_RTOL, "mode %d, %d support vectors" % (mode, n_sv)
        assert np.isfinite(got[valid.astype(bool)]).all()
    ctx.set_svr_mode(0)


def test_config_outside_the_factored_tables_falls_back_to_dense(ctx, oracle):
    # a 64-base arm is beyond FACT_MAX_LEN: no factored tables for this configuration
    cfg = panel.Config(162, 157, 5, 30, [16, 64, 20, 25], [64, 16, 25, 20])
    _genome, regions = synthetic_regions(oracle, cfg, 2, 25, 50, 991)
    model = random_model(oracle, cfg, 40, 77, os.path.join(tmpdir(), "m.model"))
    ctx.set_config(cfg)
    ctx.load_svr_model(model)
    assert ctx.svr_factored_available() == 0
    want = oracle_svr(oracle, cfg, regions, model)
    ctx.set_svr_mode(0)
    _o, valid, _l, got, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
    assert valid.any() and rel_err(got, want) <= SVR_RTOL
    ctx.set_svr_mode(2)
    with pytest.raises(mg.MgError):
        ctx.score_regions(regions, mg.MG_WANT_SVR)
    ctx.set_svr_mode(0)
