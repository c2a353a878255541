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


def test_random_configs_match_oracle(ctx, oracle):
    """Random capture ranges / increments / arm-sum sets (incl. increments that do not divide the range and a
    region clamped at the chromosome start): validity and features bit-exact, scores within tolerance."""
    rng = np.random.default_rng(77)
    all_sums = list(range(36, 50))
    for trial in range(6):
        sums = tuple(sorted(rng.choice(all_sums, int(rng.integers(1, 4)), replace=False).tolist()))
        max_cap = int(rng.integers(120, 260))
        min_cap = int(max(max(sums) + 1, max_cap - rng.integers(0, 30)))
        inc = int(rng.choice([0, 1, 3, 5, 7]))
        cfg = small_config(sums, max_cap, min_cap, inc)
        genome, regions = synthetic_regions(oracle, cfg, 2, 5, 40, 600 + trial)
        edge = panel.cut_region(genome, int(rng.integers(20, 60)), int(rng.integers(70, 110)), cfg, 0, "edge")
        edge.lrc = rng.uniform(0, 0.3, 44)
        regions.append(edge)
        model = random_model(oracle, cfg, 20, 900 + trial, os.path.join(tmpdir(), "m.model"))
        ctx.set_config(cfg)
        ctx.load_svr_model(model)
        offs, valid, lo, sv, ft = ctx.score_regions(regions, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR | mg.MG_WANT_FEATURES)
        _o, valid2, _l, sv2, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)  # workspace path (factored kernel when it fits)
        h = oracle.svm_load_model(model)
        for i, r in enumerate(regions):
            a, b = offs[i], offs[i + 1]
            wv, wl, ws, wf = oracle.grid_region(r, cfg, h, want_logistic=True, want_svr=True, want_feats=True)
            tag = (trial, sums, max_cap, min_cap, inc, i)
            assert b - a == wv.size, tag
            assert np.array_equal(valid[a:b], wv) and np.array_equal(valid2[a:b], wv), tag
            ok = wv.astype(bool)
            assert np.array_equal(ft[a:b][ok], wf[ok]), tag
            assert rel_err(lo[a:b], wl) <= 1e-12, tag
            assert rel_err(sv[a:b], ws) <= SVR_RTOL and rel_err(sv2[a:b], ws) <= SVR_RTOL, tag
        oracle.svm_free(h)
