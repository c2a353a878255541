"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle.

Bars (BASELINE.json north_star):
  * enumeration / static validity ............ bit-exact
  * the 192 features ......................... bit-exact (each is one IEEE division)
  * logistic score ........................... <= 1e-12 relative (north star allows 1e-6;
                                               only pow() differs: CUDA vs glibc, <= 2 ulp)
  * SVR score ................................ <= 1e-9 relative (north star allows 1e-6;
                                               FP64 DMMA contraction vs libsvm's sequential sums)
"""
import os

import numpy as np
import pytest

import mipgen_b200 as mg
from mipgen_b200 import panel
from helpers import (small_config, synthetic_regions, mutate, random_model, calibrated_model, rel_err, tmpdir,
                     read_model_dense)

pytestmark = pytest.mark.gpu

LOGISTIC_RTOL = 1e-12
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


def test_library_reports_version(ctx):
    assert b"sm_100a" in ctx.lib.mg_version()


def test_long_range_content_bit_exact(ctx, oracle):
    rng = np.random.default_rng(5)
    for n in (2400, 3111, 2739):
        g = panel.lcg_genome(n, 100 + n)
        g = mutate(g, rng, 4)
        want = oracle.long_range_content(g, 1001, 1001 + n - 2001)
        got = ctx.long_range_content(g, 1001, 1001 + n - 2001)
        assert np.array_equal(want, got)


def _random_candidates(rng, n, genome, edge=False):
    cands = []
    for _ in range(n):
        e, l = int(rng.integers(16, 31)), int(rng.integers(16, 31))
        t = int(rng.integers(60, 260))
        p = int(rng.integers(0, len(genome) - 400))
        ext, tgt, lig = genome[p:p + e], genome[p + e:p + e + t], genome[p + e + t:p + e + t + l]
        c = dict(ext=ext, lig=lig, tgt=tgt, ext_copy=int(rng.choice([1, 1, 1, 2, 3, 10, 100, 101])),
                 lig_copy=int(rng.choice([1, 1, 1, 2, 3, 10, 100, 101])))
        if edge:
            which = int(rng.integers(0, 6))
            if which == 0:
                c["tgt"] = mutate(tgt, rng, 3)                       # N / IUPAC / '-' / lower case in the insert
            elif which == 1:
                c["ext"] = mutate(ext, rng, 1, b"N")                # N in an arm -> invalid
            elif which == 2:
                c["lig"] = mutate(lig, rng, 1, b"-")                # '-' in an arm -> invalid via mip_seq
            elif which == 3:
                c["lig"] = mutate(lig, rng, 2, b"RYKMacgt")          # IUPAC in an arm: scored, never matches
            elif which == 4:
                c["lig"] = b"R" + lig[1:]                            # unknown junction key -> score 0 / no one-hot
            else:
                c["ext_copy"], c["lig_copy"] = 0, 1                  # log10(0) = -inf
        cands.append(c)
    return cands


@pytest.mark.parametrize("edge", [False, True])
def test_explicit_candidates_match_oracle(ctx, oracle, edge):
    rng = np.random.default_rng(11 + edge)
    genome = panel.lcg_genome(20000, 77)
    cands = _random_candidates(rng, 700, genome, edge)
    lrc = rng.uniform(0, 0.3, (len(cands), 44))
    lo, _sv, ft = ctx.score_candidates(cands, lrc, mg.MG_WANT_LOGISTIC | mg.MG_WANT_FEATURES)
    want_lo = np.array([oracle.get_score(c["ext"], c["lig"], c["tgt"], ext_copy=c["ext_copy"], lig_copy=c["lig_copy"]) for c in cands])
    want_ft = np.array([oracle.get_parameters(c["ext"], c["lig"], c["tgt"], lrc[i], ext_copy=c["ext_copy"], lig_copy=c["lig_copy"])
                        for i, c in enumerate(cands)])
    assert np.array_equal(want_ft, ft, equal_nan=True), "features must be bit-exact"
    assert rel_err(lo, want_lo) <= LOGISTIC_RTOL
    if edge:
        assert (want_lo == -1000.0).any() and np.array_equal(lo == -1000.0, want_lo == -1000.0)


def test_explicit_empty_and_ragged(ctx, oracle):
    lo, sv, ft = ctx.score_candidates([], None, mg.MG_WANT_LOGISTIC | mg.MG_WANT_FEATURES)
    assert lo.size == 0 and ft.shape == (0, 192)
    # strings shorter than the constructor lengths (substr clamps near a sequence end)
    c = dict(ext=b"ACGTACGTACGTACGTAC", lig=b"GGCATCGATCGATCGATCGA", tgt=b"ACGT" * 25, ext_len=20, lig_len=22, scan_size=104)
    lo, _s, ft = ctx.score_candidates([c], np.zeros((1, 44)), mg.MG_WANT_LOGISTIC | mg.MG_WANT_FEATURES)
    want = oracle.get_parameters(c["ext"], c["lig"], c["tgt"], np.zeros(44), ext_len=20, lig_len=22, scan_size=104)
    assert np.array_equal(want, ft[0])
    assert rel_err(lo, [oracle.get_score(c["ext"], c["lig"], c["tgt"], ext_len=20, lig_len=22, scan_size=104)]) <= LOGISTIC_RTOL


def _grid_case(oracle, cfg, regions, model_path):
    h = oracle.svm_load_model(model_path) if model_path else None
    out = []
    for r in regions:
        out.append(oracle.grid_region(r, cfg, h, want_logistic=True, want_svr=h is not None, want_feats=True))
    if h:
        oracle.svm_free(h)
    return out


def test_region_grid_matches_oracle(ctx, oracle):
    cfg = small_config((40, 43, 45), 162, 152, 5)
    rng = np.random.default_rng(3)
    genome, regions = synthetic_regions(oracle, cfg, 3, 30, 70, 21)
    # a region hard against the chromosome start (scan start clamp, bounds skips) and one with junk
    edge = panel.cut_region(genome, 150, 190, cfg, 0, "edge")
    edge.lrc = rng.uniform(0, 0.3, 44)
    regions.append(edge)
    dirty = regions[1]
    dirty.seq = mutate(dirty.seq, rng, 12)
    # copy tables on one region
    r0 = regions[0]
    r0.copies = rng.choice([0, 1, 1, 1, 2, 5, 100, 101], size=(len(cfg.oligo_sizes), len(r0.seq))).astype(np.int32)
    d = tmpdir()
    model = random_model(oracle, cfg, 150, 9, os.path.join(d, "m.model"), sparse_tail=True)
    ctx.set_config(cfg)
    ctx.load_svr_model(model)
    offs, valid, lo, sv, ft = ctx.score_regions(regions, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR | mg.MG_WANT_FEATURES)
    want = _grid_case(oracle, cfg, regions, model)
    for i, (wv, wl, ws, wf) in enumerate(want):
        a, b = offs[i], offs[i + 1]
        assert b - a == wv.size == cfg.grid_size(regions[i])
        assert np.array_equal(valid[a:b], wv), "static validity (mipgen.cpp:429,443,444) must match"
        ok = wv.astype(bool)
        assert np.array_equal(ft[a:b][ok], wf[ok]), "features must be bit-exact (region %d)" % i
        assert rel_err(lo[a:b], wl) <= LOGISTIC_RTOL
        assert rel_err(sv[a:b], ws) <= SVR_RTOL
    assert not valid.all() and valid.any()


def test_factored_svr_matches_dense_and_oracle(ctx, oracle):
    """The factored kernel (distinct arms/inserts, product of block factors) and the dense DMMA
    contraction are two evaluations of the same FP64 decision function."""
    rng = np.random.default_rng(17)
    for cfg in (small_config((40, 43, 45), 162, 152, 5), panel.Config()):
        genome, regions = synthetic_regions(oracle, cfg, 3, 30, 80, 77)
        edge = panel.cut_region(genome, 150, 190, cfg, 0, "edge")   # clamped at the chromosome start
        edge.lrc = rng.uniform(0, 0.3, 44)
        regions.append(edge)
        regions[1].seq = mutate(regions[1].seq, rng, 12)            # N / IUPAC / '-' : zero rows, odd run counts
        regions[0].copies = rng.choice([0, 1, 1, 1, 2, 5, 100, 101], size=(len(cfg.oligo_sizes), len(regions[0].seq))).astype(np.int32)
        d = tmpdir()
        model = random_model(oracle, cfg, 150, 9, os.path.join(d, "m.model"), sparse_tail=True)
        ctx.set_config(cfg)
        ctx.load_svr_model(model)
        assert ctx.svr_factored_available() > 0
        ctx.set_svr_mode(1)
        _o, v1, _l, dense, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
        ctx.set_svr_mode(2)
        _o, v2, _l, fact, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
        _o, v3, _l, fact_f, _f = ctx.score_regions(regions, mg.MG_WANT_SVR | mg.MG_WANT_FEATURES)
        ctx.set_svr_mode(0)
        assert np.array_equal(v1, v2) and np.array_equal(v1, v3)
        assert rel_err(fact, dense) <= 1e-11
        assert np.array_equal(fact, fact_f, equal_nan=True)
        h = oracle.svm_load_model(model)
        want = np.concatenate([oracle.grid_region(r, cfg, h, want_logistic=False, want_svr=True)[2] for r in regions])
        oracle.svm_free(h)
        assert rel_err(fact, want) <= SVR_RTOL
        assert np.isfinite(dense[v1.astype(bool)]).all() and np.isnan(dense[~v1.astype(bool)]).all()


def test_row_table_workspace_follows_the_configuration(ctx, oracle):
    """The K-feat -> K-svr row-table workspace is reused between calls; its per-work-item stride depends on the arm table, so a
    context that goes from a narrow arm table to a wide one with the same number of work items must re-size it (a run of the
    drop-in CLI does exactly this: one mg_set_config per batch)."""
    d = tmpdir()
    narrow = small_config((45,), 162, 152, 5)
    wide = panel.Config(162, 152, 5, 30, *panel.default_arm_pairs((40, 41, 42, 43, 44, 45)))
    _g, regions = synthetic_regions(oracle, wide, 3, 60, 120, 91)
    for cfg in (narrow, wide, narrow):
        model = random_model(oracle, cfg, 64, 3, os.path.join(d, "m%d.model" % len(cfg.ext_len)))
        ctx.set_config(cfg)
        ctx.load_svr_model(model)
        assert ctx.svr_factored_available() > 0
        _o, valid, _l, got, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
        h = oracle.svm_load_model(model)
        want = np.concatenate([oracle.grid_region(r, cfg, h, want_logistic=False, want_svr=True)[2] for r in regions])
        oracle.svm_free(h)
        assert rel_err(got, want) <= SVR_RTOL


def test_region_grid_chunked_svr_equals_feature_path(ctx, oracle):
    """SVR through the chunked workspace path == SVR computed while features are kept."""
    cfg = small_config((40, 45))
    _g, regions = synthetic_regions(oracle, cfg, 4, 40, 90, 31)
    d = tmpdir()
    ctx.set_config(cfg)
    ctx.load_svr_model(random_model(oracle, cfg, 100, 4, os.path.join(d, "m.model")))
    _o, v1, _l, s1, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
    _o, v2, _l, s2, _f = ctx.score_regions(regions, mg.MG_WANT_SVR | mg.MG_WANT_FEATURES)
    assert np.array_equal(v1, v2) and np.array_equal(s1, s2, equal_nan=True)


def test_empty_inputs(ctx, oracle):
    ctx.set_config(small_config())
    offs, valid, lo, _s, _f = ctx.score_regions([], mg.MG_WANT_LOGISTIC)
    assert offs.tolist() == [0] and valid.size == 0 and lo.size == 0


def test_svr_predict_matches_libsvm_order(ctx, oracle):
    cfg = small_config((40, 45))
    d = tmpdir()
    path = random_model(oracle, cfg, 333, 8, os.path.join(d, "m.model"), sparse_tail=True)  # not a multiple of 64
    ctx.load_svr_model(path)
    n_sv, gamma, rho = ctx.model_info()
    assert n_sv == 333 and gamma == 1.0 / 192
    sv, _a, _g = read_model_dense(path)
    rng = np.random.default_rng(2)
    X = sv[rng.integers(0, sv.shape[0], 300)] + rng.normal(0, 0.02, (300, 192))
    X[5] = 0.0                       # the all-zero vector of an invalid candidate (SVMipv4.cpp:63-68)
    X[6, 190] = -np.inf              # log10(0) copy feature: every kernel value becomes 0
    h = oracle.svm_load_model(path)
    want = oracle.svm_predict_rows(h, X)
    oracle.svm_free(h)
    got = ctx.svr_predict(X)
    assert rel_err(got, want) <= SVR_RTOL
    assert got[6] == -rho
    direct = ctx.svr_predict(X[:64], direct=True)
    assert rel_err(direct[np.arange(64) != 6], want[:64][np.arange(64) != 6]) <= 1e-14


def test_svr_rows_are_position_independent(ctx, oracle):
    """Same feature vector => same score bits wherever the row sits in a tile."""
    cfg = small_config((40, 45))
    d = tmpdir()
    path = random_model(oracle, cfg, 128, 6, os.path.join(d, "m.model"))
    ctx.load_svr_model(path)
    sv, _a, _g = read_model_dense(path)
    row = sv[3] * 0.9
    X = np.tile(row, (200, 1))
    got = ctx.svr_predict(X)
    assert np.all(got == got[0])


def last_pair_threshold(cfg, valid, score):
    """A score threshold that the LAST pair of the first arm-sum list beats about half the
    time: previous_best_score is the best of the last evaluated pair only (mipgen.cpp:495),
    and the synthetic model is length dominated, so a fixed 2.2 may never fire."""
    sums = [e + l for e, l in zip(cfg.ext_len, cfg.lig_len)]
    last = max(i for i, s in enumerate(sums) if s == sums[0])
    g = score.reshape(-1, len(cfg.captures), cfg.n_pairs, 2)
    v = valid.reshape(g.shape).astype(bool)
    best = np.nanmax(np.where(v[:, :, last, :], g[:, :, last, :], np.nan), axis=-1)
    return float(np.nanmedian(best))


def test_replay_of_score_dependent_skips(ctx, oracle):
    """mg_tile_replay over GPU scores == the oracle's replay over oracle scores, with
    thresholds placed so the optimal-score shortcuts (mipgen.cpp:430,434) really fire."""
    cfg = small_config((40, 42, 45), 162, 152, 5)
    _g, regions = synthetic_regions(oracle, cfg, 2, 130, 160, 41)
    d = tmpdir()
    _v, _l, _s, feats = oracle.grid_region(regions[0], cfg, None, want_logistic=False, want_feats=True)
    sample = feats[np.isfinite(feats[:, 0])][::37]
    path = calibrated_model(oracle, cfg, 96, 13, os.path.join(d, "cal.model"), sample)
    ctx.set_config(cfg)
    ctx.load_svr_model(path)
    offs, valid, lo, sv, _f = ctx.score_regions(regions, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    h = oracle.svm_load_model(path)
    fired = 0
    for i, r in enumerate(regions):
        wv, wl, ws, _ = oracle.grid_region(r, cfg, h, want_logistic=True, want_svr=True)
        a, b = offs[i], offs[i + 1]
        cases = [(1, sv[a:b], ws, 2.2), (1, sv[a:b], ws, last_pair_threshold(cfg, wv, ws)),
                 (0, lo[a:b], wl, 0.98), (0, lo[a:b], wl, last_pair_threshold(cfg, wv, wl)),
                 (2, lo[a:b], wl, last_pair_threshold(cfg, wv, wl))]
        for method, score_g, score_o, upper in cases:
            for heuristic in (True, False):
                got = mg.tile_replay(cfg, r, valid[a:b], score_g, method, heuristic, upper)
                want = oracle.tile_replay(r, cfg, wv, score_o, method, heuristic, upper)
                assert np.array_equal(got, want)
                fired += want.size < wv.sum()
    oracle.svm_free(h)
    assert fired >= 4, "no pruning fired: the replay path would be untested"
