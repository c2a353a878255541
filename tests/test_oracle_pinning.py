"""Pin the C oracle: (1) against golden vectors produced by the compiled, unmodified reference
(tests/golden/, generator tools/make_golden.py) and (2), where oracle/_ref is built, directly
against the reference objects on fresh random inputs.  Integer/feature outputs and the scores
must be BIT-EXACT: the oracle restates the same double arithmetic in the same order."""
import os

import numpy as np
import pytest

from mipgen_b200 import panel
from mipgen_b200.panel import Config, Region
from oracle_api import Oracle, Ref, have_ref
from helpers import small_config, synthetic_regions, mutate, GOLDEN_DIR, random_model, tmpdir


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "reference_vectors.npz"), allow_pickle=True)


def test_survey_known_answer(oracle):
    """SURVEY.md Appendix D (measured on the compiled reference during the survey)."""
    g = panel.lcg_genome(2400, 12345)
    seq = g[1000:1400]
    s, e, l, cap = 1150, 20, 24, 162
    t = s + cap - e - l - 1
    ext, lig, tgt = seq[s - e - 1001:s - 1001], seq[t + 1 - 1001:t + 1 + l - 1001], seq[s - 1001:t + 1 - 1001]
    assert ext == b"AGCTCGAAAATTGAAATTCA" and lig == b"TCATTAAAATAGTACCTTATGCAA"
    assert "%.17g" % oracle.get_score(ext, lig, tgt, ext_copy=1, lig_copy=3) == "0.37788108837548862"
    mext = oracle.reverse_comp(seq[t + 1 - 1001:t + 1 + e - 1001])
    mlig = oracle.reverse_comp(seq[s - l - 1001:s - 1001])
    assert mext == b"ATAAGGTACTATTTTAATGA" and mlig == b"TGAATTTCAATTTTCGAGCTTGCC"
    assert "%.17g" % oracle.get_score(mext, mlig, oracle.reverse_comp(tgt), ext_copy=1, lig_copy=3) == "0.51124752005484275"
    lrc = oracle.long_range_content(g, 1001, 1400)
    assert lrc[0] == 0.5083333333333333 and lrc[22] == 0.49166666666666664
    p = oracle.get_parameters(ext, lig, tgt, lrc, ext_copy=1, lig_copy=3)
    assert p[21] == 20 and p[151] == 118 and p[173] == 24 and p[187] == 1 and p[190] == 0
    assert abs(p[191] - 0.477121) < 1e-6 and p[0] == 0.45 and p[15] == 0.3


def test_golden_long_range_content(oracle, golden):
    g = panel.lcg_genome(2400, int(golden["kat_genome_seed"][0]))
    assert np.array_equal(oracle.long_range_content(g, 1001, 1400), golden["kat_lrc"])


def test_golden_explicit_candidates(oracle, golden):
    n = len(golden["cand_ext"])
    for i in range(n):
        ext, lig, tgt = golden["cand_ext"][i], golden["cand_lig"][i], golden["cand_tgt"][i]
        ec, lc = (int(x) for x in golden["cand_copies"][i])
        lo = oracle.get_score(ext, lig, tgt, ext_copy=ec, lig_copy=lc)
        ft = oracle.get_parameters(ext, lig, tgt, golden["cand_lrc"][i], ext_copy=ec, lig_copy=lc)
        assert np.array_equal(np.array([lo]), golden["cand_logistic"][i:i + 1], equal_nan=True), i
        assert np.array_equal(ft, golden["cand_feat"][i], equal_nan=True), i
    assert (golden["cand_logistic"] == -1000.0).any()


def golden_regions(golden):
    e, l = golden["cfg_ext"].tolist(), golden["cfg_lig"].tolist()
    mx, mn, inc, ov = (int(x) for x in golden["cfg_caps"])
    cfg = Config(mx, mn, inc, ov, e, l)
    regs = []
    for i in range(2):
        sf, ef, a, b = (int(x) for x in golden["r%d_coords" % i])
        r = Region(sf, ef, a, b, golden["r%d_seq" % i].tobytes(), golden["r%d_lrc" % i])
        if "r%d_copies" % i in golden.files:
            r.copies = golden["r%d_copies" % i]
        regs.append(r)
    return cfg, regs


def test_golden_region_grids_and_svr(oracle, golden):
    cfg, regs = golden_regions(golden)
    h = oracle.svm_load_model(os.path.join(GOLDEN_DIR, "golden_svr.model"))
    for i, r in enumerate(regs):
        v, lo, sv, ft = oracle.grid_region(r, cfg, h, want_logistic=True, want_svr=True, want_feats=True)
        assert np.array_equal(v, golden["r%d_valid" % i])
        assert np.array_equal(lo, golden["r%d_logistic" % i], equal_nan=True)
        assert np.array_equal(sv, golden["r%d_svr" % i], equal_nan=True)
        assert np.array_equal(ft[golden["r%d_feat_rows" % i]], golden["r%d_feat" % i], equal_nan=True)
    pred = oracle.svm_predict_rows(h, golden["svr_X"])
    assert np.array_equal(pred, golden["svr_pred"])
    # the "%.17g" text round trip of mipgen::predict_value is the identity on doubles
    for k in (0, 3, 17):
        assert oracle.predict_value(h, golden["svr_X"][k]) == pred[k]
    oracle.svm_free(h)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_equals_compiled_reference_on_fresh_inputs(oracle):
    ref = Ref()
    rng = np.random.default_rng(77)
    cfg = small_config((40, 42, 45), 162, 152, 5)
    genome, regions = synthetic_regions(oracle, cfg, 2, 40, 140, 5150)
    regions[0].seq = mutate(regions[0].seq, rng, 15)
    regions.append(panel.cut_region(genome, 120, 170, cfg, 0, "edge"))  # clamped at the chromosome start
    regions[-1].lrc = rng.uniform(0, 0.2, 44)
    regions[1].copies = rng.choice([0, 1, 2, 7, 100, 101, 1000], size=(len(cfg.oligo_sizes), len(regions[1].seq))).astype(np.int32)
    d = tmpdir()
    path = random_model(oracle, cfg, 40, 12, os.path.join(d, "m.model"), sparse_tail=True)
    ho, hr = oracle.svm_load_model(path), ref.svm_load_model(path)
    for r in regions:
        if r.flank_seq is not None:
            assert np.array_equal(oracle.long_range_content(r.flank_seq, r.seq_start, r.seq_stop),
                                  ref.long_range_content(r.flank_seq, r.seq_start, r.seq_stop))
        a = oracle.grid_region(r, cfg, ho, want_logistic=True, want_svr=True, want_feats=True)
        b = ref.grid_region(r, cfg, hr, want_logistic=True, want_svr=True, want_feats=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y, equal_nan=True)
    oracle.svm_free(ho)
    ref.svm_free(hr)


def test_reverse_comp_passes_other_characters_through(oracle):
    assert oracle.reverse_comp(b"ACGTNRYacgt-") == b"-tgcaYRNACGT"


def test_tile_replay_counts_match_reference_cli_rule(oracle):
    """Static grid of one capture size: every scan start yields n_pairs*2 candidates when no
    score exceeds the limit ('exactly 114 x scan starts')."""
    cfg = Config()
    _g, regs = synthetic_regions(oracle, cfg, 1, 40, 41, 3, with_lrc=False)
    v, lo, _s, _f = oracle.grid_region(regs[0], cfg, None)
    idx = oracle.tile_replay(regs[0], cfg, v, lo, 0, True, 0.98)
    assert cfg.n_pairs == 57 and idx.size == cfg.n_scan(regs[0]) * 114 == v.sum()
