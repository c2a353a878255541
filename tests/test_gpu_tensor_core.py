"""K-svr on the tensor cores (mg_set_svr_mode(3): tcgen05 split-FP16 contraction, FP32 TMEM accumulators, FP64 exponent / exp /
row sum -- k_svr_tc.cu) against the oracle (libsvm's arithmetic restated, pinned to the compiled reference) and against the
FP64 kernels.  The north star allows 1e-6 relative on scores; this form is measured at ~1e-9, asserted at 1e-7."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mipgen_b200 as mg  # noqa: E402
from mipgen_b200 import panel  # noqa: E402
from helpers import random_model, calibrated_model, read_model_dense, rel_err, small_config, synthetic_regions, mutate, tmpdir  # noqa: E402

pytestmark = pytest.mark.gpu
TC_RTOL = 1e-7


@pytest.fixture(scope="module")
def oracle():
    from oracle_api import Oracle
    return Oracle()


def oracle_svr(oracle, cfg, regions, model):
    h = oracle.svm_load_model(model)
    want = np.concatenate([oracle.grid_region(r, cfg, h, want_logistic=False, want_svr=True)[2] for r in regions])
    oracle.svm_free(h)
    return want


@pytest.mark.parametrize("n_sv", [1, 63, 64, 65, 200])
def test_tensor_core_svr_matches_oracle(oracle, n_sv):
    cfg = small_config((40, 43, 45), 162, 157, 5)
    rng = np.random.default_rng(n_sv)
    genome, regions = synthetic_regions(oracle, cfg, 3, 30, 70, 500 + n_sv)
    # edge cases in the same panel: N / IUPAC / '-' / lower case, copy numbers 0 / 1 / 2 / 100 / 101, a region clamped at the chromosome start
    regions[1].seq = mutate(regions[1].seq, rng, 4)
    regions[2].copies = rng.choice([0, 1, 1, 1, 2, 5, 100, 101], size=(len(cfg.oligo_sizes), len(regions[2].seq))).astype(np.int32)
    regions.append(panel.cut_region(genome, 150, 230, cfg, 0, "edge"))
    regions[-1].lrc = rng.uniform(0, 0.3, 44)
    model = random_model(oracle, cfg, n_sv, 70 + n_sv, os.path.join(tmpdir(), "m.model"))
    want = oracle_svr(oracle, cfg, regions, model)
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    ctx.load_svr_model(model)
    assert ctx.svr_tensor_core_available()
    ctx.set_svr_mode(3)
    _o, valid, _l, got, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
    ctx.set_svr_mode(0)
    _o, valid2, _l, fp64, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
    t = ctx.timings()
    assert t.svr_tc_mma > 0, "the tcgen05 kernel must have run"
    assert np.array_equal(valid, valid2)
    e_oracle, e_fp64 = rel_err(got, want), rel_err(got, fp64)
    print("tensor-core SVR, %d SV: max rel err vs oracle %.2e, vs the FP64 kernel %.2e (FP64 kernel vs oracle %.2e)"
          % (n_sv, e_oracle, e_fp64, rel_err(fp64, want)))
    assert e_oracle <= TC_RTOL and e_fp64 <= TC_RTOL
    ctx.close()


def test_tensor_core_mode_refuses_models_it_cannot_represent(oracle):
    """Length / junction columns must be small integers to be exact in FP16; anything else keeps the FP64 kernels."""
    cfg = small_config((40, 45))
    _g, regions = synthetic_regions(oracle, cfg, 1, 30, 40, 321)
    d = tmpdir()
    model = random_model(oracle, cfg, 20, 9, os.path.join(d, "m.model"))
    sv, alpha, gamma = read_model_dense(model)
    sv[3, 21] += 0.25      # a fractional extension-arm length
    bad = os.path.join(d, "bad.model")
    panel.write_svr_model(bad, sv, alpha, gamma, 0.1)
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    ctx.load_svr_model(bad)
    assert not ctx.svr_tensor_core_available()
    ctx.set_svr_mode(3)
    with pytest.raises(mg.MgError):
        ctx.score_regions(regions, mg.MG_WANT_SVR)
    ctx.set_svr_mode(0)
    _o, valid, _l, got, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
    assert rel_err(got, oracle_svr(oracle, cfg, regions, bad)) <= 1e-9
    ctx.close()


def test_tensor_core_selection_equals_fp64_selection(oracle):
    """F8: what matters downstream is which MIP wins.  On a calibrated model (scores straddling 1.5 / 2.2) the condense /
    collapse winners computed from tensor-core scores equal the ones from FP64 scores."""
    cfg = small_config((40, 41, 42, 43, 44, 45), 162, 152, 5)
    genome, regions = synthetic_regions(oracle, cfg, 6, 100, 260, 8800)
    d = tmpdir()
    _v, _l, _s, feats = oracle.grid_region(regions[0], cfg, None, want_logistic=False, want_feats=True)
    model = calibrated_model(oracle, cfg, 256, 5, os.path.join(d, "m.model"), feats[np.isfinite(feats[:, 0])][::97])
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    ctx.load_svr_model(model)
    pnl = ctx.panel(regions)
    pnl.score(mg.MG_WANT_SVR)
    _v, _l, fp64, _ = pnl.fetch(valid=True, svr=True)
    so, sb, po, pb = pnl.select(regions, 1, 1.5, 2.2)
    ctx.set_svr_mode(3)
    pnl.score(mg.MG_WANT_SVR)
    valid, _l, tc, _ = pnl.fetch(valid=True, svr=True)
    so2, sb2, po2, pb2 = pnl.select(regions, 1, 1.5, 2.2)
    ok = valid.astype(bool)
    err = np.abs(tc[ok] - fp64[ok]) / np.abs(fp64[ok])
    print("tensor-core vs FP64 on %d candidates: max rel %.2e, median %.2e; scan winners differing %d / %d, position winners %d / %d"
          % (ok.sum(), err.max(), np.median(err), int((sb != sb2).sum()), sb.size, int((pb != pb2).sum()), pb.size))
    assert err.max() <= TC_RTOL
    assert np.array_equal(sb, sb2) and np.array_equal(pb, pb2)
    pnl.close()
    ctx.close()
