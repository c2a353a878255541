#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref).

Needs /root/reference (to build oracle/_ref); the fixtures are committed so every other
box can pin the oracle and the CUDA path against the reference's own numbers.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mipgen_b200 import panel  # noqa: E402
from oracle_api import Ref  # noqa: E402
from helpers import small_config, synthetic_regions, mutate, GOLDEN_DIR  # noqa: E402


def main():
    ref = Ref()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    rng = np.random.default_rng(2024)

    # --- 1. known answer of SURVEY.md Appendix D (object harness) -----------------------
    g = panel.lcg_genome(2400, 12345)
    lrc = ref.long_range_content(g, 1001, 1400)

    # --- 2. explicit candidates incl. edge cases ----------------------------------------
    genome = panel.lcg_genome(30000, 4242)
    cands = []
    for i in range(400):
        e, l, t = int(rng.integers(16, 31)), int(rng.integers(16, 31)), int(rng.integers(60, 260))
        p = int(rng.integers(0, len(genome) - 400))
        ext, tgt, lig = genome[p:p + e], genome[p + e:p + e + t], genome[p + e + t:p + e + t + l]
        k = i % 8
        if k == 1:
            tgt = mutate(tgt, rng, 3)
        elif k == 2:
            ext = mutate(ext, rng, 1, b"N")
        elif k == 3:
            lig = mutate(lig, rng, 1, b"-")
        elif k == 4:
            lig = mutate(lig, rng, 2, b"RYKMacgt")
        elif k == 5:
            lig = b"R" + lig[1:]
        ec = int(rng.choice([1, 1, 1, 2, 3, 10, 100, 101, 0]))
        lc = int(rng.choice([1, 1, 1, 2, 3, 10, 100, 101]))
        cands.append((ext, lig, tgt, ec, lc))
    clrc = rng.uniform(0, 0.3, (len(cands), 44))
    c_log = np.array([ref.get_score(c[0], c[1], c[2], ext_copy=c[3], lig_copy=c[4]) for c in cands])
    c_feat = np.array([ref.get_parameters(c[0], c[1], c[2], clrc[i], ext_copy=c[3], lig_copy=c[4]) for i, c in enumerate(cands)])

    # --- 3. region grids (3 capture sizes, junk, copies) + SVR with a small model ----------
    cfg = small_config((40, 45), 162, 152, 5)
    _gen, regions = synthetic_regions(ref, cfg, 2, 30, 50, 808)
    regions[1].seq = mutate(regions[1].seq, rng, 10)
    regions[0].copies = rng.choice([0, 1, 1, 1, 2, 5, 100, 101], size=(len(cfg.oligo_sizes), len(regions[0].seq))).astype(np.int32)
    # model: SVs from the reference's own feature rows of another region
    _g2, mregs = synthetic_regions(ref, cfg, 1, 60, 61, 909)
    _v, _l, _s, mf = ref.grid_region(mregs[0], cfg, None, want_logistic=False, want_feats=True)
    mf = mf[np.isfinite(mf[:, 0])]
    sv = mf[rng.choice(mf.shape[0], 80, replace=False)]
    alpha = rng.uniform(-1, 1, 80)
    model_path = os.path.join(GOLDEN_DIR, "golden_svr.model")
    panel.write_svr_model(model_path, sv, alpha, 1.0 / 192, 0.3)
    h = ref.svm_load_model(model_path)
    out = {}
    for i, r in enumerate(regions):
        v, lo, sv_s, ft = ref.grid_region(r, cfg, h, want_logistic=True, want_svr=True, want_feats=True)
        rows = np.arange(0, v.size, 41)
        out["r%d_seq" % i] = np.frombuffer(r.seq, dtype=np.uint8)
        out["r%d_coords" % i] = np.array([r.start_flanked, r.stop_flanked, r.seq_start, r.seq_stop])
        out["r%d_lrc" % i] = r.lrc
        out["r%d_valid" % i] = v
        out["r%d_logistic" % i] = lo
        out["r%d_svr" % i] = sv_s
        out["r%d_feat_rows" % i] = rows
        out["r%d_feat" % i] = ft[rows]
        if r.copies is not None:
            out["r%d_copies" % i] = r.copies
    X = sv[:40] + rng.normal(0, 0.03, (40, 192))
    X[3] = 0.0
    pred = ref.svm_predict_rows(h, X)
    ref.svm_free(h)

    np.savez_compressed(
        os.path.join(GOLDEN_DIR, "reference_vectors.npz"),
        kat_genome_seed=np.array([12345]), kat_lrc=lrc,
        cand_ext=np.array([c[0] for c in cands], dtype=object), cand_lig=np.array([c[1] for c in cands], dtype=object),
        cand_tgt=np.array([c[2] for c in cands], dtype=object), cand_copies=np.array([[c[3], c[4]] for c in cands]),
        cand_lrc=clrc, cand_logistic=c_log, cand_feat=c_feat,
        cfg_ext=np.array(cfg.ext_len), cfg_lig=np.array(cfg.lig_len), cfg_caps=np.array([cfg.max_capture, cfg.min_capture, cfg.capture_increment, cfg.max_mip_overlap]),
        svr_X=X, svr_pred=pred, **out)
    print("wrote", os.path.join(GOLDEN_DIR, "reference_vectors.npz"), os.path.getsize(os.path.join(GOLDEN_DIR, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
