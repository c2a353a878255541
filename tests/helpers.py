"""Shared builders for the parity tests: seeded synthetic regions, models, edge cases."""
from __future__ import annotations

import os
import tempfile
from typing import List, Tuple

import numpy as np

from mipgen_b200 import panel
from mipgen_b200.panel import Config, Region

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def small_config(sums=(40, 45), max_capture=162, min_capture=162, inc=5) -> Config:
    e, l = panel.default_arm_pairs(sums)
    return Config(max_capture, min_capture, inc, 30, e, l)


def synthetic_regions(oracle, cfg: Config, n: int, len_lo: int, len_hi: int, seed: int, with_lrc=True) -> Tuple[bytes, List[Region]]:
    glen = panel.genome_length_for(n, len_hi, cfg)
    genome = panel.lcg_genome(glen, seed)
    regions = panel.make_regions(genome, n, len_lo, len_hi, cfg, seed + 1)
    if with_lrc:
        for r in regions:
            r.lrc = oracle.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    return genome, regions


def mutate(seq: bytes, rng: np.random.Generator, n_runs: int = 3, alphabet: bytes = b"NNNRYKM-acgtn") -> bytes:
    """Sprinkle N runs, IUPAC codes, '-' and lower-case letters into a sequence."""
    a = bytearray(seq)
    for _ in range(n_runs):
        pos = int(rng.integers(0, len(a)))
        ln = int(rng.integers(1, 6))
        ch = alphabet[int(rng.integers(0, len(alphabet)))]
        for i in range(pos, min(len(a), pos + ln)):
            a[i] = ch
    return bytes(a)


def random_model(oracle, cfg: Config, n_sv: int, seed: int, path: str, sparse_tail: bool = False):
    """A synthetic libsvm model whose SVs are real feature vectors of random candidates from
    a different genome seed.  Returns (sv, alpha, gamma, rho) as written."""
    rng = np.random.default_rng(seed)
    _g, regs = synthetic_regions(oracle, cfg, 2, 120, 200, seed + 1000)
    rows = []
    for r in regs:
        _v, _l, _s, feats = oracle.grid_region(r, cfg, None, want_logistic=False, want_feats=True)
        ok = np.isfinite(feats[:, 0])
        rows.append(feats[ok])
    F = np.concatenate(rows)
    sv = F[rng.choice(F.shape[0], n_sv, replace=F.shape[0] < n_sv)]
    alpha = rng.uniform(-1, 1, n_sv)
    gamma = 1.0 / 192
    panel.write_svr_model(path, sv, alpha, gamma, 0.0)
    if sparse_tail:
        # features beyond index 192 on a few SVs: libsvm adds their squares to the distance
        lines = open(path).read().split("\n")
        hdr = lines.index("SV")
        for k in range(hdr + 1, min(hdr + 6, len(lines))):
            if lines[k].strip():
                lines[k] = lines[k].rstrip() + " 200:0.125 "
        open(path, "w").write("\n".join(lines))
    return path


def calibrated_model(oracle, cfg: Config, n_sv: int, seed: int, path: str, sample: np.ndarray):
    """Scale alpha / set rho so scores straddle the 1.5 / 2.2 thresholds (pruning fires)."""
    tmp = path + ".raw"
    random_model(oracle, cfg, n_sv, seed, tmp)
    h = oracle.svm_load_model(tmp)
    raw = oracle.svm_predict_rows(h, sample)
    oracle.svm_free(h)
    # re-read what was written so the calibrated file keeps the %.8g-rounded SVs
    sv, alpha, gamma = read_model_dense(tmp)
    alpha2, rho = panel.calibrate(alpha, raw)
    panel.write_svr_model(path, sv, alpha2, gamma, rho)
    os.unlink(tmp)
    return path


def read_model_dense(path: str):
    lines = open(path).read().split("\n")
    gamma = float([l for l in lines if l.startswith("gamma")][0].split()[1])
    hdr = lines.index("SV")
    sv, alpha = [], []
    for l in lines[hdr + 1:]:
        if not l.strip():
            continue
        toks = l.split()
        alpha.append(float(toks[0]))
        row = np.zeros(192)
        for t in toks[1:]:
            i, v = t.split(":")
            if 1 <= int(i) <= 192:
                row[int(i) - 1] = float(v)
        sv.append(row)
    return np.array(sv), np.array(alpha), gamma


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    a, b = np.asarray(a, float), np.asarray(b, float)
    m = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    if not m.any():
        return 0.0
    d = np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-300)
    return float(d.max())


def tmpdir() -> str:
    return tempfile.mkdtemp(prefix="mipgen_b200_")
