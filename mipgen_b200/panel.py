"""Synthetic inputs for the MIPgen scoring hot path.

Everything here is *input generation*: a deterministic LCG genome, BED-like target
regions, the arm-pair table of the reference's defaults and a writer for libsvm
text models.  No scoring arithmetic lives here.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_A = np.uint64(6364136223846793005)
_C = np.uint64(1442695040888963407)
_BLOCK = 1 << 14


def lcg_genome(n: int, seed: int) -> bytes:
    """n bases, iid uniform over ACGT: s = s*A + C (mod 2^64); base = "ACGT"[(s>>33)&3]."""
    with np.errstate(over="ignore"):
        # affine maps s0 -> s_k for k = 1.._BLOCK
        mult = np.empty(_BLOCK, dtype=np.uint64)
        add = np.empty(_BLOCK, dtype=np.uint64)
        m, a = np.uint64(1), np.uint64(0)
        for k in range(_BLOCK):
            m = m * _A
            a = a * _A + _C
            mult[k], add[k] = m, a
        out = np.empty(n, dtype=np.uint8)
        s = np.uint64(seed)
        lut = np.frombuffer(b"ACGT", dtype=np.uint8)
        pos = 0
        while pos < n:
            states = mult * s + add
            take = min(_BLOCK, n - pos)
            out[pos:pos + take] = lut[((states[:take] >> np.uint64(33)) & np.uint64(3)).astype(np.int64)]
            s = states[-1]
            pos += take
    return out.tobytes()


def write_fasta(path: str, name: str, seq: bytes, width: int = 60) -> None:
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        for i in range(0, len(seq), width):
            f.write(seq[i:i + width] + b"\n")


def default_arm_pairs(sums: Sequence[int] = (40, 41, 42, 43, 44, 45), lig_min: int = 18,
                      ext_min: int = 16) -> Tuple[List[int], List[int]]:
    """Arm pairs in the reference's enumeration order (mipgen.cpp:243-259, 431, 438):
    arm sum descending, extension length ascending within a sum."""
    ext, lig = [], []
    for s in sorted(set(sums), reverse=True):
        es = []
        l = lig_min
        while l <= s - ext_min and l <= 30:
            e = s - l
            if e <= 30:
                es.append(e)
            l += 1
        for e in sorted(es):
            ext.append(e)
            lig.append(s - e)
    return ext, lig


@dataclass
class Region:
    """One Featurev5 worth of input (Featurev5.h:10-23), coordinates 1-based inclusive."""
    start_flanked: int
    stop_flanked: int
    seq_start: int
    seq_stop: int
    seq: bytes
    lrc: Optional[np.ndarray] = None       # float64[44]
    flank_seq: Optional[bytes] = None      # region +- (max_capture+1000): input of the lrc
    label: str = "r"
    copies: Optional[np.ndarray] = None    # int32[n_oligo_sizes, len(seq)] or None (=> all 1)
    # selection-only inputs (mipgen.cpp:606-625, 634-760); None = absent
    masked_seq: Optional[bytes] = None     # masked_chromosomal_sequence: len(seq) characters, 'N' = TRF-masked
    snp: Optional[np.ndarray] = None       # uint8[len(seq)]: 1 where chr_snp_positions has an entry
    unmappable: Optional[np.ndarray] = None  # uint8[n_captures, len(seq)]: 1 where a MIP of that capture size starting here maps ambiguously


@dataclass
class Config:
    """The knobs of mipgen.cpp that shape the candidate grid."""
    max_capture: int = 162
    min_capture: int = 162
    capture_increment: int = 5
    max_mip_overlap: int = 30
    ext_len: List[int] = field(default_factory=lambda: default_arm_pairs()[0])
    lig_len: List[int] = field(default_factory=lambda: default_arm_pairs()[1])

    @property
    def n_pairs(self) -> int:
        return len(self.ext_len)

    @property
    def captures(self) -> List[int]:
        inc = self.capture_increment or 1
        return list(range(self.max_capture, self.min_capture - 1, -inc))

    @property
    def max_sum(self) -> int:
        return max(e + l for e, l in zip(self.ext_len, self.lig_len))

    @property
    def oligo_sizes(self) -> List[int]:
        return sorted(set(self.ext_len) | set(self.lig_len))

    def first_scan_start(self, r: Region) -> int:
        return max(0, r.start_flanked - self.max_capture + self.max_sum) + 1

    def n_scan(self, r: Region) -> int:
        return max(0, r.stop_flanked - self.first_scan_start(r) + 1)

    def grid_size(self, r: Region) -> int:
        return self.n_scan(r) * len(self.captures) * self.n_pairs * 2


def cut_region(genome: bytes, start: int, stop: int, cfg: Config, flank: int = 0, label: str = "r") -> Region:
    """Slice a region out of a chromosome the way get_chr_fasta_sequence_from_genome_dir does
    (mipgen.cpp:1214-1225).  start/stop are 1-based inclusive feature coordinates."""
    sf, ef = start - flank, stop + flank
    a = max(1, sf - cfg.max_capture)
    b = min(len(genome), ef + cfg.max_capture + 15)
    seq = genome[a - 1:b]
    lo = sf - cfg.max_capture - 1 - 1000
    flank_seq = genome[lo:lo + (b - a + 1) + 2000] if lo >= 0 else None
    return Region(sf, ef, a, b, seq, None, flank_seq, label)


def make_regions(genome: bytes, n: int, len_lo: int, len_hi: int, cfg: Config, seed: int,
                 first_start: int = 3000, gap: int = 2000) -> List[Region]:
    """n regions, lengths U[len_lo, len_hi], separated by >= gap so the BED merge rule
    (mipgen.cpp:1019) never fires."""
    rng = np.random.default_rng(seed)
    out, pos = [], first_start
    for i in range(n):
        ln = int(rng.integers(len_lo, len_hi + 1))
        start, stop = pos + 1, pos + ln   # BED [pos, pos+ln) -> 1-based [pos+1, pos+ln]
        if stop + cfg.max_capture + 1100 > len(genome):
            raise ValueError("genome too short for %d regions" % n)
        out.append(cut_region(genome, start, stop, cfg, 0, "t%04d" % i))
        pos = stop + gap + int(rng.integers(0, 500))
    return out


def genome_length_for(n_regions: int, len_hi: int, cfg: Config, first_start: int = 3000, gap: int = 2000) -> int:
    return first_start + n_regions * (len_hi + gap + 500) + cfg.max_capture + 2000


def write_bed(path: str, regions: Sequence[Region], chrom: str = "chr1") -> None:
    with open(path, "w") as f:
        for r in regions:
            f.write("%s\t%d\t%d\t%s\n" % (chrom, r.start_flanked - 1, r.stop_flanked, r.label))


def write_svr_model(path: str, sv: np.ndarray, alpha: np.ndarray, gamma: float, rho: float) -> None:
    """libsvm text model in the writer's own format (svm.cpp:2644-2736): coef "%.16g",
    non-zero features "idx:%.8g"."""
    sv = np.asarray(sv, dtype=np.float64)
    n = sv.shape[0]
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "w") as f:
        f.write("svm_type epsilon_svr\nkernel_type rbf\ngamma %.17g\nnr_class 2\ntotal_sv %d\nrho %.17g\nSV\n"
                % (gamma, n, rho))
        for i in range(n):
            row = sv[i]
            nz = np.nonzero(row)[0]
            f.write("%.16g " % alpha[i] + " ".join("%d:%.8g" % (j + 1, row[j]) for j in nz) + " \n")
    os.replace(tmp, path)


def calibrate(alpha: np.ndarray, raw_scores: np.ndarray, median: float = 1.8, upper: float = 2.2,
              frac_above: float = 0.15):
    """Scale alpha and pick rho so that scores a*f - rho have the given median and a
    `frac_above` share above `upper` (the 1.5 / 2.2 thresholds and the
    optimal-score shortcuts of mipgen.cpp:430,434 must be exercised)."""
    f = raw_scores[np.isfinite(raw_scores)]
    q50, qhi = np.quantile(f, [0.5, 1.0 - frac_above])
    spread = float(qhi - q50) or 1.0
    a = (upper - median) / spread
    rho = a * float(q50) - median
    return alpha * a, rho
