"""ctypes bindings for the test-only checkers under oracle/.

liboracle.so     -- the C restatement (always buildable: gcc only)
libmipgen_ref.so -- the unmodified reference objects behind ref_harness.cpp
                    (prebuilt in oracle/_ref; rebuilt only where /root/reference exists)

Both expose the same region/config structs and grid call so a test can swap one for
the other.  Nothing in mipgen_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

from mipgen_b200.panel import Config, Region

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libmipgen_ref.so")
REF_CLI = os.path.join(ORACLE_DIR, "_ref", "mipgen")
REF_BWA_DIR = os.path.join(ORACLE_DIR, "_ref")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)
c_long_p = C.POINTER(C.c_long)


class CMip(C.Structure):
    _fields_ = [("ext", C.c_char_p), ("ext_n", C.c_int), ("lig", C.c_char_p), ("lig_n", C.c_int),
                ("tgt", C.c_char_p), ("tgt_n", C.c_int), ("ext_len", C.c_int), ("lig_len", C.c_int),
                ("scan_size", C.c_int), ("ext_copy", C.c_int), ("lig_copy", C.c_int)]


class CRegion(C.Structure):
    _fields_ = [("seq", C.c_char_p), ("seq_len", C.c_int), ("seq_start", C.c_int), ("seq_stop", C.c_int),
                ("start_flanked", C.c_int), ("stop_flanked", C.c_int), ("lrc", c_double_p),
                ("copies", c_int_p)]


class CSel(C.Structure):
    _fields_ = [("lower_score_limit", C.c_double), ("upper_score_limit", C.c_double), ("max_arm_copy", C.c_int),
                ("target_arm_copy", C.c_int), ("masked_arm_threshold", C.c_double), ("masked_seq", C.c_char_p),
                ("snp", c_ubyte_p), ("unmappable", c_ubyte_p)]


class CCfg(C.Structure):
    _fields_ = [("max_capture", C.c_int), ("min_capture", C.c_int), ("capture_increment", C.c_int),
                ("max_mip_overlap", C.c_int), ("n_pairs", C.c_int), ("ext_len", c_int_p),
                ("lig_len", c_int_p), ("n_oligo_sizes", C.c_int), ("oligo_sizes", c_int_p)]


def build_oracle(force: bool = False) -> None:
    """Compile oracle/ (and oracle/_ref when the reference sources are present)."""
    need = force or not os.path.exists(ORACLE_SO) or \
        os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "mipgen_oracle.c"))
    if need:
        subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
    if os.path.exists("/root/reference/mipgen.cpp"):
        stale = not os.path.exists(REF_SO) or not os.path.exists(REF_CLI) or \
            os.path.getmtime(REF_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "ref_harness.cpp"))
        if force or stale:
            subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, capture_output=True)


def _np_ptr(a: Optional[np.ndarray], ptr_type):
    return a.ctypes.data_as(ptr_type) if a is not None else ptr_type()


class _Keep:
    """Holds numpy buffers alive for the lifetime of a ctypes struct."""

    def __init__(self):
        self.refs = []


def c_cfg(cfg: Config, keep: _Keep) -> CCfg:
    e = np.asarray(cfg.ext_len, dtype=np.int32)
    l = np.asarray(cfg.lig_len, dtype=np.int32)
    o = np.asarray(cfg.oligo_sizes, dtype=np.int32)
    keep.refs += [e, l, o]
    return CCfg(cfg.max_capture, cfg.min_capture, cfg.capture_increment, cfg.max_mip_overlap,
                len(e), _np_ptr(e, c_int_p), _np_ptr(l, c_int_p), len(o), _np_ptr(o, c_int_p))


def c_region(r: Region, keep: _Keep) -> CRegion:
    lrc = np.ascontiguousarray(r.lrc, dtype=np.float64) if r.lrc is not None else None
    cop = np.ascontiguousarray(r.copies, dtype=np.int32) if r.copies is not None else None
    keep.refs += [lrc, cop, r.seq]
    return CRegion(r.seq, len(r.seq), r.seq_start, r.seq_stop, r.start_flanked, r.stop_flanked,
                   _np_ptr(lrc, c_double_p), _np_ptr(cop, c_int_p))


class _Lib:
    """Common surface of liboracle.so (prefix orc_) and libmipgen_ref.so (prefix ref_)."""

    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.p = prefix
        L = self.lib
        g = lambda n: getattr(L, prefix + n)
        g("get_score").restype = C.c_double
        g("get_score").argtypes = [C.POINTER(CMip)]
        g("get_parameters").restype = None
        g("get_parameters").argtypes = [C.POINTER(CMip), c_double_p, c_double_p]
        g("svm_load_model").restype = C.c_void_p
        g("svm_load_model").argtypes = [C.c_char_p]
        g("svm_free").restype = None
        g("svm_free").argtypes = [C.c_void_p]
        g("svm_nsv").restype = C.c_int
        g("svm_nsv").argtypes = [C.c_void_p]
        g("svm_predict").restype = C.c_double
        g("svm_predict").argtypes = [C.c_void_p, c_double_p, C.c_int]
        g("grid_region").restype = None
        g("grid_region").argtypes = [C.POINTER(CRegion), C.POINTER(CCfg), C.c_void_p, c_ubyte_p,
                                     c_double_p, c_double_p, c_double_p]
        if prefix == "orc_":
            L.orc_long_range_content.restype = None
            L.orc_long_range_content.argtypes = [C.c_char_p, C.c_int, C.c_int, c_double_p]
            L.orc_predict_value.restype = C.c_double
            L.orc_predict_value.argtypes = [C.c_void_p, c_double_p, C.c_int]
            L.orc_tile_replay.restype = C.c_long
            L.orc_tile_replay.argtypes = [C.POINTER(CRegion), C.POINTER(CCfg), c_ubyte_p, c_double_p,
                                          C.c_int, C.c_int, C.c_double, c_long_p, C.c_long]
            L.orc_n_positions.restype = C.c_int
            L.orc_n_positions.argtypes = [C.POINTER(CRegion), C.POINTER(CCfg)]
            L.orc_condense.restype = None
            L.orc_condense.argtypes = [C.POINTER(CRegion), C.POINTER(CCfg), C.POINTER(CSel), c_double_p, c_long_p, C.c_long, c_long_p]
            L.orc_collapse.restype = None
            L.orc_collapse.argtypes = [C.POINTER(CRegion), C.POINTER(CCfg), C.POINTER(CSel), c_double_p, c_long_p, c_long_p]
            L.orc_reverse_comp.restype = None
            L.orc_reverse_comp.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
        else:
            L.ref_long_range_content.restype = None
            L.ref_long_range_content.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, c_double_p]

    # -- per candidate (explicit, strand-oriented strings) -------------------
    def _mip(self, ext, lig, tgt, ext_len, lig_len, scan_size, ext_copy, lig_copy) -> CMip:
        return CMip(ext, len(ext), lig, len(lig), tgt, len(tgt), ext_len, lig_len, scan_size, ext_copy, lig_copy)

    def get_score(self, ext: bytes, lig: bytes, tgt: bytes, ext_len=None, lig_len=None, scan_size=None,
                  ext_copy=1, lig_copy=1) -> float:
        m = self._mip(ext, lig, tgt, ext_len or len(ext), lig_len or len(lig), scan_size or len(tgt), ext_copy, lig_copy)
        return getattr(self.lib, self.p + "get_score")(C.byref(m))

    def get_parameters(self, ext: bytes, lig: bytes, tgt: bytes, lrc: np.ndarray, ext_len=None, lig_len=None,
                       scan_size=None, ext_copy=1, lig_copy=1) -> np.ndarray:
        m = self._mip(ext, lig, tgt, ext_len or len(ext), lig_len or len(lig), scan_size or len(tgt), ext_copy, lig_copy)
        lrc = np.ascontiguousarray(lrc, dtype=np.float64)
        out = np.empty(192, dtype=np.float64)
        getattr(self.lib, self.p + "get_parameters")(C.byref(m), _np_ptr(lrc, c_double_p), _np_ptr(out, c_double_p))
        return out

    # -- long-range content --------------------------------------------------
    def long_range_content(self, flank_seq: bytes, seq_start: int, seq_stop: int) -> np.ndarray:
        out = np.empty(44, dtype=np.float64)
        if self.p == "orc_":
            self.lib.orc_long_range_content(flank_seq, len(flank_seq), seq_stop - seq_start + 2001, _np_ptr(out, c_double_p))
        else:
            self.lib.ref_long_range_content(flank_seq, len(flank_seq), seq_start, seq_stop, _np_ptr(out, c_double_p))
        return out

    # -- svm -----------------------------------------------------------------
    def svm_load_model(self, path: str):
        h = getattr(self.lib, self.p + "svm_load_model")(path.encode())
        if not h:
            raise RuntimeError("cannot load model " + path)
        return C.c_void_p(h)

    def svm_free(self, h) -> None:
        getattr(self.lib, self.p + "svm_free")(h)

    def svm_predict(self, h, x: np.ndarray) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        return getattr(self.lib, self.p + "svm_predict")(h, _np_ptr(x, c_double_p), x.size)

    def svm_predict_rows(self, h, X: np.ndarray) -> np.ndarray:
        X = np.ascontiguousarray(X, dtype=np.float64)
        f = getattr(self.lib, self.p + "svm_predict")
        return np.array([f(h, X[i].ctypes.data_as(c_double_p), X.shape[1]) for i in range(X.shape[0])])

    # -- region grid ---------------------------------------------------------
    def grid_region(self, r: Region, cfg: Config, model=None, want_logistic=True, want_svr=False, want_feats=False):
        keep = _Keep()
        cr, cc = c_region(r, keep), c_cfg(cfg, keep)
        n = cfg.grid_size(r)
        valid = np.zeros(n, dtype=np.uint8)
        logi = np.empty(n, dtype=np.float64) if want_logistic else None
        svr = np.empty(n, dtype=np.float64) if want_svr else None
        feats = np.empty((n, 192), dtype=np.float64) if want_feats else None
        getattr(self.lib, self.p + "grid_region")(C.byref(cr), C.byref(cc), model, _np_ptr(valid, c_ubyte_p),
                                                  _np_ptr(logi, c_double_p), _np_ptr(svr, c_double_p),
                                                  _np_ptr(feats, c_double_p))
        return valid, logi, svr, feats


class Oracle(_Lib):
    def __init__(self):
        build_oracle()
        super().__init__(ORACLE_SO, "orc_")

    def reverse_comp(self, s: bytes) -> bytes:
        out = C.create_string_buffer(len(s) + 1)
        self.lib.orc_reverse_comp(s, len(s), out)
        return out.value

    def predict_value(self, h, x: np.ndarray) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self.lib.orc_predict_value(h, _np_ptr(x, c_double_p), x.size)

    def tile_replay(self, r: Region, cfg: Config, valid: np.ndarray, score: np.ndarray, method: int,
                    heuristic: bool, upper: float) -> np.ndarray:
        keep = _Keep()
        cr, cc = c_region(r, keep), c_cfg(cfg, keep)
        out = np.empty(valid.size, dtype=np.int64)
        score = np.ascontiguousarray(score, dtype=np.float64)
        n = self.lib.orc_tile_replay(C.byref(cr), C.byref(cc), _np_ptr(valid, c_ubyte_p), _np_ptr(score, c_double_p),
                                     method, int(heuristic), upper, out.ctypes.data_as(c_long_p), out.size)
        return out[:n]


    def select(self, r: Region, cfg: Config, score: np.ndarray, enum_idx: np.ndarray, lower: float, upper: float,
               max_arm_copy: int = 75, target_arm_copy: int = 20, masked_arm_threshold: float = 0.5):
        """condense_mips + collapse_mips: (scan_best[n_scan,2], pos_best[n_pos,2]) as grid indices (-1: none).
        The region's masked_seq / snp / unmappable (if any) feed arm_fraction_masked / snp_count / mapping_failed."""
        keep = _Keep()
        cr, cc = c_region(r, keep), c_cfg(cfg, keep)
        snp = np.ascontiguousarray(r.snp, np.uint8) if getattr(r, "snp", None) is not None else None
        unm = np.ascontiguousarray(r.unmappable, np.uint8) if getattr(r, "unmappable", None) is not None else None
        keep.refs += [snp, unm]
        sel = CSel(lower, upper, max_arm_copy, target_arm_copy, masked_arm_threshold, getattr(r, "masked_seq", None),
                   _np_ptr(snp, c_ubyte_p), _np_ptr(unm, c_ubyte_p))
        score = np.ascontiguousarray(score, dtype=np.float64)
        enum_idx = np.ascontiguousarray(enum_idx, dtype=np.int64)
        n_scan = cfg.n_scan(r)
        n_pos = self.lib.orc_n_positions(C.byref(cr), C.byref(cc))
        sb = np.empty((n_scan, 2), dtype=np.int64)
        pb = np.empty((n_pos, 2), dtype=np.int64)
        self.lib.orc_condense(C.byref(cr), C.byref(cc), C.byref(sel), _np_ptr(score, c_double_p),
                              enum_idx.ctypes.data_as(c_long_p), enum_idx.size, sb.ctypes.data_as(c_long_p))
        self.lib.orc_collapse(C.byref(cr), C.byref(cc), C.byref(sel), _np_ptr(score, c_double_p), sb.ctypes.data_as(c_long_p),
                              pb.ctypes.data_as(c_long_p))
        return sb, pb


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class Ref(_Lib):
    """The compiled reference.  Only available where oracle/_ref has been built."""

    def __init__(self):
        build_oracle()
        if not have_ref():
            raise RuntimeError("oracle/_ref not built (needs /root/reference)")
        super().__init__(REF_SO, "ref_")
