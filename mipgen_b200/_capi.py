"""ctypes binding of include/mipgen_b200.h -- the same C-ABI a cgo/JNI/C++ caller binds.

Loading fails loudly when libmipgen_b200.so is missing; creating a Context fails
loudly when there is no CUDA device.  There is no CPU fallback anywhere.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import List, Optional, Sequence

import numpy as np

from .panel import Config, Region

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmipgen_b200.so")

MG_WANT_LOGISTIC, MG_WANT_SVR, MG_WANT_FEATURES = 1, 2, 4
MG_NFEAT, MG_NLRC = 192, 44

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)
c_int64_p = C.POINTER(C.c_int64)


class MgConfig(C.Structure):
    _fields_ = [("max_capture", C.c_int), ("min_capture", C.c_int), ("capture_increment", C.c_int),
                ("max_mip_overlap", C.c_int), ("n_pairs", C.c_int), ("ext_len", c_int_p),
                ("lig_len", c_int_p), ("n_oligo_sizes", C.c_int), ("oligo_sizes", c_int_p)]


class MgRegion(C.Structure):
    _fields_ = [("seq", C.c_char_p), ("seq_len", C.c_int), ("seq_start", C.c_int), ("seq_stop", C.c_int),
                ("start_flanked", C.c_int), ("stop_flanked", C.c_int), ("lrc", c_double_p),
                ("copies", c_int_p), ("scan_begin", C.c_int), ("scan_end", C.c_int),
                ("masked_seq", C.c_char_p), ("snp", c_ubyte_p), ("unmappable", c_ubyte_p)]


class MgCandidate(C.Structure):
    _fields_ = [("ext", C.c_char_p), ("ext_n", C.c_int), ("lig", C.c_char_p), ("lig_n", C.c_int),
                ("tgt", C.c_char_p), ("tgt_n", C.c_int), ("ext_len", C.c_int), ("lig_len", C.c_int),
                ("scan_size", C.c_int), ("ext_copy", C.c_int), ("lig_copy", C.c_int)]


class MgSelectParams(C.Structure):
    _fields_ = [("method", C.c_int), ("heuristic", C.c_int), ("lower_score_limit", C.c_double),
                ("upper_score_limit", C.c_double), ("max_arm_copy", C.c_int), ("target_arm_copy", C.c_int),
                ("masked_arm_threshold", C.c_double)]


class MgRecordMeta(C.Structure):
    _fields_ = [("chr", C.c_char_p), ("label", C.c_char_p), ("feature_start", C.c_int), ("feature_stop", C.c_int)]


MG_TEXT_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)   # mg_text_sink


class MgTileResult(C.Structure):
    _fields_ = [("scan_best", c_int64_p), ("pos_best", c_int64_p), ("scan_best_logistic", c_double_p),
                ("scan_best_svr", c_double_p), ("valid", c_ubyte_p), ("logistic", c_double_p), ("svr", c_double_p)]


class MgMipInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("strand", "ext_len", "lig_len", "scan_start", "scan_stop", "ext_start", "ext_stop",
                                        "lig_start", "lig_stop", "ext_copy", "lig_copy")]


class MgTimings(C.Structure):
    _fields_ = [("ms_feat", C.c_double), ("launches_feat", C.c_long), ("ms_svr", C.c_double),
                ("launches_svr", C.c_long), ("ms_other", C.c_double), ("launches_other", C.c_long),
                ("candidates_feat", C.c_long), ("candidates_svr", C.c_long),
                ("svr_dmma", C.c_double), ("svr_exp", C.c_double), ("svr_gather", C.c_double), ("svr_tc_mma", C.c_double)]


# every symbol include/mipgen_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("mg_version", C.c_char_p, []),
    ("mg_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("mg_destroy", None, [C.c_void_p]),
    ("mg_last_error", C.c_char_p, [C.c_void_p]),
    ("mg_set_config", C.c_int, [C.c_void_p, C.POINTER(MgConfig)]),
    ("mg_sync", C.c_int, [C.c_void_p]),
    ("mg_timer_start", C.c_int, [C.c_void_p]),
    ("mg_timer_stop", C.c_int, [C.c_void_p, c_double_p]),
    ("mg_reset_timings", C.c_int, [C.c_void_p]),
    ("mg_get_timings", C.c_int, [C.c_void_p, C.POINTER(MgTimings)]),
    ("mg_load_svr_model", C.c_int, [C.c_void_p, C.c_char_p]),
    ("mg_set_svr_model", C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int, C.c_int, C.c_double, C.c_double]),
    ("mg_set_svr_mode", C.c_int, [C.c_void_p, C.c_int]),
    ("mg_svr_factored_available", C.c_int, [C.c_void_p]),
    ("mg_svr_tensor_core_available", C.c_int, [C.c_void_p]),
    ("mg_model_info", C.c_int, [C.c_void_p, c_int_p, c_double_p, c_double_p]),
    ("mg_svr_predict", C.c_int, [C.c_void_p, c_double_p, C.c_long, C.c_long, c_double_p]),
    ("mg_svr_predict_direct", C.c_int, [C.c_void_p, c_double_p, C.c_long, C.c_long, c_double_p]),
    ("mg_long_range_content", C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, c_double_p]),
    ("mg_score_candidates", C.c_int, [C.c_void_p, C.POINTER(MgCandidate), C.c_long, c_double_p, C.c_int,
                                      c_double_p, c_double_p, c_double_p]),
    ("mg_grid_size", C.c_int64, [C.c_void_p, C.POINTER(MgRegion)]),
    ("mg_first_scan_start", C.c_int, [C.c_void_p, C.POINTER(MgRegion)]),
    ("mg_config_grid_size", C.c_int64, [C.POINTER(MgConfig), C.POINTER(MgRegion)]),
    ("mg_config_first_scan_start", C.c_int, [C.POINTER(MgConfig), C.POINTER(MgRegion)]),
    ("mg_score_regions", C.c_int, [C.c_void_p, C.POINTER(MgRegion), C.c_int, C.c_int, c_int64_p, c_ubyte_p,
                                   c_double_p, c_double_p, c_double_p]),
    ("mg_panel_create", C.c_int, [C.c_void_p, C.POINTER(MgRegion), C.c_int, C.POINTER(C.c_void_p)]),
    ("mg_panel_destroy", None, [C.c_void_p]),
    ("mg_panel_candidates", C.c_int64, [C.c_void_p]),
    ("mg_panel_row_table_bytes", C.c_int64, [C.c_void_p]),
    ("mg_panel_valid_candidates", C.c_int64, [C.c_void_p]),
    ("mg_panel_score", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("mg_panel_fetch", C.c_int, [C.c_void_p, C.c_void_p, c_ubyte_p, c_double_p, c_double_p, c_double_p]),
    ("mg_panel_device_ptrs", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    ("mg_region_scan_count", C.c_int, [C.c_void_p, C.POINTER(MgRegion)]),
    ("mg_region_position_count", C.c_int, [C.c_void_p, C.POINTER(MgRegion)]),
    ("mg_panel_select", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(MgSelectParams), c_int64_p, c_int64_p]),
    ("mg_tile_replay", C.c_int64, [C.POINTER(MgConfig), C.POINTER(MgRegion), c_ubyte_p, c_double_p, C.c_int, C.c_int,
                                   C.c_double, c_int64_p, C.c_int64]),
    ("mg_describe_candidates", C.c_int, [C.POINTER(MgConfig), C.POINTER(MgRegion), c_int64_p, C.c_int, C.POINTER(MgMipInfo)]),
    ("mg_format_mip_records", C.c_int64, [C.POINTER(MgConfig), C.POINTER(MgRegion), c_int64_p, C.c_int, c_double_p, C.c_char_p,
                                          C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int64]),
    ("mg_format_mip_record", C.c_int64, [C.POINTER(MgRegion), C.POINTER(MgMipInfo), C.c_double, C.c_char_p, C.c_char_p, C.c_int,
                                         C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int64]),
    ("mg_tile_sizes", C.c_int, [C.POINTER(MgConfig), C.POINTER(MgRegion), C.c_int, c_int64_p, c_int64_p, c_int64_p]),
    ("mg_tile_regions", C.c_int, [C.c_void_p, C.POINTER(MgRegion), C.c_int, C.c_int, C.POINTER(MgSelectParams), C.c_int64,
                                  C.POINTER(MgTileResult)]),
    ("mg_tile_regions_multi", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(MgRegion), C.c_int, C.c_int,
                                        C.POINTER(MgSelectParams), C.c_int64, C.POINTER(MgTileResult)]),
    ("mg_score_regions_multi", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(MgRegion), C.c_int, C.c_int, c_int64_p,
                                         c_ubyte_p, c_double_p, c_double_p]),
    ("mg_partition_regions", C.c_int, [C.POINTER(MgConfig), C.POINTER(MgRegion), C.c_int, C.c_int, c_int_p]),
    ("mg_panel_gather", C.c_int, [C.c_void_p, C.c_void_p, c_int64_p, C.c_int64, c_double_p, c_double_p]),
    ("mg_genome_create", C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), c_int64_p, C.c_int, C.POINTER(C.c_void_p)]),
    ("mg_genome_destroy", None, [C.c_void_p]),
    ("mg_genome_info", C.c_int, [C.c_void_p, c_int64_p, c_int64_p, c_int64_p]),
    ("mg_count_arm_copies", C.c_int, [C.c_void_p, C.POINTER(MgRegion), C.c_int, c_int_p, C.c_int, c_int_p, c_int64_p]),
    ("mg_panel_format_records", C.c_int64, [C.c_void_p, C.c_void_p, C.POINTER(MgRecordMeta), c_int64_p, C.c_int64, C.c_int, C.c_char_p,
                                            C.c_int, C.c_void_p, C.c_int64]),
    ("mg_panel_format_enumerated", C.c_int64, [C.c_void_p, C.c_void_p, C.POINTER(MgRecordMeta), C.POINTER(MgSelectParams), C.c_char_p, C.c_int,
                                               c_int64_p, MG_TEXT_SINK, C.c_void_p]),
    ("mg_tile_regions_records", C.c_int, [C.c_void_p, C.POINTER(MgRegion), C.c_int, C.c_int, C.POINTER(MgSelectParams), C.c_int64,
                                          C.POINTER(MgTileResult), C.POINTER(MgRecordMeta), C.c_char_p, C.c_int, c_int64_p, MG_TEXT_SINK,
                                          C.c_void_p]),
    ("mg_format_g", C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_void_p, c_int_p]),
    ("mg_format_capture_fastq", C.c_int64, [C.c_void_p, C.POINTER(MgRegion), C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_int64]),
    ("mg_format_oligo_fastq", C.c_int64, [C.c_void_p, C.POINTER(MgRegion), C.POINTER(C.c_char_p), C.c_int, c_int_p, C.c_int, C.c_void_p,
                                          C.c_int64]),
]

_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree library and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "mipgen_b200 has no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class MgError(RuntimeError):
    pass


def _ptr(a: Optional[np.ndarray], t):
    return a.ctypes.data_as(t) if a is not None else t()


def _c_config(cfg: Config):
    e = np.asarray(cfg.ext_len, np.int32)
    l = np.asarray(cfg.lig_len, np.int32)
    o = np.asarray(cfg.oligo_sizes, np.int32)
    c = MgConfig(cfg.max_capture, cfg.min_capture, cfg.capture_increment, cfg.max_mip_overlap, len(e),
                 _ptr(e, c_int_p), _ptr(l, c_int_p), len(o), _ptr(o, c_int_p))
    return c, (e, l, o)


def _c_regions(regions: Sequence[Region]):
    keep = []
    arr = (MgRegion * max(len(regions), 1))()
    for i, r in enumerate(regions):
        lrc = np.ascontiguousarray(r.lrc, np.float64) if r.lrc is not None else None
        cop = np.ascontiguousarray(r.copies, np.int32) if r.copies is not None else None
        snp = np.ascontiguousarray(r.snp, np.uint8) if getattr(r, "snp", None) is not None else None
        unm = np.ascontiguousarray(r.unmappable, np.uint8) if getattr(r, "unmappable", None) is not None else None
        msk = getattr(r, "masked_seq", None)
        if msk is not None and len(msk) != len(r.seq):
            raise ValueError("masked_seq must be as long as seq")
        if snp is not None and snp.size != len(r.seq):
            raise ValueError("snp must have one entry per base of seq")
        keep += [lrc, cop, r.seq, snp, unm, msk]
        arr[i] = MgRegion(r.seq, len(r.seq), r.seq_start, r.seq_stop, r.start_flanked, r.stop_flanked,
                          _ptr(lrc, c_double_p), _ptr(cop, c_int_p), getattr(r, "scan_begin", 0), getattr(r, "scan_end", 0),
                          msk, _ptr(snp, c_ubyte_p), _ptr(unm, c_ubyte_p))
    return arr, keep


def tile_sizes(cfg: Config, regions: Sequence[Region]):
    """Prefix sums of grid points, scan starts and coverable positions over the regions (mg_tile_sizes; no device)."""
    c, _k = _c_config(cfg)
    arr, _k2 = _c_regions(regions)
    n = len(regions)
    g, s, p = (np.zeros(n + 1, np.int64) for _ in range(3))
    if load_library().mg_tile_sizes(C.byref(c), arr, n, _ptr(g, c_int64_p), _ptr(s, c_int64_p), _ptr(p, c_int64_p)) != 0:
        raise MgError("mg_tile_sizes: bad config")
    return g, s, p


def partition_regions(cfg: Config, regions: Sequence[Region], n_parts: int) -> np.ndarray:
    """owner[i] = which of n_parts contexts scores region i (mg_partition_regions: LPT on grid sizes; no device)."""
    c, _k = _c_config(cfg)
    arr, _k2 = _c_regions(regions)
    owner = np.zeros(max(1, len(regions)), np.int32)
    if load_library().mg_partition_regions(C.byref(c), arr, len(regions), n_parts, _ptr(owner, c_int_p)) != 0:
        raise MgError("mg_partition_regions: bad arguments")
    return owner[:len(regions)]


class TileResult:
    """Outputs of mg_tile_regions[_multi] (numpy views over the caller-owned buffers)."""

    def __init__(self, grid_off, scan_off, pos_off, scan_best, pos_best, scan_best_logistic, scan_best_svr, valid, logistic, svr):
        self.grid_off, self.scan_off, self.pos_off = grid_off, scan_off, pos_off
        self.scan_best, self.pos_best = scan_best, pos_best
        self.scan_best_logistic, self.scan_best_svr = scan_best_logistic, scan_best_svr
        self.valid, self.logistic, self.svr = valid, logistic, svr


def tile_regions(ctxs, regions: Sequence[Region], want: int, select=None, full_grids: bool = False,
                 max_batch_candidates: int = 0) -> TileResult:
    """mg_tile_regions (one Context) / mg_tile_regions_multi (a list of Contexts on different GPUs).
    select: None or dict(method, lower, upper, heuristic=True, max_arm_copy=75, target_arm_copy=20, masked_arm_threshold=0.5)."""
    multi = isinstance(ctxs, (list, tuple))
    first = ctxs[0] if multi else ctxs
    g, s, p = tile_sizes(first.cfg, regions)
    arr, _keep = _c_regions(regions)
    ns, npos, ng = int(s[-1]), int(p[-1]), int(g[-1])
    sb = np.full((ns, 2), -1, np.int64) if select else None
    pb = np.full((npos, 2), -1, np.int64) if select else None
    sbl = np.full((ns, 2), np.nan) if select and want & MG_WANT_LOGISTIC else None
    sbs = np.full((ns, 2), np.nan) if select and want & MG_WANT_SVR else None
    v = np.empty(ng, np.uint8) if full_grids else None
    lo = np.empty(ng, np.float64) if full_grids and want & MG_WANT_LOGISTIC else None
    sv = np.empty(ng, np.float64) if full_grids and want & MG_WANT_SVR else None
    res = MgTileResult(_ptr(sb, c_int64_p), _ptr(pb, c_int64_p), _ptr(sbl, c_double_p), _ptr(sbs, c_double_p),
                       _ptr(v, c_ubyte_p), _ptr(lo, c_double_p), _ptr(sv, c_double_p))
    sp = None
    if select:
        sp = MgSelectParams(select["method"], int(select.get("heuristic", True)), select["lower"], select["upper"],
                            select.get("max_arm_copy", 75), select.get("target_arm_copy", 20), select.get("masked_arm_threshold", 0.5))
    spp = C.byref(sp) if sp is not None else None
    lib = first.lib
    t0 = time.perf_counter()
    if multi:
        hs = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        rc = lib.mg_tile_regions_multi(hs, len(ctxs), arr, len(regions), want, spp, max_batch_candidates, C.byref(res))
        if rc != 0:
            raise MgError("mg_tile_regions_multi failed (%d): %s" % (rc, "; ".join(lib.mg_last_error(c.h).decode() for c in ctxs)))
    else:
        first._check(lib.mg_tile_regions(first.h, arr, len(regions), want, spp, max_batch_candidates, C.byref(res)))
    out = TileResult(g, s, p, sb, pb, sbl, sbs, v, lo, sv)
    out.call_seconds = time.perf_counter() - t0   # the C call alone (the ctypes marshalling of the regions above is the wrapper's)
    return out


def config_grid_size(cfg: Config, r: Region) -> int:
    """Host arithmetic only (no device): size of a region's candidate grid."""
    c, _k = _c_config(cfg)
    arr, _k2 = _c_regions([r])
    return int(load_library().mg_config_grid_size(C.byref(c), arr))


def tile_replay(cfg: Config, r: Region, valid: np.ndarray, score: np.ndarray, method: int, heuristic: bool,
                upper: float) -> np.ndarray:
    """Host replay of the tile loop's score-dependent skips (mg_tile_replay; no device needed)."""
    c, _k = _c_config(cfg)
    arr, _k2 = _c_regions([r])
    out = np.empty(valid.size, np.int64)
    valid = np.ascontiguousarray(valid, np.uint8)
    score = np.ascontiguousarray(score, np.float64)
    n = load_library().mg_tile_replay(C.byref(c), arr, _ptr(valid, c_ubyte_p), _ptr(score, c_double_p), method,
                                      int(heuristic), upper, _ptr(out, c_int64_p), out.size)
    if n < 0:
        raise MgError("mg_tile_replay: bad config")
    return out[:n]


UNIVERSAL_CONSTANT = "CTTCAGCTTCCCGATATCCGACGGTAGTGT"  # mipgen.cpp:199


def universal_middle(lig_tag_length: int = 0, ext_tag_length: int = 5) -> str:
    """mipgen.cpp:200 (defaults of -lig_tag_sizes / -ext_tag_sizes)."""
    return "N" * lig_tag_length + UNIVERSAL_CONSTANT + "N" * ext_tag_length


def describe_candidates(cfg: Config, r: Region, idx: np.ndarray):
    """Geometry and arm copy numbers of region r's candidates idx (mg_describe_candidates; no device needed).
    Returns a list of dicts with the fields of mg_mip_info."""
    c, _k = _c_config(cfg)
    arr, _k2 = _c_regions([r])
    idx = np.ascontiguousarray(idx, np.int64)
    info = (MgMipInfo * max(1, idx.size))()
    if load_library().mg_describe_candidates(C.byref(c), arr, _ptr(idx, c_int64_p), idx.size, info) != 0:
        raise MgError("mg_describe_candidates: bad config or grid index")
    return [{n: getattr(info[k], n) for n, _t in MgMipInfo._fields_} for k in range(idx.size)]


def design_records(cfg: Config, r: Region, idx: np.ndarray, score: np.ndarray, chrom: str, label: str, feature_start: int,
                   feature_stop: int, first_index: int, middle: Optional[str] = None, raw: bool = False):
    """The all_mips.txt / collapsed_mips.txt lines of region r's candidates idx (grid indices, in output order),
    numbered from first_index (mg_format_mip_records; no device needed).  Returns str, or with raw=True a
    memoryview over the C buffer."""
    lib = load_library()
    c, _k = _c_config(cfg)
    arr, _k2 = _c_regions([r])
    idx = np.ascontiguousarray(idx, np.int64)
    score = np.ascontiguousarray(score, np.float64)
    mid = (middle if middle is not None else universal_middle()).encode()
    cap = max(1, idx.size) * (512 + 3 * cfg.max_capture + len(mid) + len(label) + len(chrom))
    buf = C.create_string_buffer(cap)
    n = lib.mg_format_mip_records(C.byref(c), arr, _ptr(idx, c_int64_p), idx.size, _ptr(score, c_double_p), chrom.encode(),
                                  label.encode(), feature_start, feature_stop, mid, first_index, buf, cap)
    if n < 0:
        raise MgError("mg_format_mip_records: bad config, grid index or buffer")
    return memoryview(buf)[:n] if raw else buf.raw[:n].decode()  # raw: no copy, for writing large files


class Panel:
    """Regions resident in HBM (mg_panel)."""

    def __init__(self, ctx: "Context", handle, offsets: np.ndarray, keep):
        self.ctx, self.h, self.offsets, self._keep = ctx, handle, offsets, keep

    @property
    def n_candidates(self) -> int:
        return int(self.ctx.lib.mg_panel_candidates(self.h))

    def valid_candidates(self) -> int:
        return int(self.ctx.lib.mg_panel_valid_candidates(self.h))

    def row_table_bytes(self) -> int:
        """Bytes of arm / insert row tables K-feat hands to the factored SVR per scoring pass (mg_panel_row_table_bytes)."""
        return int(self.ctx.lib.mg_panel_row_table_bytes(self.h))

    def score(self, want: int) -> None:
        """Launch the kernels for the whole panel (asynchronous on the context's stream)."""
        self.ctx._check(self.ctx.lib.mg_panel_score(self.ctx.h, self.h, want))

    def fetch(self, valid=True, logistic=False, svr=False, features=False):
        n = self.n_candidates
        v = np.empty(n, np.uint8) if valid else None
        lo = np.empty(n, np.float64) if logistic else None
        sv = np.empty(n, np.float64) if svr else None
        ft = np.empty((n, MG_NFEAT), np.float64) if features else None
        self.ctx._check(self.ctx.lib.mg_panel_fetch(self.ctx.h, self.h, _ptr(v, c_ubyte_p), _ptr(lo, c_double_p),
                                                    _ptr(sv, c_double_p), _ptr(ft, c_double_p)))
        return v, lo, sv, ft

    def fetch_into(self, valid: Optional[np.ndarray], logistic: Optional[np.ndarray], svr: Optional[np.ndarray]) -> None:
        self.ctx._check(self.ctx.lib.mg_panel_fetch(self.ctx.h, self.h, _ptr(valid, c_ubyte_p), _ptr(logistic, c_double_p),
                                                    _ptr(svr, c_double_p), c_double_p()))

    def gather(self, idx: np.ndarray, logistic: bool = False, svr: bool = False):
        """Scores of the given panel-global grid indices (mg_panel_gather); -1 yields NaN."""
        idx = np.ascontiguousarray(idx, np.int64).reshape(-1)
        lo = np.empty(idx.size) if logistic else None
        sv = np.empty(idx.size) if svr else None
        self.ctx._check(self.ctx.lib.mg_panel_gather(self.ctx.h, self.h, _ptr(idx, c_int64_p), idx.size, _ptr(lo, c_double_p),
                                                     _ptr(sv, c_double_p)))
        return lo, sv

    def format_records(self, regions: Sequence[Region], idx: np.ndarray, which: int, chrom: str = "1", first_index: int = 1,
                       middle: Optional[str] = None, out: Optional[np.ndarray] = None):
        """all_mips.txt / collapsed_mips.txt lines of the given panel-global grid indices, written on the device
        (mg_panel_format_records).  Returns a uint8 array view of the bytes."""
        idx = np.ascontiguousarray(idx, np.int64).reshape(-1)
        meta = (MgRecordMeta * max(1, len(regions)))()
        keep = []
        for i, r in enumerate(regions):
            lab = r.label.encode()
            keep.append(lab)
            meta[i] = MgRecordMeta(chrom.encode(), lab, r.start_flanked, r.stop_flanked)
        mid = (middle if middle is not None else universal_middle()).encode()
        cap = max(1, idx.size) * (512 + 3 * self.ctx.cfg.max_capture + len(mid) + 2 * len(chrom) + 32)
        if out is None or out.size < cap:
            out = np.empty(cap, np.uint8)
        n = self.ctx.lib.mg_panel_format_records(self.ctx.h, self.h, meta, _ptr(idx, c_int64_p), idx.size, which, mid, first_index,
                                                 out.ctypes.data_as(C.c_void_p), out.size)
        if n < 0:
            raise MgError("mg_panel_format_records failed (%d): %s" % (n, self.ctx.lib.mg_last_error(self.ctx.h).decode()))
        return out[:n]

    def format_enumerated(self, regions: Sequence[Region], method: int, upper: float, heuristic: bool = True, chrom: str = "1",
                          first_index: int = 1, middle: Optional[str] = None):
        """all_mips.txt of the scored panel written on the device (mg_panel_format_enumerated): the records of every candidate the
        tile loop enumerates, in its order.  Returns (bytes, records_per_region)."""
        meta = (MgRecordMeta * max(1, len(regions)))()
        keep = []
        for i, r in enumerate(regions):
            lab = r.label.encode()
            keep.append(lab)
            meta[i] = MgRecordMeta(chrom.encode(), lab, r.start_flanked, r.stop_flanked)
        mid = (middle if middle is not None else universal_middle()).encode()
        sp = MgSelectParams(method, int(heuristic), 0.0, upper, 75, 20, 0.5)
        pieces = []

        def sink(_user, text, n):
            pieces.append(C.string_at(text, n))
            return 0

        per = np.zeros(max(1, len(regions)), np.int64)
        cb = MG_TEXT_SINK(sink)
        n = self.ctx.lib.mg_panel_format_enumerated(self.ctx.h, self.h, meta, C.byref(sp), mid, first_index, _ptr(per, c_int64_p), cb, None)
        if n < 0:
            raise MgError("mg_panel_format_enumerated failed (%d): %s" % (n, self.ctx.lib.mg_last_error(self.ctx.h).decode()))
        return b"".join(pieces), per[:len(regions)]

    def select(self, regions: Sequence[Region], method: int, lower: float, upper: float, heuristic: bool = True,
               max_arm_copy: int = 75, target_arm_copy: int = 20, masked_arm_threshold: float = 0.5):
        """condense_mips + collapse_mips on the device.  Returns (scan_offsets, scan_best[n,2], pos_offsets,
        pos_best[m,2]); entries are global grid indices of the panel or -1."""
        arr, _keep = _c_regions(regions)
        n = len(regions)
        so = np.zeros(n + 1, np.int64)
        po = np.zeros(n + 1, np.int64)
        for i in range(n):
            so[i + 1] = so[i] + self.ctx.lib.mg_region_scan_count(self.ctx.h, C.byref(arr[i]))
            po[i + 1] = po[i] + self.ctx.lib.mg_region_position_count(self.ctx.h, C.byref(arr[i]))
        sb = np.empty((int(so[-1]), 2), np.int64)
        pb = np.empty((int(po[-1]), 2), np.int64)
        sp = MgSelectParams(method, int(heuristic), lower, upper, max_arm_copy, target_arm_copy, masked_arm_threshold)
        self.ctx._check(self.ctx.lib.mg_panel_select(self.ctx.h, self.h, C.byref(sp), _ptr(sb, c_int64_p), _ptr(pb, c_int64_p)))
        return so, sb, po, pb

    def close(self) -> None:
        if self.h:
            self.ctx.lib.mg_panel_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One mg_ctx on one CUDA device."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.mg_create(device, C.byref(h))
        if rc != 0:
            raise MgError("mg_create(%d) failed (%d): %s" % (device, rc, self.lib.mg_last_error(None).decode()))
        self.h = h
        self.cfg: Optional[Config] = None
        self._cfg_keep = None

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.mg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise MgError("mipgen_b200 error %d: %s" % (rc, self.lib.mg_last_error(self.h).decode()))

    # -- configuration / model ----------------------------------------------
    def set_config(self, cfg: Config) -> None:
        c, _keep = _c_config(cfg)
        self._check(self.lib.mg_set_config(self.h, C.byref(c)))
        self.cfg = cfg

    def load_svr_model(self, path: str) -> None:
        self._check(self.lib.mg_load_svr_model(self.h, path.encode()))

    def set_svr_model(self, sv: np.ndarray, alpha: np.ndarray, gamma: float, rho: float) -> None:
        sv = np.ascontiguousarray(sv, np.float64)
        alpha = np.ascontiguousarray(alpha, np.float64)
        self._check(self.lib.mg_set_svr_model(self.h, _ptr(sv, c_double_p), _ptr(alpha, c_double_p), sv.shape[0],
                                              sv.shape[1], gamma, rho))

    def set_svr_mode(self, mode: int) -> None:
        """0 auto, 1 dense DMMA contraction, 2 factored, 3 tensor cores (tcgen05, split FP16)."""
        self._check(self.lib.mg_set_svr_mode(self.h, mode))

    def svr_factored_available(self) -> int:
        return int(self.lib.mg_svr_factored_available(self.h))

    def format_g(self, values: np.ndarray):
        """printf('%g') of each value on the device (mg_format_g); returns a list of str (None where the device declines)."""
        v = np.ascontiguousarray(values, np.float64).reshape(-1)
        out = np.zeros((v.size, 32), np.uint8)
        ln = np.zeros(v.size, np.int32)
        self._check(self.lib.mg_format_g(self.h, _ptr(v, c_double_p), v.size, out.ctypes.data_as(C.c_void_p), _ptr(ln, c_int_p)))
        return [bytes(out[i, :ln[i]]).decode() if ln[i] >= 0 else None for i in range(v.size)]

    def fastq(self, regions: Sequence[Region], chrom: str = "1", oligo: bool = False) -> bytes:
        """check_copy_numbers' all_sequences.fq (oligo=False) / oligo_copy_count.fq (oligo=True) for the regions, formatted on the
        device (mg_format_capture_fastq / mg_format_oligo_fastq)."""
        arr, _keep = self._regions(regions)
        names = (C.c_char_p * max(1, len(regions)))(*[chrom.encode()] * len(regions))
        cfg = self.cfg
        if oligo:
            sizes = np.asarray(cfg.oligo_sizes, np.int32)
            call = lambda b, cap: self.lib.mg_format_oligo_fastq(self.h, arr, names, len(regions), _ptr(sizes, c_int_p), sizes.size, b, cap)
        else:
            call = lambda b, cap: self.lib.mg_format_capture_fastq(self.h, arr, names, len(regions), cfg.max_capture, cfg.min_capture,
                                                                    cfg.capture_increment, b, cap)
        need = call(None, 0)
        if need < 0:
            raise MgError("fastq sizing failed (%d): %s" % (need, self.lib.mg_last_error(self.h).decode()))
        buf = np.empty(max(1, need), np.uint8)
        n = call(buf.ctypes.data_as(C.c_void_p), buf.size)
        if n != need:
            raise MgError("fastq formatting failed (%d): %s" % (n, self.lib.mg_last_error(self.h).decode()))
        return buf[:n].tobytes()

    def svr_tensor_core_available(self) -> bool:
        return bool(self.lib.mg_svr_tensor_core_available(self.h))

    def model_info(self):
        n, g, r = C.c_int(), C.c_double(), C.c_double()
        self._check(self.lib.mg_model_info(self.h, C.byref(n), C.byref(g), C.byref(r)))
        return n.value, g.value, r.value

    # -- timings ---------------------------------------------------------------
    def sync(self) -> None:
        self._check(self.lib.mg_sync(self.h))

    def timer_start(self) -> None:
        self._check(self.lib.mg_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self.lib.mg_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def reset_timings(self) -> None:
        self._check(self.lib.mg_reset_timings(self.h))

    def timings(self) -> MgTimings:
        t = MgTimings()
        self._check(self.lib.mg_get_timings(self.h, C.byref(t)))
        return t

    # -- scoring ---------------------------------------------------------------
    def svr_predict(self, X: np.ndarray, direct: bool = False) -> np.ndarray:
        X = np.ascontiguousarray(X, np.float64)
        out = np.empty(X.shape[0], np.float64)
        f = self.lib.mg_svr_predict_direct if direct else self.lib.mg_svr_predict
        self._check(f(self.h, _ptr(X, c_double_p), X.shape[0], X.shape[1], _ptr(out, c_double_p)))
        return out

    def long_range_content(self, flank_seq: bytes, seq_start: int, seq_stop: int) -> np.ndarray:
        out = np.empty(MG_NLRC, np.float64)
        self._check(self.lib.mg_long_range_content(self.h, flank_seq, len(flank_seq), seq_stop - seq_start + 2001,
                                                   _ptr(out, c_double_p)))
        return out

    def score_candidates(self, cands: Sequence[dict], lrc: Optional[np.ndarray] = None, want: int = MG_WANT_LOGISTIC):
        """cands: dicts with ext/lig/tgt bytes (strand-oriented) and optional ext_len, lig_len,
        scan_size, ext_copy, lig_copy -- the fields of an SVMipv4 object."""
        n = len(cands)
        arr = (MgCandidate * max(n, 1))()
        for i, c in enumerate(cands):
            arr[i] = MgCandidate(c["ext"], len(c["ext"]), c["lig"], len(c["lig"]), c["tgt"], len(c["tgt"]),
                                 c.get("ext_len", len(c["ext"])), c.get("lig_len", len(c["lig"])),
                                 c.get("scan_size", len(c["tgt"])), c.get("ext_copy", 1), c.get("lig_copy", 1))
        lrc_a = np.ascontiguousarray(lrc, np.float64) if lrc is not None else None
        lo = np.empty(n, np.float64) if want & MG_WANT_LOGISTIC else None
        sv = np.empty(n, np.float64) if want & MG_WANT_SVR else None
        ft = np.empty((n, MG_NFEAT), np.float64) if want & MG_WANT_FEATURES else None
        self._check(self.lib.mg_score_candidates(self.h, arr, n, _ptr(lrc_a, c_double_p), want, _ptr(lo, c_double_p),
                                                 _ptr(sv, c_double_p), _ptr(ft, c_double_p)))
        return lo, sv, ft

    def _regions(self, regions: Sequence[Region]):
        return _c_regions(regions)

    def grid_size(self, r: Region) -> int:
        arr, _keep = self._regions([r])
        return int(self.lib.mg_grid_size(self.h, arr))

    def score_regions_multi(self, others: Sequence["Context"], regions: Sequence[Region], want: int, out=None):
        """mg_score_regions_multi over this context and `others` (one per GPU).  Returns offsets, valid, logistic, svr."""
        ctxs = [self] + list(others)
        g, _s, _p = tile_sizes(self.cfg, regions)
        arr, _keep = self._regions(regions)
        total = int(g[-1])
        if out is not None:
            valid, lo, sv = out
        else:
            valid = np.empty(total, np.uint8)
            lo = np.empty(total, np.float64) if want & MG_WANT_LOGISTIC else None
            sv = np.empty(total, np.float64) if want & MG_WANT_SVR else None
        offsets = np.zeros(len(regions) + 1, np.int64)
        hs = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        rc = self.lib.mg_score_regions_multi(hs, len(ctxs), arr, len(regions), want, _ptr(offsets, c_int64_p), _ptr(valid, c_ubyte_p),
                                             _ptr(lo, c_double_p), _ptr(sv, c_double_p))
        if rc != 0:
            raise MgError("mg_score_regions_multi failed (%d): %s" % (rc, "; ".join(self.lib.mg_last_error(c.h).decode() for c in ctxs)))
        return offsets, valid, lo, sv

    def score_regions(self, regions: Sequence[Region], want: int = MG_WANT_LOGISTIC, out=None):
        """Host-buffer call (H2D + kernels + D2H inside).  Returns offsets, valid, logistic, svr, features."""
        arr, _keep = self._regions(regions)
        n = len(regions)
        offsets = np.zeros(n + 1, np.int64)
        for i in range(n):
            offsets[i + 1] = offsets[i] + self.lib.mg_grid_size(self.h, C.byref(arr[i]))
        total = int(offsets[-1])
        if out is not None:
            valid, lo, sv, ft = out
        else:
            valid = np.empty(total, np.uint8)
            lo = np.empty(total, np.float64) if want & MG_WANT_LOGISTIC else None
            sv = np.empty(total, np.float64) if want & MG_WANT_SVR else None
            ft = np.empty((total, MG_NFEAT), np.float64) if want & MG_WANT_FEATURES else None
        self._check(self.lib.mg_score_regions(self.h, arr, n, want, _ptr(offsets, c_int64_p), _ptr(valid, c_ubyte_p),
                                              _ptr(lo, c_double_p), _ptr(sv, c_double_p), _ptr(ft, c_double_p)))
        return offsets, valid, lo, sv, ft

    def genome(self, contigs: Sequence[str]) -> "Genome":
        """Exact-match index of the given contig sequences, resident on the device (mg_genome_create; SURVEY.md 8 f4, opt-in)."""
        return Genome(self, contigs)

    def prepare_regions(self, regions: Sequence[Region]):
        """The mg_region array of `regions` and their grid offsets, built once: what a C caller holds before it calls
        mg_score_regions (the ctypes marshalling costs ~12 us per region and is the wrapper's, not the library's)."""
        arr, keep = self._regions(regions)
        n = len(regions)
        offsets = np.zeros(n + 1, np.int64)
        for i in range(n):
            offsets[i + 1] = offsets[i] + self.lib.mg_grid_size(self.h, C.byref(arr[i]))
        return arr, keep, n, offsets

    def score_regions_prepared(self, prep, want: int, out):
        """mg_score_regions on a prepared mg_region array (H2D + kernels + D2H inside); out = (valid, logistic, svr, features) buffers."""
        arr, _keep, n, offsets = prep
        valid, lo, sv, ft = out
        self._check(self.lib.mg_score_regions(self.h, arr, n, want, _ptr(offsets, c_int64_p), _ptr(valid, c_ubyte_p),
                                              _ptr(lo, c_double_p), _ptr(sv, c_double_p), _ptr(ft, c_double_p)))

    def panel(self, regions: Sequence[Region]) -> Panel:
        arr, keep = self._regions(regions)
        n = len(regions)
        offsets = np.zeros(n + 1, np.int64)
        for i in range(n):
            offsets[i + 1] = offsets[i] + self.lib.mg_grid_size(self.h, C.byref(arr[i]))
        h = C.c_void_p()
        self._check(self.lib.mg_panel_create(self.h, arr, n, C.byref(h)))
        return Panel(self, h, offsets, keep)


class Genome:
    """mg_genome: sorted 32-mers of a reference genome in HBM; counts exact occurrences of arm-sized oligos on both strands."""

    def __init__(self, ctx: Context, contigs: Sequence[str]):
        self.ctx = ctx
        raw = [c.encode() if isinstance(c, str) else bytes(c) for c in contigs]
        n = len(raw)
        seqs = (C.c_char_p * max(n, 1))(*raw)
        lens = np.asarray([len(b) for b in raw], np.int64)
        h = C.c_void_p()
        ctx._check(ctx.lib.mg_genome_create(ctx.h, seqs, _ptr(lens, c_int64_p), n, C.byref(h)))
        self.h = h

    def info(self):
        """(positions, positions with 32 valid bases ahead, short suffixes)"""
        a, b, c = (np.zeros(1, np.int64) for _ in range(3))
        self.ctx.lib.mg_genome_info(self.h, _ptr(a, c_int64_p), _ptr(b, c_int64_p), _ptr(c, c_int64_p))
        return int(a[0]), int(b[0]), int(c[0])

    def count_arm_copies(self, regions: Sequence[Region], oligo_sizes: Sequence[int]):
        """One int32 table [n_oligo_sizes][seq_len] per region, in the layout of Region.copies (mg_count_arm_copies)."""
        arr, _keep = _c_regions(regions)
        n = len(regions)
        sizes = np.asarray(list(oligo_sizes), np.int32)
        off = np.zeros(n + 1, np.int64)
        self.ctx._check(self.ctx.lib.mg_count_arm_copies(self.h, arr, n, _ptr(sizes, c_int_p), sizes.size, c_int_p(), _ptr(off, c_int64_p)))
        flat = np.zeros(max(int(off[-1]), 1), np.int32)
        self.ctx._check(self.ctx.lib.mg_count_arm_copies(self.h, arr, n, _ptr(sizes, c_int_p), sizes.size, _ptr(flat, c_int_p), _ptr(off, c_int64_p)))
        return [flat[off[i]:off[i + 1]].reshape(sizes.size, len(regions[i].seq)) for i in range(n)]

    def close(self):
        if self.h:
            self.ctx.lib.mg_genome_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
