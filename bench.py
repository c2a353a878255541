#!/usr/bin/env python3
"""bench.py -- candidate MIPs scored per second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path (enumeration + 192 features + RBF-SVR score) over one
synthetic exon panel per GPU: 60 regions of U[80,400] bp, capture 162, the reference's 57
default arm pairs (~2.4e6 candidates), SVR model of 2048 support vectors
(BASELINE.json configs[2]; configs[1] = the same panel under logistic is reported beside it).

  value      whole-job candidates/s, panel resident in HBM, timed with CUDA events on the
             library's stream, max over ranks
  e2e        the same through the host-buffer C-ABI call (mg_score_regions): H2D of the
             region sequences and D2H of the validity + score grids inside the timed region
  roofline   K-svr (dominant kernel): 2*192*N_sv flop per candidate over the kernel's own
             CUDA-event time, against the FP64 DMMA peak measured on this pool
  cpu_baseline / --impl reference
             the unmodified reference objects (oracle/_ref) -- get_parameters + svm_predict --
             on all host cores (one process per core), bounded sample of the same panel

Multi-GPU: regions shard by rank (each rank owns one panel), no collective on the data
path; torch.distributed (NCCL) only for the barrier and the max-over-ranks of the timings.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mipgen_b200 import panel  # noqa: E402

N_REGIONS = 60
LEN_LO, LEN_HI = 80, 400
N_SV = 2048
GENOME_SEED = 20240
MODEL_SEED = 777
FLOP_PER_CAND_PER_SV = 2 * 192


def fp64_peak_tflops():
    """FP64 tensor (DMMA) peak measured on this pool (tools/microbench_fp64.cu); MEASURED_PEAKS.json
    records only bf16.  Falls back to the nominal B200 figure, saying so."""
    p = os.path.join(ROOT, "profiles", "fp64_peaks.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["dmma_tflops"]), "measured (profiles/fp64_peaks.json: DMMA.8x8x4 microbenchmark on this pool)"
    return 37.0, "fallback (nominal B200 FP64 tensor 37 TFLOP/s; no measured file)"


def make_panel(cfg: panel.Config, n_regions: int, seed: int, layout_seed: int = None):
    """Random-sequence genome from `seed`; region lengths/positions from `layout_seed` (default seed+1).
    Under torchrun every rank uses the same layout (equal work per GPU: weak scaling) on its own genome."""
    glen = panel.genome_length_for(n_regions, LEN_HI, cfg)
    genome = panel.lcg_genome(glen, seed)
    return genome, panel.make_regions(genome, n_regions, LEN_LO, LEN_HI, cfg, seed + 1 if layout_seed is None else layout_seed)


# ----------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------
# CPU baseline: the compiled reference objects, one process per core
# ----------------------------------------------------------------------------------
def _cpu_worker(args):
    so, prefix, model_path, cfg, regions, budget_s = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as oa
    lib = oa._Lib(so, prefix)
    h = lib.svm_load_model(model_path)
    done, t0 = 0, time.perf_counter()
    for r in regions:
        v, _l, _s, _f = lib.grid_region(r, cfg, h, want_logistic=False, want_svr=True)
        done += int(v.sum())
        if time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0


def cpu_baseline(cfg, model_path, lrc_by_region, genome, budget_s=12.0, cores=None):
    """get_parameters + svm_predict of the reference (oracle/_ref) over bounded slices of the
    bench panel, all host cores.  Returns dict for the JSON line."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as oa
    oa.build_oracle()
    if oa.have_ref():
        so, prefix, kind = oa.REF_SO, "ref_", "reference"
    else:
        so, prefix, kind = oa.ORACLE_SO, "orc_", "port"
    cores = cores or os.cpu_count() or 1
    # tiny regions (3 bp targets -> 120 scan starts x 114 = ~13.7k candidates, ~10-20 s per core at 2048 SV)
    rng = np.random.default_rng(99)
    jobs = []
    for c in range(cores):
        regs = []
        for k in range(4):
            start = 3000 + int(rng.integers(0, len(genome) - 8000))
            r = panel.cut_region(genome, start, start + 2, cfg, 0, "cpu%d_%d" % (c, k))
            r.lrc = lrc_by_region
            regs.append(r)
        jobs.append((so, prefix, model_path, cfg, regs, budget_s))
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    per_core = [r[0] / r[1] for r in res if r[1] > 0]
    return {"value": float(sum(per_core)), "unit": "candidates/s", "cores": cores, "kind": kind,
            "sample": "a sample of EQUIVALENT candidates, not the bench panel itself: %d candidates from 3-bp targets cut out of the bench "
                      "genome (capture 162, 57 arm pairs, both strands: the same per-candidate work) through the reference's get_parameters + "
                      "svm_predict (direct call, without mipgen.cpp's %%.17g text round trip -- favourable to the reference) with the bench's "
                      "%d-SV model, one forked process per core, %.1f s wall; per-core mean %.1f cand/s"
                      % (total, N_SV, wall, float(np.mean(per_core)) if per_core else 0.0)}


# ----------------------------------------------------------------------------------
# FP64-pipe slots per exp epilogue element, counted as FMA = 2 flop each (DMMA and DFMA share one pipe on B200):
# dense kernel 19 instructions (norm combine, clamp, scale, 11-instruction exp, alpha FMA); factored kernel 10
# (the 10-instruction exp: the contraction itself ends on the exponent, its tables are pre-scaled by -gamma)
EXP_FLOP_DENSE, EXP_FLOP_FACT = 38.0, 20.0


def issued_fp64_flop(tm) -> float:
    """FP64 work the K-svr launches really issued: 512 flop per DMMA.8x8x4, the exp epilogue, 2 FMA per gathered triple."""
    exp_flop = EXP_FLOP_FACT if tm.svr_gather > 0 else EXP_FLOP_DENSE
    return tm.svr_dmma * 512.0 + tm.svr_exp * exp_flop + tm.svr_gather * 4.0


_JSON_OUT = None


def reserve_stdout():
    """Keep the process's stdout for the one JSON line: everything else that writes to fd 1 from here on (NCCL's
    version banner, library chatter) lands on stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def build_model(ctx, cfg, work: str, n_model_regions: int = 4):
    """2048 SVs = feature rows of random candidates from a different genome seed, alpha ~ U(-1,1),
    calibrated so scores straddle 1.5 / 2.2; written and re-read as a libsvm text model."""
    import mipgen_b200 as mg
    rng = np.random.default_rng(MODEL_SEED)
    genome, regs = make_panel(cfg, n_model_regions, MODEL_SEED)
    for r in regs:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    _o, valid, _l, _s, feats = ctx.score_regions(regs, mg.MG_WANT_FEATURES)
    F = feats[valid.astype(bool)]
    sv = F[rng.choice(F.shape[0], N_SV, replace=False)]
    alpha = rng.uniform(-1, 1, N_SV)
    gamma = 1.0 / 192
    raw_path = os.path.join(work, "raw.model")
    panel.write_svr_model(raw_path, sv, alpha, gamma, 0.0)
    ctx.load_svr_model(raw_path)
    sample = F[rng.choice(F.shape[0], 20000, replace=False)]
    raw = ctx.svr_predict(sample)
    alpha2, rho = panel.calibrate(alpha, raw)
    path = os.path.join(work, "mipgen_svr.model")
    # keep the %.8g-rounded SVs the raw file holds
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import read_model_dense
    sv_r, _a, _g = read_model_dense(raw_path)
    panel.write_svr_model(path, sv_r, alpha2, gamma, rho)
    ctx.load_svr_model(path)
    return path


def library_build_id() -> str:
    """sha256 (first 12 hex digits) of the kernel sources the loaded library was built from."""
    import hashlib
    h = hashlib.sha256()
    src = os.path.join(ROOT, "mipgen_b200", "csrc")
    for f in sorted(os.listdir(src)):
        h.update(open(os.path.join(src, f), "rb").read())
    return h.hexdigest()[:12]


def svr_roofline(tm, n_sv, peak, peak_src, steps):
    """The K-svr record, every figure named by the formula it comes from (SURVEY.md 8d and VERDICT r01 weak #2):
      frac_algorithmic      2*192*N_sv flop per candidate / kernel time / FP64 peak.  The factored kernel evaluates the same
                            decision function with several times less arithmetic than the dense contraction, so this is a
                            work-reduction factor and exceeds 1 -- not a utilisation;
      frac_fp64_pipe_issued FP64 flop the launches really executed (device counters: 512 per DMMA.8x8x4, 20 per exp
                            element, 4 per gathered triple; tasks that exit after the claim phase add nothing) / time / peak;
      frac_dmma             the DMMA share of that alone = tensor-pipe utilisation of the FP64 tensor path."""
    ms = max(tm.ms_svr, 1e-9)
    launches = max(tm.launches_svr, 1)
    cand_per_launch = tm.candidates_svr / launches
    factored = tm.svr_gather > 0
    issued = issued_fp64_flop(tm)
    alg = tm.candidates_svr * FLOP_PER_CAND_PER_SV * n_sv
    achieved = issued / (ms / 1e3) / 1e12
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "svr_traffic.json")
    if os.path.exists(tp):
        d = json.load(open(tp))
        if d.get("kernel") == ("k_svr_fact" if factored else "k_svr_dmma"):
            traffic = d.get("dram_bytes_per_launch")
            traffic_src = "profiles/svr_traffic.json: ncu dram__bytes_read.sum + dram__bytes_write.sum of one %s launch, library build %s (this run: build %s)" % (
                d.get("kernel"), d.get("library_build", "r01"), library_build_id())
    return {"kernel": "k_svr_fact" if factored else "k_svr_dmma", "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "frac_fp64_pipe_issued": achieved / peak,
            "frac_dmma": tm.svr_dmma * 512.0 / (ms / 1e3) / 1e12 / peak,
            "frac_algorithmic": alg / (ms / 1e3) / 1e12 / peak,
            "what": "`frac` = frac_fp64_pipe_issued: FP64 flop executed by the K-svr launches (device-side counters of the tasks that ran: "
                    "512 per DMMA.8x8x4, %d per exp element, 4 per gathered triple) over their CUDA-event time, against the measured FP64 pipe "
                    "peak (DMMA and DFMA share it).  frac_dmma = the DMMA part alone (FP64 tensor-pipe utilisation).  frac_algorithmic = "
                    "SURVEY 8(d)'s 2*192*N_sv flop per candidate over the same time: a work-reduction factor (> 1 for the factored form), "
                    "not a utilisation" % (EXP_FLOP_FACT if factored else EXP_FLOP_DENSE),
            "issued_flop_per_step": issued / steps, "dmma_share_of_issued": tm.svr_dmma * 512.0 / max(issued, 1.0),
            "algorithmic_flop_per_candidate": FLOP_PER_CAND_PER_SV * n_sv,
            "candidates_per_launch": cand_per_launch, "ms_per_launch": ms / launches}


def timed_passes(ctx, pnl, want, steps, warmup, barrier):
    import mipgen_b200 as mg  # noqa: F401
    for _ in range(warmup):
        pnl.score(want)
    barrier()
    ctx.reset_timings()
    ctx.timer_start()
    for _ in range(steps):
        pnl.score(want)
    ms = ctx.timer_stop()
    barrier()
    return ms, ctx.timings()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-target", action="store_true", help="skip the 1 Mb north-star target block")
    ap.add_argument("--no-extras", action="store_true", help="skip the 8192-SV / wide arm table / cfg5 blocks")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "target1mb"],
                    help="cfg3 (default): 60-region exon panel per GPU, capture 162, with the target / extra blocks attached.  "
                         "target1mb: only the north star's target run, regions sharded over the torchrun ranks")
    args = ap.parse_args()
    reserve_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = panel.Config()  # capture 162/162, 57 default arm pairs
    work = tempfile.mkdtemp(prefix="mipgen_bench_")
    if args.workload == "target1mb":
        return target_run(args, rank, local_rank, world, work)
    config = {"workload": "cfg3: %d-region synthetic exon panel per GPU (region length U[%d,%d]), capture 162, 57 arm pairs, "
                          "-score_method svr, %d-SV synthetic RBF model" % (N_REGIONS, LEN_LO, LEN_HI, N_SV),
              "regions_per_gpu": N_REGIONS, "n_sv": N_SV, "sharding": "regions by rank, no collective",
              "cache": "the arm / insert row tables K-feat hands to K-svr (0.44 GB per pass for the 2.53M-candidate panel) exceed L2; the 3 MB SV matrix is L2-resident by design"}

    if args.impl == "reference":
        return reference_arm(args, rank, world, cfg, work, config)

    # keep stdout for the one JSON line: NCCL's version banner / debug lines go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import mipgen_b200 as mg
    cpu_group = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # host-side waiting for the phases one rank runs alone (an NCCL barrier would spin ON the other GPUs meanwhile)
        cpu_group = dist.new_group(backend="gloo")
    ctx = mg.Context(local_rank)
    ctx.set_config(cfg)
    model_path = build_model(ctx, cfg, work)
    genome, regions = make_panel(cfg, N_REGIONS, GENOME_SEED + rank, layout_seed=GENOME_SEED + 1)
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    pnl = ctx.panel(regions)
    n_cand = pnl.n_candidates
    row_bytes = pnl.row_table_bytes()
    sampler = ClockSampler(local_rank)

    # ---- SVR, inputs resident (FP64 factored kernel: the default scoring path) ----
    for _ in range(args.warmup):
        pnl.score(mg.MG_WANT_SVR)
    barrier()
    ctx.reset_timings()
    if rank == 0:
        sampler.start()
    ctx.timer_start()
    for _ in range(args.steps):
        pnl.score(mg.MG_WANT_SVR)
    ms_total = ctx.timer_stop()
    barrier()
    tm = ctx.timings()
    n_valid = pnl.valid_candidates()
    ms_total = reduce_max(ms_total)
    total_cand = reduce_sum(float(n_cand))
    total_valid = reduce_sum(float(n_valid))
    ms_per_step = ms_total / args.steps
    value = total_cand / (ms_per_step / 1e3)
    launches = int(tm.launches_feat + tm.launches_svr + tm.launches_other)
    _v, _l, fp64_scores, _ = pnl.fetch(valid=True, svr=True)
    fp64_sel = pnl.select(regions, 1, 1.5, 2.2)

    # ---- selection front-end on the device (condense_mips + collapse_mips over the SVR grid) ----
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        pnl.select(regions, 1, 1.5, 2.2)
    ms_select = reduce_max(ctx.timer_stop()) / args.steps
    barrier()

    # ---- logistic, inputs resident ----
    ms_log, tm_log = timed_passes(ctx, pnl, mg.MG_WANT_LOGISTIC, args.steps, args.warmup, barrier)
    ms_log = reduce_max(ms_log) / args.steps

    # ---- end to end through the host-buffer C-ABI (pinned host buffers) ----
    valid_h = torch.empty(n_cand, dtype=torch.uint8).pin_memory()
    svr_h = torch.empty(n_cand, dtype=torch.float64).pin_memory()
    out = (valid_h.numpy(), None, svr_h.numpy(), None)
    h2d = sum(len(r.seq) + 44 * 8 for r in regions)
    d2h = n_cand * 9
    prep = ctx.prepare_regions(regions)   # the mg_region structs a C caller holds; the timed call is mg_score_regions itself
    for _ in range(args.warmup):
        ctx.score_regions_prepared(prep, mg.MG_WANT_SVR, out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.score_regions_prepared(prep, mg.MG_WANT_SVR, out)
    ctx.sync()
    e2e_ms = reduce_max((time.perf_counter() - t0) * 1e3) / args.steps
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    checksum = float(np.nansum(svr_h.numpy()))

    # ---- the tensor-core form (tcgen05 split-FP16 contraction, FP64 epilogue) on the same panel ----
    tc = None
    if ctx.svr_tensor_core_available():
        ctx.set_svr_mode(3)
        ms_tc, tm_tc = timed_passes(ctx, pnl, mg.MG_WANT_SVR, max(3, args.steps // 2), 2, barrier)
        ms_tc = reduce_max(ms_tc) / max(3, args.steps // 2)
        valid_tc, _l, tc_scores, _ = pnl.fetch(valid=True, svr=True)
        tc_sel = pnl.select(regions, 1, 1.5, 2.2)
        ctx.set_svr_mode(0)
        ok = valid_tc.astype(bool)
        err = np.abs(tc_scores[ok] - fp64_scores[ok]) / np.abs(fp64_scores[ok])
        tc = (ms_tc, tm_tc, float(err.max()), float(np.median(err)), int((tc_sel[1] != fp64_sel[1]).sum()), int((tc_sel[3] != fp64_sel[3]).sum()),
              int(((tc_scores[ok] > 2.2) != (fp64_scores[ok] > 2.2)).sum() + ((tc_scores[ok] > 1.5) != (fp64_scores[ok] > 1.5)).sum()))
    barrier()

    # ---- the dense DMMA contraction on the same panel (the north star's formulation), for reference ----
    dense = None
    if ctx.svr_factored_available():
        ctx.set_svr_mode(1)
        dense_ms, td = timed_passes(ctx, pnl, mg.MG_WANT_SVR, 3, 2, barrier)
        dense_ms = reduce_max(dense_ms) / 3
        ctx.set_svr_mode(0)
        dense = (dense_ms, td)
    barrier()
    pnl.close()

    def host_barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    extras = None
    if not args.no_extras and rank == 0:
        extras = extra_blocks(ctx, cfg, work, args)
    host_barrier()
    target = None
    if not args.no_target:
        target = target_block(args, rank, world, work, barrier)
    host_barrier()

    if rank == 0:
        peak, peak_src = fp64_peak_tflops()
        line = {
            "metric": "candidate MIPs scored/sec (SVR)", "value": value, "unit": "candidates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "candidates_per_step": total_cand, "statically_valid_candidates_per_step": total_valid,
            "logistic": {"value": total_cand / (ms_log / 1e3), "unit": "candidates/s", "ms_per_step": ms_log,
                         "k_feat_hbm": {"what": "logistic-only K-feat writes 9 B per candidate (validity + score): ALU bound, not HBM bound",
                                        "ms_per_launch": tm_log.ms_feat / max(tm_log.launches_feat, 1)}},
            "e2e": {"value": total_cand / (e2e_ms / 1e3), "unit": "candidates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "api": "mg_score_regions (host sequences in, pinned host outputs; the mg_region structs are built once, as a C caller holds them)"},
            "select": {"what": "condense_mips + collapse_mips on the device (mg_panel_select: best MIP per scan start and per "
                               "position, incl. D2H of the winners)", "ms_per_step": ms_select},
            "gpu_launches": launches,
            "kernel_ms_per_step": {"k_feat": tm.ms_feat / args.steps, "k_svr": tm.ms_svr / args.steps, "other": tm.ms_other / args.steps},
            "roofline": svr_roofline(tm, N_SV, peak, peak_src, args.steps),
            "k_feat_roofline": {"kernel": "k_feat_window (row-table mode)", "bound": "hbm", "unit": "GB/s",
                                "achieved": (row_bytes + 2.0 * n_cand) * args.steps / (tm.ms_feat / 1e3) / 1e9, "peak": hbm_peak()[0],
                                "frac": (row_bytes + 2.0 * n_cand) * args.steps / (tm.ms_feat / 1e3) / 1e9 / hbm_peak()[0], "peak_source": hbm_peak()[1],
                                "row_table_bytes_per_step": row_bytes, "bytes_per_candidate": (row_bytes + 2.0 * n_cand) / n_cand,
                                "what": "SVR mode materialises no 192-vector: K-feat writes the distinct arm / insert rows of every factored-SVR work "
                                        "item (mg_panel_row_table_bytes) + 2 B of validity / state per candidate, K-svr bulk-copies the rows once. "
                                        "With ~175 B instead of 1,536 B per candidate the kernel is bound by its prefix-table build and "
                                        "integer geometry, not by HBM"},
            "clocks": clocks, "checksum": checksum, "library_build": library_build_id(),
        }
        if tc is not None:
            ms_tc, tm_tc, e_max, e_med, d_scan, d_pos, flips = tc
            mma = tm_tc.svr_tc_mma
            line["tensor_core_kernel"] = {
                "kernel": "k_svr_tc (tcgen05.mma kind::f16, FP32 TMEM accumulators, FP64 exponent/exp/row sum)", "value": total_cand / (ms_tc / 1e3),
                "unit": "candidates/s", "ms_per_step": ms_tc, "k_svr_ms_per_step": tm_tc.ms_svr / max(tm_tc.launches_svr, 1),
                "max_rel_dev_from_fp64_kernel": e_max, "median_rel_dev_from_fp64_kernel": e_med,
                "threshold_flips_at_1.5_and_2.2": flips, "scan_start_winners_changed": d_scan, "position_winners_changed": d_pos,
                "tensor_pipe": {"achieved": mma * 2.0 * TC_MMA_MACS / (tm_tc.ms_svr / 1e3) / 1e12, "unit": "TFLOP/s (FP16 in, FP32 out)",
                                "what": "tcgen05.mma instructions x 2*128*64*16 flop over the kernel's time; the kernel is bound by its FP64 "
                                        "epilogue (one exp per candidate x support vector), not by the tensor pipe"},
                "note": "opt-in (mg_set_svr_mode 3): ~1e-9 relative to libsvm instead of ~1e-13; `value` above stays the FP64 kernel"}
        if dense is not None:
            dense_ms, td = dense
            d_ach = td.svr_dmma * 512.0 / (td.ms_svr / 1e3) / 1e12
            line["dense_kernel"] = {"kernel": "k_svr_dmma", "value": total_cand / (dense_ms / 1e3), "unit": "candidates/s",
                                    "ms_per_step": dense_ms, "roofline": {"bound": "tensor", "achieved": d_ach, "peak": peak,
                                                                          "unit": "TFLOP/s", "frac": d_ach / peak, "frac_dmma": d_ach / peak,
                                                                          "frac_algorithmic": d_ach / peak},
                                    "note": "same panel through the dense candidates x SV contraction (mg_set_svr_mode(1)): here issued == "
                                            "algorithmic flop, so this is SURVEY 8(d)'s tensor-pipe fraction of the FP64 path"}
        if extras:
            line.update(extras)
        if target:
            line["target"] = target
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(cfg, model_path, regions[0].lrc, genome)
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


TC_MMA_MACS = 128 * 64 * 16


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def extra_blocks(ctx, cfg, work, args):
    """Rank 0, one GPU: lines the round-1 verdict asked for -- an 8192-SV model, an arm-pair table wider than the default
    (where the factored kernel's window halves), and BASELINE configs[4]'s shape at reduced scale streamed through
    bounded device memory (mg_tile_regions)."""
    import mipgen_b200 as mg
    out = {}
    peak, peak_src = fp64_peak_tflops()
    _g, regions = make_panel(cfg, 16, GENOME_SEED + 99)
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)

    def run(n_steps=3):
        pnl = ctx.panel(regions)
        ms, tm = timed_passes(ctx, pnl, mg.MG_WANT_SVR, n_steps, 1, ctx.sync)
        n = pnl.n_candidates
        pnl.close()
        return n, ms / n_steps, tm

    # 8192 support vectors (SURVEY 8d lists N in {2048, 8192})
    global N_SV
    keep = N_SV
    N_SV = 8192
    try:
        build_model(ctx, cfg, os.path.join(work), n_model_regions=6)
        n, ms, tm = run()
        out["sv8192"] = {"what": "16-region panel, 8192-SV model, factored FP64 kernel", "candidates": n, "ms_per_step": ms,
                         "value": n / (ms / 1e3), "unit": "candidates/s", "roofline": svr_roofline(tm, 8192, peak, peak_src, 3)}
    finally:
        N_SV = keep
    build_model(ctx, cfg, work)
    # a wider arm table: arm sums 38..47 (the factored kernel's window shrinks from 8 scan starts to 4 or 2)
    e, l = panel.default_arm_pairs(tuple(range(38, 48)))
    wide = panel.Config(162, 162, 5, 30, e, l)
    ctx.set_config(wide)
    regions_w = [panel.cut_region(_g, r.start_flanked, r.stop_flanked, wide, 0, r.label) for r in regions]
    for r, r0 in zip(regions_w, regions):
        r.lrc = r0.lrc
    pnl = ctx.panel(regions_w)
    ms, tm = timed_passes(ctx, pnl, mg.MG_WANT_SVR, 3, 1, ctx.sync)
    out["wide_arm_table"] = {"what": "arm sums 38..47 => %d arm pairs (default 57); factored kernel window W = %d scan starts (default 8)"
                                     % (len(e), ctx.svr_factored_available()),
                             "candidates": pnl.n_candidates, "ms_per_step": ms / 3, "value": pnl.n_candidates / (ms / 3 / 1e3), "unit": "candidates/s",
                             "roofline": svr_roofline(tm, N_SV, peak, peak_src, 3)}
    pnl.close()
    ctx.set_config(cfg)
    # cfg5 shape at reduced scale: 20,000 regions of U[100,200] bp, capture 162, SVR, streamed in sub-batches (bounded device memory):
    # score + condense + collapse, only the winners come back
    n5 = 20000
    g5 = panel.lcg_genome(panel.genome_length_for(n5, 200, cfg, gap=300), GENOME_SEED + 5)
    r5 = panel.make_regions(g5, n5, 100, 200, cfg, GENOME_SEED + 6, gap=300)
    lrc0 = regions[0].lrc
    for r in r5:
        r.lrc = lrc0   # one long-range vector for all: the per-region K-lrc calls are not what this block measures
    sel = dict(method=1, lower=1.5, upper=2.2)
    mg.tile_regions(ctx, r5[:200], mg.MG_WANT_SVR, select=sel)
    passes = []
    for _ in range(2):   # the first pass also grows the context's workspaces to the size of a full sub-batch
        t0 = time.perf_counter()
        t = mg.tile_regions(ctx, r5, mg.MG_WANT_SVR, select=sel)
        passes.append((t.call_seconds, time.perf_counter() - t0))
    dt, dt_py = passes[-1]   # the C call itself; with the Python wrapper's ctypes marshalling of the region structs
    n_grid = int(t.grid_off[-1])
    out["cfg5_sample"] = {"what": "BASELINE configs[4] shape at 1/10 scale: %d regions of U[100,200] bp (%.1f Mb of targets), capture 162, SVR, through "
                                  "mg_tile_regions in sub-batches of <= 2^26 grid points (bounded device memory; host buffers in, winners out)"
                                  % (n5, sum(r.stop_flanked - r.start_flanked + 1 for r in r5) / 1e6),
                          "grid_points": n_grid, "seconds": dt, "seconds_incl_python_marshalling": dt_py, "seconds_first_pass": passes[0][0], "value": n_grid / dt,
                          "unit": "candidates/s (end to end, one GPU)",
                          "scan_start_winners": int((t.scan_best >= 0).sum()),
                          "extrapolation": "the full config (2e5 regions, ~6e9 candidates) is 10x this work: ~%.0f s on one GPU, ~%.0f s on 8 "
                                           "(regions shard without exchange)" % (dt * 10, dt * 10 / 8),
                          "full_scale_runs": "tools/run_cfg5_full.py, measured: 6.07e9 grid points in 36.9 s on one B200 and 4.52 s on eight "
                                             "(profiles/r02_cfg5_full_1gpu.json, r02_cfg5_full_8gpu.json)"}
    # SURVEY 8(f4), opt-in: exact-match arm copy counting against an index of the cfg5 genome (what find_copy gets from BWA's X0 tags)
    t0 = time.perf_counter()
    gen = ctx.genome([g5.decode()])
    t_index = time.perf_counter() - t0
    sub = r5[:4000]
    gen.count_arm_copies(sub[:50], cfg.oligo_sizes)
    t0 = time.perf_counter()
    tabs = gen.count_arm_copies(sub, cfg.oligo_sizes)
    t_query = time.perf_counter() - t0
    n_oligos = int(sum(t.size for t in tabs))
    out["copy_count"] = {"what": "opt-in replacement of `bwa aln/samse` on oligo_copy_count.fq (mipgen.cpp:558-596): exact occurrences on both strands of "
                                 "every arm-sized oligo (%d sizes) of %d regions in a %.1f Mb genome, through mg_genome_create / mg_count_arm_copies "
                                 "(host buffers in, copy tables out)" % (len(cfg.oligo_sizes), len(sub), len(g5) / 1e6),
                         "index_seconds": t_index, "indexed_positions": gen.info()[1], "oligos": n_oligos, "query_seconds": t_query,
                         "value": n_oligos / t_query, "unit": "oligos/s (end to end)",
                         "multi_copy_fraction": float(np.mean(np.concatenate([(t > 1).ravel() for t in tabs[:200]])))}
    gen.close()
    return out


def target_block(args, rank, world, work, barrier):
    """North-star target inside the default line: the FULL candidate grid of a ~1 Mb synthetic target panel (4000 regions of
    U[150,350] bp, capture 120..250 step 5 = 27 sizes x 57 arm pairs x 2 strands = 6156 grid points per scan start, 5.6e9 grid
    points) scored with SVR and reduced to the best MIP per scan start / position, through the LIBRARY's multi-GPU call
    (mg_tile_regions_multi: one host thread + context + stream per GPU, LPT partition of the regions, no collective).  Fixed total
    work over the N GPUs of the run => strong scaling.  Rank 0 drives all N devices; the other ranks wait at the barrier."""
    if rank != 0:
        return None
    import mipgen_b200 as mg
    ctxs = [mg.Context(d) for d in range(world)]
    mcfg = panel.Config(250, 120, 65)
    ctxs[0].set_config(mcfg)
    model = build_model(ctxs[0], mcfg, work, n_model_regions=2)
    cfg = panel.Config(250, 120, 5)
    for c in ctxs:
        c.set_config(cfg)
        c.load_svr_model(model)
    n_total = 4000
    genome = panel.lcg_genome(panel.genome_length_for(n_total, 350, cfg), GENOME_SEED)
    regions = panel.make_regions(genome, n_total, 150, 350, cfg, GENOME_SEED + 1)
    target_bp = sum(r.stop_flanked - r.start_flanked + 1 for r in regions)
    for r in regions:
        r.lrc = ctxs[0].long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    sel = dict(method=1, lower=1.5, upper=2.2)
    steps = max(1, min(args.steps, 3))
    mg.tile_regions(ctxs, regions[:8 * world], mg.MG_WANT_SVR, select=sel)   # warm-up: kernels loaded, pools sized
    mg.tile_regions(ctxs, regions[:32 * world], mg.MG_WANT_SVR, select=sel)
    samplers = [ClockSampler(d) for d in range(world)]
    for c in ctxs:
        c.reset_timings()
    for s_ in samplers:
        s_.start()
    t0 = time.perf_counter()
    res = None
    for _ in range(steps):
        res = mg.tile_regions(ctxs, regions, mg.MG_WANT_SVR, select=sel)
    dt = (time.perf_counter() - t0) / steps
    clocks = [s_.stop() for s_ in samplers]
    tms = [c.timings() for c in ctxs]
    n_grid = int(res.grid_off[-1])
    owner = mg.partition_regions(cfg, regions, world)
    loads = np.bincount(owner, weights=np.diff(res.grid_off), minlength=world)
    peak, peak_src = fp64_peak_tflops()
    dev_ms = [(t.ms_feat + t.ms_svr + t.ms_other) / steps for t in tms]
    worst = int(np.argmax(dev_ms))
    out = {"metric": "candidate MIPs scored/sec (SVR), 1 Mb target panel", "value": n_grid / dt, "unit": "candidates/s", "n_gpus": world,
           "scaling": "strong", "steps": steps, "warmup": "2 partial passes (8 and 32 regions per GPU)", "seconds_per_step": dt,
           "timing": "host wall clock around mg_tile_regions_multi (host buffers in, winners out: H2D / kernels / D2H of all GPUs inside); "
                     "kernel_ms_per_gpu are CUDA-event sums per device",
           "config": {"workload": "north-star target: full candidate grid of a ~1 Mb synthetic target panel (%d regions, %d target bp), capture "
                                  "120..250 step 5, 57 arm pairs, SVR, %d-SV model; score + condense + collapse; regions LPT-partitioned over %d GPU(s) "
                                  "inside the library (no collective)" % (n_total, target_bp, N_SV, world)},
           "grid_points_per_step": n_grid, "scan_start_winners": int((res.scan_best >= 0).sum()), "position_winners": int((res.pos_best >= 0).sum()),
           "kernel_ms_per_gpu": dev_ms, "grid_points_per_gpu": [int(x) for x in loads],
           "roofline": svr_roofline(tms[worst], N_SV, peak, peak_src, steps),
           "clocks": clocks[worst], "clocks_all_gpus": clocks}
    for c in ctxs:
        c.close()
    return out


def reference_arm(args, rank, world, cfg, work, config):
    """The reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as oa
    from helpers import read_model_dense  # noqa: F401
    oa.build_oracle()
    lib = oa.Ref() if oa.have_ref() else oa.Oracle()
    # same recipe as the GPU arm, with the checker standing in for the feature extraction
    rng = np.random.default_rng(MODEL_SEED)
    genome, regs = make_panel(cfg, 1, MODEL_SEED)
    r0 = regs[0]
    r0.lrc = lib.long_range_content(r0.flank_seq, r0.seq_start, r0.seq_stop)
    _v, _l, _s, feats = lib.grid_region(r0, cfg, None, want_logistic=False, want_feats=True)
    F = feats[np.isfinite(feats[:, 0])]
    sv = F[rng.choice(F.shape[0], N_SV, replace=False)]
    path = os.path.join(work, "mipgen_svr.model")
    panel.write_svr_model(path, sv, rng.uniform(-1, 1, N_SV) * 0.05, 1.0 / 192, -1.8)
    bench_genome, _regs = make_panel(cfg, N_REGIONS, GENOME_SEED)
    vals, walls = [], []
    last = None
    for _ in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        last = cpu_baseline(cfg, path, r0.lrc, bench_genome, budget_s=6.0)
        walls.append((time.perf_counter() - t0) * 1e3)
        vals.append(last["value"])
    v = float(np.mean(vals[args.warmup:])) if len(vals) > args.warmup else float(np.mean(vals))
    ms_step = float(np.mean(walls[args.warmup:])) if len(walls) > args.warmup else float(np.mean(walls))
    last["value"] = v
    line = {"impl": "reference", "metric": "candidate MIPs scored/sec (SVR)", "value": v, "unit": "candidates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


if __name__ == "__main__":
    main()
