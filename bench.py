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
            "sample": "%d candidates (3-bp targets of the bench genome, capture 162, 57 arm pairs, both strands) through "
                      "get_parameters + svm_predict with the bench's %d-SV model, one forked process per core, %.1f s wall; "
                      "per-core mean %.1f cand/s" % (total, N_SV, wall, float(np.mean(per_core)) if per_core else 0.0)}


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "target1mb"],
                    help="cfg3 (default): 60-region exon panel per GPU, capture 162.  target1mb: the north star's target run -- "
                         "4000 regions (~1 Mb), capture sweep 120..250 step 5 (27 sizes), SVR, regions sharded over the ranks")
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
              "cache": "feature rows in flight (3.9 GB for the 2.53M-candidate panel; chunks of <= 4M candidates) exceed L2; the 3 MB SV matrix is L2-resident by design"}

    if args.impl == "reference":
        return reference_arm(args, rank, world, cfg, work, config)

    # keep stdout for the one JSON line: NCCL's version banner / debug lines go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import mipgen_b200 as mg
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
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
    sampler = ClockSampler(local_rank)

    # ---- SVR, inputs resident ----
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

    # ---- selection front-end on the device (condense_mips + collapse_mips over the SVR grid) ----
    pnl.select(regions, 1, 1.5, 2.2)
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        pnl.select(regions, 1, 1.5, 2.2)
    ms_select = reduce_max(ctx.timer_stop()) / args.steps
    barrier()

    # ---- logistic, inputs resident ----
    for _ in range(args.warmup):
        pnl.score(mg.MG_WANT_LOGISTIC)
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        pnl.score(mg.MG_WANT_LOGISTIC)
    ms_log = reduce_max(ctx.timer_stop()) / args.steps
    barrier()

    # ---- end to end through the host-buffer C-ABI (pinned host buffers) ----
    valid_h = torch.empty(n_cand, dtype=torch.uint8).pin_memory()
    svr_h = torch.empty(n_cand, dtype=torch.float64).pin_memory()
    out = (valid_h.numpy(), None, svr_h.numpy(), None)
    h2d = sum(len(r.seq) + 44 * 8 for r in regions)
    d2h = n_cand * 9
    for _ in range(args.warmup):
        ctx.score_regions(regions, mg.MG_WANT_SVR, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.score_regions(regions, mg.MG_WANT_SVR, out=out)
    ctx.sync()
    e2e_ms = reduce_max((time.perf_counter() - t0) * 1e3) / args.steps
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    checksum = float(np.nansum(svr_h.numpy()))

    # ---- the dense DMMA contraction on the same panel (the north star's formulation), for reference ----
    dense = None
    if ctx.svr_factored_available():
        ctx.set_svr_mode(1)
        for _ in range(2):
            pnl.score(mg.MG_WANT_SVR)
        barrier()
        ctx.reset_timings()
        ctx.timer_start()
        for _ in range(3):
            pnl.score(mg.MG_WANT_SVR)
        dense_ms = reduce_max(ctx.timer_stop()) / 3
        td = ctx.timings()
        ctx.set_svr_mode(0)
        dense = (dense_ms, td)
    barrier()

    if rank == 0:
        peak, peak_src = fp64_peak_tflops()
        svr_ms_per_launch = tm.ms_svr / max(tm.launches_svr, 1)
        cand_per_launch = tm.candidates_svr / max(tm.launches_svr, 1)
        dense_equiv = cand_per_launch * FLOP_PER_CAND_PER_SV * N_SV / (svr_ms_per_launch / 1e3) / 1e12
        factored = tm.svr_gather > 0
        issued_flop = issued_fp64_flop(tm)
        achieved = issued_flop / (tm.ms_svr / 1e3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "svr_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        line = {
            "metric": "candidate MIPs scored/sec (SVR)", "value": value, "unit": "candidates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "candidates_per_step": total_cand, "statically_valid_candidates_per_step": total_valid,
            "logistic": {"value": total_cand / (ms_log / 1e3), "unit": "candidates/s", "ms_per_step": ms_log},
            "e2e": {"value": total_cand / (e2e_ms / 1e3), "unit": "candidates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "api": "mg_score_regions (host buffers, pinned outputs)"},
            "select": {"what": "condense_mips + collapse_mips on the device (mg_panel_select: best MIP per scan start and per "
                               "position, incl. D2H of the winners)", "ms_per_step": ms_select},
            "gpu_launches": launches,
            "kernel_ms_per_step": {"k_feat": tm.ms_feat / args.steps, "k_svr": tm.ms_svr / args.steps, "other": tm.ms_other / args.steps},
            "roofline": {"kernel": "k_svr_fact" if factored else "k_svr_dmma", "bound": "tensor", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "what": "FP64 pipe (DMMA.8x8x4 + the exp/gather DFMAs share it): flop actually issued by the K-svr launches "
                                 "(512 per DMMA, %d per exp element, 4 per gathered triple) over their CUDA-event time"
                                 % (EXP_FLOP_FACT if factored else EXP_FLOP_DENSE),
                         "issued_flop_per_step": issued_flop / args.steps, "dmma_share_of_issued": tm.svr_dmma * 512.0 / issued_flop,
                         "algorithmic_flop_per_candidate": FLOP_PER_CAND_PER_SV * N_SV,
                         "dense_equivalent_tflops": dense_equiv,
                         "note": ("the factored kernel evaluates the same decision function with ~%.1fx less FP64 work than the dense "
                                  "2*192*N_sv contraction, so the dense-equivalent rate exceeds the pipe's peak"
                                  % (cand_per_launch * FLOP_PER_CAND_PER_SV * N_SV * tm.launches_svr / issued_flop)) if factored else None,
                         "candidates_per_launch": cand_per_launch, "ms_per_launch": svr_ms_per_launch},
            "clocks": clocks, "checksum": checksum,
        }
        if dense is not None:
            dense_ms, td = dense
            d_ach = td.svr_dmma * 512.0 / (td.ms_svr / 1e3) / 1e12
            line["dense_kernel"] = {"kernel": "k_svr_dmma", "value": total_cand / (dense_ms / 1e3), "unit": "candidates/s",
                                    "ms_per_step": dense_ms, "roofline": {"bound": "tensor", "achieved": d_ach, "peak": peak,
                                                                          "unit": "TFLOP/s", "frac": d_ach / peak},
                                    "note": "same panel through the dense candidates x SV contraction (mg_set_svr_mode(1))"}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(cfg, model_path, regions[0].lrc, genome)
        emit(line)
    pnl.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def target_run(args, rank, local_rank, world, work):
    """North-star target: the FULL candidate grid of a ~1 Mb synthetic target panel (4000 regions of
    U[150,350] bp, capture 120..250 step 5 = 27 sizes x 57 arm pairs x 2 strands = 6156 grid points per scan
    start, ~5.6e9 grid points) scored with SVR.  Fixed total work, regions LPT-sharded over the ranks
    (strong scaling), no collective on the data path."""
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import mipgen_b200 as mg
    from mipgen_b200 import shard
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = mg.Context(local_rank)
    # model: SVs from a trimmed capture sweep so they span the same scan sizes
    mcfg = panel.Config(250, 120, 65)
    ctx.set_config(mcfg)
    build_model(ctx, mcfg, work, n_model_regions=2)
    cfg = panel.Config(250, 120, 5)
    ctx.set_config(cfg)
    n_total = 4000
    glen = panel.genome_length_for(n_total, 350, cfg)
    genome = panel.lcg_genome(glen, GENOME_SEED)
    regions = panel.make_regions(genome, n_total, 150, 350, cfg, GENOME_SEED + 1)
    target_bp = sum(r.stop_flanked - r.start_flanked + 1 for r in regions)
    costs = [cfg.grid_size(r) for r in regions]
    mine = [regions[i] for i in shard.lpt_assign(costs, world)[rank]]
    for r in mine:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    pnl = ctx.panel(mine)
    n_cand = pnl.n_candidates

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    for _ in range(warm):
        pnl.score(mg.MG_WANT_SVR)
    barrier()
    ctx.reset_timings()
    ctx.timer_start()
    for _ in range(steps):
        pnl.score(mg.MG_WANT_SVR)
    ms = ctx.timer_stop()
    barrier()
    tm = ctx.timings()
    n_valid = pnl.valid_candidates()
    ms = reduce(ms, dist.ReduceOp.MAX if world > 1 else None) / steps
    total = reduce(float(n_cand), dist.ReduceOp.SUM if world > 1 else None)
    total_valid = reduce(float(n_valid), dist.ReduceOp.SUM if world > 1 else None)
    if rank == 0:
        peak, peak_src = fp64_peak_tflops()
        issued = issued_fp64_flop(tm)
        emit(({
            "metric": "candidate MIPs scored/sec (SVR)", "value": total / (ms / 1e3), "unit": "candidates/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "north-star target: full candidate grid of a ~1 Mb synthetic target panel (%d regions, %d target bp), "
                                   "capture 120..250 step 5, 57 arm pairs, SVR, %d-SV model; regions LPT-sharded over %d rank(s)"
                                   % (n_total, target_bp, N_SV, world)},
            "grid_points_per_step": total, "statically_valid_per_step": total_valid,
            "valid_candidates_per_s": total_valid / (ms / 1e3),
            "kernel_ms_rank0": {"k_feat": tm.ms_feat / steps, "k_svr": tm.ms_svr / steps},
            "roofline": {"kernel": "k_svr_fact", "bound": "tensor", "achieved": issued / (tm.ms_svr / 1e3) / 1e12, "peak": peak,
                         "unit": "TFLOP/s", "frac": issued / (tm.ms_svr / 1e3) / 1e12 / peak, "peak_source": peak_src, "traffic": None},
        }))
    pnl.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


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
