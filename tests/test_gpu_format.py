"""Design-file records written on the device (mg_panel_format_records, k_format.cu: print_details of mipgen.cpp:765-794) against
the host formatter, whose output tests/test_design_records.py pins byte for byte to the reference CLI's all_mips.txt /
collapsed_mips.txt; and the device's printf("%g") (exact 128-bit integer rounding) against the C library's."""
import os

import numpy as np
import pytest

import mipgen_b200 as mg
from mipgen_b200 import panel
from helpers import small_config, synthetic_regions, calibrated_model, mutate, tmpdir

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    from oracle_api import Oracle
    return Oracle()


def test_device_percent_g_equals_libc():
    rng = np.random.default_rng(5)
    vals = [0.0, -0.0, 1.0, -1000.0, 0.5, 123456.5, 1234565.0, 999999.5, 9999995.0, 0.0001, 0.00001, 0.000099999949, 1e5, 1e6, 1e-5,
            2.5e-7, 0.377881088375488, 0.9999995, 0.99999949999, 1.5, 2.2, 1e14, 9.99999e14, 1e-12, float("nan"), float("inf"), -float("inf")]
    vals += list(rng.uniform(0, 1, 200000))                      # logistic-like
    vals += list(rng.normal(1.8, 0.6, 200000))                   # SVR-like
    vals += list(10.0 ** rng.uniform(-12, 15, 100000) * rng.choice([-1, 1], 100000))
    # exact decimal ties and their neighbours: k + 0.5 scaled by powers of ten and two
    ties = (rng.integers(100000, 999999, 20000) + 0.5) * 10.0 ** rng.integers(-8, 6, 20000)
    vals += list(ties) + list(np.nextafter(ties, np.inf)) + list(np.nextafter(ties, -np.inf))
    v = np.array(vals, np.float64)
    ctx = mg.Context(0)
    got = ctx.format_g(v)
    bad = 0
    for x, g in zip(v, got):
        want = "%g" % x
        if g is None:
            assert not (1e-12 <= abs(x) < 1e15), x
            continue
        if g != want:
            bad += 1
            assert bad < 5, (x.hex(), g, want)
    assert bad == 0
    assert ctx.format_g(np.array([1e-13, 1e15, 5e-324])) == [None, None, None]
    ctx.close()


def test_device_records_equal_host_records(oracle):
    cfg = small_config((40, 43, 45), 162, 152, 5)
    rng = np.random.default_rng(3)
    genome, regions = synthetic_regions(oracle, cfg, 4, 60, 150, 4242)
    regions[1].seq = mutate(regions[1].seq, rng, 3, alphabet=b"RYKMacgt")   # IUPAC / lower case pass through reverse complementing
    regions[2].copies = rng.choice([0, 1, 1, 1, 2, 7, 100, 101], size=(len(cfg.oligo_sizes), len(regions[2].seq))).astype(np.int32)
    regions.append(panel.cut_region(genome, 150, 240, cfg, 0, "edge_region_with_a_long_label"))
    regions[-1].lrc = rng.uniform(0, 0.3, 44)
    d = tmpdir()
    _v, _l, _s, feats = oracle.grid_region(regions[0], cfg, None, want_logistic=False, want_feats=True)
    model = calibrated_model(oracle, cfg, 64, 5, os.path.join(d, "m.model"), feats[np.isfinite(feats[:, 0])][::53])
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    ctx.load_svr_model(model)
    pnl = ctx.panel(regions)
    pnl.score(mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    valid, lo, sv, _ = pnl.fetch(valid=True, logistic=True, svr=True)
    offs = pnl.offsets
    for which, score, method, upper, middle in ((0, lo, 0, 0.9, None), (1, sv, 1, 2.2, mg.universal_middle(4, 4))):
        idx_all, want = [], b""
        first = 17
        for i, r in enumerate(regions):
            a, b = offs[i], offs[i + 1]
            enum_idx = mg.tile_replay(cfg, r, valid[a:b], score[a:b], method, True, upper)
            want += bytes(mg.design_records(cfg, r, enum_idx, score[a:b], "chr7_alt", r.label, r.start_flanked, r.stop_flanked,
                                            first + len(idx_all), middle=middle, raw=True))
            idx_all += list(enum_idx + a)
        got = pnl.format_records(regions, np.array(idx_all), which, chrom="chr7_alt", first_index=first, middle=middle)
        assert len(idx_all) > 20000
        assert bytes(got) == want, "device-written records differ from print_details' bytes"
        # the whole all_mips.txt in one call: K-replay finds the enumerated grid points on the device (with and without the
        # logistic heuristic's early exits, mipgen.cpp:494)
        for heuristic in (True, False):
            if not heuristic:
                idx_all, want = [], b""
                for i, r in enumerate(regions):
                    a, b = offs[i], offs[i + 1]
                    enum_idx = mg.tile_replay(cfg, r, valid[a:b], score[a:b], method, False, upper)
                    want += bytes(mg.design_records(cfg, r, enum_idx, score[a:b], "chr7_alt", r.label, r.start_flanked, r.stop_flanked,
                                                    first + len(idx_all), middle=middle, raw=True))
                    idx_all += list(enum_idx + a)
            text, per_region = pnl.format_enumerated(regions, method, upper, heuristic=heuristic, chrom="chr7_alt", first_index=first, middle=middle)
            assert int(per_region.sum()) == len(idx_all)
            assert text == want, "device-enumerated all_mips.txt differs (method %d, heuristic %s)" % (method, heuristic)
    # regions with selection-only inputs are refused (their flags need design_mip)
    regions[0].snp = np.zeros(len(regions[0].seq), np.uint8)
    pnl2 = ctx.panel(regions)
    pnl2.score(mg.MG_WANT_LOGISTIC)
    with pytest.raises(mg.MgError):
        pnl2.format_records(regions, np.array([0, 1]), 0)
    pnl.close()
    pnl2.close()
    ctx.close()


def test_device_formatter_throughput_on_the_bench_panel():
    """Every statically valid candidate of a 12-region slice of the bench panel as an all_mips.txt record."""
    import time
    import bench
    cfg = panel.Config()
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    _g, regions = bench.make_panel(cfg, 12, bench.GENOME_SEED)
    pnl = ctx.panel(regions)
    pnl.score(mg.MG_WANT_LOGISTIC)
    valid, lo, _sv, _ = pnl.fetch(valid=True, logistic=True)
    idx = np.nonzero(valid)[0]
    out = np.empty(idx.size * 420, np.uint8)
    pnl.format_records(regions, idx[:1000], 0, out=out)
    ctx.reset_timings()
    t0 = time.perf_counter()
    got = pnl.format_records(regions, idx, 0, out=out)
    dt = time.perf_counter() - t0
    t = ctx.timings()
    a, b = pnl.offsets[0], pnl.offsets[1]
    first = idx[idx < b]
    want = bytes(mg.design_records(cfg, regions[0], first - a, lo[a:b], "1", regions[0].label, regions[0].start_flanked, regions[0].stop_flanked, 1, raw=True))
    assert bytes(got[:len(want)]) == want
    print("device formatter: %d records, %.1f MB in %.3f s end to end (%.2f GB/s incl. D2H into pageable memory); kernels %.2f ms (%.1f GB/s)"
          % (idx.size, got.size / 1e6, dt, got.size / dt / 1e9, t.ms_other, got.size / (t.ms_other / 1e3) / 1e9))
    assert got.size / (t.ms_other / 1e3) / 1e9 > 5.0
    pnl.close()
    ctx.close()


def test_fastq_files_for_bwa_equal_the_reference_loops():
    """check_copy_numbers' two FASTQ files (mipgen.cpp:804-838) formatted on the device against a literal Python restatement of the
    loops: capture sizes descending, MIP starts from start_flanked - capture while < stop_flanked (kept if > 0 and the read fits),
    then every oligo size over relative starts 0 .. len - size - 1.  Coordinates crossing 9,999 -> 10,000 change the record
    length inside a group; a region at the chromosome start exercises the start > 0 filter."""
    cfg = panel.Config(170, 150, 5)
    genome = panel.lcg_genome(40000, 77)
    regions = [panel.cut_region(genome, 9990, 10080, cfg, 0, "digits"), panel.cut_region(genome, 60, 140, cfg, 0, "edge"),
               panel.cut_region(genome, 20000, 20130, cfg, 0, "plain")]
    ctx = mg.Context(0)
    ctx.set_config(cfg)

    def want_capture():
        out = []
        for r in regions:
            for cap in cfg.captures:
                start = r.start_flanked - cap
                while start < r.stop_flanked:
                    if start > 0 and start + cap - 1 <= r.seq_stop:
                        out.append(b"@%d_chrA_%d\n%s\n+\n%s\n" % (cap, start, r.seq[start - r.seq_start:start - r.seq_start + cap], b"#" * cap))
                    start += 1
        return b"".join(out)

    def want_oligo():
        out = []
        for r in regions:
            for size in cfg.oligo_sizes:
                for rel in range(0, len(r.seq) - size):
                    a = r.seq_start + rel
                    out.append(b"@chrchrA:%d-%d\n%s\n+\n%s\n" % (a, a + size - 1, r.seq[rel:rel + size], b"#" * size))
        return b"".join(out)

    got_c, got_o = ctx.fastq(regions, "chrA"), ctx.fastq(regions, "chrA", oligo=True)
    assert got_c == want_capture() and len(got_c) > 100000
    assert got_o == want_oligo() and len(got_o) > 100000
    ctx.close()
