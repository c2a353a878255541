"""Python mirror of the deterministic rules of the stub tools (oracle/stub_bwa.sh with MIPGEN_STUB_RULES=1,
oracle/stub_trf.sh, oracle/stub_tabix.sh): builds the per-region copy tables, unmappable MIP starts, masked
sequence and SNP flags that the reference CLI ends up with when it runs against those stubs, so that the
oracle / the device can be given the same selection inputs (mipgen.cpp:558-596, 796-873, 1045-1084, 875-978)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from mipgen_b200.panel import Config, Region


def arm_copy(start: int, stop: int) -> int:
    k = (31 * start + stop - start) % 97
    return 101 if k == 0 else 30 if k == 1 else 3 if k < 5 else 2 if k < 9 else 1


def copy_table(cfg: Config, r: Region) -> np.ndarray:
    """copy_chr_start_stop as find_copy fills it from the stub's SAM: reads exist for relative starts
    0 .. len(seq) - size - 1 (mipgen.cpp:826-828); every other key reads as 0 (mipgen.cpp:612-613)."""
    sizes = cfg.oligo_sizes
    t = np.zeros((len(sizes), len(r.seq)), np.int32)
    for k, size in enumerate(sizes):
        for i in range(0, len(r.seq) - size):
            t[k, i] = arm_copy(r.seq_start + i, r.seq_start + i + size - 1)
    return t


def unmappable_table(cfg: Config, r: Region) -> np.ndarray:
    """unmappable_positions[capture][chr] restricted to this region (mipgen.cpp:808-823, 857-866)."""
    caps = cfg.captures
    t = np.zeros((len(caps), len(r.seq)), np.uint8)
    for ci, cap in enumerate(caps):
        for start in range(r.start_flanked - cap, r.stop_flanked):
            if start > 0 and start + cap - 1 <= r.seq_stop and (start + cap) % 53 == 0:
                i = start - r.seq_start
                if 0 <= i < len(r.seq):
                    t[ci, i] = 1
    return t


def masked_sequence(r: Region) -> bytes:
    a = bytearray(r.seq)
    for i in range(len(a)):
        if (r.seq_start + i) % 211 < 17:
            a[i] = ord("N")
    return bytes(a)


def snp_positions(genome: bytes, regions: Sequence[Region]) -> List[Tuple[int, str, str]]:
    """A deterministic set of variants inside / around the regions: (1-based position, ref, alt)."""
    out = []
    comp = {"A": "C", "C": "G", "G": "T", "T": "A"}
    for r in regions:
        for x in range(r.seq_start, r.seq_stop + 1):
            if x % 37 == 5:
                ref = chr(genome[x - 1])
                if x % 5 == 0:
                    out.append((x, ref + chr(genome[x]), ref))        # deletion: marks position x+1 (mipgen.cpp:961-967)
                elif x % 7 == 0:
                    out.append((x, ref, "N"))                          # unsupported allele: snp_failed
                else:
                    out.append((x, ref, comp.get(ref, "A")))
    return sorted(set(out))


def write_vcf(path: str, snps: Sequence[Tuple[int, str, str]], chrom: str = "1") -> None:
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\n")
        for pos, ref, alt in snps:
            f.write("%s\t%d\t.\t%s\t%s\t.\t.\t.\n" % (chrom, pos, ref, alt))


def snp_flags(r: Region, snps: Sequence[Tuple[int, str, str]]) -> np.ndarray:
    """chr_snp_positions restricted to the region: an indel of reference length L marks positions pos+1 .. pos+L-1,
    anything else marks pos (parse_vcf, mipgen.cpp:961-971)."""
    f = np.zeros(len(r.seq), np.uint8)
    for pos, ref, _alt in snps:
        marked = range(pos + 1, pos + len(ref)) if len(ref) > 1 else [pos]
        for x in marked:
            i = x - r.seq_start
            if 0 <= i < len(r.seq):
                f[i] = 1
    return f


def decorate(cfg: Config, genome: bytes, regions: Sequence[Region], copies=True, unmappable=True, masked=True, snps=None):
    """Attach the stub-rule inputs to the regions in place."""
    for r in regions:
        if copies:
            r.copies = copy_table(cfg, r)
        if unmappable:
            r.unmappable = unmappable_table(cfg, r)
        if masked:
            r.masked_seq = masked_sequence(r)
        if snps is not None:
            r.snp = snp_flags(r, snps)
    return regions
