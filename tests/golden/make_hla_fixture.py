"""
BASELINE config 1 fixture: the reference's own test data (test/S55_test_realigned.bam + the HLA-A allele
VCF) as a handful of windows, scored by the reference's own compiled code.

Run in the BUILD container (needs /root/reference and oracle/_ref); the GPU box only reads the committed
tests/golden/hla_window_ref.npz.

What the fixture holds
  * reads: the records of the BAM that fall into the window slices of four HLA-A windows (exon 1 at
    6:29910331 and the three best-covered ones), decoded, filtered and quality-trimmed by platypus_b200/reads.py (mirror of htslibWrapper.pyx:328-406,
    cwindow.pyx:208-236, 332-481, 560-595), capped to keep the file small;
  * haplotypes: the REF string and the first distinct ALT strings of the VCF records at that position
    (each ALT is a whole-window HLA allele, the way --HLATyping=1 --source= uses them), flanked by a
    consensus of the reads themselves (the BAM's reference FASTA is not shipped); the flank is
    shorter than the reads so that some of them hang over the haplotype ends and HLA mode clips them;
  * reference outputs: mapAndAlignReadToHaplotype scores from the reference's calign.pyx for every
    (read, haplotype) pair in default mode, in HLA mode (read clipped to the haplotype as
    chaplotype.pyx:647-655 does, hashes of the unclipped read) and with doCalculateFlankScore=1.
"""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from platypus_b200 import reads as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PLATYPUS_REFERENCE", "/root/reference")
BAM = os.path.join(REF, "test", "S55_test_realigned.bam")
VCF = os.path.join(REF, "test", "HLA_A_alignment_hapsREF.vcf.gz")
# 1-based VCF positions of the windows used: exon 1 (the position SURVEY §8d names) and the three windows
# of the file with the deepest coverage
POSITIONS = [29910331, 29911899, 29912836, 29913011]
FLANK = 200
MAX_READS = 64
MAX_HAPS = 8


def consensus(reads, lo, hi):
    """Majority base per reference position in [lo, hi) from the M-aligned bases of the reads."""
    counts = np.zeros((hi - lo, 256), np.int32)
    for r in reads:
        ref_pos = r.pos + (r.cigar[0][1] if r.cigar and r.cigar[0][0] == 4 else 0)
        q = 0
        for op, n in r.cigar:
            if op in (0, 7, 8):
                for k in range(n):
                    p = ref_pos + k
                    if lo <= p < hi:
                        counts[p - lo, r.seq[q + k]] += 1
                ref_pos += n
                q += n
            elif op in (1, 4):
                q += n
            elif op in (2, 3):
                ref_pos += n
    out = bytearray()
    for i in range(hi - lo):
        c = counts[i]
        out.append(int(np.argmax(c)) if c.sum() > 0 else ord("N"))
    return bytes(out)


def main():
    cw = O.ref_calign()
    assert cw is not None, "reference calign.pyx not built"
    refs, recs = R.decode_bam(BAM)
    print("BAM:", len(recs), "records on", refs[recs[0].chrom_id])
    buf = R.ReadBuffer()       # per-sample buffer (cwindow.pyx:560-595)
    for r in recs:
        buf.add(r)
    print("buffer: %d good, %d bad reads; filter counts %s" % (len(buf.reads), len(buf.bad_reads), buf.counts))
    windows = []
    for POS in POSITIONS:
        alleles = []
        ref_allele = None
        with gzip.open(VCF, "rt") as f:
            for line in f:
                if line.startswith("#"):
                    continue
                c = line.split("\t")
                if int(c[1]) != POS:
                    continue
                ref_allele = c[3].encode()
                a = c[4].encode()
                if a != ref_allele and a not in alleles:
                    alleles.append(a)
        assert ref_allele is not None
        win_start = POS - 1                       # 0-based
        win_end = win_start + len(ref_allele)
        hap_start = win_start - FLANK
        good = R.window_slice(buf.reads, win_start, win_end)        # cwindow.pyx:208-236
        bad = R.window_slice(buf.bad_reads, win_start, win_end)
        good = good[::max(1, len(good) // MAX_READS)][:MAX_READS]
        bad = bad[:12]
        left = consensus(buf.reads, hap_start, win_start)
        right = consensus(buf.reads, win_end, win_end + FLANK)
        haps = [left + a + right for a in [ref_allele] + alleles[:MAX_HAPS - 1]]
        print("window [%d, %d): %d good, %d bad reads, %d haplotypes (%d alleles in the VCF), N in flanks %d" %
              (win_start, win_end, len(good), len(bad), len(haps), len(alleles) + 1, left.count(b"N") + right.count(b"N")))
        windows.append((win_start, win_end, hap_start, haps, good, bad))

    modes = {"default": (0, 0), "hla": (1, 0), "flank": (0, 1)}
    all_haps, all_reads, scores = [], [], {k: [] for k in modes}
    meta = []
    for win_start, win_end, hap_start, haps, good, bad in windows:
        rd = good + bad
        meta.append((win_start, win_end, hap_start, len(haps), len(good), len(bad)))
        all_haps += haps
        all_reads += rd
        sc = {k: np.zeros((len(haps), len(rd)), np.int32) for k in modes}
        for hi_, hap in enumerate(haps):
            go = O.gap_open(hap)
            for ri, r in enumerate(rd):
                seq, qual, pos = r.seq, bytes(r.qual), r.pos
                for name, (hla, flank) in modes.items():
                    s_, q_, p_ = seq, qual, pos
                    if hla:   # chaplotype.pyx:647-655
                        o1 = max(0, hap_start - pos)
                        o2 = max(0, pos + len(seq) - win_start - len(hap))
                        s_, q_, p_ = seq[o1:len(seq) - o2], qual[o1:len(seq) - o2], pos + o1
                    if len(s_) < 7:
                        continue
                    sc[name][hi_, ri] = cw.map_and_align(s_, q_, p_, hap_start, hap, go, 3, 2, FLANK, flank, 0, seq)
        for k in modes:
            scores[k].append(sc[k].reshape(-1))

    def pack(lst):
        off = np.zeros(len(lst) + 1, np.int64)
        np.cumsum([len(b) for b in lst], out=off[1:])
        return off, np.frombuffer(b"".join(lst), np.uint8).copy()
    ho, hs = pack(all_haps)
    ro, rs = pack([r.seq for r in all_reads])
    _, qs = pack([bytes(r.qual) for r in all_reads])
    out = {k: np.concatenate(v) for k, v in scores.items()}
    np.savez_compressed(
        os.path.join(HERE, "hla_window_ref.npz"), hap_off=ho, hap=hs, read_off=ro, read=rs, qual=qs,
        read_pos=np.array([r.pos for r in all_reads], np.int32), read_end=np.array([r.end for r in all_reads], np.int32),
        read_mapq=np.array([r.mapq for r in all_reads], np.uint8),
        read_qcfail=np.array([1 if r.flag & R.F_QCFAIL else 0 for r in all_reads], np.uint8),
        windows=np.array(meta, np.int32),      # win_start, win_end, hap_start, n_haps, n_good, n_bad
        score_default=out["default"], score_hla=out["hla"], score_flank=out["flank"])
    for k, v in out.items():
        print(k, "scores:", len(v), "range", int(v.min()), int(v.max()), "median", int(np.median(v)))
    print("hla_window_ref.npz written:", os.path.getsize(os.path.join(HERE, "hla_window_ref.npz")), "bytes")


if __name__ == "__main__":
    main()
