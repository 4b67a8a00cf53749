"""
Golden vectors for the --HLATyping haplotype selection (src/cython/variantFilter.pyx:655-736
getAllHLAHaplotypesInRegion).

Run in the BUILD container (needs /root/reference and oracle/_ref/n1_ref, built by oracle/build.py from the reference's own
source lines of that function, computeBestScoreForHaplotype and computeBestScoreForGenotype).  The inputs are
tests/cases.py hla_window_case(seed); this file stores what the REFERENCE returns for them:
  haps        the variant index of every haplotype the function returns, in order (a haplotype may appear twice)
  hap_score   computeBestScoreForHaplotype of every FILE_VAR haplotype
  gt_score    computeBestScoreForGenotype(best haplotype, it) of every FILE_VAR haplotype
  ref_seq / hap_start   the window's reference haplotype
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEEDS = list(range(24))


def main():
    ref = O.ref_l3()
    assert ref is not None and hasattr(ref, "hla_haplotypes"), "needs /root/reference (oracle/build.py)"
    haps, hs, gs, refs = [], [], [], []
    hap_off, fv_off, ref_off, hap_start = [0], [0], [0], []
    for seed in SEEDS:
        c = cases.hla_window_case(seed)
        o = c["opts"]
        r = ref.hla_haplotypes(c["genome"], c["win_start"], c["win_end"], c["variants"], c["per_ind"], c["max_read_len"],
                               o["original_max_haplotypes"], o["coverage_sampling_level"])
        assert r["file_vars"] == [i for i, v in enumerate(c["variants"]) if v[4] == 2]
        haps += r["haps"]
        hap_off.append(len(haps))
        hs += r["hap_scores"]
        gs += r["gt_scores"]
        fv_off.append(len(hs))
        refs.append(np.frombuffer(r["ref_seq"], np.uint8))
        ref_off.append(ref_off[-1] + len(r["ref_seq"]))
        hap_start.append(r["hap_start"])
    np.savez_compressed(os.path.join(HERE, "n1_hla_ref.npz"), seeds=np.asarray(SEEDS, np.int32), haps=np.asarray(haps, np.int32),
                        hap_off=np.asarray(hap_off, np.int64), hap_score=np.asarray(hs, np.float64),
                        gt_score=np.asarray(gs, np.float64), fv_off=np.asarray(fv_off, np.int64),
                        ref_seq=np.concatenate(refs), ref_off=np.asarray(ref_off, np.int64),
                        hap_start=np.asarray(hap_start, np.int32))
    print("windows", len(SEEDS), "haplotypes returned", len(haps), "scored", len(hs))


if __name__ == "__main__":
    main()
