"""
Golden vectors for SURVEY §8f row N1 (haplotype construction + the haplotype selection loop).

Run in the BUILD container (needs /root/reference and oracle/_ref/n1_ref, built by oracle/build.py from the
reference's own getFilteredHaplotypes / computeBestScoreForGenotype / isHaplotypeValid source lines).  The inputs
are tests/cases.py n1_window_case(seed, drop); this file stores what the REFERENCE returns for them:
  ref_seq / hap_start   the reference haplotype (Haplotype(..., variants=()))
  sel_mask, hap_seq     the variant sets getFilteredHaplotypes returns, in order, and Haplotype.cHaplotypeSequence of each
  hap_score             computeBestScoreForHaplotype of the reference haplotype (ref_hap_score) and of every selected one
  trial_mask / score    every trial set the rounds score (in the order they are scored) with the reference's
                        computeBestScoreForGenotype value
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import select_oracle as S  # noqa: E402
from tests import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEEDS = list(range(72)) + list(range(100, 112)) + list(range(100, 112))
DROP = [0] * 72 + [1] * 12 + [2] * 12      # cases.n1_window_case(seed, drop): all reads / the first individual's reads removed


def main():
    ref = O.ref_l3()
    assert ref is not None, "needs /root/reference (oracle/build.py)"
    d = {k: [] for k in ("ref_seq", "hap_start", "sel_mask", "hap_seq", "trial_mask", "trial_score", "hap_score", "ref_hap_score")}
    ref_off, sel_off, hs_off, trial_off = [0], [0], [0], [0]
    opts = {k: [] for k in ("max_haplotypes", "original_max_haplotypes", "max_variants", "filter_by_coverage", "coverage_sampling_level")}
    for seed, drop in zip(SEEDS, DROP):
        c = cases.n1_window_case(seed, drop)
        o = c["opts"]
        args = (c["genome"], c["win_start"], c["win_end"], c["variants"], c["per_ind"], c["max_read_len"], o["max_haplotypes"],
                o["original_max_haplotypes"], o["max_variants"], o["filter_by_coverage"], o["coverage_sampling_level"])
        r = ref.select_haplotypes(*args)
        # the trial sets are those the restated loop visits; their scores come from the reference
        w = cases.n1_select_window(c, r["ref_seq"], r["hap_start"])
        tr = []
        got = S.select_haplotypes(w, o["max_haplotypes"], o["original_max_haplotypes"], o["max_variants"], o["filter_by_coverage"],
                                  o["coverage_sampling_level"], trace=tr)
        assert [g[0] for g in got] == r["selected"], seed
        sets = [vs for rnd in tr for (vs, _) in rnd]
        scores = ref.select_haplotypes(*args, 0, sets)["scores"] if sets else []
        assert scores == [s for rnd in tr for (_, s) in rnd], seed
        # computeBestScoreForHaplotype (variantFilter.pyx:212-234) of the reference haplotype and of every selected one
        hs = ref.select_haplotypes(*args, 0, None, [()] + r["selected"])["hap_scores"]
        assert hs == S.best_score_haplotypes(w, [()] + [tuple(w.vars[i] for i in s_) for s_ in r["selected"]]), seed
        d["ref_hap_score"].append(hs[0])
        d["hap_score"] += hs[1:]
        d["ref_seq"].append(np.frombuffer(r["ref_seq"], np.uint8))
        ref_off.append(ref_off[-1] + len(r["ref_seq"]))
        d["hap_start"].append(r["hap_start"])
        d["sel_mask"] += cases.masks_of(r["selected"])
        sel_off.append(sel_off[-1] + len(r["selected"]))
        for sq in r["hap_seqs"]:
            d["hap_seq"].append(np.frombuffer(sq, np.uint8))
            hs_off.append(hs_off[-1] + len(sq))
        d["trial_mask"] += cases.masks_of(sets)
        d["trial_score"] += scores
        trial_off.append(trial_off[-1] + len(sets))
        for k in opts:
            opts[k].append(o[k])
    np.savez_compressed(
        os.path.join(HERE, "n1_ref.npz"), seeds=np.asarray(SEEDS, np.int32), drop=np.asarray(DROP, np.int32), ref_seq=np.concatenate(d["ref_seq"]),
        ref_off=np.asarray(ref_off, np.int64), hap_start=np.asarray(d["hap_start"], np.int32),
        sel_mask=np.asarray(d["sel_mask"], np.uint64), sel_off=np.asarray(sel_off, np.int64),
        hap_seq=np.concatenate(d["hap_seq"]), hap_seq_off=np.asarray(hs_off, np.int64),
        trial_mask=np.asarray(d["trial_mask"], np.uint64), trial_score=np.asarray(d["trial_score"], np.float64),
        trial_off=np.asarray(trial_off, np.int64), hap_score=np.asarray(d["hap_score"], np.float64),
        ref_hap_score=np.asarray(d["ref_hap_score"], np.float64), **{"opt_" + k: np.asarray(v, np.int32) for k, v in opts.items()})
    print("windows", len(SEEDS), "selected", sel_off[-1], "trials", trial_off[-1])


if __name__ == "__main__":
    main()
