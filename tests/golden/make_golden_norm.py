#!/usr/bin/env python3
"""Golden files for cohort score normalisation: run the UNMODIFIED reference script
utils/adaptive_score_normalization.py on small synthetic score files.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_norm.py

The script is a top-level program with two hard-coded input paths (:17-18) and four output paths derived from
them (:75-78).  It is executed as it is with runpy; np.genfromtxt / np.savetxt are wrapped only to redirect
those paths to the fixture files written here (c7_raw_scores.tsv, c7_cohort_scores.tsv) and to
tests/golden/c7_raw_scores.tsv_{znorm,tnorm,snorm,asnorm1}.tsv.  ASnorm_topN stays the script's 500, so the
cohort has 510 utterances (the top-N slice is a strict subset).
"""
import os
import runpy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SCRIPT = "/root/reference/utils/adaptive_score_normalization.py"
RAW = os.path.join(HERE, "c7_raw_scores.tsv")
COHORT = os.path.join(HERE, "c7_cohort_scores.tsv")


def write_inputs():
    rng = np.random.default_rng(7007)
    enrols = ["spk%02d" % i for i in range(4)]
    tests = ["seg%02d" % i for i in range(3)]
    cohort = ["coh%03d" % i for i in range(510)]
    with open(COHORT, "w") as f:
        f.write("modelid\tsegmentid\tside\tLLR\n")
        for k, idn in enumerate(enrols + tests):
            sc = (rng.normal(-0.8 + 0.05 * k, 0.25 + 0.01 * k, len(cohort))).astype(np.float32)
            sc[5] = sc[17]                                     # ties, also across the top-N boundary candidates
            sc[100:104] = sc.min()
            for c, s in zip(cohort, sc):
                f.write("%s\t%s\ta\t%s\n" % (idn, c, str(s)))
    with open(RAW, "w") as f:
        f.write("modelid\tsegmentid\tside\tLLR\n")
        for i in range(40):
            e, t = enrols[int(rng.integers(4))], tests[int(rng.integers(3))]
            s = np.float32(rng.normal(0.6 if i % 4 == 0 else -0.8, 0.2))
            f.write("%s\t%s.sph\ta\t%s\n" % (e, t, str(s)))


def main():
    write_inputs()
    real_gen, real_save = np.genfromtxt, np.savetxt

    def gen(fname, *a, **k):
        fname = COHORT if "cohort" in os.path.basename(fname) else RAW
        return real_gen(fname, *a, **k)

    def save(fname, *a, **k):
        suffix = fname[fname.rindex("_before_norm.tsv") + len("_before_norm.tsv"):]
        return real_save(RAW + suffix, *a, **k)

    np.genfromtxt, np.savetxt = gen, save
    try:
        runpy.run_path(REF_SCRIPT, run_name="__main__")
    finally:
        np.genfromtxt, np.savetxt = real_gen, real_save
    for s in ("znorm", "tnorm", "snorm", "asnorm1"):
        p = RAW + "_%s.tsv" % s
        print(p, os.path.getsize(p))


if __name__ == "__main__":
    main()
