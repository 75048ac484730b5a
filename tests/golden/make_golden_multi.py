#!/usr/bin/env python
"""Golden outputs for MOTIF COLLECTIONS, produced by the REFERENCE ITSELF (build container only).

The reference scans one PFM (pair) per run; a collection means one run per motif.  For every case below the
reference's own main() (same harness as make_golden.py: oracle/bio_shim for Bio, the reference's compiled
_pwm.c) is run once per motif pair with single-PFM files and the stdouts are concatenated -- that is what
``rnascan -p <multi-PFM> -q <multi-PFM> ...`` must print.  The same motifs are also written as multi-PFM
files with the reference's pfmutil.write_multi_pfm.

    make -C oracle ref && python tests/golden/make_golden_multi.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets up the reference + shims)
import rnascan.pfmutil as ref_pfmutil  # noqa: E402

N_MOTIFS = 8
MULTI = os.path.join(mg.INP, "multi")


def author_motifs():
    os.makedirs(MULTI, exist_ok=True)
    rng = np.random.default_rng(20261018)
    seq_ids, struct_ids, seq_pfms, struct_pfms = [], [], [], []
    for k in range(N_MOTIFS):
        W = int(rng.integers(7, 13))
        for kind, letters, ids, pfms in (("mseq", "ACGU", seq_ids, seq_pfms), ("mstruct", "BEHLMRT", struct_ids, struct_pfms)):
            rows = rng.dirichlet(0.4 * np.ones(len(letters)), size=W)
            pfm = {l: [round(float(v), 4) for v in rows[:, i]] for i, l in enumerate(letters)}
            name = "%s%d" % (kind, k)
            ref_pfmutil.write_pfm(pfm, os.path.join(MULTI, name + ".txt"))
            ids.append(name)
            pfms.append(pfm)
    ref_pfmutil.write_multi_pfm(seq_ids, seq_pfms, os.path.join(MULTI, "multi_seq.pfm"))
    ref_pfmutil.write_multi_pfm(struct_ids, struct_pfms, os.path.join(MULTI, "multi_struct.pfm"))
    return seq_ids, struct_ids


def M(name):
    return os.path.join("tests", "golden", "inputs", "multi", name)


P = mg.P
# name -> (argv with {seq} / {struct} placeholders for the single-PFM runs, extra flags of OUR multi run)
CASES = {
    "multi_rna_mixed": (["-p", "{seq}", "-C", "0.01", "-m", "0.5", P("mixed.fa")], []),
    "multi_ss_fasta": (["-q", "{struct}", "-C", "0.01", "-m", "0", P("mixed_struct.fa")], []),
    "multi_rnass_fasta": (["-p", "{seq}", "-q", "{struct}", "-C", "0.01", "-u", "-m", " -1", P("mixed.fa"),
                           P("mixed_struct.fa")], []),
    # averaged profiles: the reference on py >= 3.6 pairs columns by position -> --reference-compat
    "multi_ss_avg": (["-q", "{struct}", "-C", "0.01", "-B", P("bg_struct_example.txt"), "-m", "0.5", P("profiles_mixed")],
                     ["--reference-compat"]),
    "multi_ss_avg_nohits": (["-q", "{struct}", "-C", "0.01", "-B", P("bg_struct_example.txt"), "-m", "100", P("profiles_mixed")],
                            ["--reference-compat"]),
    "multi_rnass_avg": (["-p", "{seq}", "-q", "{struct}", "-C", "0.01", "-u", "-m", " -6", P("mixed.fa"),
                         P("profiles_mixed")], ["--reference-compat"]),
}


def main():
    os.chdir(mg.REPO)
    seq_ids, struct_ids = author_motifs()
    cdir = os.path.join(HERE, "cli")
    meta = {}
    for name, (template, extra) in CASES.items():
        outs, code = [], 0
        for s_id, q_id in zip(seq_ids, struct_ids):
            argv = [a.replace("{seq}", M(s_id + ".txt")).replace("{struct}", M(q_id + ".txt")) for a in template]
            out, err, c = mg.run_cli(argv)
            outs.append(out)
            code = code or c
        with open(os.path.join(cdir, name + ".stdout"), "w") as fh:
            fh.write("".join(outs))
        multi_argv = [a.replace("{seq}", M("multi_seq.pfm")).replace("{struct}", M("multi_struct.pfm")) for a in template]
        meta[name] = {"argv": multi_argv + extra, "exit": code, "runs": len(outs)}
        print("%-20s exit=%s runs=%d stdout_lines=%d" % (name, code, len(outs), "".join(outs).count("\n")))
    with open(os.path.join(cdir, "multi_cases.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
