#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REFERENCE ITSELF.

Run only in the build container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

How the reference is made to run here:
  * ``Bio`` is provided by oracle/bio_shim (a restatement of Biopython <= 1.77);
  * ``rnascan.BioAddons.motifs._pwm`` is the reference's own _pwm.c compiled into
    oracle/_ref/_pwm.so (the package ``__path__`` is extended, nothing is written to
    /root/reference);
  * pandas >= 2 dropped the positional ``axis`` of ``DataFrame.drop`` used at
    rnascan.py:243 -- a 3-line wrapper restores it.
Everything else is the unmodified reference code: ``rnascan.rnascan.main()`` is invoked
with a patched ``sys.argv`` and its stdout (hits.tab / --bgonly dict) is stored verbatim.

Outputs (all committed):
  inputs/            copies of the reference's example/ and tests/ data + authored fixtures
  cli/<case>.stdout  stdout of the reference CLI for that case; cli/cases.json = argv
  api.json           PSSM tables, dense scores, backgrounds from the reference's functions
"""
import contextlib
import io
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
INP = os.path.join(HERE, "inputs")
sys.path.insert(0, os.path.join(REPO, "oracle", "bio_shim"))
sys.path.insert(0, "/root/reference")

import pandas as pd  # noqa: E402

_orig_drop = pd.DataFrame.drop


def _drop(self, labels=None, *args, **kw):          # rnascan.py:243 on pandas >= 2
    if args:
        kw["axis"] = args[0]
    return _orig_drop(self, labels, **kw)


pd.DataFrame.drop = _drop

import rnascan.BioAddons.motifs as _motifs_pkg  # noqa: E402

_motifs_pkg.__path__.append(os.path.join(REPO, "oracle", "_ref"))
import rnascan.rnascan as ms  # noqa: E402
from rnascan.BioAddons.Alphabet import ContextualSecondaryStructure  # noqa: E402
from Bio.Alphabet import IUPAC  # noqa: E402
from Bio.Seq import Seq  # noqa: E402
from Bio.SeqRecord import SeqRecord  # noqa: E402


# ----------------------------------------------------------------------------- fixtures
def author_fixtures():
    rng = np.random.default_rng(20161018)
    recs = [
        ("rec1 first record, DNA letters", "ACGTTGCATGCATGCCGTAGCTAGCTAGGATCGATCGTAGCTAGCTAGCTAGCATCG"),
        ("rec2 lower/mixed case with N", "acguuGCAUNNNgcaugcauGCUAGCUAGCuagcuagcuaGCUAGCNAUCGAUCG"),
        ("rec3", "UUU"),                               # shorter than any motif
        ("rec4 IUPAC codes and gap", "ACGURYKMACGU-ACGUACGUUUUGCUCUGUAUAUAGGCUCUUUUCAGAGCC"),
        ("rec5", ""),                                  # empty record
        ("rec6 long",
         "".join(rng.choice(list("ACGU"), size=400, p=[.27, .22, .22, .29]))),
    ]
    with open(os.path.join(INP, "mixed.fa"), "w") as fh:
        for title, seq in recs:
            fh.write(">%s\n" % title)
            for k in range(0, len(seq), 60):
                fh.write(seq[k:k + 60] + "\n")
    # matching one-hot structure contexts (same ids/lengths); lower case + one unknown
    with open(os.path.join(INP, "mixed_struct.fa"), "w") as fh:
        for title, seq in recs:
            n = len(seq)
            st, cur = [], "E"
            for _ in range(n):
                if rng.random() > 0.8:
                    cur = str(rng.choice(list("EHTBLRM")))
                st.append(cur)
            st = "".join(st)
            if title.startswith("rec2"):
                st = st[:10] + st[10:20].lower() + st[20:30] + "X" + st[31:]
            fh.write(">%s\n%s\n" % (title, st))
    # averaged profiles for mixed.fa records (PO + B,E,H,L,M,R,T), some exact zeros
    pdir = os.path.join(INP, "profiles_mixed")
    os.makedirs(pdir, exist_ok=True)
    for title, seq in recs:
        rid, n = title.split()[0], len(seq)
        if n == 0:
            continue
        w = rng.dirichlet(0.3 * np.ones(7), size=n)
        w[w < 0.02] = 0.0
        w /= w.sum(axis=1, keepdims=True)
        with open(os.path.join(pdir, "structure.%s.txt" % rid), "w") as fh:
            fh.write("PO\tB\tE\tH\tL\tM\tR\tT\n")
            for i in range(n):
                fh.write(str(i) + "".join("\t" + str(float(v)) for v in w[i]) + "\n")
    # example staged per SURVEY.md H10
    edir = os.path.join(INP, "profiles_example")
    os.makedirs(edir, exist_ok=True)
    shutil.copyfile(os.path.join(INP, "HIST2H3C_3p_end_structure.txt"),
                    os.path.join(edir, "structure.hg19_dna.txt"))
    bg = {}
    with open(os.path.join(INP, "3p_UTR_background_structural_context.txt")) as fh:
        for ln in fh:
            k, v = ln.split()
            bg[k] = float(v)
    with open(os.path.join(INP, "bg_struct_example.txt"), "w") as fh:
        fh.write(repr(bg) + "\n")
    open(os.path.join(INP, "empty.fa"), "w").close()
    with open(os.path.join(INP, "bg_seq_custom.txt"), "w") as fh:
        fh.write(repr({"A": 0.3, "C": 0.2, "G": 0.2, "U": 0.3}) + "\n")
    # (added later; drawn after everything above so that the earlier fixtures stay byte-identical)
    # many records: ordering of hits across records, Match_ID, descriptions with tabs/blanks, DNA letters,
    # lower case, runs of N, records shorter than the motif, and matching one-hot structure contexts
    import gzip
    many = []
    for r in range(300):
        n = int(rng.integers(2, 600))
        seq = rng.choice(list("ACGU"), size=n, p=[.27, .22, .22, .29])
        if r % 7 == 0:
            a = int(rng.integers(0, max(1, n - 5)))
            seq[a:a + int(rng.integers(1, 30))] = "N"
        seq = "".join(seq)
        if r % 5 == 0:
            seq = seq.replace("U", "T")
        if r % 11 == 0:
            seq = seq.lower()
        title = "tx%d" % r + ("" if r % 3 else " gene=G%d  note with  blanks" % (r // 3)) + ("\textra" if r % 13 == 0 else "")
        many.append((title, seq))
    with open(os.path.join(INP, "many.fa"), "w") as fh:
        for title, seq in many:
            fh.write(">%s\n" % title)
            for k in range(0, len(seq), 70):
                fh.write(seq[k:k + 70] + "\n")
    with open(os.path.join(INP, "many_struct.fa"), "w") as fh:
        for title, seq in many:
            st, cur = [], "E"
            for _ in range(len(seq)):
                if rng.random() > 0.8:
                    cur = str(rng.choice(list("EHTBLRM")))
                st.append(cur)
            fh.write(">%s\n%s\n" % (title, "".join(st)))
    with open(os.path.join(INP, "mixed.fa"), "rb") as src, gzip.GzipFile(os.path.join(INP, "mixed.fa.gz"), "wb", mtime=0) as dst:
        dst.write(src.read())


# ----------------------------------------------------------------------------- CLI cases
def P(name):
    return os.path.join("tests", "golden", "inputs", name)


CLI_CASES = {
    # --- sequence only (mode RNA)
    "rna_test_default": ["-p", P("test_seq_pfm.txt"), "-m", "0", P("test.fa")],
    "rna_test_uniform_all": ["-p", P("test_seq_pfm.txt"), "-u", "-m", " -inf", P("test.fa")],
    "rna_example_uniform": ["-p", P("SLBP_pfm_assembled_normalized_seq.txt"), "-u",
                            P("HIST2H3C_3p_end.fa")],
    "rna_example_bg_all": ["-p", P("SLBP_pfm_assembled_normalized_seq.txt"), "-m", " -inf",
                           P("HIST2H3C_3p_end.fa")],
    "rna_mixed_pc": ["-p", P("test_seq_pfm.txt"), "-C", "0.01", "-m", "0.5", P("mixed.fa")],
    "rna_mixed_all": ["-p", P("test_seq_pfm.txt"), "-m", " -inf", "-c", "2", P("mixed.fa")],
    "rna_mixed_bgfile": ["-p", P("test_seq_pfm.txt"), "-b", P("bg_seq_custom.txt"), "-m", "0",
                         P("mixed.fa")],
    "rna_mixed_slbp": ["-p", P("SLBP_pfm_assembled_normalized_seq.txt"), "-m", "-5",
                       P("mixed.fa")],
    "rna_testseq": ["-p", P("test_seq_pfm.txt"), "-m", "-1", "-t", "AGTTCCGGTCCGGCAGAGATCGCG"],
    "rna_bgonly": ["-p", P("test_seq_pfm.txt"), "--bgonly", P("mixed.fa")],
    "rna_nohits": ["-p", P("test_seq_pfm.txt"), "-m", "100", P("mixed.fa")],
    "rna_empty_fasta": ["-p", P("test_seq_pfm.txt"), "-u", P("empty.fa")],
    "rna_allrecords_hit": ["-p", P("test_seq_pfm.txt"), "-m", "1.5", P("test.fa"), ],
    # --- structure only (mode SS)
    "ss_mixed_all": ["-q", P("test_struct_pfm.txt"), "-m", " -inf", P("mixed_struct.fa")],
    "ss_mixed_thr": ["-q", P("test_struct_pfm.txt"), "-u", "-m", "1.0", P("mixed_struct.fa")],
    "ss_mixed_slbp": ["-q", P("SLBP_pfm_assembled_normalized_struct.txt"), "-C", "0.001",
                      "-m", "-20", P("mixed_struct.fa")],
    "ss_bgonly": ["-q", P("test_struct_pfm.txt"), "--bgonly", P("mixed_struct.fa")],
    # --- sequence + one-hot structure FASTA (mode RNASS)
    "rnass_fasta_all": ["-p", P("test_seq_pfm.txt"), "-q", P("test_struct_pfm.txt"),
                        "-m", " -inf", P("mixed.fa"), P("mixed_struct.fa")],
    "rnass_fasta_thr": ["-p", P("test_seq_pfm.txt"), "-q", P("test_struct_pfm.txt"), "-u",
                        "-m", "0", P("mixed.fa"), P("mixed_struct.fa")],
    "rnass_fasta_nohits": ["-p", P("test_seq_pfm.txt"), "-q", P("test_struct_pfm.txt"), "-u",
                           "-m", "50", P("mixed.fa"), P("mixed_struct.fa")],
    "rnass_testseq": ["-p", P("test_seq_pfm.txt"), "-q", P("test_struct_pfm.txt"), "-m", "-2",
                      "-t", "AGUUCCGGUCCGG,EEELLLHHHRRRE"],
    # --- many records / compressed input (added later)
    "rna_many_thr": ["-p", P("test_seq_pfm.txt"), "-C", "0.05", "-m", "2", P("many.fa")],
    "rna_many_slbp": ["-p", P("SLBP_pfm_assembled_normalized_seq.txt"), "-C", "0.01", "-m", "0", P("many.fa")],
    "ss_many_thr": ["-q", P("test_struct_pfm.txt"), "-m", "1.5", P("many_struct.fa")],
    "rnass_many_thr": ["-p", P("test_seq_pfm.txt"), "-q", P("test_struct_pfm.txt"), "-m", "0.5",
                       P("many.fa"), P("many_struct.fa")],
    "rna_gz_input": ["-p", P("test_seq_pfm.txt"), "-m", "0", P("mixed.fa.gz")],
}
# Averaged-structure CLI runs of the reference on py>=3.6 use the label-MISALIGNED
# np.dot.  They are stored to document the divergence; the canonical
# (label-aligned) vectors come from the function-level section below.
CLI_CASES_MISALIGNED = {
    "rnass_avg_example_misaligned": [
        "-p", P("SLBP_pfm_assembled_normalized_seq.txt"),
        "-q", P("SLBP_pfm_assembled_normalized_struct.txt"), "-u", "-m", " -inf",
        P("HIST2H3C_3p_end.fa"), P("profiles_example")],
}


def run_cli(argv):
    out, err = io.StringIO(), io.StringIO()
    old = sys.argv
    sys.argv = ["rnascan"] + argv
    code = 0
    try:
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
            try:
                ms.main()
            except SystemExit as e:
                code = e.code or 0
    finally:
        sys.argv = old
    return out.getvalue(), err.getvalue(), code


# ----------------------------------------------------------------------------- API level
def jf(x):
    """JSON-able float list keeping inf/nan (python json writes Infinity/NaN)."""
    return [float(v) for v in x]


def api_vectors():
    rna, ss = IUPAC.IUPACUnambiguousRNA(), ContextualSecondaryStructure()
    A = {}
    # backgrounds (the reference's only numeric KAT is the first one)
    A["bg_test_fa"] = dict(ms.compute_background([P("test.fa")], rna, verbose=False))
    A["bg_mixed_fa"] = dict(ms.compute_background([P("mixed.fa")], rna, verbose=False))
    A["bg_mixed_struct_fa"] = dict(ms.compute_background([P("mixed_struct.fa")], ss,
                                                         verbose=False))
    with open(P("bg_struct_example.txt")) as fh:
        bg_struct_example = eval(fh.read())
    # PSSMs
    pssm_cases = {
        "test_seq_uniform": (P("test_seq_pfm.txt"), 0, rna, None),
        "test_seq_bg_test_fa": (P("test_seq_pfm.txt"), 0, rna, A["bg_test_fa"]),
        "test_seq_pc": (P("test_seq_pfm.txt"), 0.01, rna, A["bg_mixed_fa"]),
        "test_struct_uniform": (P("test_struct_pfm.txt"), 0, ss, None),
        "test_struct_bg": (P("test_struct_pfm.txt"), 0.5, ss, A["bg_mixed_struct_fa"]),
        "slbp_seq_uniform": (P("SLBP_pfm_assembled_normalized_seq.txt"), 0, rna, None),
        "slbp_struct_examplebg": (P("SLBP_pfm_assembled_normalized_struct.txt"), 0, ss,
                                  bg_struct_example),
        "slbp_struct_uniform_pc": (P("SLBP_pfm_assembled_normalized_struct.txt"), 0.01, ss,
                                   None),
    }
    A["pssm"] = {}
    pssms = {}
    for name, (f, pc, alpha, bg) in pssm_cases.items():
        pm = ms.pfm2pssm(f, pc, alpha, bg)
        pssms[name] = pm
        A["pssm"][name] = {"file": f, "pseudocount": pc,
                           "alphabet": alpha.letters, "background": bg,
                           "values": {l: jf(pm[l]) for l in alpha.letters}}
    # dense calculate() outputs through the reference class (C path and Python path)
    A["calculate"] = {}
    seqs = {
        "K1": "UUUUGCUCUGUAUAUA",
        "mixedcase_amb": "acguuGCAUNNNgcaugcauGCUAGCUAGCTTTtttACGU-ACG",
        "one_window": "ACGU",
        "too_short": "ACG",
    }
    for sname, s in seqs.items():
        for pname in ("test_seq_uniform", "test_seq_bg_test_fa", "test_seq_pc"):
            r = pssms[pname].calculate(s)
            r = np.atleast_1d(np.asarray(r))
            A["calculate"]["%s|%s" % (pname, sname)] = {
                "seq": s, "dtype": str(r.dtype), "scores": jf(r)}
    sseqs = {"ss1": "EEELLLHHHRRREEMMBBTT", "ss_lower_x": "EEEllLHHHRRXEEMMBBTTeh",
             "ss_one": "EHTB"}
    for sname, s in sseqs.items():
        for pname in ("test_struct_uniform", "test_struct_bg"):
            r = pssms[pname].calculate(s)
            r = np.atleast_1d(np.asarray(r, dtype=np.float64))
            A["calculate"]["%s|%s" % (pname, sname)] = {
                "seq": s, "dtype": "float64", "scores": jf(r)}
    # averaged structure, CANONICAL = label-aligned: the reference function is run on a
    # PSSM re-keyed in the profile's column order B,E,H,L,M,R,T so that its positional
    # np.dot pairs equal labels (this is what py2.7/py3.5 dict ordering produced).
    A["averaged"] = {}

    def rekey(pm, order="BEHLMRT"):
        from collections import OrderedDict
        return OrderedDict((l, list(pm[l])) for l in order)

    avg_cases = {
        "example_examplebg": ("slbp_struct_examplebg",
                              P("profiles_example/structure.hg19_dna.txt")),
        "example_uniform_pc": ("slbp_struct_uniform_pc",
                               P("profiles_example/structure.hg19_dna.txt")),
        "mixed_rec6_nonfinite": ("slbp_struct_examplebg",
                                 P("profiles_mixed/structure.rec6.txt")),
        "mixed_rec1_test": ("test_struct_bg", P("profiles_mixed/structure.rec1.txt")),
        "mixed_rec3_short": ("test_struct_bg", P("profiles_mixed/structure.rec3.txt")),
    }
    for name, (pname, f) in avg_cases.items():
        for thr in (float("-inf"), 6.0, 0.0):
            with np.errstate(all="ignore"):
                df = ms.scan_averaged_structure(f, {"m": rekey(pssms[pname])}, thr)
            rows = [] if df.shape[0] == 0 else \
                [[int(r.Start), int(r.End), float(r.LogOdds)] for r in df.itertuples()]
            A["averaged"]["%s|%r" % (name, thr)] = {
                "pssm": pname, "profile": f, "threshold": thr, "rows": rows}
    # combine() on the example: sequence hits (uniform bg, m=6) x aligned struct hits
    seq_pssm = {"SLBP_pfm_assembled_normalized_seq": pssms["slbp_seq_uniform"]}

    class _A(object):
        minscore, debug, cores = 6.0, True, 1
    with contextlib.redirect_stderr(io.StringIO()):
        seq_df = ms.scan_main(P("HIST2H3C_3p_end.fa"), seq_pssm, rna, None, _A())
    with np.errstate(all="ignore"):
        st_df = ms.scan_averaged_structure(
            P("profiles_example/structure.hg19_dna.txt"),
            {"SLBP_pfm_assembled_normalized_struct": rekey(pssms["slbp_struct_examplebg"])},
            6.0)
    ms._add_sequence_id(st_df, "hg19_dna", "")
    cols = st_df.columns.tolist()
    st_df = st_df[cols[-2:] + cols[:-2]]
    comb = ms.combine(seq_df, st_df)
    ms._add_match_id(comb)
    buf = io.StringIO()
    comb.to_csv(buf, sep="\t", index=False)
    A["combine_example_aligned_tsv"] = buf.getvalue()
    A["pfmutil"] = pfmutil_vectors()
    return A


def pfmutil_vectors():
    """Outputs of the reference's rnascan/pfmutil.py (pure Python, imported as is)."""
    import tempfile
    from rnascan import pfmutil as pu
    V = {}
    struct = pu.read_pfm(P("SLBP_pfm_assembled_normalized_struct.txt"))
    seq = pu.read_pfm(P("test_seq_pfm.txt"))
    V["read_struct"] = struct
    V["format_struct"] = pu.format_pfm(struct)
    V["format_seq"] = pu.format_pfm(seq)
    V["norm_seq"] = pu.norm_pfm(seq)
    V["is_normalized"] = [pu.is_normalized(seq), pu.is_normalized(pu.norm_pfm(seq)),
                          pu.is_normalized(struct)]
    V["from_IUPAC"] = pu.pfm_from_IUPAC("ACGURYSWKMBDHVN")
    V["from_string"] = pu.pfm_from_string("EHLLRT", pu.FULL_STRUCT_ALPHABET)
    pwm = pu.pfm_to_pwm(pu.norm_pfm(seq), 20)
    V["to_pwm_seq_20"] = pwm
    V["scan_fwd_seq"] = {"seq": "UUUUGCUCUGUAUAUAGGCAUCG", "scores": pu.pwm_scan_fwd(pwm, "UUUUGCUCUGUAUAUAGGCAUCG")}
    spwm = pu.pfm_to_pwm(pu.norm_pfm(pu.read_pfm(P("test_struct_pfm.txt"))), 7)
    V["scan_fwd_struct"] = {"seq": "EEELLLHHHRRREEMMBBTT", "scores": pu.pwm_scan_fwd(spwm, "EEELLLHHHRRREEMMBBTT"),
                            "pwm": spwm}
    V["reduce_struct"] = pu.reduce_pfm_alphabet(struct)
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "multi.txt")
        pu.write_multi_pfm(["m1", "m2"], [seq, struct], f)
        V["multi_text"] = open(f).read()
        V["multi_iter"] = [[i, {k: list(v) for k, v in pfm.items()}] for i, pfm in pu.multi_pfm_iter(f)]
    return V


def dotbracket_vectors():
    """Annotations by the reference's own C++ tool (oracle/_ref/parse_secondary_structure, built
    from /root/reference/scripts/parse_secondary_structure.cpp) for seeded random balanced
    structures and a few hand-written ones."""
    import random
    import subprocess
    import tempfile
    rnd = random.Random(20170106)

    def rand_struct(n):
        s, depth = [], 0
        while len(s) < n:
            r, rem = rnd.random(), n - len(s)
            if depth >= rem:
                s.append(")"); depth -= 1
            elif r < 0.35:
                s.append(".")
            elif r < 0.70 and rem > depth + 1:
                s.append("("); depth += 1
            elif depth > 0:
                s.append(")"); depth -= 1
            else:
                s.append(".")
        return "".join(s)
    structs = [rand_struct(rnd.randint(2, 160)) for _ in range(400)]
    structs += ["..((...))..", "((..((...))..((...))..))", "(((...)))", "((.))..((.))", "(.)", "....",
                "(((..)).(..))", "((((...)).((...)).))", ".((..)).", "(.(.(.).).)", "((.(...).(...).))..(..)",
                "((((((..((((........)))).(((((.......))))).....(((((.......))))))))))).."]
    structs = [s for s in structs if "." in s and s.count("(") == s.count(")")]
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "in.txt"), os.path.join(d, "out.txt")
        with open(fi, "w") as fh:
            fh.write("\n".join(structs) + "\n")
        subprocess.check_call([os.path.join(REPO, "oracle", "_ref", "parse_secondary_structure"), fi, fo])
        with open(fo) as fh:
            ann = fh.read().split("\n")[:-1]
    assert len(ann) == len(structs)
    return [[s, a] for s, a in zip(structs, ann)]


def averaging_vectors():
    """struct_pfm_from_aligned + norm_pfm of the reference (average_structure.py:28-42, pfmutil.py:136-151)
    on seeded aligned annotation strings."""
    import random
    from rnascan import average_structure as av
    from rnascan import pfmutil as pu
    rnd = random.Random(7)
    cases = []
    for L, nfrag, flen in ((40, 9, 20), (120, 30, 50), (7, 3, 7)):
        seqs = []
        for k in range(nfrag):
            start = (k * (L - flen)) // max(1, nfrag - 1)          # first at 0, last flush with the end
            body = "".join(rnd.choice("BEHLMRT") for _ in range(min(flen, L - start)))
            seqs.append("-" * start + body + "-" * (L - start - len(body)))
        counts = av.struct_pfm_from_aligned(seqs)
        cases.append({"sequences": seqs, "counts": counts, "profile": pu.norm_pfm(counts),
                      "text": pu.format_pfm(pu.norm_pfm(counts))})
    return cases


def main():
    os.chdir(REPO)
    author_fixtures()
    cdir = os.path.join(HERE, "cli")
    os.makedirs(cdir, exist_ok=True)
    meta = {}
    for group, cases in (("aligned_irrelevant", CLI_CASES), ("misaligned", CLI_CASES_MISALIGNED)):
        for name, argv in cases.items():
            out, err, code = run_cli(argv)
            with open(os.path.join(cdir, name + ".stdout"), "w") as fh:
                fh.write(out)
            meta[name] = {"argv": argv, "exit": code, "group": group,
                          "stderr_lines": [l for l in err.splitlines()
                                           if "seconds" not in l and "minutes" not in l]}
            print("%-32s exit=%s stdout_lines=%d" % (name, code, out.count("\n")))
    with open(os.path.join(cdir, "cases.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    A = api_vectors()
    with open(os.path.join(HERE, "api.json"), "w") as fh:
        json.dump(A, fh, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "dotbracket.json"), "w") as fh:
        json.dump(dotbracket_vectors(), fh, indent=0)
    with open(os.path.join(HERE, "averaging.json"), "w") as fh:
        json.dump(averaging_vectors(), fh, indent=0)
    print("api.json written: %d pssm, %d calculate, %d averaged" %
          (len(A["pssm"]), len(A["calculate"]), len(A["averaged"])))


if __name__ == "__main__":
    main()
