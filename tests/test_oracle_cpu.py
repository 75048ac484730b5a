"""The oracle is pinned here: against the golden vectors produced by running the REFERENCE's
own code in the build container (tests/golden/make_golden.py), against the reference's only
numeric known-answer test (tests/motif_scan_test.py:33-43) and, when present, against the
reference's own compiled kernel oracle/_ref/_pwm.so on random inputs."""
import importlib.machinery
import importlib.util
import math
import os

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INP = os.path.join(REPO, "tests", "golden", "inputs")


def same(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and \
        np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def load_ref_pwm():
    path = os.path.join(REPO, "oracle", "_ref", "_pwm.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/_pwm.so not built (reference tree absent)")
    loader = importlib.machinery.ExtensionFileLoader("_pwm", path)
    mod = importlib.util.module_from_spec(importlib.util.spec_from_loader("_pwm", loader))
    loader.exec_module(mod)
    return mod


def test_background_known_answer_of_the_reference(oracle):
    # /root/reference/tests/motif_scan_test.py:33-43
    recs = oracle.parse_fasta(os.path.join(INP, "test.fa"))
    bg = oracle.background([oracle.preprocess_rna(s) for _, _, s in recs], oracle.RNA_LETTERS)
    want = {"A": 0.19444444, "C": 0.13888888, "U": 0.52777777, "G": 0.13888888}
    for k, v in want.items():
        assert abs(bg[k] - v) < 1e-3
    assert list(bg) == list("GAUC")


def test_backgrounds_match_golden(oracle, golden_api):
    for fa, key, letters, prep in (("test.fa", "bg_test_fa", oracle.RNA_LETTERS, True),
                                   ("mixed.fa", "bg_mixed_fa", oracle.RNA_LETTERS, True),
                                   ("mixed_struct.fa", "bg_mixed_struct_fa", oracle.SS_LETTERS, False)):
        seqs = [s for _, _, s in oracle.parse_fasta(os.path.join(INP, fa))]
        if prep:
            seqs = [oracle.preprocess_rna(s) for s in seqs]
        bg = oracle.background(seqs, letters)
        assert bg == golden_api[key]


def test_pssm_matches_golden(oracle, golden_api):
    for name, g in golden_api["pssm"].items():
        counts = oracle.read_pfm_table(os.path.join(REPO, g["file"]))
        bg = g["background"]
        if bg is not None:
            bg = {l: bg[l] for l in g["alphabet"]}      # json sorted the keys; restore letters order
        pssm = oracle.pfm_to_pssm(counts, g["alphabet"], g["pseudocount"], bg)
        for letter in g["alphabet"]:
            assert same(pssm[letter], g["values"][letter]), (name, letter)


def _golden_table(golden_api, name, order):
    v = golden_api["pssm"][name]["values"]
    return np.array([v[l] for l in order], np.float64).T.copy()


def test_calculate_matches_golden(oracle, golden_api):
    for key, g in golden_api["calculate"].items():
        pname = key.split("|")[0]
        if g["dtype"] == "float32":
            got = oracle.seq_scores(g["seq"], _golden_table(golden_api, pname, "ACGU"))
            want = np.array(g["scores"], np.float32)
            assert got.dtype == np.float32
        else:
            got = oracle.alpha_scores(g["seq"], _golden_table(golden_api, pname, "BEHLMRT"), "BEHLMRT")
            want = np.array(g["scores"], np.float64)
        assert same(got, want), key


def test_survey_known_answers(oracle, golden_api):
    # float32 window scores of UUUUGCUCUGUAUAUA, uniform background
    got = oracle.seq_scores("UUUUGCUCUGUAUAUA", _golden_table(golden_api, "test_seq_uniform", "ACGU"))
    want = [-0.129, 1.276, -0.612, -0.866, -2.281, 0.629, 0.655, 0.48, -5.624, -0.277, -3.537, -0.206, -3.537]
    assert [float(oracle.round3(v)) for v in got] == pytest.approx(want, abs=1e-6)
    assert (oracle.search_hits(got, 0.0) + 1).tolist() == [2, 6, 7, 8]
    # K3: the single SLBP hit of the example at m = 6
    _, _, seq = oracle.parse_fasta(os.path.join(INP, "HIST2H3C_3p_end.fa"))[0]
    seq = oracle.preprocess_rna(seq)
    sc = oracle.seq_scores(seq, _golden_table(golden_api, "slbp_seq_uniform", "ACGU"))
    hits = oracle.search_hits(sc, 6.0)
    assert len(sc) == 219 and hits.tolist() == [212]
    assert seq[212:230] == "AAAGGCUCUUUUCAGAGC" and str(oracle.round3(sc[212])) == "14.259"


def _read_profile(path):
    import pandas as pd
    t = pd.read_csv(path, sep="\t")
    return t[list("BEHLMRT")].to_numpy(np.float64)


def test_averaged_matches_golden(oracle, golden_api):
    for key, g in golden_api["averaged"].items():
        tab = _golden_table(golden_api, g["pssm"], "BEHLMRT")
        prof = _read_profile(os.path.join(REPO, g["profile"]))
        with np.errstate(all="ignore"):
            sc = oracle.profile_scores(prof, tab)
            py = oracle.profile_scores_py(prof[:40], tab) if prof.shape[0] >= tab.shape[0] else np.zeros(0)
        assert same(sc[:len(py)], py)                       # C loop == numpy restatement
        rows = oracle.averaged_rows("m", sc, tab.shape[0], g["threshold"])
        assert [[r[1], r[2]] for r in rows] == [[r[0], r[1]] for r in g["rows"]], key
        assert same([r[4] for r in rows], [r[2] for r in g["rows"]]), key


def test_k4_structure_hits_of_the_example(oracle, golden_api):
    rows = golden_api["averaged"]["example_examplebg|6.0"]["rows"]
    assert [r[0] for r in rows] == [10, 212, 213, 214]
    assert rows[2][2] == pytest.approx(20.370341, abs=1e-5)


def test_c_port_equals_reference_compiled_pwm(oracle):
    ref = load_ref_pwm()
    rng = np.random.default_rng(7)
    letters = np.frombuffer(b"ACGUTacgutNnRY-", np.uint8)
    for trial in range(60):
        m = int(rng.integers(1, 30))
        n = int(rng.integers(0, 400))
        seq = letters[rng.integers(0, len(letters), size=n)].tobytes().decode()
        M = rng.normal(size=(m, 4)) * 3
        if trial % 7 == 0:
            M[rng.integers(0, m), rng.integers(0, 4)] = -np.inf
        if n - m + 1 <= 0:
            continue
        want = ref.calculate(seq, M)
        got = oracle.seq_scores(seq, M)
        assert want.dtype == np.float32 and same(got, want)
        assert same(oracle.seq_scores(seq, M, threads=True), want)


def test_reference_pwm_argument_errors():
    ref = load_ref_pwm()
    with pytest.raises(ValueError):
        ref.calculate("ACGU", np.zeros((4, 4), np.float32))
    with pytest.raises(ValueError):
        ref.calculate("ACGU", np.zeros((4, 5)))
    with pytest.raises(ValueError):
        ref.calculate("ACGU", np.zeros(4))


def test_search_semantics(oracle):
    s = np.array([1.0, np.nan, -np.inf, 6.0, 6.0000001, np.inf], np.float64)
    assert oracle.search_hits(s, 6.0).tolist() == [4, 5]
    assert oracle.search_hits(s, float("-inf")).tolist() == [0, 3, 4, 5]      # NaN and -inf never pass


def test_round3_follows_dtype(oracle):
    x = np.float32(0.6949999928474426)
    assert isinstance(oracle.round3(x), np.float32)
    assert oracle.round3(2.0005) == round(2.0005, 3)


def test_combine_rows(oracle):
    seq = [("a", 1, 4, np.float32(1.5)), ("a", 3, 6, np.float32(2.0)), ("b", 1, 4, np.float32(0.5))]
    st = [("a", 3, 6, 0.25), ("b", 2, 5, 1.0)]
    out = oracle.combine_rows(seq, st)
    assert out == [("a", 3, 6, np.float32(2.0), 0.25, 2.25)]
    assert math.isclose(out[0][5], 2.25)
