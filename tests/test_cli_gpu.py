"""Drop-in parity of the scan API and the CLI on the GPU: stdout of ``rnascan`` (hits.tab,
--bgonly dicts) must be byte-identical to what the reference's own main() printed for the same
arguments (tests/golden/cli/*.stdout, produced by tests/golden/make_golden.py), stderr
messages included; API-level results must equal the golden vectors of the reference's
functions."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "tests", "golden", "cli")
INP = os.path.join(REPO, "tests", "golden", "inputs")

with open(os.path.join(CLI, "cases.json")) as _fh:
    CASES = json.load(_fh)
ALIGNED = sorted(k for k, v in CASES.items() if v["group"] != "misaligned")


def run_cli(argv):
    from rnascan_b200 import rnascan as ms
    out, err = io.StringIO(), io.StringIO()
    code = 0
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        try:
            ms.main(list(argv))
        except SystemExit as e:
            code = e.code or 0
    lines = [l for l in err.getvalue().splitlines() if "seconds" not in l and "minutes" not in l]
    return out.getvalue(), lines, code


def same(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and \
        np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


@pytest.mark.parametrize("name", ALIGNED)
def test_cli_output_is_byte_identical_native_writer(name, in_repo, monkeypatch):
    """Same golden files with the OTHER way of printing.  main() formats FASTA results through the C++ hits.tab
    formatter (rs_host_format_hits*) from NATIVE_WRITER_MIN_ROWS = 1 row on; here the threshold is out of reach,
    so every result is assembled as DataFrames and printed by pandas, as the reference does."""
    from rnascan_b200 import rnascan as ms
    monkeypatch.setattr(ms, "NATIVE_WRITER_MIN_ROWS", 10 ** 12)
    test_cli_output_is_byte_identical(name, in_repo)


def test_directory_results_print_natively_without_building_a_frame(tmp_path, in_repo, monkeypatch):
    for order in ("hit_first", "none_first"):
        _directory_native_case(order, tmp_path / order, monkeypatch)


def _directory_native_case(order, tmp_path, monkeypatch):
    """Structure-only scan of a profile directory: when the first file has a hit the text pandas would print for the
    reference's concatenated per-file frames -- Start / End as FLOATS as soon as some file has no hit -- comes from
    the native writer and no DataFrame is built; with a first file without hits (another column order) the frames do
    the printing.  Same bytes either way.  The combined FASTA + directory mode never builds the structure frame."""
    import shutil
    from rnascan_b200 import rnascan as ms
    monkeypatch.setattr(ms, "NATIVE_WRITER_MIN_ROWS", 1)
    work = tmp_path / "profiles"
    os.makedirs(work)
    example = os.path.join(INP, "profiles_example", "structure.hg19_dna.txt")
    none = sorted(os.listdir(os.path.join(INP, "profiles_mixed")))[0]
    shutil.copy(example, work / "structure.a_hit.txt")
    shutil.copy(os.path.join(INP, "profiles_mixed", none), work / "structure.b_none.txt")
    shutil.copy(example, work / "structure.c_hit.txt")
    listing = sorted(str(work / f) for f in os.listdir(work))
    if order == "none_first":
        listing = [listing[1], listing[0], listing[2]]
    monkeypatch.setattr(ms, "_profile_files", lambda directory: list(listing))
    built = []
    real = ms._averaged_dir_frame
    monkeypatch.setattr(ms, "_averaged_dir_frame", lambda *a, **k: built.append(1) or real(*a, **k))
    argv = ["-q", os.path.join(INP, "SLBP_pfm_assembled_normalized_struct.txt"), "-B",
            os.path.join(INP, "bg_struct_example.txt"), "-m", "2", str(work)]
    native, err1, code = run_cli(argv)
    assert code == 0 and native.count("\n") == 1 + 16
    assert len(built) == (0 if order == "hit_first" else 1)
    monkeypatch.setattr(ms, "NATIVE_WRITER_MIN_ROWS", 10 ** 12)
    frames, err2, _ = run_cli(argv)
    assert native == frames and err1 == err2
    if order == "hit_first":
        first = native.splitlines()[1].split("\t")
        assert first[0] == "a_hit" and first[3] == "10.0" and first[4] == "27.0"     # float Start / End
    else:
        assert native.splitlines()[0].split("\t")[:2] == ["Sequence", "LogOdds"]      # the other column order
    # combined mode over the same directory: the joint writer prints, the structure frame is never built
    monkeypatch.setattr(ms, "NATIVE_WRITER_MIN_ROWS", 1)
    monkeypatch.setattr(ms, "_profile_files", lambda directory: [str(work / "structure.a_hit.txt")])
    shutil.copy(example, work / "structure.hg19_dna.txt")
    monkeypatch.setattr(ms, "_profile_files", lambda directory: [str(work / "structure.hg19_dna.txt")])
    del built[:]
    cargv = ["-p", os.path.join(INP, "SLBP_pfm_assembled_normalized_seq.txt"), "-q",
             os.path.join(INP, "SLBP_pfm_assembled_normalized_struct.txt"), "-u", "-m", " -inf",
             os.path.join(INP, "HIST2H3C_3p_end.fa"), str(work)]
    joint, _, code = run_cli(cargv)
    assert code == 0 and joint.count("\n") > 3 and not built
    monkeypatch.setattr(ms, "NATIVE_WRITER_MIN_ROWS", 10 ** 12)
    assert run_cli(cargv)[0] == joint and built


def test_fasta_run_does_not_import_pandas(in_repo):
    """A FASTA scan reads its PFM without pandas and prints through the native writer: pandas (1-3 s of import)
    is never loaded -- and the output is the golden one."""
    import subprocess
    import sys
    name = "rna_mixed_all"
    code = ("import sys, warnings; warnings.simplefilter('ignore'); from rnascan_b200 import rnascan as ms\n"
            "try:\n    ms.main(%r)\nexcept SystemExit:\n    pass\n"
            "sys.stderr.write('PANDAS_LOADED=%%s\\n' %% ('pandas' in sys.modules))\n" % (CASES[name]["argv"],))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    with open(os.path.join(CLI, name + ".stdout")) as fh:
        assert out.stdout == fh.read()
    assert "PANDAS_LOADED=False" in out.stderr, out.stderr[-500:]


@pytest.mark.parametrize("name", ALIGNED)
def test_cli_output_is_byte_identical(name, in_repo):
    case = CASES[name]
    with open(os.path.join(CLI, name + ".stdout")) as fh:
        want = fh.read()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out, err_lines, code = run_cli(case["argv"])
    assert code == case["exit"]
    assert out == want
    assert err_lines == case["stderr_lines"]


with open(os.path.join(CLI, "multi_cases.json")) as _fh:
    MULTI_CASES = json.load(_fh)


def test_motif_collections_print_what_one_reference_run_per_motif_prints(in_repo):
    """-p / -q naming multi-PFM files (pfmutil.write_multi_pfm layout): stdout is the concatenation of the
    reference's stdout for one run per motif pair (golden files made by looping the reference's own main(),
    tests/golden/make_golden_multi.py).  Directories of averaged profiles go through ONE batched scan."""
    import warnings
    from rnascan_b200 import rnascan as ms
    for name, case in sorted(MULTI_CASES.items()):
        with open(os.path.join(CLI, name + ".stdout")) as fh:
            want = fh.read()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out, _, code = run_cli(case["argv"])
        ms.REFERENCE_COMPAT = False
        assert code == case["exit"], name
        assert out == want, name
        assert out.count("Match_ID") == case["runs"], name


def test_scan_many_equals_scan_main_per_motif(in_repo):
    """scan_many(directory, motifs) == [scan_main(directory, motif) for every motif] (frames), and with
    sequence motifs the windows are restricted to those whose sequence score passes as well."""
    import argparse
    from rnascan_b200 import rnascan as ms
    from rnascan_b200.BioAddons.Alphabet import ContextualSecondaryStructure
    structure, rna = ContextualSecondaryStructure(), ms.IUPAC.IUPACUnambiguousRNA()
    directory = os.path.join(INP, "profiles_mixed")
    bg = {"B": 0.0163, "E": 0.2721, "H": 0.1530, "L": 0.2046, "M": 0.0196, "R": 0.1970, "T": 0.1374}
    blocks_q = ms.pfm_blocks(os.path.join(INP, "multi", "multi_struct.pfm"))
    blocks_p = ms.pfm_blocks(os.path.join(INP, "multi", "multi_seq.pfm"))
    assert len(blocks_q) == len(blocks_p) == 8 and ms.pfm_blocks(os.path.join(INP, "test_seq_pfm.txt")) is None
    struct_pssms = [ms.load_motif(b, 0.01, structure, bg) for b in blocks_q]
    frames = ms.scan_many(directory, struct_pssms, 0.5)
    ns = argparse.Namespace(minscore=0.5, debug=False)
    for pssm, frame in zip(struct_pssms, frames):
        want = ms.scan_main(directory, pssm, structure, None, ns)
        assert list(frame.columns) == list(want.columns) and len(frame) == len(want) > 0
        for col in want.columns:
            assert frame[col].tolist() == want[col].tolist()
    seq_pssms = [ms.load_motif(b, 0.01, rna, None) for b in blocks_p]
    both = ms.scan_many(directory, struct_pssms, -6.0, seq_file=os.path.join(INP, "mixed.fa"), seq_pssms=seq_pssms)
    alone = ms.scan_many(directory, struct_pssms, -6.0)
    assert sum(len(f) for f in both) > 0
    for a, b in zip(both, alone):
        assert "LogOdds.Seq" in a.columns and len(a) <= len(b) and (a["LogOdds.Seq"] > -6.0).all()
        keys = set(zip(b["Sequence_ID"], b["Start"]))
        assert all(k in keys for k in zip(a["Sequence_ID"], a["Start"]))


def test_generated_cases_print_what_the_reference_printed(tmp_path, in_repo):
    """80-odd randomly generated CLI runs (all modes incl. -t, thresholds from -inf up, zero cells, odd records, shuffled
    PFM headers, --bgonly): inputs are regenerated from their seeds, stdout must equal what the REFERENCE printed
    for them in the build container (tests/golden/fuzz, made by tests/golden/make_golden_fuzz.py)."""
    import warnings
    import test_differential_cpu as diff
    from rnascan_b200 import rnascan as ms
    fuzz = os.path.join(REPO, "tests", "golden", "fuzz")
    with open(os.path.join(fuzz, "cases.json")) as fh:
        cases = json.load(fh)
    assert len(cases) >= 60
    for seed, case in sorted(cases.items(), key=lambda kv: int(kv[0])):
        root = str(tmp_path / seed)
        paths = diff.draw_inputs(np.random.default_rng(9000 + int(seed)), root)
        argv = [a.format(**paths) if a.startswith("{") else a for a in case["argv"]]
        ms._BATCH_CACHE.clear()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out, _, code = run_cli(argv)
        ms.REFERENCE_COMPAT = False
        with open(os.path.join(fuzz, seed + ".stdout")) as fh:
            want = fh.read()
        assert code == case["exit"], (seed, case["mode"], argv)
        assert out == want, (seed, case["mode"], argv)


def test_cores_flag_does_not_change_the_result(in_repo):
    base = CASES["rna_mixed_all"]["argv"]
    a = run_cli(base)[0]
    b = run_cli([x for x in base if x not in ("-c", "2")] + ["-c", "7"])[0]
    assert a == b


def test_averaged_cli_is_label_aligned_not_the_py3_misaligned_variant(in_repo):
    """on python >= 3.6 the reference pairs profile columns B,E,H,L,M,R,T with PSSM
    columns E,H,T,B,L,R,M positionally.  The canonical semantics here are label-aligned; the
    misaligned output is kept as a golden file only to document the divergence."""
    case = CASES["rnass_avg_example_misaligned"]
    out, _, code = run_cli(case["argv"])
    with open(os.path.join(CLI, "rnass_avg_example_misaligned.stdout")) as fh:
        mis = fh.read()
    assert code == 0 and out != mis
    rows, mrows = out.splitlines(), mis.splitlines()
    assert rows[0] == mrows[0] and len(rows) == len(mrows) == 220
    cols = rows[0].split("\t")
    i_seq, i_str, i_start = cols.index("LogOdds.Seq"), cols.index("LogOdds.Struct"), cols.index("Start")
    best = max(rows[1:], key=lambda r: float(r.split("\t")[i_str])).split("\t")
    assert best[i_start] == "213" and float(best[i_str]) > 10.0              # the SLBP stem-loop
    # sequence column is unaffected by the alignment question
    assert [r.split("\t")[i_seq] for r in rows[1:]] == [r.split("\t")[i_seq] for r in mrows[1:]]


def test_reference_compat_reproduces_the_py3_reference_byte_for_byte(in_repo):
    """--reference-compat pairs profile and PSSM columns by position, as the unmodified reference does on
    Python >= 3.6: stdout equals what the reference's own main() printed (golden file)."""
    case = CASES["rnass_avg_example_misaligned"]
    out, err_lines, code = run_cli(list(case["argv"]) + ["--reference-compat"])
    with open(os.path.join(CLI, "rnass_avg_example_misaligned.stdout")) as fh:
        want = fh.read()
    assert code == case["exit"] and out == want
    assert err_lines == case["stderr_lines"]
    from rnascan_b200 import rnascan as ms
    ms.REFERENCE_COMPAT = False


def test_averaged_combined_native_writer_equals_dataframe_path(in_repo, monkeypatch):
    """FASTA + profile directory (mode RNASS, averaged): the array/native-writer path and the
    DataFrame + merge path print the same bytes, at -m -inf and at a threshold."""
    from rnascan_b200 import rnascan as ms
    base = CASES["rnass_avg_example_misaligned"]["argv"]
    for extra in ([], ["-m", "-3"]):
        argv = [a for a in base if a not in ("-m", " -inf")] + (extra or ["-m", " -inf"])
        monkeypatch.setattr(ms, "NATIVE_WRITER_MIN_ROWS", 10 ** 9)
        frames = run_cli(argv)[0]
        monkeypatch.setattr(ms, "NATIVE_WRITER_MIN_ROWS", 0)
        native = run_cli(argv)[0]
        assert frames == native and frames.count("\n") > 5


def test_profile_pack_gives_identical_output(tmp_path, in_repo):
    for mode in ("struct", "combined"):
        _pack_case(tmp_path / mode, mode)


def _pack_case(tmp_path, mode):
    """--pack leaves rnascan_b200.pack in the profile directory; scans of the unchanged directory map it
    (quantised filter rows + exact float64 rows) and print the same bytes; a changed file invalidates it; a
    pack can stand for the text files altogether."""
    import shutil
    from rnascan_b200 import pack
    os.makedirs(tmp_path)
    work = tmp_path / "profiles"
    shutil.copytree(os.path.join(INP, "profiles_mixed"), work)
    shutil.copy(os.path.join(INP, "profiles_example", "structure.hg19_dna.txt"), work / "structure.hg19_dna.txt")
    argv = ["-q", os.path.join(INP, "SLBP_pfm_assembled_normalized_struct.txt"), "-B",
            os.path.join(INP, "bg_struct_example.txt"), "-m", "2"]
    if mode == "combined":
        argv += ["-p", os.path.join(INP, "SLBP_pfm_assembled_normalized_seq.txt"), "-b",
                 os.path.join(INP, "bg_seq_custom.txt"), os.path.join(INP, "HIST2H3C_3p_end.fa")]
        argv[argv.index("2")] = " -2"
    argv += [str(work)]
    plain, err0, code = run_cli(argv)
    assert code == 0 and plain.count("\n") > 3
    assert not os.path.exists(pack.pack_path(str(work)))
    written, _, _ = run_cli(argv + ["--pack"])
    pk = pack.read(str(work))
    assert written == plain and pk is not None and pk.q8 is not None
    assert sorted(pk.names) == sorted(f for f in os.listdir(work) if f != pack.NAME)
    mapped, err1, _ = run_cli(argv)
    assert mapped == plain and err1 == err0
    # a touched file makes the pack stale: the text is parsed again (same output), --pack refreshes it
    victim = work / "structure.rec2.txt"
    os.utime(victim, ns=(1, 1))
    assert not pack.matches(pack.read(str(work)), [str(work / n) for n in pk.names])
    assert run_cli(argv + ["--pack"])[0] == plain
    assert pack.matches(pack.read(str(work)), [str(work / n) for n in pack.read(str(work)).names])
    # the pack alone
    for f in os.listdir(work):
        if f != pack.NAME:
            os.remove(work / f)
    alone, _, code = run_cli(argv)
    assert code == 0 and alone == plain


def test_pack_scan_picks_the_smallest_filter_form_the_threshold_allows(tmp_path, in_repo):
    """A pack holds the 8-byte and the 4-byte filter rows.  The 4-byte form goes to the device when its guard band
    is at most 0.4 of a positive threshold, the 8-byte form otherwise; the output does not depend on the choice."""
    import shutil
    from rnascan_b200 import pack
    work = tmp_path / "profiles"
    shutil.copytree(os.path.join(INP, "profiles_mixed"), work)
    # a motif close to the background: few, small positive log-odds, so the 4-bit form's band (scale / 15 times the
    # sum of the positive entries, about 0.1 here) is well below 0.4 x 0.3
    import ast
    bg = ast.literal_eval(open(os.path.join(INP, "bg_struct_example.txt")).read())
    pfm = tmp_path / "near_bg.txt"
    with open(pfm, "w") as fh:
        fh.write("PO\t" + "\t".join("BEHLMRT") + "\n")
        for w, up in enumerate(("E", "H", "L")):
            row = {c: bg[c] * (1.12 if c == up else 1.0) for c in "BEHLMRT"}
            tot = sum(row.values())
            fh.write("%d\t%s\n" % (w, "\t".join(repr(row[c] / tot) for c in "BEHLMRT")))
    base = ["-q", str(pfm), "-B", os.path.join(INP, "bg_struct_example.txt"), "-C", "0"]
    forms = {}
    for thr in (" -3", "0.1", "0.3"):
        argv = base + ["-m", thr, str(work)]
        plain = run_cli(argv)[0]
        run_cli(argv + ["--pack"])
        pk = pack.read(str(work))
        assert pk.q4 is not None and pk.q4.shape == (pk.q8.shape[0], 4)
        stats = tmp_path / "stats.json"
        mapped = run_cli(argv + ["--stats", str(stats)])[0]
        assert mapped == plain
        st = json.loads(stats.read_text().splitlines()[-1])
        assert st["profile_source"] == "pack"
        forms[thr] = st["profile_filter_form"]
    assert forms[" -3"] == "q8" and forms["0.3"] == "q4", forms


def test_stats_option_writes_one_json_line(tmp_path, in_repo):
    """--stats FILE: phases, sizes and throughput of the run as one JSON line; stdout is unchanged."""
    base = CASES["rna_mixed_all"]["argv"]
    plain = run_cli(base)[0]
    path = tmp_path / "stats.json"
    from rnascan_b200 import rnascan as ms
    ms._BATCH_CACHE.clear()                            # a cached input is not parsed again: no ingest phase
    out, _, code = run_cli(list(base) + ["--stats", str(path)])
    assert code == 0 and out == plain
    lines = path.read_text().splitlines()
    assert len(lines) == 1
    st = json.loads(lines[0])
    assert st["mode"] == "RNA" and st["records"] > 0 and st["symbols"] > 0
    assert 0 < st["scored_positions"] <= st["symbols"]
    assert st["total_s"] > 0 and "scan_s" in st["phases_s"] and "ingest_fasta_s" in st["phases_s"]
    argv = ["-q", os.path.join(INP, "SLBP_pfm_assembled_normalized_struct.txt"), "-B",
            os.path.join(INP, "bg_struct_example.txt"), "-m", "2", os.path.join(INP, "profiles_mixed"),
            "--stats", str(path)]
    assert run_cli(argv)[2] == 0
    st = json.loads(path.read_text().splitlines()[1])
    assert st["mode"] == "SS" and st["profile_rows"] > 0 and st["profile_source"] == "text"


def test_every_position_structure_scan_through_the_cli(tmp_path, in_repo, oracle):
    """BASELINE config 3 in small: one-hot structure FASTA, -m -inf, every scorable window is a row of
    hits.tab (native writer); rows are rebuilt here from the oracle's dense scores."""
    from rnascan_b200 import synth
    from rnascan_b200 import rnascan as ms
    rng = np.random.default_rng(33)
    lengths = synth.record_lengths(60_000, 40, rng)
    codes, off = synth.struct_codes(lengths, rng)
    text = synth.to_text(codes, "struct").decode()
    recs = text.split("\n")[:-1]
    fa = tmp_path / "contexts.fa"
    with open(fa, "w") as fh:
        for k, r in enumerate(recs):
            fh.write(">s%d structure %d\n" % (k, k))
            for a in range(0, len(r), 70):
                fh.write(r[a:a + 70] + "\n")
    pfm = os.path.join(INP, "test_struct_pfm.txt")
    out, err, code = run_cli(["-q", pfm, "-u", "-C", "0.01", "-m", " -inf", str(fa)])
    assert code == 0 and "Processed 40 sequences" in err
    from rnascan_b200.BioAddons.Alphabet import ContextualSecondaryStructure
    pm = ms.pfm2pssm(pfm, 0.01, ContextualSecondaryStructure(), None)
    tab = np.array([pm[c] for c in "BEHLMRT"]).T.copy()
    W = tab.shape[0]
    want = ["Sequence_ID\tDescription\tMotif_ID\tStart\tEnd\tSequence\tLogOdds\tMatch_ID"]
    k = 0
    for r_i, r in enumerate(recs):
        sc = oracle.alpha_scores(r, tab, "BEHLMRT")
        for i, v in enumerate(sc.tolist()):
            if v > float("-inf"):
                k += 1
                want.append("s%d\ts%d structure %d\ttest_struct_pfm\t%d\t%d\t%s\t%r\t%d"
                            % (r_i, r_i, r_i, i + 1, i + W, r[i:i + W], round(v, 3), k))
    assert k > 50_000
    assert out.splitlines() == want


# ----------------------------------------------------------------------------- API level
def _pssm(golden_api, name):
    from rnascan_b200 import rnascan as ms
    from rnascan_b200.seq import IUPAC
    from rnascan_b200.BioAddons.Alphabet import ContextualSecondaryStructure
    g = golden_api["pssm"][name]
    alpha = IUPAC.IUPACUnambiguousRNA() if g["alphabet"] == "GAUC" else ContextualSecondaryStructure()
    bg = g["background"]
    if bg is not None:
        bg = {l: bg[l] for l in g["alphabet"]}
    return ms.pfm2pssm(g["file"], g["pseudocount"], alpha, bg), alpha


def test_calculate_matches_the_reference_class(golden_api, in_repo):
    for key, g in golden_api["calculate"].items():
        pm, _ = _pssm(golden_api, key.split("|")[0])
        r = pm.calculate(g["seq"])
        if len(g["scores"]) == 1:
            assert np.ndim(r) == 0                       # scalar when there is one window
        r = np.atleast_1d(np.asarray(r))
        if g["dtype"] == "float32":
            assert r.dtype == np.float32 or len(r) == 0
        assert same(r, g["scores"]), key


def test_pwm_module_signature_and_errors(in_repo):
    from rnascan_b200.BioAddons.motifs import _pwm
    out = _pwm.calculate("ACGUN", np.zeros((2, 4)))
    assert out.dtype == np.float32 and out.shape == (4,) and np.isnan(out[3])
    assert _pwm.calculate(sequence="AC", matrix=np.ones((2, 4)))[0] == 2.0
    with pytest.raises(ValueError):
        _pwm.calculate("ACGU", np.zeros((4, 4), np.float32))
    with pytest.raises(ValueError):
        _pwm.calculate("ACGU", np.zeros((4, 5)))
    with pytest.raises(ValueError):
        _pwm.calculate("ACGU", np.zeros(4))
    with pytest.raises(TypeError):
        _pwm.calculate(1234, np.zeros((4, 4)))


def test_compute_background_matches_the_reference(golden_api, in_repo, capsys):
    from rnascan_b200 import rnascan as ms
    from rnascan_b200.seq import IUPAC
    from rnascan_b200.BioAddons.Alphabet import ContextualSecondaryStructure
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bg = ms.compute_background([os.path.join(INP, "test.fa")], IUPAC.IUPACUnambiguousRNA(), verbose=False)
        assert dict(bg) == golden_api["bg_test_fa"] and list(bg) == list("GAUC")
        # the reference's own known-answer test, tests/motif_scan_test.py:33-43
        for k, v in {"A": 0.1944, "C": 0.1388, "U": 0.5277, "G": 0.1388}.items():
            assert abs(bg[k] - v) < 1e-3
        bg = ms.compute_background(os.path.join(INP, "mixed_struct.fa"), ContextualSecondaryStructure(), False)
        assert dict(bg) == golden_api["bg_mixed_struct_fa"] and list(bg) == list("EHTBLRM")
    capsys.readouterr()


def test_scan_averaged_structure_matches_the_reference(golden_api, in_repo):
    from rnascan_b200 import rnascan as ms
    for key, g in golden_api["averaged"].items():
        pm, _ = _pssm(golden_api, g["pssm"])
        df = ms.scan_averaged_structure(g["profile"], {"m": pm}, g["threshold"])
        rows = [] if df.shape[0] == 0 else [[int(r.Start), int(r.End), float(r.LogOdds)] for r in df.itertuples()]
        assert [r[:2] for r in rows] == [r[:2] for r in g["rows"]], key
        assert same([r[2] for r in rows], [r[2] for r in g["rows"]]), key
        if rows:
            assert list(df.columns) == ["Motif_ID", "Start", "End", "Sequence", "LogOdds"]
            assert set(df["Sequence"]) == {"."} and set(df["Motif_ID"]) == {"m"}


def test_combined_example_matches_the_reference(golden_api, in_repo, capsys):
    """K5: combine() of the example's sequence hits (uniform background) and label-aligned
    averaged-structure hits at m = 6 -- one row, Start 213."""
    from rnascan_b200 import rnascan as ms
    seq_pm, rna = _pssm(golden_api, "slbp_seq_uniform")
    st_pm, _ = _pssm(golden_api, "slbp_struct_examplebg")

    class Args(object):
        minscore, debug, cores = 6.0, False, 1
    seq_df = ms.scan_main(os.path.join(INP, "HIST2H3C_3p_end.fa"),
                          {"SLBP_pfm_assembled_normalized_seq": seq_pm}, rna, None, Args())
    st_df = ms.scan_main(os.path.join(INP, "profiles_example"),
                         {"SLBP_pfm_assembled_normalized_struct": st_pm}, None, None, Args())
    comb = ms.combine(seq_df, st_df)
    ms._add_match_id(comb)
    buf = io.StringIO()
    comb.to_csv(buf, sep="\t", index=False)
    assert buf.getvalue() == golden_api["combine_example_aligned_tsv"]
    capsys.readouterr()


def test_pwm_scan_fwd_matches_the_reference(golden_api):
    from rnascan_b200 import pfmutil as pu
    V = golden_api["pfmutil"]
    got = pu.pwm_scan_fwd(V["to_pwm_seq_20"], V["scan_fwd_seq"]["seq"])
    assert same(got, V["scan_fwd_seq"]["scores"]) and isinstance(got[0], float)
    got = pu.pwm_scan_fwd(V["scan_fwd_struct"]["pwm"], V["scan_fwd_struct"]["seq"])
    assert same(got, V["scan_fwd_struct"]["scores"])
    with pytest.raises(KeyError):
        pu.pwm_scan_fwd(V["to_pwm_seq_20"], "ACGUNACGU")
    assert pu.pwm_scan_fwd(V["to_pwm_seq_20"], "AC") == []


def test_search_generator_semantics(golden_api, in_repo):
    pm, _ = _pssm(golden_api, "test_seq_uniform")
    hits = list(pm.search("UUUUGCUCUGUAUAUA", threshold=0.0, both=False))
    assert [p for p, _ in hits] == [1, 5, 6, 7]
    assert all(isinstance(s, np.float32) for _, s in hits)
    assert list(pm.search("ACG", threshold=-100.0, both=False)) == []
    # NaN / -inf windows are never reported, even at -inf
    hits = list(pm.search("ACGUNACGU", threshold=float("-inf"), both=False))
    assert [p for p, _ in hits] == [0, 5]
