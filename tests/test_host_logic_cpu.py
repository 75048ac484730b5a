"""HOST logic of the CLI / scan API on a machine without a GPU: the device entry points are
replaced by the CPU oracle (tests/oracle_backend.py) so that packing, hit mapping, frame
assembly, messages and TSV text are compared with the reference's golden outputs.  The same
cases run through the real kernels in tests/test_cli_gpu.py (-m gpu)."""
import os
import warnings

import numpy as np

import pytest

import test_cli_gpu as gpu_cases
from oracle_backend import install

FUNCS = [getattr(gpu_cases, n) for n in dir(gpu_cases) if n.startswith("test_")
         and n not in ("test_cli_output_is_byte_identical", "test_cli_output_is_byte_identical_native_writer",
                       "test_pwm_module_signature_and_errors", "test_fasta_run_does_not_import_pandas")]


@pytest.mark.parametrize("name", gpu_cases.ALIGNED)
def test_cli_native_writer_against_golden(name, in_repo, monkeypatch):
    from rnascan_b200 import rnascan as ms
    install(monkeypatch)
    monkeypatch.setattr(ms, "NATIVE_WRITER_MIN_ROWS", 10 ** 12)       # the other way of printing: DataFrames
    _call(gpu_cases.test_cli_output_is_byte_identical, name=name, in_repo=in_repo)


@pytest.mark.parametrize("name", gpu_cases.ALIGNED)
def test_cli_host_logic_against_golden(name, in_repo, monkeypatch):
    install(monkeypatch)
    gpu_cases.test_cli_output_is_byte_identical.__wrapped__(name, in_repo) \
        if hasattr(gpu_cases.test_cli_output_is_byte_identical, "__wrapped__") else \
        _call(gpu_cases.test_cli_output_is_byte_identical, name=name, in_repo=in_repo)


def _call(fn, **available):
    import inspect
    names = inspect.signature(fn).parameters
    return fn(**{k: available[k] for k in names})


@pytest.mark.parametrize("fn", FUNCS, ids=[f.__name__ for f in FUNCS])
def test_api_host_logic_against_golden(fn, in_repo, golden_api, capsys, monkeypatch, tmp_path, oracle):
    install(monkeypatch)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _call(fn, in_repo=in_repo, golden_api=golden_api, capsys=capsys, monkeypatch=monkeypatch,
              tmp_path=tmp_path, oracle=oracle)


# ----------------------------------------------------------------------------- two ranks (gloo)
_RANK_WORKER = r"""
import contextlib, io, json, os, sys, warnings
sys.path.insert(0, %(repo)r)
sys.path.insert(0, os.path.join(%(repo)r, "tests"))
os.chdir(%(repo)r)
import oracle_backend
from rnascan_b200 import rnascan as ms, shard


class Patch(object):
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


oracle_backend.install(Patch())
rank, size = shard.init("gloo")
cases = json.load(open("tests/golden/cli/cases.json"))
extra = %(extra)r                     # name -> {"argv": [...], "want_file": path}: stdout only
for name, case in extra.items():
    cases[name] = {"argv": case["argv"], "stderr_lines": None, "want_file": case["want_file"]}
bad = []
for name in list(%(names)r) + sorted(extra):
    ms._BATCH_CACHE.clear()
    out, err = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            ms.main(list(cases[name]["argv"]))
        except SystemExit:
            pass
    ms.REFERENCE_COMPAT = False
    if rank == 0:
        want = open(cases[name].get("want_file") or "tests/golden/cli/%%s.stdout" %% name).read()
        lines = [l for l in err.getvalue().splitlines() if "seconds" not in l and "minutes" not in l]
        if out.getvalue() != want or (cases[name]["stderr_lines"] is not None and lines != cases[name]["stderr_lines"]):
            bad.append(name)
    elif out.getvalue() or err.getvalue():
        bad.append(name + ":rank%%d-wrote-output" %% rank)
import torch.distributed as dist
allbad = [None] * size
dist.all_gather_object(allbad, bad)
if rank == 0:
    print(json.dumps({"bad": [b for part in allbad for b in part], "size": size}))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("nproc", [2, 3])
def test_cli_sharded_over_ranks_matches_golden(tmp_path, nproc):
    """torchrun with 2 and 3 ranks (gloo): records are split over ranks (rec6 of mixed.fa is cut
    with an overlap), counts are all-reduced, rank 0 prints -- output identical to one process."""
    import json
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = ["rna_mixed_all", "rna_mixed_pc", "rna_bgonly", "rna_test_default", "ss_mixed_all", "ss_mixed_thr",
             "ss_bgonly", "rnass_fasta_all", "rnass_fasta_thr", "rna_empty_fasta", "rna_nohits",
             "rna_example_bg_all"]
    # profile directories (files split over ranks), motif collections, the reference's column pairing, and a
    # directory that holds a pack (every rank maps only its rows)
    import shutil
    golden = os.path.join(repo, "tests", "golden")
    extra = {}
    with open(os.path.join(golden, "cli", "multi_cases.json")) as fh:
        for name, case in json.load(fh).items():
            extra[name] = {"argv": case["argv"], "want_file": os.path.join(golden, "cli", name + ".stdout")}
    with open(os.path.join(golden, "cli", "cases.json")) as fh:
        mis = json.load(fh)["rnass_avg_example_misaligned"]
    extra["avg_compat"] = {"argv": mis["argv"] + ["--reference-compat"],
                           "want_file": os.path.join(golden, "cli", "rnass_avg_example_misaligned.stdout")}
    packed = tmp_path / "packed"
    shutil.copytree(os.path.join(golden, "inputs", "profiles_mixed"), packed)
    argv = list(extra["multi_ss_avg"]["argv"])
    argv[argv.index(os.path.join("tests", "golden", "inputs", "profiles_mixed"))] = str(packed)
    from rnascan_b200 import device, pack, rnascan as ms
    files = ms._profile_files(str(packed))
    rows, lengths = ms._read_profiles_packed(files)
    hp = device.HostProfile(rows)
    sep = np.zeros(rows.shape[0], np.uint8)
    sep[np.cumsum(lengths + 1) - 1] = 0xFF
    assert hp.make_q8(sep)
    pack.write(str(packed), files, rows, lengths, hp.stats(), hp.q8, hp.q8_scale)
    extra["multi_ss_avg_from_pack"] = {"argv": argv, "want_file": extra["multi_ss_avg"]["want_file"]}
    script = tmp_path / "rank_worker.py"
    script.write_text(_RANK_WORKER % {"repo": repo, "names": names, "extra": extra})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=%d" % nproc, "--master-addr", "127.0.0.1", "--master-port",
                          str(29540 + nproc), str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res == {"bad": [], "size": nproc}


_NOISY_WORKER = r"""
import json, os, sys, warnings
sys.path.insert(0, %(repo)r)
sys.path.insert(0, os.path.join(%(repo)r, "tests"))
os.chdir(%(repo)r)
import oracle_backend
from rnascan_b200 import rnascan as ms, shard


class Patch(object):
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


oracle_backend.install(Patch())
real_init = shard.init


def noisy_init(*a, **k):
    os.write(1, b"NCCL version 0.0.0 (a native library writing to file descriptor 1)\n")
    return real_init("gloo")


shard.init = noisy_init
warnings.simplefilter("ignore")
ms.main(%(argv)r)
"""


def test_pack_round_trips_both_quantised_forms(tmp_path):
    """rnascan_b200.pack version 2: float64 rows, the 8-byte and the 4-byte filter rows, all page aligned and
    mapped back bit for bit; a pack of another version reads as absent (the text is parsed again)."""
    from rnascan_b200 import pack
    rng = np.random.default_rng(5)
    lengths = np.array([5, 1, 9], np.int64)
    n = int((lengths + 1).sum())
    rows = rng.random((n, 7))
    q8 = rng.integers(0, 256, (n, 8), dtype=np.uint8)
    q4 = rng.integers(0, 256, (n, 4), dtype=np.uint8)
    d = tmp_path / "p"
    d.mkdir()
    names = ["structure.r%d.txt" % i for i in range(3)]
    pack.write(str(d), None, rows, lengths, (1.0, 0, 0, 1.0), q8, 0.5, names=names, q4=q4)
    pk = pack.read(str(d))
    assert pk.header["version"] == 2 and pk.q8_scale == 0.5
    for off in ("off_q8", "off_q4", "off_rows"):
        assert pk.header[off] % 4096 == 0 and pk.header[off] > 0
    assert np.array_equal(pk.rows, rows) and np.array_equal(pk.q8, q8) and np.array_equal(pk.q4, q4)
    pack.write(str(d), None, rows, lengths, (1.0, 0, 0, 1.0), None, None, names=names, q4=q4)
    pk = pack.read(str(d))
    assert pk.q8 is None and pk.q4 is None and np.array_equal(pk.rows, rows)
    raw = bytearray(open(pack.pack_path(str(d)), "rb").read())
    raw[:8] = b"RSB200P0"
    open(pack.pack_path(str(d)), "wb").write(bytes(raw))
    assert pack.read(str(d)) is None


def test_pack_mapping_is_cached_per_process_and_dropped_when_the_file_changes(tmp_path):
    from rnascan_b200 import pack
    rng = np.random.default_rng(6)
    lengths = np.array([4, 7], np.int64)
    n = int((lengths + 1).sum())
    rows = rng.random((n, 7))
    q8 = rng.integers(0, 256, (n, 8), dtype=np.uint8)
    q4 = rng.integers(0, 256, (n, 4), dtype=np.uint8)
    d = tmp_path / "p"
    d.mkdir()
    names = ["structure.a.txt", "structure.b.txt"]
    pack.write(str(d), None, rows, lengths, (1.0, 0, 0, 1.0), q8, 0.5, names=names, q4=q4)
    first = pack.read(str(d))
    assert pack.read(str(d)) is first                       # same mapping: no new page faults, no munmap
    assert first.page_locked("q4") is None                  # first request: never copied (one-shot runs)
    second = first.page_locked("q4")                        # second request: a pinned copy -- or None without CUDA
    assert second is None or (second.is_pinned() and np.array_equal(second.numpy(), q4))
    assert first.page_locked("q4") is second
    pack.write(str(d), None, rows + 1.0, lengths, (8.0, 0, 0, 2.0), q8, 0.5, names=names, q4=q4)
    fresh = pack.read(str(d))
    assert fresh is not first and np.array_equal(fresh.rows, rows + 1.0)
    os.remove(pack.pack_path(str(d)))
    assert pack.read(str(d)) is None


_NO_PANDAS_WORKER = r"""
import os, sys, warnings
sys.path.insert(0, %(repo)r)
sys.path.insert(0, os.path.join(%(repo)r, "tests"))
os.chdir(%(repo)r)
import oracle_backend
from rnascan_b200 import rnascan as ms


class Patch(object):
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


oracle_backend.install(Patch())
warnings.simplefilter("ignore")
try:
    ms.main(%(argv)r)
except SystemExit:
    pass
sys.stderr.write("PANDAS_LOADED=%%s\n" %% ("pandas" in sys.modules))
"""


@pytest.mark.parametrize("name", ["rna_mixed_all", "ss_mixed_thr", "rnass_fasta_all"])
def test_fasta_runs_never_import_pandas(name, tmp_path):
    """FASTA scans with hits read their PFMs without pandas (_read_pfm_counts) and print through the native writer:
    pandas is not even imported -- and stdout is the golden file.  (Host logic; the kernels are the oracle's.)"""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(repo, "tests", "golden", "cli", "cases.json")) as fh:
        argv = json.load(fh)[name]["argv"]
    script = tmp_path / "worker.py"
    script.write_text(_NO_PANDAS_WORKER % {"repo": repo, "argv": argv})
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    with open(os.path.join(repo, "tests", "golden", "cli", name + ".stdout")) as fh:
        assert out.stdout == fh.read()
    assert "PANDAS_LOADED=False" in out.stderr, out.stderr[-800:]


def test_native_library_output_stays_out_of_hits_tab_under_torchrun(tmp_path):
    """NCCL prints its version banner on file descriptor 1 while the process group comes up; under torchrun
    the CLI points fd 1 at stderr for the run and writes hits.tab to a private duplicate of the real stdout
    (found by timing the sharded CLI on 8 GPUs: the first line of hits.tab was the banner)."""
    import json
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(repo, "tests", "golden", "cli", "cases.json")) as fh:
        case = json.load(fh)["rna_mixed_all"]
    script = tmp_path / "noisy_worker.py"
    script.write_text(_NOISY_WORKER % {"repo": repo, "argv": case["argv"]})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29561", str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    with open(os.path.join(repo, "tests", "golden", "cli", "rna_mixed_all.stdout")) as fh:
        assert out.stdout == fh.read()
    assert out.stderr.count("a native library writing to file descriptor 1") == 2
