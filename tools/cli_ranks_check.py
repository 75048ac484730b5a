"""Run golden CLI cases under torchrun on real GPUs (one rank per GPU) and compare rank 0's
stdout with the reference's golden output.  Usage: torchrun --nproc-per-node N tools/cli_ranks_check.py"""
import contextlib, io, json, os, sys, warnings
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.chdir(REPO)
from rnascan_b200 import rnascan as ms, shard

rank, size = shard.init()
cases = json.load(open("tests/golden/cli/cases.json"))
names = ["rna_mixed_all", "rna_mixed_pc", "rna_bgonly", "rna_test_default", "ss_mixed_all", "ss_mixed_thr",
         "ss_bgonly", "rnass_fasta_all", "rnass_fasta_thr", "rna_empty_fasta", "rna_nohits", "rna_example_bg_all"]
multi = json.load(open("tests/golden/cli/multi_cases.json"))
cases.update(multi)
names += sorted(multi) + ["rnass_avg_example_misaligned+compat"]
cases["rnass_avg_example_misaligned+compat"] = {"argv": cases["rnass_avg_example_misaligned"]["argv"] + ["--reference-compat"]}
bad = []
for name in names:
    ms._BATCH_CACHE.clear()
    out, err = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            ms.main(list(cases[name]["argv"]))
        except SystemExit:
            pass
    if rank == 0:
        want = open("tests/golden/cli/%s.stdout" % name.split("+")[0]).read()
        ms.REFERENCE_COMPAT = False
        if out.getvalue() != want:
            bad.append(name)
import torch.distributed as dist
allbad = [None] * size
dist.all_gather_object(allbad, bad)
if rank == 0:
    print(json.dumps({"cli_ranks": size, "cases": len(names), "bad": [b for p in allbad for b in p]}))
dist.destroy_process_group()
