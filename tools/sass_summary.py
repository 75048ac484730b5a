"""SASS evidence: per kernel of librnascan_b200.so, how often the Blackwell-specific mnemonics occur.
    python tools/sass_summary.py > profiles/sass_summary.txt         (no GPU needed: cuobjdump -sass)
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine, 1-D),
SYNCS = mbarrier operations, F64 = DADD/DFMA/DMUL."""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rnascan_b200", "librnascan_b200.so")
text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "F64", "instructions"]
pat = {"UTCHMMA": re.compile(r"\bUTCHMMA"), "LDTM": re.compile(r"\bLDTM"), "UTCBAR": re.compile(r"\bUTCBAR"),
       "UBLKCP": re.compile(r"\bUBLKCP"), "SYNCS": re.compile(r"\bSYNCS"), "F64": re.compile(r"\b(DADD|DFMA|DMUL)\b")}
counts, fn = collections.OrderedDict(), None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn and re.search(r"/\*[0-9a-f]{4}\*/", line):
        counts[fn]["instructions"] += 1
        for k, p in pat.items():
            if p.search(line):
                counts[fn][k] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass rnascan_b200/librnascan_b200.so (sm_100a): Blackwell-specific mnemonics per kernel family")
print("# (all template instantiations of a family added up; `inst.` = how many instantiations)")
print("# %-34s %6s %s" % ("kernel family", "inst.", " ".join("%8s" % k for k in keys)))
fam, tot = collections.OrderedDict(), collections.Counter()
for (fn, c), name in zip(counts.items(), demangled):
    base = re.sub(r"^void ", "", re.sub(r"[<(].*", "", name))
    fam.setdefault(base, [0, collections.Counter()])
    fam[base][0] += 1
    fam[base][1].update(c)
    tot.update(c)
for base, (n, c) in fam.items():
    print("%-36s %6d %s" % (base, n, " ".join("%8d" % c[k] for k in keys)))
print("%-36s %6d %s" % ("TOTAL", len(counts), " ".join("%8d" % tot[k] for k in keys)))
print()
print("# the W = 7 instantiations the benchmark configurations run")
for (fn, c), name in zip(counts.items(), demangled):
    name = re.sub(r"\(.*", "", name)
    if re.search(r"fused_filter_kernel<7>|filter_q8_kernel<7>|kmer_scan_kernel<7>|dense_w_kernel<7, 7, [01]>|batched_tc_kernel|"
                 r"resolve_kernel|hist_kernel<2>|profile_exact_kernel|kmer_finish_kernel<4>|order_kernel", name):
        print("%-48s %s" % (name[:48], " ".join("%8d" % c[k] for k in keys)))
