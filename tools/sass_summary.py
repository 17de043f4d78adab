"""Per-kernel SASS opcode summary of libadc_b200.so (cuobjdump -sass): counts of the Blackwell-specific instructions that show which
hardware path each kernel uses -- UTCHMMA (tcgen05.mma), LDTM / STTM (tensor-memory loads / stores), UTMALDG / UTMASTG / UTMAREDG (TMA bulk
tensor load / store / reduce-add), UTCBAR (tcgen05.commit), SYNCS (mbarrier), FFMA2 (packed fp32 FMA), HMMA (legacy mma.sync: must be 0).
Usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "aerial_depth_completion_b200", "libadc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "FFMA2", "FFMA", "HMMA", "DFMA")
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[kern]["_n"] += 1
        for o in OPS:
            if op == o or (o != "FFMA" and op.startswith(o)):
                counts[kern][o] += 1
                total[o] += 1
                break
demangle = subprocess.run(["cu++filt"] + list(counts.keys()), capture_output=True, text=True).stdout.splitlines() if counts else []
names = dict(zip(counts.keys(), demangle)) if len(demangle) == len(counts) else {k: k for k in counts}
print("# cuobjdump -sass aerial_depth_completion_b200/libadc_b200.so (sm_100a): %d kernels" % len(counts))
print("# totals: " + "  ".join("%s %d" % (o, total[o]) for o in OPS))
print("%-92s %7s " % ("kernel", "instr") + " ".join("%8s" % o for o in OPS))
for k, c in sorted(counts.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000000 + kv[1]["UTMALDG"] * 1000 + kv[1]["FFMA2"])):
    nm = re.sub(r"\(.*", "", names[k].replace("(int)", "").replace("(bool)", "")).replace("void ", "")[:92]
    print("%-92s %7d " % (nm, c["_n"]) + " ".join("%8d" % c[o] for o in OPS))
