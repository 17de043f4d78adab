"""Per-instruction view of one kernel launch of an `ncu --set full --import-source on` report: excessive shared-memory wavefronts
(bank conflicts) and stall samples by opcode.  Usage: ncu_source_hot.py report.ncu-rep <kernel regex> <launch skip>"""
import collections, csv, subprocess, sys

def f(v):
    try:
        return int(float(v))
    except ValueError:
        return 0

def dump(rep, kregex, skip):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex, "--launch-skip", str(skip),
                          "--launch-count", "1", "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = txt.splitlines()
    print(lines[0][:140])
    rows = list(csv.reader(lines[1:]))
    h = rows[0]
    rows = [r for r in rows[1:] if len(r) == len(h) and r[0] != h[0]]
    iS, iN, iE, iW, iX = (h.index(k) for k in ("Source", "# Samples", "L1 Wavefronts Shared Excessive", "L1 Wavefronts Shared", "Instructions Executed"))
    tot = sum(f(r[iN]) for r in rows)
    exc = sorted(((f(r[iE]), f(r[iW]), f(r[iX]), r[iS][:90]) for r in rows if f(r[iE]) > 0), reverse=True)
    print("samples %d; instructions with excessive shared wavefronts: %d (sum %d)" % (tot, len(exc), sum(e[0] for e in exc)))
    for e in exc[:10]:
        print("   excessive %d of %d wavefronts, executed %d: %s" % e)
    c = collections.Counter()
    for r in rows:
        sp = r[iS].split()
        op = sp[1] if sp and sp[0].startswith('@') and len(sp) > 1 else (sp[0] if sp else '?')
        c[op.split('.')[0]] += f(r[iN])
    print("stall samples by opcode:", c.most_common(10))

if __name__ == "__main__":
    dump(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
