"""profiles/r<round>_nconv_traffic.json from an `ncu --set full` report of one training step: DRAM bytes (read + write) and duration per
launch of the NConv kernels, keyed like bench.py's roofline table (kernel:call@level).  Usage: ncu_traffic.py report.ncu-rep out.json
Capture: STEPS=3 MARK=2 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:nconv_ -o rep python tools/prof_bench_step.py"""
import csv, json, os, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = rows[0]
col = {k: h.index(k) for k in ("Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                               "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed")}
units = rows[1]
scale = lambda u: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
# The capture holds exactly ONE steady-state training step (tools/prof_bench_step.py with MARK: cudaProfilerStart before the step), so
# the kernels appear in the pyramid's issue order (adc_nconv.cu): forward calls 0..9 (nconv7, call 10, is the 1x1 head with its own
# kernel), then the backward in reverse: weight gradients of calls 9..0, data gradients of calls 9..1 (nconv1 has none).
fwd_calls = ["nconv1@L0", "nconv2@L0", "nconv3@L0", "nconv2@L1", "nconv3@L1", "nconv2@L2", "nconv2@L3", "nconv4@L2", "nconv5@L1", "nconv6@L0"]
bwd_calls = list(reversed(fwd_calls))
seen = {"fwd": 0, "bwd_weight": 0, "bwd_data": 0}
res = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    kind = "fwd" if "nconv_fwd" in name else ("bwd_weight" if "nconv_bwd_weight" in name else ("bwd_data" if "nconv_bwd_data" in name else None))
    if kind is None:
        continue
    i = seen[kind]
    seen[kind] += 1
    calls = fwd_calls if kind == "fwd" else (bwd_calls if kind == "bwd_weight" else bwd_calls[:-1])
    if i >= len(calls):
        continue
    key = "nconv_%s:%s" % (kind, calls[i])
    rd = float(r[col["dram__bytes_read.sum"]]) * scale(units[col["dram__bytes_read.sum"]])
    wr = float(r[col["dram__bytes_write.sum"]]) * scale(units[col["dram__bytes_write.sum"]])
    res[key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "ncu_us": float(r[col["gpu__time_duration.sum"]]),
                "grid": r[col["Grid Size"]], "kernel": name.split("(")[0][-60:],
                "fma_pipe_pct": float(r[col["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]]),
                "source": "ncu --set full --clock-control none --profile-from-start off, tools/prof_bench_step.py MARK=2 (B=%s, 228x304, one "
                          "steady-state step), " % os.environ.get("B", "32") + rep.split("/")[-1]}
assert seen == {"fwd": 10, "bwd_weight": 10, "bwd_data": 9}, "the capture is not exactly one training step: %s" % seen
json.dump(res, open(out, "w"), indent=1)
for k, v in res.items():
    print("%-32s %-12s %7.1f us  dram %6.1f MB (rd %6.1f wr %6.1f)  fma %4.1f%%" % (k, v["grid"], v["ncu_us"], v["dram_bytes_per_launch"] / 1e6, v["dram_read"] / 1e6, v["dram_write"] / 1e6, v["fma_pipe_pct"]))
