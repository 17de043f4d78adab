"""Summarise an ncu --csv launch list: per kernel (name, grid) mean duration, DRAM bytes and pipe utilisation."""
import collections, csv, sys

def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    data = collections.OrderedDict()
    for row in csv.DictReader(lines):
        data.setdefault((row['ID'], row['Kernel Name'], row['Grid Size'], row['Block Size']), {})[row['Metric Name']] = (row['Metric Value'], row['Metric Unit'])
    return data

def num(m, name):
    if name not in m: return float('nan')
    v, u = m[name]; v = float(v.replace(',', ''))
    scale = {'nsecond': 1e-3, 'ns': 1e-3, 'usecond': 1, 'us': 1, 'msecond': 1e3, 'ms': 1e3, 'second': 1e6, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1)
    return v * scale

def main():
    data = load(sys.argv[1])
    agg = collections.OrderedDict()
    for (id_, k, g, b), m in data.items():
        key = (k.split('(')[0].replace('void ', '').replace('adc::', '')[:48], g, b)
        agg.setdefault(key, []).append(m)
    tot = sum(num(m, 'gpu__time_duration.sum') for ms in agg.values() for m in ms)
    print("%-48s %-14s %-12s %4s %9s %6s %8s %8s %6s %6s" % ("kernel", "grid", "block", "n", "mean_us", "share", "rd_MB", "wr_MB", "sm%", "fma%"))
    for key, ms in sorted(agg.items(), key=lambda kv: -sum(num(m, 'gpu__time_duration.sum') for m in kv[1])):
        n = len(ms); t = sum(num(m, 'gpu__time_duration.sum') for m in ms)
        f = lambda name: sum(num(m, name) for m in ms) / n
        print("%-48s %-14s %-12s %4d %9.1f %5.1f%% %8.1f %8.1f %6.1f %6.1f" % (key[0], key[1], key[2], n, t / n, 100 * t / tot, f('dram__bytes_read.sum'), f('dram__bytes_write.sum'),
              f('sm__throughput.avg.pct_of_peak_sustained_elapsed'), f('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active')))
    print("total %.1f us over %d launches" % (tot, sum(len(v) for v in agg.values())))

if __name__ == '__main__':
    main()
