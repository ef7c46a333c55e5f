"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <tag> [--launches gpurun_out/launches.csv] [--rep name=gpurun_out/x.ncu-rep ...]"""
import collections
import csv
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
           'launch__occupancy_limit_registers', 'sm__cycles_active.avg', 'lts__t_bytes.sum',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
           'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active']


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]; kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(',', '')); u = r[mu]
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        a = agg.setdefault(r[kn].split('(')[0], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, 'w') as f:
        f.write('kernel,launches,total_us,avg_us,share_pct\n')
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('%s,%d,%.1f,%.2f,%.2f\n' % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
    print('wrote', out)


def report(path, out):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in ('Kernel Name', 'Grid Size', 'Block Size') or h in METRICS]
    with open(out, 'w') as f:
        w = csv.writer(f)
        w.writerow(['%s [%s]' % (hdr[i], units[i]) if units[i] else hdr[i] for i in cols])
        for r in rows[2:]:
            w.writerow([r[i].split('(')[0] if hdr[i] == 'Kernel Name' else r[i] for i in cols])
    print('wrote', out)


if __name__ == '__main__':
    tag = sys.argv[1]
    args = sys.argv[2:]
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    i = 0
    while i < len(args):
        if args[i] == '--launches':
            launches(args[i + 1], os.path.join(ROOT, 'profiles', tag + '_launches.csv')); i += 2
        elif args[i] == '--rep':
            name, path = args[i + 1].split('=')
            report(path, os.path.join(ROOT, 'profiles', '%s_ncu_%s.csv' % (tag, name))); i += 2
        else:
            i += 1
