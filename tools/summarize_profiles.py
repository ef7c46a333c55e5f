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


def _mb(row, hdr, name):
    for i, h in enumerate(hdr):
        if h.startswith(name):
            v = float(row[i].replace(',', ''))
            unit = h[h.index('[') + 1:h.index(']')] if '[' in h else 'byte'
            return v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    raise KeyError(name)


def traffic(tag):
    """profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch (per call for the three-kernel
    back-projection) from the tag's ncu --set full summaries; bench.py copies these into roofline.traffic."""
    import json
    out = {}
    pg = os.path.join(ROOT, 'profiles', tag + '_ncu_gemm.csv')
    if os.path.exists(pg):
        rows = list(csv.reader(open(pg))); hdr = rows[0]
        b = [_mb(r, hdr, 'dram__bytes_read.sum') + _mb(r, hdr, 'dram__bytes_write.sum') for r in rows[1:] if r and 'gemm_split' in r[0]]
        g = [r for r in rows[1:] if r and 'gemm_split' in r[0]]
        ti = [i for i, h in enumerate(hdr) if h.startswith('gpu__time_duration.sum')][0]
        pi = [i for i, h in enumerate(hdr) if h.startswith('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')][0]
        tw = sum(float(r[ti]) * float(r[pi]) for r in g) / sum(float(r[ti]) for r in g)
        out['gemm'] = dict(kernel='tc2::gemm_split_bf16_persistent_kernel', launches_captured=len(b), dram_bytes_per_launch=sum(b) / len(b),
                           tensor_pipe_active_pct_time_weighted=tw,
                           dram_bytes_per_step=sum(b), source='profiles/%s_ncu_gemm.csv (ncu --set full, one bench step: %d GEMM launches)' % (tag, len(b)))
    per = collections.OrderedDict()
    for name in ('label', 'icp'):                      # ncu picks the byte unit per report: one summary file per report
        per = collections.OrderedDict() if name == 'label' else per
        pl = os.path.join(ROOT, 'profiles', '%s_ncu_%s.csv' % (tag, name))
        if os.path.exists(pl):
            rows = list(csv.reader(open(pl))); hdr = rows[0]
            for r in rows[1:]:
                if r:
                    per.setdefault(r[0].strip(), []).append(_mb(r, hdr, 'dram__bytes_read.sum') + _mb(r, hdr, 'dram__bytes_write.sum'))
    if per:
        avg = {k: sum(v) / len(v) for k, v in per.items()}
        surf = [k for k in avg if k.startswith('surface_')]
        if surf:
            out['surface_backproject'] = dict(kernels={k: avg[k] for k in surf}, dram_bytes_per_launch=sum(avg[k] for k in surf),
                                              source='profiles/%s_ncu_label.csv (mask + scan + emit of one 512-frame call)' % tag)
        icp = [k for k in avg if 'icp_p2p' in k]
        if icp:
            nreg = 1184 if tag.startswith('r01') else 3552
            out['icp_p2p'] = dict(kernel=icp[0], dram_bytes_per_launch=avg[icp[0]], registrations_per_launch=nreg,
                                  source='profiles/%s_ncu_icp.csv (captured with %d registrations per launch; bench.py scales it to '
                                         'its own launch size)' % (tag, nreg))
    pa = os.path.join(ROOT, 'profiles', tag + '_ncu_adds.csv')
    if os.path.exists(pa):
        rows = list(csv.reader(open(pa))); hdr = rows[0]
        b = [_mb(r, hdr, 'dram__bytes_read.sum') + _mb(r, hdr, 'dram__bytes_write.sum') for r in rows[1:] if r and 'add_metric' in r[0]]
        if b:
            mixed = b[:-1] if len(b) > 1 else b           # the last captured launch is the all-symmetric run
            out['add_metric'] = dict(kernel='add_metric_kernel', dram_bytes_per_launch=sum(mixed) / len(mixed), instances_per_launch=12500,
                                     dram_bytes_per_launch_all_symmetric=b[-1],
                                     source='profiles/%s_ncu_adds.csv (12 500 instances per launch, dataset mix of symmetric classes; last row: all symmetric)' % tag)
    with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('wrote profiles/traffic.json', {k: round(v['dram_bytes_per_launch'] / 1e6, 1) for k, v in out.items()}, 'MB per launch')


if __name__ == '__main__':
    tag = sys.argv[1]
    args = sys.argv[2:]
    outdir = os.path.join(ROOT, 'profiles')
    if '--outdir' in args:
        outdir = args[args.index('--outdir') + 1]
    os.makedirs(outdir, exist_ok=True)
    i = 0
    while i < len(args):
        if args[i] == '--launches':
            launches(args[i + 1], os.path.join(outdir, tag + '_launches.csv')); i += 2
        elif args[i] == '--traffic':
            traffic(tag); i += 1
        elif args[i] == '--rep':
            name, path = args[i + 1].split('=')
            report(path, os.path.join(outdir, '%s_ncu_%s.csv' % (tag, name))); i += 2
        elif args[i] == '--outdir':
            i += 2
        else:
            i += 1
