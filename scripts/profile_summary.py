"""Summarises gpurun_out/<tag>_launches.csv and <tag>_kernels.ncu-rep into profiles/<tag>_summary.md."""
import csv, collections, subprocess, sys, json
tag = sys.argv[1]
out = []
lines = [l for l in open(f'gpurun_out/{tag}_launches.csv') if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    agg.setdefault(row['Kernel Name'].split('(')[0][:70], []).append(float(row['Metric Value'].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
out.append(f"# ncu evidence {tag}\n\n## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, `python bench.py --steps 2 --warmup 1`)\n")
out.append("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
for k, v in agg.items():
    out.append(f"| `{k}` | {len(v)} | {sum(v)/1e6:.3f} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/tot*100:.1f}% |")
raw = subprocess.run(['ncu', '-i', f'gpurun_out/{tag}_kernels.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h, u = rr[0], rr[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size']
out.append(f"\n## `ncu --set full --clock-control none` (`bench.py --frames 60 --steps 1 --warmup 1`: 60 frames of 1920x1080 per launch)\n")
for v in rr[2:]:
    name = v[h.index('Kernel Name')].split('(')[0]
    out.append(f"\n### `{name}`\n\n| metric | value | unit |\n|---|---:|---|")
    for k in keys:
        if k in h: out.append(f"| {k} | {v[h.index(k)]} | {u[h.index(k)]} |")
    st = [(h[i].split('stalled_')[1], float(v[i].replace(',', ''))) for i in range(len(h)) if 'pcsamp_warps_issue_stalled' in h[i] and 'not_issued' not in h[i]]
    s = sum(x[1] for x in st) or 1
    out.append("| top stall reasons | " + ", ".join(f"{n} {x/s*100:.0f}%" for n, x in sorted(st, key=lambda x: -x[1])[:6]) + " | |")
open(f'profiles/{tag}_summary.md', 'w').write("\n".join(out) + "\n")
print("\n".join(out))
