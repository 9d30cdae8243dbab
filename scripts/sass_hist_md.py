"""profiles/r02_sass_hist.md: executed-instruction histogram per output pixel (warp-instruction lanes = warp
instructions x 32 / pixels) of the pixel kernels, round 1 vs round 2, from the ncu source pages."""
import collections
import csv
import subprocess

PX = 60 * 1920 * 1080          # every capture: bench.py --frames 60, one launch


def kernel_rows(rep, kern):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    sections, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}; sections.append(cur); continue
        if r and r[0] == 'Address':
            cur['hdr'] = r; continue
        if cur is not None and len(r) > 10:
            cur['rows'].append(r)
    sec = [s for s in sections if kern in s['name']][0]
    return sec['hdr'], sec['rows']


def hist(rep, kern):
    hdr, data = kernel_rows(rep, kern)
    iA, iE, iT, iS = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
    h = collections.Counter()
    for r in data:
        t = r[iA].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        h[op] += int(r[iE])
    tot = sum(h.values())
    thr = sum(int(r[iT]) for r in data)
    return {k: v * 32 / PX for k, v in h.items()}, tot * 32 / PX, thr / PX


cols = [("r1 warp_fast_kernel", "gpurun_out/r01f_kernels.ncu-rep", "warp_fast"),
        ("r1 crop_resize_rows", "gpurun_out/r01f_kernels.ncu-rep", "crop_resize_rows"),
        ("r2 warp_fast_kernel", "gpurun_out/r02h_twokernel.ncu-rep", "warp_fast"),
        ("r2 crop_resize_rows", "gpurun_out/r02h_twokernel.ncu-rep", "crop_resize_rows"),
        ("r2 warp_fused_kernel (3-byte tile)", "gpurun_out/r02h_kernels.ncu-rep", "warp_fused"),
        ("r2 warp_fused_kernel (final: BGRx tile)", "gpurun_out/r02i_kernels.ncu-rep", "warp_fused")]
res = [(n,) + hist(rep, k) for n, rep, k in cols]
ops = collections.Counter()
for _, h, _, _ in res:
    for k, v in h.items():
        ops[k] = max(ops[k], v)
lines = ["# Executed instructions per output pixel, pixel kernels, round 1 vs round 2\n",
         "Source: `ncu --set full --import-source on`, `bench.py --frames 60 --steps 1` (60 frames of 1920x1080 per launch, c2 workload);",
         "`scripts/sass_hist_md.py` over `gpurun_out/r01f_kernels.ncu-rep`, `r02h_twokernel.ncu-rep`, `r02h_kernels.ncu-rep`, `r02i_kernels.ncu-rep`.",
         "Unit: warp-instruction lanes per output pixel = executed warp instructions x 32 / (60 x 1920 x 1080); the last",
         "row is thread instructions per pixel (lanes that were actually active).  The fused kernel does the work of the",
         "two kernels to its left (it computes only the pixels inside the crop rectangle, plus one row / column of overlap per tile).\n",
         "| opcode | " + " | ".join(n for n, *_ in res) + " |", "|---|" + "---:|" * len(res)]
for op, _ in ops.most_common(26):
    lines.append(f"| {op} | " + " | ".join(f"{h.get(op, 0):.2f}" for _, h, _, _ in res) + " |")
lines.append("| **all (lanes / px)** | " + " | ".join(f"**{t:.1f}**" for _, _, t, _ in res) + " |")
lines.append("| thread instructions / px | " + " | ".join(f"{t:.1f}" for _, _, _, t in res) + " |")
open('profiles/r02_sass_hist.md', 'w').write("\n".join(lines) + "\n")
print("\n".join(lines))
