"""Per-instruction executed counts of one kernel from an ncu report (source page), normalised per pixel.
usage: sass_prof.py <ncu-rep> <kernel substring> [pixels] -> prints summary + writes /tmp/sass_prof.txt"""
import csv, collections, subprocess, sys
rep, kern = sys.argv[1:3]
px = float(sys.argv[3]) if len(sys.argv) > 3 else 60 * 1920 * 1080
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; sections.append(cur); continue
    if r and r[0] == 'Address': cur['hdr'] = r; continue
    if cur is not None and len(r) > 10: cur['rows'].append(r)
sec = [s for s in sections if kern in s['name']][0]
hdr = sec['hdr']; data = sec['rows']
iA = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iS = hdr.index('# Samples')
tot = sum(int(r[iE]) for r in data)
print(sec['name'][:70]); print('warp instr', tot, 'thread instr/px', sum(int(r[iT]) for r in data) / px, 'lanes/px', tot * 32 / px)
with open('/tmp/sass_prof.txt', 'w') as fo:
    for i, r in enumerate(data):
        e = int(r[iE]); fo.write(f'{i:5d} {e*32/px:6.3f} {int(r[iT])/max(e,1):5.1f} {int(r[iS]):6d}  {r[iA]}\n')
h = collections.Counter(); hs = collections.Counter()
for r in data:
    t = r[iA].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    h[op] += int(r[iE]); hs[op] += int(r[iS])
ts = sum(hs.values())
for op, n in h.most_common(24): print(f'{op:10s} {n*32/px:7.2f} lanes/px  {100*n/tot:5.1f}%  samples {100*hs[op]/ts:5.1f}%')
