"""Per-source-line executed-instruction histogram: joins an ncu SASS source page with nvdisasm -g line info.
usage: line_hist.py <ncu-rep> <nvdisasm.txt> <mangled kernel substring> [pixels]"""
import csv, collections, re, sys, subprocess
rep, dis, kern = sys.argv[1:4]
px = float(sys.argv[4]) if len(sys.argv) > 4 else 60*1920*1080
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 10]
iA = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iS = hdr.index('# Samples')
# nvdisasm: collect (line info) per instruction in order within the kernel's section
lines = open(dis).read().splitlines()
start = next(i for i,l in enumerate(lines) if l.startswith('.text.') and kern in l and l.endswith(':'))
cur = ('?',0); seq = []
for l in lines[start+1:]:
    if l.startswith('//-----') or (l.startswith('.text.') and l.endswith(':')): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): seq.append(cur)
print('sass instr: ncu', len(data), 'nvdisasm', len(seq))
agg = collections.Counter(); aggT = collections.Counter(); smp = collections.Counter()
for r, ln in zip(data, seq):
    agg[ln] += int(r[iE]); aggT[ln] += int(r[iT]); smp[ln] += int(r[iS])
tot = sum(agg.values())
src_cache = {}
def src(fn, n):
    import glob
    if fn not in src_cache:
        c = glob.glob('/root/repo/meshflow_b200/csrc/'+fn)
        src_cache[fn] = open(c[0]).read().splitlines() if c else []
    s = src_cache[fn]
    return s[n-1].strip()[:90] if 0 < n <= len(s) else ''
for ln, n in agg.most_common(40):
    print(f"{ln[0]:14s}:{ln[1]:4d} lanes/px {n*32/px:6.1f} thr/px {aggT[ln]/px:6.1f} smp {smp[ln]:6d} | {src(*ln)}")
