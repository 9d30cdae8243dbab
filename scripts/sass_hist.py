import csv, collections, re, sys, subprocess
rep=sys.argv[1]; px=float(sys.argv[2]) if len(sys.argv)>2 else 60*1920*1080
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]; data=rows[2:]
iA=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iT=hdr.index('Thread Instructions Executed'); iS=hdr.index('# Samples')
agg=collections.Counter(); aggT=collections.Counter(); samp=collections.Counter(); tot=0; totT=0
for r in data:
    if len(r)<=iT: continue
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iA].strip())
    op=m.group(2).split('.')[0] if m else r[iA][:10]
    n=int(r[iE]); agg[op]+=n; aggT[op]+=int(r[iT]); samp[op]+=int(r[iS]); tot+=n; totT+=int(r[iT])
print('total warp instr', tot, 'thread instr/pixel %.1f'%(totT/px), 'warp-level lanes/pixel %.1f'%(tot*32/px))
for op,n in agg.most_common(18):
    print(f"{op:10s} warp-instr {n/1e6:9.1f}M  lanes/px {n*32/px:7.1f}  thread-instr/px {aggT[op]/px:7.1f}  samples {samp[op]}")
big=sorted([r for r in data if len(r)>iS],key=lambda r:-int(r[iS]))[:12]
for r in big: print(r[iS], r[iE], r[iA][:90])
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rr=list(csv.reader(raw.splitlines())); h,u,v=rr[0],rr[1],rr[2]
for k in ['gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']:
    if k in h: print(k, v[h.index(k)], u[h.index(k)])
