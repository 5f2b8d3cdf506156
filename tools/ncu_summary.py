#!/usr/bin/env python
"""Summarises an ncu report exported with `--page raw --csv` and `--page source --csv`: key metrics, and the
instruction / stall-sample split between the regions of the kernel delimited by BAR.SYNC (phase 1 = element
work, phase 2 = pull).  usage: ncu_summary.py raw.csv src.csv"""
import csv, sys
raw, src = sys.argv[1], sys.argv[2]
rows=list(csv.reader(open(raw)))
hdr=rows[0]; units=rows[1]; vals=rows[2]
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","launch__registers_per_thread","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","smsp__sass_inst_executed_op_local_ld.sum","smsp__sass_inst_executed_op_local_st.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__cycles_elapsed.max","smsp__warps_eligible.avg.per_cycle_active","launch__waves_per_multiprocessor","lts__t_sector_hit_rate.pct","l1tex__t_sector_hit_rate.pct","lts__t_bytes.sum","lts__t_sectors_op_write.sum","lts__t_sectors_op_read.sum","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers"]
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print(w, vals[i], units[i])
rows=list(csv.reader(open(src)))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
data=rows[2:]
stall_cols=[h for h in hdr if h.startswith('stall_') and '(' not in h]
regions=[]; cur={'inst':0,'samp':0,'start':0,'ops':{},'st':{}}
for k,r in enumerate(data):
    s=r[idx['Source']]
    inst=int(r[idx['Instructions Executed']] or 0); samp=int(r[idx['# Samples']] or 0)
    cur['inst']+=inst; cur['samp']+=samp
    t=s.split()
    op=(t[1] if t and t[0].startswith('@') and len(t)>1 else (t[0] if t else '')).split('.')[0]
    cur['ops'][op]=cur['ops'].get(op,0)+inst
    for sc in stall_cols: cur['st'][sc]=cur['st'].get(sc,0)+int(r[idx[sc]] or 0)
    if 'BAR.SYNC' in s or 'EXIT' in s:
        cur['end']=k; regions.append(cur); cur={'inst':0,'samp':0,'start':k+1,'ops':{},'st':{}}
regions.append(cur)
T=sum(r['inst'] for r in regions); S=sum(r['samp'] for r in regions)
print("total warp instructions", T)
for r in regions:
    if r['inst']==0 and r['samp']==0: continue
    top=sorted(r['ops'].items(), key=lambda x:-x[1])[:8]
    st=sorted(r['st'].items(), key=lambda x:-x[1])[:6]
    print("sass lines %d-%s: inst %.1f%% (%.3g) samples %.1f%%"%(r['start'],r.get('end'),100*r['inst']/T,r['inst'],100*r['samp']/S), top, st)
