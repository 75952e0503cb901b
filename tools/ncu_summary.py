"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv --print-source cuda,sass`."""
import csv, collections, sys
raw, src = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
rows=list(csv.reader(open(raw)))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','launch__occupancy_limit','launch__waves','smsp__average_warps_issue_stalled','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__grid_size','launch__block_size','dram__throughput']
for vals in rows[2:]:
    print("="*60)
    for i,h in enumerate(hdr):
        if any(h==w or h.startswith(w+'.') or (w.endswith('stalled') and h.startswith(w)) or (w.startswith('launch__occ') and h.startswith(w)) for w in want):
            if 'stalled' in h and float(vals[i] or 0) < 0.3: continue
            print("%-90s %-16s %s"%(h, units[i], vals[i]))
if src:
    rows=list(csv.reader(open(src)))
    # several kernels: split on header rows
    hi=[i for i,r in enumerate(rows) if len(r)>3 and r[0]=='Line No']
    for n,start in enumerate(hi):
        end = hi[n+1]-2 if n+1<len(hi) else len(rows)
        h=rows[start]
        iI=h.index('Instructions Executed'); iSm=h.index('# Samples')
        agg=collections.OrderedDict()
        for r in rows[start+1:end]:
            if len(r)<=iI: continue
            try: ln=int(r[0])
            except: continue
            a=agg.setdefault((ln, r[1][:100]),[0,0])
            try: a[0]+=int(r[iI]); a[1]+=int(r[iSm])
            except: pass
        tot=sum(v[0] for v in agg.values()) or 1; tots=sum(v[1] for v in agg.values()) or 1
        print("-"*60); print(rows[start-2][:2] if start>=2 else '', "total inst", tot, "samples", tots)
        for (ln,s),v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:24]:
            print("%5d inst %5.1f%% smp %5.1f%%  %s"%(ln, 100*v[0]/tot, 100*v[1]/tots, s))
