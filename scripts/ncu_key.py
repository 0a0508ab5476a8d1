"""Key metrics of an ncu report: python scripts/ncu_key.py rep.ncu-rep"""
import csv, subprocess, sys
txt=subprocess.run(["ncu","-i",sys.argv[1],"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines())); H,U=rows[0],rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','lts__t_sector_hit_rate.pct']
stalls=[h for h in H if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for V in rows[2:]:
    print("==",V[H.index('Kernel Name')][:80])
    for w in want:
        if w in H: print(f"  {w:75s} {V[H.index(w)]:>16s} {U[H.index(w)]}")
    st=sorted(((float(V[H.index(h)]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for h in stalls),reverse=True)
    print("  stalls:", ", ".join(f"{n}={v:.2f}" for v,n in st[:7]))
