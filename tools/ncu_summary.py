import csv,collections,sys,subprocess
rep=sys.argv[1]; moves=float(sys.argv[2])
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','smsp__warps_eligible.avg.per_cycle_active','smsp__average_warp_latency_per_inst_issued.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum','launch__shared_mem_per_block_dynamic','sm__cycles_active.avg']
for i,h in enumerate(hdr):
    if h in want: print(f"{h:75s} {units[i]:12s} {vals[i]}")
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}; data=rows[2:]
tot=sum(int(r[ix['Instructions Executed']]) for r in data)
print('total warp inst',tot,'per move (warp-level)',tot/moves)
tthr=sum(int(r[ix['Thread Instructions Executed']]) for r in data)
print('thread inst per move', tthr/moves)
cls=collections.Counter(); stall=collections.Counter()
for r in data:
    t=r[ix['Source']].split()
    op=t[1] if t[0].startswith('@') else t[0]
    op=op.split('.')[0]
    cls[op]+=int(r[ix['Instructions Executed']])
    stall[op]+=int(r[ix['# Samples']])
ts=sum(stall.values())
for k,v in cls.most_common(22): print(f"{k:10s} {v/tot*100:5.1f}%  {v/moves:8.2f}/move   samples {stall[k]/ts*100:5.1f}%")
for name in ['stall_long_sb','stall_short_sb','stall_math','stall_wait','stall_not_selected','stall_selected','stall_branch_resolving','stall_no_inst','stall_dispatch','stall_mio','stall_lg','stall_barrier']:
    print(name, sum(int(r[ix[name]]) for r in data)/ts*100)
