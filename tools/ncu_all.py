import csv,sys,subprocess,io
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rr=list(csv.reader(io.StringIO(raw)))
h=rr[0]
want=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for r in rr[2:]:
    d=dict(zip(h,r))
    print('----')
    for w in want:
        if w in d: print(f'{w:70s} {d[w]}')
    st=[]
    for k,v in d.items():
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
            try: st.append((float(v),k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
            except: pass
    print('stalls:', ', '.join(f'{k}={v:.2f}' for v,k in sorted(st,reverse=True)[:8]))
