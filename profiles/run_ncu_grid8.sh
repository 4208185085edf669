#!/bin/bash
OUT=gpurun_out/${1:-grid8}
mkdir -p $OUT
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'solve_reg_gemm' -s 2 -c 2 -f -o $OUT/prof python profiles/split_cfg3_vec.py > $OUT/ncu.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv --print-source cuda,sass > $OUT/src.csv 2>/dev/null
python profiles/summarize_lines.py $OUT/src.csv "" 2.0 > $OUT/lines.txt 2>&1; rm -f $OUT/src.csv
python - <<PY
import csv
rows=list(csv.reader(open("$OUT/raw.csv")))
hdr=rows[0]; u=dict(zip(hdr,rows[1]))
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print(d['Kernel Name'][:70], d['launch__grid_size'])
    for k in ['gpu__time_duration.sum','smsp__inst_executed.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sectors_op_write.sum','lts__t_sectors_op_atom.sum','lts__t_sectors_op_red.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum']:
        if k in d: print('  ',k,d[k],u[k])
    for k in hdr:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and float(d[k] or 0)>0.4: print('   ',k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),d[k])
PY
head -40 $OUT/lines.txt
