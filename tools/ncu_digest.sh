#!/bin/bash
# usage: tools/ncu_digest.sh <report.ncu-rep> <out-prefix>   -- raw-page metrics + per-line digests as small text files
rep=$1; out=$2
KEYS='gpu__time_duration.sum|sm__inst_executed.sum|smsp__issue_active.avg.pct|sm__pipe_fp64_cycles_active.avg.pct|sm__inst_executed_pipe_fp64|lts__t_sector_hit_rate.pct|lts__t_sectors_op_read.sum|lts__t_sectors_op_write.sum|dram__bytes_read.sum|dram__bytes_write.sum|dram__throughput|gpu__dram_throughput|sm__warps_active.avg.pct|smsp__average_warps_issue_stalled|launch__registers_per_thread|launch__grid_size|launch__block_size|launch__shared_mem|launch__cluster|l1tex__data_bank_conflicts_pipe_lsu_mem_shared|smsp__inst_executed_pipe|sm__throughput|lts__throughput|sm__pipe_tensor|smsp__cycles_active.avg|sm__cycles_elapsed.max'
ncu -i $rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys,re
r=list(csv.reader(sys.stdin))
h,u,v=r[0],r[1],r[2]
pat=re.compile(r'$KEYS')
print('kernel:', v[h.index('Kernel Name')] if 'Kernel Name' in h else '?')
for i,k in enumerate(h):
    if pat.search(k): print('%-90s %-12s %s' % (k,u[i],v[i]))
" > ${out}_ncu_raw.txt
ncu -i $rep --page source --csv --print-source cuda,sass > /tmp/_src.csv 2>/dev/null
python3 tools/ncu_lines.py /tmp/_src.csv 40 > ${out}_hot_lines_by_stall.txt 2>/dev/null
python3 tools/ncu_phase.py /tmp/_src.csv 40 > ${out}_hot_lines_by_inst.txt 2>/dev/null
rm -f /tmp/_src.csv
