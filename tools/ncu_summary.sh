#!/bin/bash
# Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.txt:
#   per-launch duration, DRAM bytes, DRAM/SM throughput %, occupancy, registers, stall mix.
# Usage: tools/ncu_summary.sh gpurun_out/prof_env_d2_r01.ncu-rep profiles/ncu_env_d2_r01.txt
REP=$1; OUT=$2
PAT='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__bytes_read.sum$|dram__bytes_write.sum$|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|launch__occupancy_limit|smsp__inst_executed.sum$|sm__inst_executed_pipe_fp64|sm__pipe_fp64_cycles_active|smsp__pcsamp_warps_issue_stalled|l1tex__data_bank_conflicts_pipe_lsu_mem_shared|lts__t_sector_hit_rate|sm__cycles_elapsed.max|smsp__cycles_active.avg|smsp__issue_active.avg.pct|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|sm__pipe_fmaheavy|shared_mem|launch__shared'
ncu -i "$REP" --page raw --csv 2>/dev/null | python3 -c "
import csv, sys, re
pat = re.compile(r'''$PAT''')
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('=== launch', d.get('ID'), d.get('Kernel Name','')[:110], 'grid', d.get('Grid Size'), 'block', d.get('Block Size'))
    for k, u in zip(hdr, units):
        if pat.search(k) and d.get(k, '') not in ('', 'n/a'):
            print('  %-75s %s %s' % (k, d[k], u))
" > "$OUT"
wc -l "$OUT"
