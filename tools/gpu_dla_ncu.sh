mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:dla_reg_kernel -s 2 -c 1 -o gpurun_out/prof_dla -f python tools/ncu_dla.py > gpurun_out/ncu_dla.log 2>&1
ncu -i gpurun_out/prof_dla.ncu-rep --page raw --csv > gpurun_out/prof_dla_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_dla_raw.csv')))
h=rows[0]
keys=["Kernel Name","Grid Size","gpu__time_duration.sum","launch__registers_per_thread","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","dram__bytes_read.sum","dram__bytes_write.sum","smsp__inst_executed.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__t_sector_hit_rate.pct","smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    for k in keys:
        if k in h: print(k.split('.')[0][:70], '=', r[h.index(k)])
    print('---')
PY
