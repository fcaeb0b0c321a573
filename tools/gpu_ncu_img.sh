mkdir -p gpurun_out
for img in 0 1; do
UB200_IMG=$img timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_op_write.sum,smsp__cycles_active.avg,sm__cycles_active.max,sm__cycles_active.avg -k regex:"fwd16|bwd16|wgrad16" -s 6 -c 3 --csv --log-file gpurun_out/ncu_img_$img.csv python tools/ncu_step.py c2_ipw_mslr10k 4 > /dev/null 2>&1
done
python - <<'PY'
import csv
for img in (0,1):
    rows=[r for r in csv.reader(open('gpurun_out/ncu_img_%d.csv'%img)) if len(r)>10]
    h=rows[0]
    ik,im,iv=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value')
    print('IMG=%d'%img)
    for r in rows[1:]:
        print('  %-14s %-34s %s'%(r[ik].split('(')[0].split('::')[-1], r[im], r[iv]))
PY
