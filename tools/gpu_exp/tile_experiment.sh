#!/bin/bash
# The shared-memory tile experiment (lidar_tile_kernel, F110_LIDAR_TILE=1) against the default lidar kernel:
# bench at three batch sizes and on the 8000^2 map, then one ncu pass per kernel for time, L1 / L2 hit rates, DRAM bytes.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
for mode in 0 1; do
  export F110_LIDAR_TILE=$mode
  for cfg in "--envs 4096" "--envs 32768" "--envs 32768 --map-upsample 4"; do
    python bench.py --no-e2e --no-cpu-baseline --steps 60 --warmup 10 $cfg 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('tile=$mode', '$cfg', 'step_ms %.4f' % d['ms_per_step'], 'lidar_ms %.4f' % d['roofline']['kernel_ms'], 'env-steps/s %.3e' % d['value'])"
  done
done
for mode in 0 1; do
  for cfg in "--envs 32768" "--envs 32768 --map-upsample 4"; do
    tag=$(echo "$cfg" | tr -d ' -')
    F110_LIDAR_TILE=$mode ncu --metrics $M --clock-control none -k regex:"lidar_kernel|lidar_tile_kernel" -s 6 -c 2 --csv --log-file gpurun_out/tile_ncu_${mode}_${tag}.csv \
      python bench.py --no-e2e --no-cpu-baseline --steps 6 --warmup 3 $cfg > /dev/null 2>&1
  done
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob('gpurun_out/tile_ncu_*.csv')):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = next(r for r in rows if r[0] == 'ID')
    vals = {}
    for r in rows:
        if r[0] == 'ID': continue
        d = dict(zip(hdr, r))
        if d['ID'] != '1': continue         # second captured launch
        vals[d['Metric Name']] = (d['Metric Value'], d['Metric Unit'])
        k = d['Kernel Name'].split('(')[0][-40:]
    print(f.split('/')[-1], k, {m.split('.')[0].replace('__','.'): v[0] + ' ' + v[1] for m, v in vals.items()})
PY
