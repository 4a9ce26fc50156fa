# bash tools/gpu_exp/prof_lidar.sh [envs]  -- plain bench, then one ncu --set full capture of the lidar kernel (3 launches)
E_=${1:-4096}
mkdir -p gpurun_out
CMD="python bench.py --envs $E_ --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$E_.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lidar_kernel" -s 8 -c 3 -o gpurun_out/prof_lidar_$E_ $CMD > gpurun_out/ncu_$E_.log 2>&1
tail -2 gpurun_out/ncu_$E_.log
