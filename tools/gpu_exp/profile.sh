# One ncu pass per gpurun call (the plain command runs first and must exit 0).  Usage, from the repo root on the box:
#   bash tools/gpu_exp/profile.sh launches   launch list of the default bench workload (gpu__time_duration per launch)
#   bash tools/gpu_exp/profile.sh step       full capture of lidar_kernel + dynamics_kernel, A = 1
#   bash tools/gpu_exp/profile.sh post       full capture of post_kernel, A = 2
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
case "$1" in
launches)
  $CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
  tail -1 gpurun_out/ncu1.log ;;
step)
  $CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lidar_kernel|dynamics_kernel" -s 8 -c 4 -o gpurun_out/prof_step $CMD > gpurun_out/ncu2.log 2>&1
  tail -1 gpurun_out/ncu2.log ;;
post)
  CMD2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --agents 2"
  $CMD2 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 4 -c 1 -o gpurun_out/prof_post $CMD2 > gpurun_out/ncu4.log 2>&1
  tail -1 gpurun_out/ncu4.log ;;
esac
