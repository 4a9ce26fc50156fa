# launch list + full captures of the step kernels for the default bench workload (run under gpurun)
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lidar_kernel|dynamics_kernel" -s 8 -c 4 -o gpurun_out/prof_step $CMD > gpurun_out/ncu2.log 2>&1
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --agents 2"
$CMD2 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 4 -c 1 -o gpurun_out/prof_post $CMD2 > gpurun_out/ncu4.log 2>&1
tail -1 gpurun_out/ncu2.log; tail -1 gpurun_out/ncu4.log
