CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dynamics_kernel -s 4 -c 1 -o gpurun_out/prof_k1 $CMD > gpurun_out/ncu5.log 2>&1
tail -1 gpurun_out/ncu5.log
