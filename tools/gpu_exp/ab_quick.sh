#!/bin/bash
# quicker A/B than ab_libs.sh: bench only (two runs at C3, one at 32768 envs) for every tools/gpu_exp/libs/*.so
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for lib in tools/gpu_exp/libs/*.so; do
  name=$(basename $lib .so)
  export F110_B200_LIB=$PWD/$lib
  out=""
  for envs in 4096 4096 32768; do
    python bench.py --no-e2e --no-cpu-baseline --envs $envs --steps 200 > gpurun_out/q.json 2>/dev/null
    out="$out $(python -c "import json;d=json.load(open('gpurun_out/q.json'));print('%.4f/%.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))")"
  done
  echo "$name step/lidar ms at 4096, 4096, 32768 envs:$out"
done
