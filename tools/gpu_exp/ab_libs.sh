#!/bin/bash
# A/B of library builds: every tools/gpu_exp/libs/*.so (built by hand with -D overrides; selected through F110_B200_LIB)
# runs the bit-exact scan tests, the N=1 launch floor and the C3 bench at three batch sizes.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for lib in tools/gpu_exp/libs/*.so; do
  name=$(basename $lib .so)
  export F110_B200_LIB=$PWD/$lib
  t=$(python -m pytest tests -m gpu -x -q -k "scan or golden or batch" 2>&1 | tail -1)
  fl=$(python tools/gpu_exp/floor_probe.py 2>/dev/null | grep warm | awk '{printf "%s ", $(NF-1)}')
  for envs in 512 4096 32768; do
    python bench.py --no-e2e --no-cpu-baseline --envs $envs --steps 100 > gpurun_out/ab_$name.$envs.json 2>/dev/null
  done
  python - <<PY
import json
r=[json.load(open("gpurun_out/ab_$name.%d.json"%e))["roofline"]["kernel_ms"] for e in (512,4096,32768)]
print("$name [$t] floor_us[$fl] lidar_ms 512/4096/32768: %.4f %.4f %.4f" % tuple(r))
PY
done
