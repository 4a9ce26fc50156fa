#!/bin/bash
# A/B of library builds at the small end: N = 1 launch floor (floor_probe.py) and the C3 bench at 512 / 4096 envs.
cd "$(dirname "$0")/../.."
for lib in tools/gpu_exp/libs/${1:-*}.so; do
  name=$(basename $lib .so)
  export F110_B200_LIB=$PWD/$lib
  fl=$(python tools/gpu_exp/floor_probe.py 2>/dev/null | grep warm | awk '{printf "%s ", $(NF-1)}')
  echo "$name N=1 warm floor_us [$fl]"
  for envs in 512 4096; do
    python bench.py --no-e2e --no-cpu-baseline --envs $envs --steps 60 --warmup 10 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$name', $envs, 'step_ms %.4f' % d['ms_per_step'], 'kernels', {k: round(v,4) for k,v in d['roofline']['all_kernels_ms'].items()})"
  done
done
