#!/bin/bash
# A/B of library builds on the two-car shape (post_kernel): every tools/gpu_exp/libs/*.so runs 8192 two-car envs and 4096 four-car envs.
cd "$(dirname "$0")/../.."
for lib in tools/gpu_exp/libs/${1:-*}.so; do
  name=$(basename $lib .so)
  export F110_B200_LIB=$PWD/$lib
  for cfg in "2 8192" "4 4096"; do
    set -- $cfg
    python bench.py --no-e2e --no-cpu-baseline --agents $1 --envs $2 --steps 60 --warmup 10 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$name', 'agents $1 envs $2', 'step_ms %.4f' % d['ms_per_step'], 'kernels', {k: round(v,4) for k,v in d['roofline']['all_kernels_ms'].items()})"
  done
done
