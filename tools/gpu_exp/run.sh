(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/tests.log
cat gpurun_out/tests.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
P="import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], '%.3e'%d['value'], d['ms_per_step'], d['roofline']['all_kernels_ms'], d['roofline']['frac'])"
$B 2>&1 | tail -1 | python -c "$P" default
$B 2>&1 | tail -1 | python -c "$P" default
