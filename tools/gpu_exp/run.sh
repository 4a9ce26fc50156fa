(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
B="python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e"
P="import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(sys.argv[1], 'env-steps/s %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'], {k:round(v,4) for k,v in r['all_kernels_ms'].items()}, 'frac %.3f'%r['frac'])"
$B --agents 2 2>&1 | tail -1 | python -c "$P" A2_4096
$B --agents 2 --envs 8192 2>&1 | tail -1 | python -c "$P" A2_8192
python tools/c5_rollout.py --envs 8192 --reward --graph
