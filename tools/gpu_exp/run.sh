(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
P="import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(sys.argv[1], 'env-steps/s %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'], {k:round(v,4) for k,v in r['all_kernels_ms'].items()}, 'frac %.3f'%r['frac'])"
$B 2>&1 | tail -1 | python -c "$P" A1_4096
$B --envs 512 2>&1 | tail -1 | python -c "$P" A1_512
$B --envs 32768 --steps 30 2>&1 | tail -1 | python -c "$P" A1_32768
