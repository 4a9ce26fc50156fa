(timeout 900 python -m pytest tests -m gpu -x -q -k "host_vec or step_host" 2>&1 | tail -3)
for c in 2 3 4 6 8; do python bench.py --steps 100 --warmup 10 --no-cpu-baseline --host-chunks $c 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks',$c,'e2e %.3e'%d['e2e']['value'], d['e2e']['ms_per_step'])"; done
