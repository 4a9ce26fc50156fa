(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/tests.log
cat gpurun_out/tests.log
python bench.py --steps 100 --warmup 10 > gpurun_out/bench.log 2>&1; tail -c 1800 gpurun_out/bench.log
for c in 1 2 8; do python bench.py --steps 100 --warmup 10 --no-cpu-baseline --host-chunks $c 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks',$c,d['e2e'])"; done
