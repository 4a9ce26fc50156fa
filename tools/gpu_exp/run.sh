B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
for h in 48 64 96 160; do F110_B200_LIB=$PWD/tools/gpu_exp/lib_h$h.so $B > gpurun_out/b_h$h.log 2>&1; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/b_*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["all_kernels_ms"].items()}, "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e, open(f).read()[-300:])
PY
