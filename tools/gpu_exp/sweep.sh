#!/bin/bash
# The rows of profiles/rNN_sweeps.md: bench.py over batch size, beam count and map size (one B200, under gpurun).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
row() {   # label, bench flags...
  label="$1"; shift
  python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e "$@" > gpurun_out/sweep.json 2> gpurun_out/sweep.err || { echo "| $label | FAILED |"; return; }
  python - "$label" <<'PY'
import json, sys
d = json.load(open("gpurun_out/sweep.json")); r = d["roofline"]; k = r["all_kernels_ms"]; g = r["gather_roofline"]
print("| %s | %.3e | %.3e | %.4f | %.4f / %.4f / %.4f | %.2f | %.3f | %.3f |" % (sys.argv[1], d["value"], d["rays_per_s"], d["ms_per_step"],
      k["dynamics"], k["lidar"], k["post"], r["lookups_per_ray"], r["frac"], g["frac_of_whole_map"]))
PY
}
echo "| configuration | env-steps/s | rays/s | ms/step | K1 / K2 / K3 ms | L̄ | frac (HBM) | gather |"
echo "|---|---|---|---|---|---|---|---|"
row "C3: 4096 envs, A=1, B=1080, Shanghai"
row "512 envs" --envs 512
row "32768 envs, A=1, B=1080" --envs 32768
row "C4: 262144 envs, A=1, B=1080" --envs 262144
row "C4: 32768 envs, B=270" --envs 32768 --beams 270
row "C4: 32768 envs, B=540" --envs 32768 --beams 540
row "C4: 32768 envs, B=2160" --envs 32768 --beams 2160
row "C4: 32768 envs, B=4320" --envs 32768 --beams 4320
row "C4: 32768 envs, Shanghai x2 (4000^2, 128 MB)" --envs 32768 --map-upsample 2
row "C4: 32768 envs, Shanghai x4 (8000^2, 512 MB)" --envs 32768 --map-upsample 4
row "C4 at full size: 262144 envs, B=270" --config c4 --beams 270
row "C4 at full size: 262144 envs, B=2160" --config c4 --beams 2160
row "C4 at full size: 262144 envs, B=4320" --config c4 --beams 4320
row "C4 at full size: 262144 envs, Shanghai x2 (4000^2)" --config c4 --map-upsample 2
row "C4 at full size: 262144 envs, Shanghai x4 (8000^2)" --config c4 --map-upsample 4
row "C5 shape: 8192 envs, A=2" --envs 8192 --agents 2
row "4096 envs, A=4" --envs 4096 --agents 4
python - <<'PY'
import json
g = json.load(open("gpurun_out/sweep.json"))["roofline"]["gather_roofline"]
print("gather probes (GB/s of useful 8-byte cells): whole 32 MB map %.0f, 4 MiB window %.0f, 2 GiB HBM-resident %.0f" % (g["whole_map_gbs"], g["touched_window_gbs"], g["hbm_2gib_gbs"]))
PY
