(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3)
python __graft_entry__.py smoke 2>&1 | tail -2
B="python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e"
P="import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('SWEEP', sys.argv[1], '| %.3e | %.3e | %.4f | %.4f / %.4f / %.4f | %.2f | %.3f | %.3f |'%(d['value'], d['rays_per_s'], d['ms_per_step'], r['all_kernels_ms']['dynamics'], r['all_kernels_ms']['lidar'], r['all_kernels_ms']['post'], r['lookups_per_ray'], r['frac'], r['gather_roofline']['frac_of_whole_map']))"
$B 2>&1 | tail -1 | python -c "$P" "C3: 4096 envs, A=1, B=1080, Shanghai"
$B --envs 32768 2>&1 | tail -1 | python -c "$P" "32768 envs, A=1, B=1080"
$B --envs 262144 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "$P" "C4: 262144 envs, A=1, B=1080"
for b in 270 540 2160 4320; do $B --envs 32768 --beams $b --steps 20 2>&1 | tail -1 | python -c "$P" "C4: 32768 envs, B=$b"; done
$B --envs 32768 --map-upsample 2 --steps 20 2>&1 | tail -1 | python -c "$P" "C4: 32768 envs, Shanghai x2 (4000^2, 128 MB)"
$B --envs 32768 --map-upsample 4 --steps 20 2>&1 | tail -1 | python -c "$P" "C4: 32768 envs, Shanghai x4 (8000^2, 512 MB)"
$B --agents 2 --envs 8192 2>&1 | tail -1 | python -c "$P" "C5 shape: 8192 envs, A=2"
python tools/c5_rollout.py --envs 8192
python tools/c5_rollout.py --envs 8192 --reward
python tools/c5_rollout.py --envs 8192 --reward --graph
