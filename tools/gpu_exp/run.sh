(timeout 900 python -m pytest tests -m gpu -x -q -s -k "shaped" 2>&1 | tail -4)
python tools/gpu_exp/reward_probe.py
