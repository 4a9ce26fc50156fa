(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python tools/c5_rollout.py --envs 8192
python tools/c5_rollout.py --envs 8192 --graph
python tools/c5_rollout.py --envs 8192 --opponent const
python tools/c5_rollout.py --envs 1024
python tools/c5_rollout.py --envs 1024 --graph
