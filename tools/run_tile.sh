timeout 300 python -m pytest tests/test_gpu_operators.py -x -q 2>&1 | tail -3
echo "== tile"; timeout 200 python tools/probe.py --levels 9 --reps 5 --cycles 3 2>&1 | tail -21
echo "== no tile"; MGB_TILE=0 timeout 200 python tools/probe.py --levels 9 --reps 5 --cycles 3 2>&1 | tail -12
