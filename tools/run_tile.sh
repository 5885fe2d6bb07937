timeout 600 python -m pytest tests/test_gpu_operators.py -x -q -k "restrict or vcycle" 2>&1 | tail -2
for l in 9 8; do timeout 100 python tools/probe_tile.py --levels $l --reps 10 --which rr | awk '{print $1,$2,$3,$4}'; done
