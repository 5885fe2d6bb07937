timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for l in 10 9 8 7; do timeout 100 python tools/probe_tile.py --levels $l --reps 5 --which norm,rr | awk '{print $1,$2,$3,$4}'; done
timeout 100 python tools/probe.py --levels 9 --reps 5 --cycles 20 2>&1 | grep -E "vcycle"
