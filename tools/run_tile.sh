for l in 10 9 8; do timeout 100 python tools/probe_tile.py --levels $l --reps 5 --which hs,norm,rr,pc | awk '{print $1,$2,$3,$4,$5,$6}'; done
timeout 100 python tools/probe.py --levels 9 --reps 5 --cycles 20 2>&1 | grep -E "vcycle"
timeout 100 python tools/probe.py --levels 10 --reps 2 --cycles 5 2>&1 | grep -E "vcycle"
