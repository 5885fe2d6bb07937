timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
for f in 1 0; do for l in 9 8; do MGB_TILE_FIXED=$f timeout 100 python tools/probe_tile.py --levels $l --reps 10 --which norm,rr | awk -v f=$f '{print "fixed=" f,$1,$2,$3,$4}'; done; done
