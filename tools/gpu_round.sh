mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02j_pytest.log; tail -3 gpurun_out/r02j_pytest.log
L=gpurun_out/r02j_cycle.log; rm -f $L
python tools/cycle_case.py --cycles 10 >> $L 2>&1
MGB_TILE_MIN_PLANE=10000 python tools/cycle_case.py --cycles 10 >> $L 2>&1
MGB_TILE_PROLONG_MIN_PLANE=60000 python tools/cycle_case.py --cycles 10 >> $L 2>&1
MGB_TILE_SWEEP_MIN_PLANE=60000 python tools/cycle_case.py --cycles 10 >> $L 2>&1
MGB_TILE_MIN_PLANE=10000 MGB_TILE_PROLONG_MIN_PLANE=10000 python tools/cycle_case.py --cycles 10 >> $L 2>&1
MGB_TAIL_POINTS=40000 python tools/cycle_case.py --cycles 10 >> $L 2>&1
cat $L
