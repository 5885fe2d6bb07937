mkdir -p gpurun_out
R=r02z
timeout 600 python -m pytest tests/test_gs_lex.py -m gpu -x -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${R}_pytest.log; tail -4 gpurun_out/${R}_pytest.log
L=gpurun_out/${R}_gslex.log; rm -f $L
for mode in 2 3 4 5; do echo "== MGB_GSLEX_TILE=$mode" >> $L; MGB_GSLEX_TILE=$mode timeout 300 python tools/bench_gslex.py --n 65 129 257 513 >> $L 2>&1; done
cat $L
