# Single-GPU evidence of a round (run through gpurun): smoke, GPU tests, bench line, reference arm,
# ncu launch list and one --set full capture of the tile kernels, RB-GS and per-kernel probes.
#   gpurun -- 'R=r03a bash tools/gpu_round.sh'     -> gpurun_out/$R_*
mkdir -p gpurun_out
R=${R:-rXX}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/${R}_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${R}_pytest.log; tail -4 gpurun_out/${R}_pytest.log
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc $?"; tail -c 400 gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_n1.json 2> gpurun_out/${R}_bench_reference_n1.err; echo "ref rc $?"
python tools/cycle_case.py --cycles 2 > gpurun_out/${R}_case.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python tools/cycle_case.py --cycles 2 > gpurun_out/${R}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_tile' --launch-skip 16 -c 14 -o gpurun_out/${R}_prof -f python tools/cycle_case.py --cycles 2 > gpurun_out/${R}_ncu_full.log 2>&1; echo "ncu rc $?"
python tools/bench_rbgs.py --n 257 --iters 200 > gpurun_out/${R}_rbgs_257.json 2>&1
python tools/bench_rbgs.py --n 513 --iters 40 > gpurun_out/${R}_rbgs_513.json 2>&1
python tools/bench_gslex.py > gpurun_out/${R}_gslex.log 2>&1
python tools/probe.py > gpurun_out/${R}_probe.log 2>&1
ls -la gpurun_out/${R}_*
