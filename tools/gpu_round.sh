mkdir -p gpurun_out
R=r02p
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${R}_pytest.log; tail -5 gpurun_out/${R}_pytest.log
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc $?"; tail -c 400 gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_n1.json 2> gpurun_out/${R}_bench_reference_n1.err; echo "ref rc $?"
python tools/cycle_case.py --cycles 2 > gpurun_out/${R}_case.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python tools/cycle_case.py --cycles 2 > gpurun_out/${R}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_tile|k_half_sweep_pipe|k_residual_restrict|k_prolong_correct8' --launch-skip 60 -c 40 -o gpurun_out/${R}_prof -f python tools/cycle_case.py --cycles 2 > gpurun_out/${R}_ncu_full.log 2>&1; echo "ncu rc $?"; tail -3 gpurun_out/${R}_ncu_full.log
ls -la gpurun_out/${R}_*
