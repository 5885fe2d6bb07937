set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02a_pytest.log
python tools/cycle_case.py --lu 8 > gpurun_out/r02a_case.log 2>&1
python tools/cycle_case.py --lu 1 >> gpurun_out/r02a_case.log 2>&1
python tools/cycle_case.py --coarse 3 9 9 --levels 8 --lu 4 >> gpurun_out/r02a_case.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc $?" >> gpurun_out/r02a_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
python tools/cycle_case.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_tile_prolong_one|k_tile<' -s 12 -c 8 -o gpurun_out/r02a_prof python tools/cycle_case.py > gpurun_out/r02a_ncu_full.log 2>&1
python tools/cycle_case.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 260 --csv --log-file gpurun_out/r02a_launches.csv python tools/cycle_case.py > gpurun_out/r02a_ncu_list.log 2>&1
tail -3 gpurun_out/r02a_pytest.log; cat gpurun_out/r02a_case.log; tail -c 1500 gpurun_out/r02a_bench.err
