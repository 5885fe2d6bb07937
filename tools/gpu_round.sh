set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02b_pytest.log
python tools/cycle_case.py --lu 8 > gpurun_out/r02b_case.log 2>&1
python tools/cycle_case.py --lu 1 --levels 5 >> gpurun_out/r02b_case.log 2>&1
python tools/cycle_case.py --coarse 3 9 9 --levels 8 --lu 4 >> gpurun_out/r02b_case.log 2>&1
python tools/cycle_case.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'^k_tile$|k_tile_prolong_one' -s 12 -c 8 -o gpurun_out/r02b_prof python tools/cycle_case.py > gpurun_out/r02b_ncu_full.log 2>&1
python tools/cycle_case.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 200 --csv --log-file gpurun_out/r02b_launches.csv python tools/cycle_case.py > gpurun_out/r02b_ncu_list.log 2>&1
tail -3 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_case.log; tail -5 gpurun_out/r02b_ncu_full.log
