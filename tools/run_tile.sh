set -x
python bench.py > gpurun_out/bench_r01b.log 2>&1; tail -1 gpurun_out/bench_r01b.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01b.log 2>&1; tail -1 gpurun_out/bench_ref_r01b.log
python tools/probe.py --levels 9 --reps 2 --cycles 1 > gpurun_out/probe_plain_r01b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_half_sweep|k_tile|k_prolong_correct8" -c 10 -o gpurun_out/prof_r01b python tools/probe.py --levels 9 --reps 2 --cycles 1 > gpurun_out/ncu_r01b.log 2>&1
