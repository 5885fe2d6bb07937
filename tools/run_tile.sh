set -x
python bench.py > gpurun_out/bench_r01d.log 2>&1; tail -1 gpurun_out/bench_r01d.log | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01d.log 2>&1; tail -1 gpurun_out/bench_ref_r01d.log | cut -c1-200
python tools/probe_tile.py --levels 9 --reps 2 --which hs,pc,norm,rr > gpurun_out/pt_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_half_sweep_pipe|k_tile|k_prolong_correct8" -c 12 -o gpurun_out/prof_r01d python tools/probe_tile.py --levels 9 --reps 2 --which hs,pc,norm,rr > gpurun_out/ncu_r01d.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
