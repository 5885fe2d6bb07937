set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01e.log 2>&1; tail -1 gpurun_out/bench_r01e.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01e.log 2>&1
python tools/bench_rbgs.py --n 257 --iters 100 > gpurun_out/rbgs_257.json; python tools/bench_rbgs.py --n 513 --iters 30 > gpurun_out/rbgs_513.json
python bench.py --problem strong1025 --gpus 1 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/strong1.log
python tools/probe_tile.py --levels 9 --reps 2 --which hs,pc,norm,rr > gpurun_out/pt_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_half_sweep_pipe|k_prolong_correct8" -c 12 -o gpurun_out/prof_r01e python tools/probe_tile.py --levels 9 --reps 2 --which hs,pc,norm,rr > gpurun_out/ncu_r01e.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
