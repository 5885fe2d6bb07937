timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 100 python tools/probe.py --levels 9 --reps 5 --cycles 20 2>&1 | grep -E "vcycle|half"
python tools/bench_rbgs.py --n 513 --iters 30 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rbgs513', d['us_per_full_sweep'], d['rbgs_gbs'], d['frac_of_measured_peak'], d['frac_of_8TBs_nominal'])"
python bench.py --no-cpu-baseline | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['e2e']['seconds_per_solve'])"
