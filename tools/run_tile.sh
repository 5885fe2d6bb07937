timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 100 python tools/probe.py --levels 9 --reps 5 --cycles 20 2>&1 | grep -E "vcycle|half|prolong|resid"
timeout 100 python tools/probe.py --levels 10 --reps 2 --cycles 5 2>&1 | grep -E "vcycle|half|prolong|resid"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
