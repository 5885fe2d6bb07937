mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "operators or fullsize" > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02k_pytest.log; tail -3 gpurun_out/r02k_pytest.log
python tools/probe.py > gpurun_out/r02k_probe.log 2>&1; cat gpurun_out/r02k_probe.log
