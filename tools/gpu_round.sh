set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s -k "dropin or fmg or lu" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02c_pytest.log
tail -30 gpurun_out/r02c_pytest.log
