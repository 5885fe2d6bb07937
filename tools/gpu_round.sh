mkdir -p gpurun_out
R=r02F
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${R}_pytest.log; tail -3 gpurun_out/${R}_pytest.log
