mkdir -p gpurun_out
python -m pytest tests/test_gpu_vtk.py -m gpu -x -q > gpurun_out/r02G_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02G_pytest.log; tail -3 gpurun_out/r02G_pytest.log
