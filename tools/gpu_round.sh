mkdir -p gpurun_out
R=r02v
python -m pytest tests/test_gpu_vtk.py tests/test_gpu_dropin.py tests/test_gs_lex.py -m gpu -x -q -s > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${R}_pytest.log; grep -v "^$" gpurun_out/${R}_pytest.log | tail -25
