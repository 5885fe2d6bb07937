mkdir -p gpurun_out /tmp/dd && cd /tmp/dd && ln -sf /dev/null diff2.vtk
EXE=$GRAFT_REPO_ROOT/multigrid_parallel_b200/compat/_build/test_mg_3d_gpu
for t in 16 16 16 4 4; do
  OMP_NUM_THREADS=$t MGB_VTK_TIMING=1 $EXE 3 9 2 2>&1 | grep -E "Overall time|writeOutputData" | tr '\n' ' '; echo " threads=$t"
done
ulimit -l; nproc
