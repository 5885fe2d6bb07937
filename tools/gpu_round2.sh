mkdir -p gpurun_out
N=${NGPU:-8}
R=r02C
OLD=$PWD/multigrid_parallel_b200/libmgb_oldfence.so
run() { # name, env...
  name=$1; shift
  env "$@" DIST_CHECK_CASES=0,1,2 DIST_CHECK_REPEAT=3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_check.py > gpurun_out/${R}_${name}.log 2>&1
  echo "$name rc $? ok-lines $(grep -c '^ok' gpurun_out/${R}_${name}.log) $(grep -m1 'AssertionError' gpurun_out/${R}_${name}.log | cut -c1-150)"
}
run default MGB_X=0
run nopdl MGB_PDL=0
run oldfence MGB_LIB=$OLD
run oldfence_nopdl MGB_LIB=$OLD MGB_PDL=0
run notailsmem MGB_TAIL_SMEM=0
