mkdir -p gpurun_out
N=${NGPU:-8}
R=r02D
DIST_CHECK_CASES=0,2,3,6 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_check.py > gpurun_out/${R}_dist$N.log 2>&1; echo "dist_check rc $?" >> gpurun_out/${R}_dist$N.log
echo "ok-lines $(grep -c '^ok' gpurun_out/${R}_dist$N.log)"; tail -2 gpurun_out/${R}_dist$N.log
