mkdir -p gpurun_out
N=${NGPU:-2}
R=r02w
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_check.py > gpurun_out/${R}_dist$N.log 2>&1; echo "dist_check rc $?" >> gpurun_out/${R}_dist$N.log
tail -2 gpurun_out/${R}_dist$N.log
L=gpurun_out/${R}_sweep$N.log; rm -f $L
for pdl in 1 0; do
echo "== MGB_PDL=$pdl" >> $L
MGB_PDL=$pdl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/dist_sweep.py --problem weak --cases 16:-1 >> $L 2>&1
MGB_PDL=$pdl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/dist_sweep.py --problem strong1025 --cases 16:-1 >> $L 2>&1
done
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" $L
