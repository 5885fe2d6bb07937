mkdir -p gpurun_out
N=${NGPU:-2}
L=gpurun_out/r02s_ab$N.log; rm -f $L
OLD=$PWD/multigrid_parallel_b200/libmgb_oldfence.so
PER=$((1048576 / N))
for rep in 1 2; do
for lib in new old; do
echo "== lib $lib" >> $L
if [ $lib = old ]; then export MGB_LIB=$OLD; else unset MGB_LIB; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/dist_sweep.py --problem weak --cases 16:-1,16:$PER >> $L 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/dist_sweep.py --problem strong1025 --cases 16:-1,16:$PER >> $L 2>&1
done
done
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" $L
