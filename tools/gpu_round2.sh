set -x
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_check.py > gpurun_out/r02h_dist$N.log 2>&1; echo "dist_check rc $?" >> gpurun_out/r02h_dist$N.log
tail -2 gpurun_out/r02h_dist$N.log
rm -f gpurun_out/r02h_stages$N.log
for p in weak strong1025; do timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/dist_stages.py --problem $p >> gpurun_out/r02h_stages$N.log 2>&1; done
cat gpurun_out/r02h_stages$N.log | grep -v "^\*\|OMP_NUM\|^$\|NCCL version"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err; echo "bench rc $?" >> gpurun_out/r02h_bench_n$N.err
tail -c 300 gpurun_out/r02h_bench_n$N.err
