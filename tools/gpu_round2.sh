# Multi-GPU evidence of a round (gpurun --gpus N -- 'NGPU=N R=r03a bash tools/gpu_round2.sh'):
# bitwise parity of the partitioned solver (tests/dist_check.py), the bench line with its parity /
# strong1025 / config5 objects, per-stage tables.
mkdir -p gpurun_out
N=${NGPU:-2}
R=${R:-rXX}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_check.py > gpurun_out/${R}_dist$N.log 2>&1; echo "dist_check rc $?" >> gpurun_out/${R}_dist$N.log
tail -2 gpurun_out/${R}_dist$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench_n$N.json 2> gpurun_out/${R}_bench_n$N.err; echo "bench rc $?" >> gpurun_out/${R}_bench_n$N.err
tail -c 200 gpurun_out/${R}_bench_n$N.err
rm -f gpurun_out/${R}_stages$N.log
for p in weak strong1025; do timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/dist_stages.py --problem $p >> gpurun_out/${R}_stages$N.log 2>&1; done
grep "on $N GPU" gpurun_out/${R}_stages$N.log
