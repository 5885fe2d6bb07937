set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/dist_check.py > gpurun_out/r02f_dist2.log 2>&1; echo "dist_check rc $?" >> gpurun_out/r02f_dist2.log
tail -2 gpurun_out/r02f_dist2.log
python tools/dist_stages.py > gpurun_out/r02f_stages.log 2>&1
for p in weak strong1025 config5; do timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/dist_stages.py --problem $p >> gpurun_out/r02f_stages.log 2>&1; done
cat gpurun_out/r02f_stages.log | grep -v "^\*\|OMP_NUM\|^$"
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err; echo "bench rc $?" >> gpurun_out/r02f_bench_n2.err
