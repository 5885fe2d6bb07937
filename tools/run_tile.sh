run() { env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py "${ARGS[@]}" 2>&1 | grep -E "metric" | tail -1; }
ARGS=(--problem strong1025 --gpus 8 --steps 10 --warmup 3)
echo "strong p2p graph"; run A=1 | tee gpurun_out/strong8_p2p_graph.log | cut -c1-220
ARGS=(--gpus 8 --steps 10 --warmup 3)
echo "weak p2p graph"; run A=1 | tee gpurun_out/weak8_p2p_graph.log | cut -c1-220
