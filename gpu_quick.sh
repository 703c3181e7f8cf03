N=$1
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/bench_full_n$N.err | tee gpurun_out/bench_full_n$N.json | cut -c1-1200
echo "wall $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/bench_full_n$N.err
