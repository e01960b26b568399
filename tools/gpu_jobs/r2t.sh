set -x
mkdir -p gpurun_out/r2t
N=$1; W=$2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --steps 3 --warmup 2 > gpurun_out/r2t/bench_${W}_n$N.json 2> gpurun_out/r2t/bench_${W}_n$N.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2t/bench_${W}_n$N.err
nvidia-smi --query-gpu=memory.used --format=csv | head -3
exit 0
