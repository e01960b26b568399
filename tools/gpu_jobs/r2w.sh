set -x
mkdir -p gpurun_out/r2w
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2w/bench_n$N.json 2> gpurun_out/r2w/bench_n$N.err; echo "bench N=$N rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload c4s --steps 3 --warmup 2 > gpurun_out/r2w/bench_c4s_n8.json 2> gpurun_out/r2w/bench_c4s_n8.err; echo "c4s rc=$?"
exit 0
