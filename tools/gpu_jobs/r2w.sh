set -x
mkdir -p gpurun_out/r2w2
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2w2/bench_n$N.json 2> gpurun_out/r2w2/bench_n$N.err; echo "bench N=$N rc=$?"
done
exit 0
