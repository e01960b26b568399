set -x
mkdir -p gpurun_out/r2q
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2q/bench_n$N.json 2> gpurun_out/r2q/bench_n$N.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2q/bench_n$N.err
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu-baseline > gpurun_out/r2q/bench_n1.json 2> gpurun_out/r2q/bench_n1.err; echo "bench rc=$?"
exit 0
