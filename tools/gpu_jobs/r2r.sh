set -x
mkdir -p gpurun_out/r2r
N=$1
CUDA_VISIBLE_DEVICES=0 python -m pytest tests -m gpu -x -q > gpurun_out/r2r/pytest_gpu.log 2>&1; echo "pytest rc=$?"
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu-baseline > gpurun_out/r2r/bench_n1.json 2> gpurun_out/r2r/bench_n1.err; echo "bench rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2r/bench_n$N.json 2> gpurun_out/r2r/bench_n$N.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2r/bench_n$N.err
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu-baseline --mode block > gpurun_out/r2r/bench_n1_block.json 2> gpurun_out/r2r/bench_n1_block.err; echo "bench rc=$?"
exit 0
