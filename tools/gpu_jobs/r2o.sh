set -x
mkdir -p gpurun_out/r2o
python -m pytest tests -m gpu -x -q > gpurun_out/r2o/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py --no-cpu-baseline > gpurun_out/r2o/bench_n1.json 2> gpurun_out/r2o/bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2o/bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2o/bench_n2.json 2> gpurun_out/r2o/bench_n2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2o/bench_n2.err
