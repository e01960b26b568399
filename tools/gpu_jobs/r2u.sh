set -x
mkdir -p gpurun_out/r2u
python bench.py --no-cpu-baseline --no-k-sweep > gpurun_out/r2u/bench_n1.json 2> gpurun_out/r2u/bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2u/bench_n1.err
python -m pytest tests -m gpu -x -q > gpurun_out/r2u/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2u/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2u/bench_n2.json 2> gpurun_out/r2u/bench_n2.err; echo "bench rc=$?"
SPB_PART_L2MAX=4 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-cpu-baseline --steps 2 > gpurun_out/r2u/bench_n2_twopass.json 2> gpurun_out/r2u/bench_n2_twopass.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2u/bench_n2_twopass.err
exit 0
