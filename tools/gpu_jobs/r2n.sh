set -x
mkdir -p gpurun_out/r2n
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2n/bench_n$N.json 2> gpurun_out/r2n/bench_n$N.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2n/bench_n$N.err
[ "$2" = "test" ] && python -m pytest tests/test_gpu_parity.py -q -x -k multi_rank > gpurun_out/r2n/pytest_mg$N.log 2>&1; echo "pytest mg rc=$?"
tail -5 gpurun_out/r2n/pytest_mg$N.log
