set -x
mkdir -p gpurun_out/r2v
python bench.py --no-cpu-baseline --no-k-sweep > gpurun_out/r2v/bench_n1.json 2> gpurun_out/r2v/bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2v/bench_n1.err
python -m pytest tests -m gpu -x -q > gpurun_out/r2v/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2v/pytest_gpu.log
python bench.py --no-cpu-baseline --no-k-sweep --mode block > gpurun_out/r2v/bench_n1_block.json 2> gpurun_out/r2v/bench_n1_block.err; echo "bench rc=$?"
python bench.py --no-cpu-baseline --no-k-sweep --workload c2 > gpurun_out/r2v/bench_c2.json 2> gpurun_out/r2v/bench_c2.err; echo "bench rc=$?"
exit 0
