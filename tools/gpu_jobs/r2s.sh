set -x
mkdir -p gpurun_out/r2s
python -m pytest tests -m gpu -x -q > gpurun_out/r2s/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2s/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/r2s/bench_n1.json 2> gpurun_out/r2s/bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2s/bench_n1.err
