set -x
mkdir -p gpurun_out/r2m
python -m pytest tests -m gpu -x -q > gpurun_out/r2m/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py --no-cpu-baseline > gpurun_out/r2m/bench_n1.json 2> gpurun_out/r2m/bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2m/bench_n1.err
