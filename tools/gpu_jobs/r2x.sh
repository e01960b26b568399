set -x
mkdir -p gpurun_out/r2x
python -m pytest tests -m gpu -x -q > gpurun_out/r2x/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2x/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/r2x/bench_n1.json 2> gpurun_out/r2x/bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2x/bench_n1.err
python tools/tune_env.py --workload c3 --k 101 --out gpurun_out/r2x/k101.jsonl "-" "SPB_D_EXACT=1" > gpurun_out/r2x/k101.log 2>&1; echo "tune rc=$?"
exit 0
