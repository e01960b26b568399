set -x
mkdir -p gpurun_out/r2y
python tools/tune_env.py --workload c3 --k 301 --out gpurun_out/r2y/k301.jsonl "-" "SPB_D_EXACT=1" > gpurun_out/r2y/k301.log 2>&1; echo "tune rc=$?"
python tools/tune_env.py --workload c3 --k 101 --out gpurun_out/r2y/k101.jsonl "-" "SPB_D_EXACT=1" > gpurun_out/r2y/k101.log 2>&1; echo "tune rc=$?"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants" > gpurun_out/r2y/pytest_variants.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2y/pytest_variants.log
exit 0
