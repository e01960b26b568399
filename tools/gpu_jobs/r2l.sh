set -x
mkdir -p gpurun_out/r2l
python -m pytest tests -m gpu -x -q > gpurun_out/r2l/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py > gpurun_out/r2l/bench_default.json 2> gpurun_out/r2l/bench_default.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2l/launches_c3_snp.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2l/ncu_launch.log 2>&1; echo "ncu list rc=$?"
for kn in k_roll_partition k_partition2 k_bucket_dedup; do
  ncu --set full --clock-control none --import-source on -k regex:$kn -s 1 -c 1 -o gpurun_out/r2l/full_$kn python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2l/ncu_$kn.log 2>&1; echo "ncu $kn rc=$?"
done
ls -la gpurun_out/r2l
