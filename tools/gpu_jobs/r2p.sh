set -x
mkdir -p gpurun_out/r2p
python bench.py --no-cpu-baseline > gpurun_out/r2p/bench_n1.json 2> gpurun_out/r2p/bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2p/bench_n1.err
python -m pytest tests -m gpu -x -q > gpurun_out/r2p/pytest_gpu.log 2>&1; echo "pytest rc=$?"
for kn in k_assign_ids k_junction_emit k_origin_pairs k_emit_darts; do
ncu --set full --clock-control none --import-source on -k regex:$kn -s 1 -c 1 -o gpurun_out/r2p/full_$kn python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-digest > gpurun_out/r2p/ncu_$kn.log 2>&1; echo "ncu rc=$?"
done
