set -x
mkdir -p gpurun_out/r2z
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z/smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/r2z/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2z/pytest_gpu.log
python bench.py > gpurun_out/r2z/bench_default.json 2> gpurun_out/r2z/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z/bench_reference.json 2> gpurun_out/r2z/bench_reference.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2z/launches_c3_snp.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-k-sweep --no-digest > gpurun_out/r2z/ncu_launch.log 2>&1; echo "ncu list rc=$?"
for kn in k_roll_partition k_bucket_dedup k_verify_links k_classify_pieces k_edge_write; do
  ncu --set full --clock-control none --import-source on -k regex:$kn -s 1 -c 1 -o gpurun_out/r2z/full_$kn python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-k-sweep --no-digest > gpurun_out/r2z/ncu_$kn.log 2>&1; echo "ncu $kn rc=$?"
done
cuobjdump -sass superpang_b200/libsuperpang_b200.so 2>/dev/null | grep -E "UBLKCP|SYNCS|UTMA|LDGSTS" | sort | uniq -c | sort -rn | head -12 > gpurun_out/r2z/sass_async_copy_mnemonics.txt
exit 0
