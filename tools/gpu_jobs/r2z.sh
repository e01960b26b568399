set -x
mkdir -p gpurun_out/r2z
python bench.py > gpurun_out/r2z/bench_default.json 2> gpurun_out/r2z/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z/bench_reference.json 2> gpurun_out/r2z/bench_reference.err; echo "ref rc=$?"
python bench.py --no-cpu-baseline --no-k-sweep --workload c2 > gpurun_out/r2z/bench_c2.json 2> gpurun_out/r2z/bench_c2.err; echo "c2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2z/launches_c3_snp.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-k-sweep --no-digest > gpurun_out/r2z/ncu_launch.log 2>&1; echo "ncu list rc=$?"
for kn in k_roll_partition k_bucket_dedup k_classify_pieces; do
  ncu --set full --clock-control none --import-source on -k regex:$kn -s 1 -c 1 -o /tmp/full_$kn python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-k-sweep --no-digest > gpurun_out/r2z/ncu_$kn.log 2>&1; echo "ncu $kn rc=$?"
  ncu -i /tmp/full_$kn.ncu-rep --page raw --csv > gpurun_out/r2z/ncu_full_${kn}_raw.csv 2>/dev/null
  rm -f /tmp/full_$kn.ncu-rep
done
cuobjdump -sass superpang_b200/libsuperpang_b200.so 2>/dev/null | grep -oE "UBLKCP[A-Z0-9_.]*|SYNCS[A-Z0-9_.]*|UTMA[A-Z0-9_.]*" | sort | uniq -c | sort -rn | head -12 > gpurun_out/r2z/sass_async_copy_mnemonics.txt
du -sh gpurun_out/r2z
exit 0
