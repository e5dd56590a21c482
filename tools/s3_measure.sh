# session-3 evidence: bench line, launch list of the same command, DRAM traffic per step, ncu --set full of the kernels
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_s3_final.json 2> gpurun_out/bench_s3_final.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s3_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s3_ncu_bench.log 2>&1; echo "launch list rc=$?"
for p in 0 1; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/s3_traffic_$p.csv python tools/profile_recon.py 1024 1 $p > gpurun_out/s3_traffic_$p.log 2>&1; echo "traffic $p rc=$?"
done
# full sets: 256 streams; the replay launches come after the 16 (dense) / 32 (realistic) recording launches
timeout 600 ncu --set full --import-source on --clock-control none -k regex:recon_band -s 17 -c 2 -o gpurun_out/s3_band_dense -f \
  python tools/profile_recon.py 256 1 0 > gpurun_out/s3_full_dense.log 2>&1; echo "full dense rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:recon_ -s 34 -c 4 -o gpurun_out/s3_map_record_real -f \
  python tools/profile_recon.py 256 1 1 > gpurun_out/s3_full_real.log 2>&1; echo "full realistic rc=$?"
ls -la gpurun_out/s3_*
