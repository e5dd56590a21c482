#!/bin/bash
# round-2 final evidence on one box: DRAM traffic per step (stamped), bench line, reference arm, launch list of the same command,
# BASELINE.md table, ncu --set full of the default dense kernel and of the realistic kernel pair
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in 0 1; do
  timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_traffic_$p.csv python tools/profile_recon.py 1024 1 $p > gpurun_out/r02_traffic_$p.log 2>&1; echo "traffic $p rc=$?"
done
python tools/ncu_traffic.py gpurun_out/r02_traffic_0.csv dense_1024 1 16 16 | tee gpurun_out/r02_traffic_summary.txt
python tools/ncu_traffic.py gpurun_out/r02_traffic_1.csv realistic_1024 2 32 16 | tee -a gpurun_out/r02_traffic_summary.txt
cp profiles/r02_traffic.json gpurun_out/r02_traffic.json
python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 python tools/baseline_table.py > gpurun_out/r02_baseline_table.json 2> gpurun_out/r02_baseline_table.err; echo "table rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:recon_band_kernel -s 17 -c 2 -o gpurun_out/r02_band_tile16_dense -f \
    python tools/profile_recon.py 256 1 0 > gpurun_out/r02_band_tile16_dense.log 2>&1; echo "full dense rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:recon_ -s 34 -c 4 -o gpurun_out/r02_map_record_real -f \
  python tools/profile_recon.py 256 1 1 > gpurun_out/r02_full_real.log 2>&1; echo "full realistic rc=$?"
ls -la gpurun_out/r02_*
