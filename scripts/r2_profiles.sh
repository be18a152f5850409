#!/bin/bash
# round-2 evidence run on one B200: launch list of the bench command, --set full captures of the three hot kernels,
# DRAM traffic of the pooling / GEMM kernels on the three rigs
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_raw.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-config4 --no-config5 --no-variants > gpurun_out/bench_under_ncu_r2.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_tile_kernel|tile_build_kernel|ygemm_compact_kernel" -s 3 -c 3 -o gpurun_out/r2_fwd_kernels -f python scripts/quick_time.py MultiviewC 4 0 > gpurun_out/ncu_r2_fwd.log 2>&1
echo "full capture rc=$?"
for wl in MultiviewC MultiviewX Wildtrack; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"pool_tile_kernel|ygemm_compact_kernel|tile_build_kernel" -s 3 -c 3 --csv --log-file gpurun_out/r2_traffic_$wl.csv python scripts/quick_time.py $wl 4 0 > /dev/null 2>&1
done
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"pool_tile_kernel|ygemm_compact_bf16_kernel" -s 2 -c 2 --csv --log-file gpurun_out/r2_traffic_bf16_MultiviewC.csv python scripts/quick_time.py MultiviewC 4 4 > /dev/null 2>&1
ls -la gpurun_out/r2_*
