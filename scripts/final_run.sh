#!/bin/bash
# round-end evidence run on one B200: GPU tests, the default bench line, the ncu launch list of the same command and a
# --set full capture of the dominant kernel
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 400 python bench.py > gpurun_out/bench_r1_list.json 2> gpurun_out/bench_r1_list.err; tail -c 600 gpurun_out/bench_r1_list.json; tail -3 gpurun_out/bench_r1_list.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1_list.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pool_list_kernel|qlist_build_kernel" -s 2 -c 2 -o gpurun_out/r1_pool_list -f python scripts/quick_time.py MultiviewC 4 0 > gpurun_out/ncu_pool_list.log 2>&1
ls -la gpurun_out/r1_pool_list.ncu-rep
