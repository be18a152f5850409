#!/bin/bash
# round 2, run 1: new tile pooling -- sanitizer on a small case, the new parity tests, A/B bench (tile vs list pooling)
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/sanitize.log 2>&1; echo "sanitize rc=$?"; tail -5 gpurun_out/sanitize.log
timeout 900 python -m pytest tests/test_gpu_frame_parity.py -x -q -m gpu -s 2>&1 | tail -40
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_tile.json 2> gpurun_out/bench_tile.err; tail -c 1500 gpurun_out/bench_tile.json; tail -3 gpurun_out/bench_tile.err
VFA_POOL_TILE=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_list.json 2> gpurun_out/bench_list.err; tail -c 600 gpurun_out/bench_list.json
