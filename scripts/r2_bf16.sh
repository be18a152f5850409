#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_frame_parity.py -x -q -m gpu -s -k "bf16_mma or channels_last_output" 2>&1 | tail -8
echo "--- timing: flags 0 vs BF16_MMA (4)"
timeout 120 python scripts/quick_time.py MultiviewC 4 0 2>&1 | tail -1
timeout 120 python scripts/quick_time.py MultiviewC 4 4 2>&1 | tail -1
