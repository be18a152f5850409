timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
for ch in 1 2 4; do
timeout 200 $TR bench.py --gpus 2 --mode views --view-chunk $ch --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/views2_c$ch.err | tee gpurun_out/views2_c$ch.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('views chunk', $ch, d['value'], d['ms_per_step'])"
done
timeout 200 $TR bench.py --gpus 2 --mode views --view-chunk 1 --backward --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/views2_bwd.err | tee gpurun_out/views2_bwd.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('views bwd', d['value'], d['ms_per_step'])"
timeout 200 python bench.py --mode views --view-chunk 4 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/views1.err | tee gpurun_out/views1.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('views 1gpu', d['value'], d['ms_per_step'])"
tail -3 gpurun_out/views2_c1.err
