nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -s -k "fp16" 2>&1 | grep -E "^E  |passed|failed|tc probe" | head -12
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
python tools/profile_run.py dist 2>&1 | tail -1
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err)
tail -4 gpurun_out/r2_bench2.err; wc -c gpurun_out/r2_bench2.json
