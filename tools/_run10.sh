nvidia-smi --query-gpu=name --format=csv,noheader | wc -l; nproc; free -g | head -2
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_host_api.py -m gpu -q 2>&1 | tail -3
N=$(nvidia-smi --query-gpu=name --format=csv,noheader | wc -l)
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err)
tail -4 gpurun_out/r2_bench_n$N.err; wc -c gpurun_out/r2_bench_n$N.json
