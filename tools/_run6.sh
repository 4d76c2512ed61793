bash tools/ab_variants.sh within 1000 nostage stage4 stage16 2>&1
timeout 600 python -m pytest tests/test_gpu_tensor_filter.py tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -4
for s in 0 100 200 334; do echo -n "subpass $s: "; NDA_SUBPASS_FRAMES=$s python bench.py --steps 100 --no-extras --sustain-s 0.2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'launches', d['gpu_launches'])"; done
python tools/profile_run.py dist > gpurun_out/r02_plain_dist.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:distances_kernel -c 2 -f -o gpurun_out/r02_prof_dist python tools/profile_run.py dist > gpurun_out/r02_ncu_dist.log 2>&1
cat gpurun_out/r02_plain_dist.log
