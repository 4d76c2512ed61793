timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
(time python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err); tail -3 gpurun_out/r2_bench1.err
(time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench1_ref.json 2>/dev/null); cat gpurun_out/r2_bench1_ref.json | cut -c1-300
bash tools/r02_profile.sh > gpurun_out/r02_profile.log 2>&1; tail -12 gpurun_out/r02_profile.log
python tools/profile_run.py dist > gpurun_out/r02_plain_dist.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:distances_kernel -c 2 -f -o gpurun_out/r02_prof_dist python tools/profile_run.py dist > gpurun_out/r02_ncu_dist.log 2>&1
