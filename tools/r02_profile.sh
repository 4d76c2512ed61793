#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE B200; every ncu run follows the same command exiting 0 without ncu):
#   r02_prof_tc    --set full: pack / within_tc / rdf_tc (ungrouped + grouped) / distances (config 4 at full size)
#   r02_prof_fp32  --set full: the packed-FP32 within / RDF kernels (NDA_FILTER=fold2a1), the north-star design
#   r02_launches_step.csv  every launch of two resident bench steps with its duration and DRAM bytes
# Summaries for profiles/: python tools/ncu_summary.py gpurun_out/r02_prof_tc.ncu-rep profiles/r02_prof_tc   (etc.)
mkdir -p gpurun_out
python tools/profile_run.py all 1000 > gpurun_out/r02_plain_tc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"distances_kernel|within_tc_kernel|rdf_tc_kernel|pack_kernel" -c 14 -f \
    -o gpurun_out/r02_prof_tc python tools/profile_run.py all 1000 > gpurun_out/r02_ncu_tc.log 2>&1
NDA_FILTER=fold2a1 python tools/profile_run.py all 200 > gpurun_out/r02_plain_fp32.log 2>&1 &&
NDA_FILTER=fold2a1 ncu --set full --clock-control none --import-source on -k regex:"rdf_kernel|within_kernel" -c 6 -f \
    -o gpurun_out/r02_prof_fp32 python tools/profile_run.py all 200 > gpurun_out/r02_ncu_fp32.log 2>&1
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/r02_launches_step.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
cat gpurun_out/r02_plain_tc.log gpurun_out/r02_plain_fp32.log
ls -la gpurun_out/*.ncu-rep
