#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_swin.py tests/test_gpu_medformer_ops.py tests/test_gpu_medformer.py -q -m gpu --timeout 120 --timeout-method thread > gpurun_out/r2c_tests.log 2>&1; tail -15 gpurun_out/r2c_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-cudnn --workload swin_unetr_amos_128 > gpurun_out/r2c_bench_swin_unetr_amos_128.json 2> gpurun_out/r2c_swin.err; tail -c 300 gpurun_out/r2c_bench_swin_unetr_amos_128.json; tail -3 gpurun_out/r2c_swin.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-cudnn --workload medformer_bcv_96 > gpurun_out/r2c_bench_medformer_bcv_96.json 2>/dev/null; tail -c 300 gpurun_out/r2c_bench_medformer_bcv_96.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/r2c_launches_swin_unetr.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-cudnn --workload swin_unetr_amos_128 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2c_launches_swin_unetr.csv 3 30 > gpurun_out/r2c_launch_summary_swin_unetr.txt; head -24 gpurun_out/r2c_launch_summary_swin_unetr.txt
