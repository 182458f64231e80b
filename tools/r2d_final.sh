#!/bin/bash
# final pass of the round: full GPU suite, smoke, one bench line per workload, launch lists of the three model families
mkdir -p gpurun_out
set -x
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/r2d_gpu_tests.log 2>&1; tail -4 gpurun_out/r2d_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; tail -3 gpurun_out/r2d_smoke.log
timeout 400 python bench.py > gpurun_out/r2d_bench_resunet_acdc_128.json 2> gpurun_out/r2d_bench.err; tail -c 300 gpurun_out/r2d_bench_resunet_acdc_128.json
for w in resunet_iso_128 resunet_kits_160 medformer_bcv_96 swin_unetr_amos_128; do
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-cudnn --workload $w > gpurun_out/r2d_bench_$w.json 2>/dev/null; tail -c 200 gpurun_out/r2d_bench_$w.json; echo
done
timeout 200 python tools/aug_bench.py > gpurun_out/r2d_aug_bench.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2d_launches_resunet_acdc128.csv python bench.py --steps 6 --warmup 1 --no-cpu --no-cudnn > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2d_launches_resunet_acdc128.csv 7 30 > gpurun_out/r2d_launch_summary_resunet_acdc128.txt; head -12 gpurun_out/r2d_launch_summary_resunet_acdc128.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2d_launches_swin_unetr.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-cudnn --workload swin_unetr_amos_128 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2d_launches_swin_unetr.csv 3 30 > gpurun_out/r2d_launch_summary_swin_unetr.txt; head -12 gpurun_out/r2d_launch_summary_swin_unetr.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/r2d_launches_medformer.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-cudnn --workload medformer_bcv_96 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2d_launches_medformer.csv 3 30 > gpurun_out/r2d_launch_summary_medformer.txt; head -14 gpurun_out/r2d_launch_summary_medformer.txt
