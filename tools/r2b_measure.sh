#!/bin/bash
# Round-2 (second session) measurement pass on one B200: full GPU suite, bench lines, augmentation bench, launch lists.
# Everything lands in gpurun_out/; the summaries worth keeping are copied to profiles/ afterwards.
mkdir -p gpurun_out
set -x
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/r2b_gpu_tests.log 2>&1; tail -3 gpurun_out/r2b_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; tail -3 gpurun_out/r2b_smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_resunet_acdc_128.json 2> gpurun_out/r2b_bench_resunet.err; tail -c 600 gpurun_out/r2b_bench_resunet_acdc_128.json
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-cudnn --workload swin_unetr_amos_128 > gpurun_out/r2b_bench_swin_unetr_amos_128.json 2> gpurun_out/r2b_bench_swin.err; tail -c 400 gpurun_out/r2b_bench_swin_unetr_amos_128.json
B200SEG_WINATTN_MMA=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-cudnn --workload swin_unetr_amos_128 > gpurun_out/r2b_bench_swin_unetr_amos_128_cudacore_attention.json 2>/dev/null; tail -c 300 gpurun_out/r2b_bench_swin_unetr_amos_128_cudacore_attention.json
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-cudnn --workload medformer_bcv_96 > gpurun_out/r2b_bench_medformer_bcv_96.json 2>/dev/null; tail -c 300 gpurun_out/r2b_bench_medformer_bcv_96.json
timeout 200 python tools/aug_bench.py > gpurun_out/r2b_aug_bench.json 2> gpurun_out/r2b_aug_bench.err; cat gpurun_out/r2b_aug_bench.json | head -80
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2b_launches_swin_unetr.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-cudnn --workload swin_unetr_amos_128 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2b_launches_swin_unetr.csv 3 30 > gpurun_out/r2b_launch_summary_swin_unetr.txt; head -20 gpurun_out/r2b_launch_summary_swin_unetr.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:win_attn_fwd_mma -c 1 -o gpurun_out/r2b_ncu_win_attn_fwd_mma -f python bench.py --steps 1 --warmup 1 --no-cpu --no-cudnn --workload swin_unetr_amos_128 > /dev/null 2>&1
ncu -i gpurun_out/r2b_ncu_win_attn_fwd_mma.ncu-rep --page details > gpurun_out/r2b_ncu_full_win_attn_fwd_mma.txt 2>&1; grep -E "Duration|Executed Ipc Active|Issue Slots Busy|Registers Per|Theoretical Occ|Achieved Occ" gpurun_out/r2b_ncu_full_win_attn_fwd_mma.txt | head
timeout 200 ncu --set full --clock-control none -k regex:aug_ -c 6 -o gpurun_out/r2b_ncu_aug -f python tools/aug_bench.py > /dev/null 2>&1
ncu -i gpurun_out/r2b_ncu_aug.ncu-rep --page details > gpurun_out/r2b_ncu_full_aug.txt 2>&1; grep -E "aug_.*kernel|Duration|DRAM Throughput|Memory Throughput" gpurun_out/r2b_ncu_full_aug.txt | head -30
ls -la gpurun_out | head -40
