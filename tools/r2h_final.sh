#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/r2h_gpu_tests.log 2>&1; tail -4 gpurun_out/r2h_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -3 gpurun_out/r2h_smoke.log
timeout 400 python bench.py > gpurun_out/r2h_bench_resunet_acdc_128.json 2> gpurun_out/r2h_bench.err; tail -c 200 gpurun_out/r2h_bench_resunet_acdc_128.json; echo
for w in resunet_iso_128 resunet_kits_160 medformer_bcv_96 swin_unetr_amos_128; do
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-cudnn --workload $w > gpurun_out/r2h_bench_$w.json 2>/dev/null
done
for w in resunet_acdc_128 resunet_iso_128 resunet_kits_160 medformer_bcv_96 swin_unetr_amos_128; do python -c "
import json
d=json.loads(open('gpurun_out/r2h_bench_$w.json').read().strip().splitlines()[-1])
print('$w', 'ms', round(d['ms_per_step'],2), 'val', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'clk', d['clocks']['samples'], d.get('torch_cudnn_same_gpu',{}).get('ms_per_step_by_variant'))"; done
timeout 100 python tools/layer_times.py > gpurun_out/r2h_layer_times.txt 2>&1; cat gpurun_out/r2h_layer_times.txt
B200SEG_CONV_ROW_ALL=1 timeout 100 python tools/layer_times.py > gpurun_out/r2h_layer_times_conv_row_all.txt 2>&1; cut -c1-77 gpurun_out/r2h_layer_times_conv_row_all.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2h_launches_resunet_acdc128.csv python bench.py --steps 6 --warmup 1 --no-cpu --no-cudnn > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2h_launches_resunet_acdc128.csv 7 30 > gpurun_out/r2h_launch_summary_resunet_acdc128.txt; head -8 gpurun_out/r2h_launch_summary_resunet_acdc128.txt
