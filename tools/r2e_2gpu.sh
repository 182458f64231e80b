#!/bin/bash
mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 200 python -m pytest tests/test_gpu_ddp_nccl.py -q -m gpu -s > gpurun_out/r2e_ddp_nccl_test.log 2>&1; tail -6 gpurun_out/r2e_ddp_nccl_test.log
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-cudnn > gpurun_out/r2e_bench_resunet_1gpu.json 2>/dev/null; tail -c 250 gpurun_out/r2e_bench_resunet_1gpu.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-cudnn > gpurun_out/r2e_bench_resunet_2gpu.json 2> gpurun_out/r2e_2gpu.err; tail -c 250 gpurun_out/r2e_bench_resunet_2gpu.json; tail -3 gpurun_out/r2e_2gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-cudnn --workload swin_unetr_amos_128 > gpurun_out/r2e_bench_swin_2gpu.json 2>/dev/null; tail -c 250 gpurun_out/r2e_bench_swin_2gpu.json
python - <<'PY'
import json
def val(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        return None
a, b = val('gpurun_out/r2e_bench_resunet_1gpu.json'), val('gpurun_out/r2e_bench_resunet_2gpu.json')
if a and b:
    print("resunet: 1 GPU %.2f ms, 2 GPUs %.2f ms, efficiency %.3f; clocks %s" % (a['ms_per_step'], b['ms_per_step'], b['value'] / (2 * a['value']), b['clocks']))
PY
