"""GPU (needs >= 2 devices; skipped on a 1-GPU box): multi-rank parity ON HARDWARE over NCCL — the gradient a
DistributedDataParallel-wrapped b200seg model holds after backward equals the mean of the per-rank gradients of
the same model run without DDP (reference wrap: train_ddp.py:352-358), in fp32 and under AMP, and the
parameters of all ranks are identical after one optimiser step.  Run with `gpurun --gpus 2`; the log of that run is
committed under profiles/."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    import b200seg
    from b200seg.train import get_optimizer
    from oracle import unet3d as ounet
    from oracle.synth import make_volume
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    scale = [[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]]
    kernel = [[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]]
    base, classes = 16, 4
    sd = ounet.make_state_dict(ounet.unet_param_shapes(1, base, classes, kernel, "BasicBlock"), seed=5)
    w = torch.tensor([0.5, 1, 1, 1])
    img, lab = make_volume(1, 16, 64, 64, classes, seed=100 + rank)      # every rank its own sample
    img, lab = img.to(dev), lab.to(dev)
    out = {}
    for amp in (False, True):
        def build():
            n = b200seg.UNet(1, base, scale=scale, kernel_size=kernel, num_classes=classes, block="BasicBlock", norm="in")
            n.load_state_dict(sd)
            return n.to(dev)
        S = 1024.0 if amp else 1.0
        # (1) local gradients without DDP, averaged over ranks by an explicit all-reduce
        plain = build()
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            loss = b200seg.DiceCELoss(weight=w)(plain(img), lab)
        (loss * S).backward()
        mean = {}
        for k, p in plain.named_parameters():
            g = p.grad.detach().clone()
            dist.all_reduce(g)
            mean[k] = g / world
        # (2) the DDP bucket path (the reference's flags)
        net = DDP(build(), device_ids=[rank], find_unused_parameters=True)
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            loss = b200seg.DiceCELoss(weight=w)(net(img), lab)
        (loss * S).backward()
        worst = 0.0
        for k, p in net.module.named_parameters():
            assert p.grad is not None, k
            den = mean[k].abs().max().item() + 1e-30
            worst = max(worst, (p.grad - mean[k]).abs().max().item() / den)
        # (3) one optimiser step: parameters must stay identical across ranks
        opt = get_optimizer(net.module)
        for p in net.module.parameters():
            p.grad.div_(S)
        opt.step()
        drift = 0.0
        for p in net.module.parameters():
            lo, hi = p.detach().clone(), p.detach().clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            drift = max(drift, (hi - lo).abs().max().item())
        out[amp] = (worst, drift)
    if rank == 0:
        ret.update({"fp32": out[False], "amp": out[True]})
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_nccl_gradient_is_rank_mean():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    print("NCCL world-2 DDP: max-norm rel |ddp grad - rank mean| fp32 %.2e, AMP %.2e; parameter drift across ranks "
          "after AdamW %.1e / %.1e" % (ret["fp32"][0], ret["amp"][0], ret["fp32"][1], ret["amp"][1]))
    # the two backward passes differ only by the summation order of the kernels' fp32/fp64 atomics (~1e-7)
    assert ret["fp32"][0] < 1e-5 and ret["amp"][0] < 1e-5
    assert ret["fp32"][1] == 0.0 and ret["amp"][1] == 0.0
