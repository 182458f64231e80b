"""CPU, world_size=2 over gloo: the N>1 path of the drop-in.  The module is wrapped exactly like
train_ddp.py:352-358 (SyncBatchNorm.convert + DistributedDataParallel(find_unused_parameters=True)), with the
C-ABI ops emulated in PyTorch (tests/emu_ops.py, tests/emu_medformer.py), for the UNet and for a narrow MedFormer
with its deep-supervision head; checks that
  * custom autograd Functions keep every parameter reachable (DDP's unused-parameter walk finds none),
  * the all-reduced gradient is the MEAN of the per-rank gradients (Dice / CE normalisers are per-rank, §8e),
  * parameters stay bit-identical across ranks after an optimiser step."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, kind="unet"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import b200seg
    import emu_ops
    from oracle import unet3d as ounet
    from oracle.synth import make_volume

    class MP:
        def setattr(self, obj, name, val):
            setattr(obj, name, val)
    torch.Tensor.is_cuda = property(lambda self: True)
    torch.manual_seed(0)
    scale, kernel = [[1, 2, 2], [1, 2, 2], [2, 2, 2], [1, 1, 1]], [[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]]
    if kind == "unet":
        emu_ops.install(MP())
        net = b200seg.UNet(1, 4, scale=scale, kernel_size=kernel, num_classes=3, block="BasicBlock", norm="in")
        shapes = ounet.unet_param_shapes(1, 4, 3, kernel, "BasicBlock")
        img, lab = make_volume(1, 4, 16, 16, 3, seed=100 + rank)                  # each rank its own patch
    else:                                                                         # a narrow MedFormer, deep supervision on
        import emu_medformer
        emu_medformer.install(MP())
        net = b200seg.MedFormer(1, 3, 16, map_size=[2, 2, 2], conv_num=[1, 0, 0, 0, 0, 0, 1, 1], trans_num=[0, 1, 1, 1, 1, 1, 0, 0],
                                chan_num=[32, 32, 64, 64, 64, 32, 32, 16], num_heads=[1, 1, 2, 2, 2, 1, 1, 1], fusion_depth=1,
                                fusion_dim=64, fusion_heads=2, kernel_size=kernel, scale=[[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]],
                                aux_loss=True)
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        img, lab = make_volume(1, 8, 32, 32, 3, seed=100 + rank)
    net.load_state_dict(ounet.make_state_dict(shapes, seed=3))
    net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(net)                      # no-op: InstanceNorm only
    ddp = torch.nn.parallel.DistributedDataParallel(net, find_unused_parameters=True)
    ce_dice = b200seg.DiceCELoss(weight=torch.tensor([0.5, 1.0, 2.0]))

    def crit(res, lab):                                                           # train_ddp.py:186-191
        return sum(0.5 * ce_dice(r, lab) for r in res) if isinstance(res, (list, tuple)) else ce_dice(res, lab)
    # local (un-reduced) gradient of this rank, through the bare module
    net.zero_grad()
    crit(net(img), lab).backward()
    local = [p.grad.clone() for p in net.parameters()]
    net.zero_grad()
    crit(ddp(img), lab).backward()
    reduced = [p.grad.clone() for p in net.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, [g.double() for g in local])
    mean = [sum(gs[i] for gs in gathered) / world for i in range(len(local))]
    err = max(((r.double() - m).abs().max() / (m.abs().max() + 1e-30)).item() for r, m in zip(reduced, mean))
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, eps=1e-5, weight_decay=0.05)
    opt.step()
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    allp = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(allp, flat)
    same = all(torch.equal(allp[0], x) for x in allp)
    q.put((rank, err, same, all(g is not None for g in reduced)))
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["unet", "medformer"])
def test_ddp_world2_gloo_gradient_is_rank_mean(kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if kind == "medformer" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, kind)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same, all_grads in res:
        assert err < 1e-5, (rank, err)
        assert same and all_grads
