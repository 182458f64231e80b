"""CPU: host-side mirror of the reference interface — state_dict contract, factory dispatch, errors."""
import types

import pytest
import torch

import b200seg
from oracle import unet3d as ounet
from util import load_golden


def _args(**kw):
    a = types.SimpleNamespace(dimension="3d", model="resunet", in_chan=1, base_chan=8, classes=4,
                              down_scale=[[2, 2, 2]] * 4, kernel_size=[[3, 3, 3]] * 5, block="BasicBlock", norm="in")
    for k, v in kw.items():
        setattr(a, k, v)
    return a


@pytest.mark.parametrize("name", ["resunet_iso", "resunet_acdc", "unet_single"])
def test_state_dict_contract_matches_reference(name):
    g = load_golden(name)
    cfg = g["cfg"]
    net = b200seg.UNet(1, cfg["base"], scale=cfg["scale"], kernel_size=cfg["kernel"], num_classes=cfg["classes"],
                       block=cfg["block"], norm="in")
    sd = net.state_dict()
    assert list(sd) == g["keys"]                       # keys AND registration order (EMA zips params)
    shapes = ounet.unet_param_shapes(1, cfg["base"], cfg["classes"], cfg["kernel"], cfg["block"])
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    assert len(list(net.buffers())) == 0
    net.load_state_dict(ounet.make_state_dict(shapes, seed=1))          # strict load works
    assert [n for n, _ in net.named_parameters()] == g["keys"]


def test_full_size_resunet_has_reference_param_count():
    net = b200seg.get_model(_args(base_chan=32))
    assert sum(p.numel() for p in net.parameters()) == 40560612       # 40.56 M, 45 tensors
    assert len(net.state_dict()) == 45


def test_factory_dispatch_and_errors():
    assert isinstance(b200seg.get_model(_args(model="unet", block="SingleConv")), b200seg.UNet)
    with pytest.raises(ValueError):
        b200seg.get_model(_args(dimension="4d"))
    with pytest.raises(ValueError):
        b200seg.get_model(_args(model="resunet"), pretrain=True)     # model/utils.py:77-78
    with pytest.raises(ValueError):
        b200seg.get_model(_args(norm="bn"))
    with pytest.raises(ValueError):
        b200seg.get_model(_args(block="Bottleneck"))


def test_no_cpu_fallback():
    net = b200seg.get_model(_args())
    with pytest.raises(b200seg.B200SegError):
        net(torch.zeros(1, 1, 32, 32, 32))
    with pytest.raises(b200seg.B200SegError):
        b200seg.DiceLoss()(torch.zeros(1, 3, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4, dtype=torch.long))


def test_same_seed_same_init_as_holder_modules():
    torch.manual_seed(3)
    a = b200seg.get_model(_args())
    torch.manual_seed(3)
    b = b200seg.get_model(_args())
    assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))


def test_get_model_dispatches_medformer_with_reference_argument_list():
    """model/utils.py:95 — the MedFormer branch of the factory, driven by a config namespace like the BCV YAML."""
    import types
    import b200seg
    args = types.SimpleNamespace(
        dimension='3d', model='medformer', in_chan=1, classes=14, base_chan=32, map_size=[3, 3, 3],
        conv_block='BasicBlock', conv_num=[2, 0, 0, 0, 0, 0, 2, 2], trans_num=[0, 2, 4, 6, 4, 2, 0, 0],
        num_heads=[1, 4, 8, 10, 8, 4, 1, 1], fusion_depth=2, fusion_dim=320, fusion_heads=10, expansion=4, attn_drop=0.,
        proj_drop=0., proj_type='depthwise', norm='in', act='relu',
        kernel_size=[[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]],
        down_scale=[[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]], aux_loss=True)
    net = b200seg.get_model(args)
    assert isinstance(net, b200seg.MedFormer) and len(net.state_dict()) == 300
    import pytest
    args.map_size = [2, 6, 6]                      # ACDC YAML: 72 tokens -> loud, not silent
    with pytest.raises(ValueError):
        b200seg.get_model(args)
    args.map_size, args.act = [3, 3, 3], 'gelu'
    with pytest.raises(ValueError):
        b200seg.get_model(args)
