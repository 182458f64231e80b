"""CPU: the algorithmic-work accounting bench.py reports (roofline numerators) reproduces SURVEY.md §8(a)/(d)."""
import bench


def _fwd_gflop(name):
    scale, kernel, classes, _, (B, D, H, W) = bench.WORKLOADS[name]
    L = bench.conv_layers(scale, kernel, classes, B, D, H, W)
    return bench.conv_flops(L, B) / 1e9, L


def test_resunet_forward_flops_match_survey():
    iso, _ = _fwd_gflop("resunet_iso_128")
    acdc, _ = _fwd_gflop("resunet_acdc_128")
    kits, _ = _fwd_gflop("resunet_kits_160")
    assert abs(iso - 2590.1) < 0.5          # ResUNet-3D isotropic lists, 1x128^3
    assert abs(acdc - 3155.9) < 0.5         # literal ACDC lists
    assert abs(kits - 5058.6) < 1.0         # 2x160x160x80 per rank-step


def test_step_flops_subtract_only_the_stem_dgrad():
    fwd, L = _fwd_gflop("resunet_acdc_128")
    step = (3 * bench.conv_flops(L, 1) - bench.conv_flops(L[:1], 1)) / 1e9
    assert abs(step - 9466.4) < 1.0         # 3*3155.9 - 1.2
    assert len(L) == 44                     # 44 conv3d in the ResUNet (incl. the 1x1x1 head)


def test_workload_table_is_consistent():
    for name, wl in bench.WORKLOADS.items():
        assert len(wl[3]) == wl[2], name                       # one CE weight per class
        if bench.is_medformer(wl):
            assert wl[0]["aux_loss"] and bench.metric_of(wl) == bench.METRIC_MEDFORMER
        elif wl[0] == "swin":
            assert wl[1] == 48 and bench.metric_of(wl) == bench.METRIC_SWIN     # feature_size of config/*/swin_unetr_3d.yaml
        else:
            assert len(wl[0]) == 4 and len(wl[1]) == 5, name   # 4 scales, 5 kernel sizes (unet.py:35-45)
