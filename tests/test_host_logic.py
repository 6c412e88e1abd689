"""CPU: host-side logic of the engine (plans, gradient routing, drop-in surface, loud failure on CPU)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_matches_reference(golden):
    import aide_b200
    torch.manual_seed(2)
    f, u = aide_b200.fuseunet(num_classes=2), aide_b200.UNet(num_classes=2)
    g = golden["init"]
    assert list(f.state_dict().keys()) == g["fuse_keys"]
    assert list(u.state_dict().keys()) == g["unet_keys"]
    assert {k: tuple(v.shape) for k, v in f.state_dict().items()} == g["fuse_shapes"]
    assert {k: tuple(v.shape) for k, v in u.state_dict().items()} == g["unet_shapes"]
    # identical initialisation under the same seed (same RNG consumption order as the reference)
    assert f.modal1_downblock1.block.conv1.weight.double().sum().item() == g["fuse_first_w_sum"]
    assert len(f.state_dict()) == 226 and len(u.state_dict()) == 156


def test_plan_structure():
    from aide_b200 import engine as E
    pf, pu = E.plan_fuseunet(2), E.plan_unet(2)
    assert len(pf.units) == 32 and len(pu.units) == 22
    flops = lambda p: sum(2 * 9 * u.cin * u.cout * (256 >> u.level) ** 2 for u in p.units)
    assert abs(flops(pf) / 1e9 - 116.19) < 0.05                 # 116.19 GFLOP of 3x3 convs per slice
    assert abs(flops(pu) / 1e9 - 130.70) < 0.1
    names = {u.conv for u in pf.units}
    assert "up_block1.bilinear_up.1" in names and "modal2_downblock5.block.conv2" in names
    # concat order: decoder cat((upsampled, skip)) -> up-conv writes channel 0, skip lives at C (netblocks.py:145)
    up1 = next(u for u in pf.units if u.name == "up_block1.up")
    assert up1.dst == ("cat1", 0) and (up1.cin, up1.cout) == (1024, 512)
    d4a = next(u for u in pf.units if u.name == "modal1_downblock4.block.2")
    d4b = next(u for u in pf.units if u.name == "modal2_downblock4.block.2")
    assert d4a.dst == ("cat1", 512) and d4b.dst == ("cat1", 768)   # encoder cat((y_modal1, x_modal2)) (fuseunet.py:73)
    # modal-2 reads only its own half of the pooled concat (fuseunet.py:54-55)
    b2 = next(u for u in pf.units if u.name == "modal2_downblock2.block.1")
    a2 = next(u for u in pf.units if u.name == "modal1_downblock2.block.1")
    assert b2.src == ("p1", 32) and b2.cin == 32 and a2.src == ("p1", 0) and a2.cin == 64


def test_backward_routing():
    from aide_b200 import engine as E
    pf = E.plan_fuseunet(2)
    bp = E.BackwardPlan(pf)
    # modal-2 level-1 output: skip (U4 conv1) + pooled into modal-1 level 2 AND modal-2 level 2
    d, p = bp.sources["modal2_downblock1.block.2"]
    assert [(k, o.name, off) for k, o, off, _ in d] == [("unit", "up_block4.block.1", 96)]
    assert sorted((o.name, off) for _, o, off, _ in p) == [("modal1_downblock2.block.1", 32), ("modal2_downblock2.block.1", 0)]
    # bottom of the encoder feeds the upsample only
    d, p = bp.sources["modal1_downblock5.block.2"]
    assert [k for k, *_ in d] == ["ups"] and not p
    d, p = bp.sources["up_block4.block.2"]
    assert [k for k, *_ in d] == ["head"]
    gl = E.GradLayout(pf)
    assert gl.total >= 26675074 and gl.off["last_conv1.bias"][0] == gl.off["last_conv1.weight"][0] + 128
    for u in pf.units:
        assert gl.off[u.bn + ".weight"][0] == gl.off[u.bn + ".bias"][0] + u.cout


def test_no_cpu_fallback():
    import aide_b200
    net = aide_b200.UNet(num_classes=2)
    with pytest.raises(RuntimeError, match="CUDA only"):
        net(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA only"):
        aide_b200.CEMDiceLossImage()(torch.zeros(1, 2, 8, 8), torch.zeros(1, 8, 8, dtype=torch.long))
    # learned_bilinear=True keeps the reference's module tree: ConvTranspose2d at index 0, BatchNorm at index 1
    lb = aide_b200.fuseunet(learned_bilinear=True)
    assert isinstance(lb.up_block1.bilinear_up[0], torch.nn.ConvTranspose2d)
    assert tuple(lb.up_block1.bilinear_up[0].weight.shape) == (1024, 512, 2, 2)
    with pytest.raises(RuntimeError, match="CUDA only"):
        lb(torch.zeros(1, 3, 32, 32), torch.zeros(1, 3, 32, 32))


def test_dropin_packages_import():
    """The reference scripts do `from models_twomodalinputs import fuseunet`, `from utils import ...`
    (trainchaos_proposed_30cases1labeled.py:20-23); aide_b200/dropin provides those names."""
    sys.path.insert(0, os.path.join(ROOT, "aide_b200", "dropin"))
    try:
        for m in ("models_twomodalinputs", "models_singlemodalinput", "utils"):
            sys.modules.pop(m, None)
        from models_twomodalinputs import fuseunet
        from models_singlemodalinput import UNet, UNetsa
        from utils import (CrossEntropyLoss2d, DiceLoss, MulticlassDiceLoss, CEMDiceLoss, CEMDiceLossImage, PolyLR,
                           MulticlassDice_fn, MulticlassAccuracy_fn, Dice_fn, MulticlassMSELoss)
        import aide_b200
        assert fuseunet is aide_b200.fuseunet and UNet is aide_b200.UNet
        with pytest.raises(NotImplementedError):
            UNetsa()
    finally:
        sys.path.pop(0)
        for m in ("models_twomodalinputs", "models_singlemodalinput", "utils"):
            sys.modules.pop(m, None)


def test_polylr_matches_formula():
    import aide_b200
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=0.1)
    sch = aide_b200.PolyLR(opt, max_epoch=10, power=0.9)
    for e in range(1, 4):
        opt.step()
        sch.step()
        assert abs(opt.param_groups[0]["lr"] - 0.1 * (1 - e / 10) ** 0.9) < 1e-12


def test_conv_planner_returns_valid_tilings_for_every_layer():
    """Host-only: the tile planner of the tcgen05 conv kernel (aide_conv3x3_plan_info) must return a feasible tiling
    for every tensor-core layer of both networks at the batches the step uses (B and 4B) and at 256 / 320 inputs:
    TMEM columns <= 512, dynamic shared memory <= 227 KB, cout tile divides cout, at least two weight stages."""
    import ctypes as C
    from aide_b200 import engine as E, lib
    out = (C.c_int * 9)()
    checked = 0
    for plan in (E.plan_fuseunet(2), E.plan_unet(2)):
        for S in (256, 320):
            for B in (4, 8, 32):
                for u in plan.units:
                    if u.first:
                        continue
                    h = S >> u.level
                    for fmt in (1, 2, 3):
                        for cin, cout in ((u.cin, u.cout), (u.cout, u.cin)):          # forward and dgrad roles
                            if cin % 32 or cout % 32:
                                continue
                            rc = lib.aide_conv3x3_plan_info(fmt, cin, cout, B, h, h, out)
                            assert rc == 0, (plan.kind, S, B, u.name, fmt)
                            bn, mb, nacc, nbuf, rb, a_st, b_st, smem, stack = list(out)
                            assert cout % bn == 0 and bn in (32, 64, 128, 256) and mb in (1, 2, 4)
                            assert nbuf * mb * nacc * bn * (1 + stack) <= 512
                            assert smem <= 227 * 1024 and a_st >= 1 and b_st >= 2 and rb in (64, 128)
                            assert not stack or (fmt != 2 and 2 * bn <= 256)
                            # same answer when asked again (the launch and the workspace query must agree)
                            again = (C.c_int * 9)()
                            lib.aide_conv3x3_plan_info(fmt, cin, cout, B, h, h, again)
                            assert list(again) == list(out)
                            checked += 1
    assert checked > 1000


def test_stat_rows_scale_with_batch_groups():
    """The stacked pseudo-label forward slices the conv's BatchNorm partial-statistics rows per view: the row count
    must be an exact multiple of the batch (rows are image-major for every conv kernel)."""
    from aide_b200 import lib
    for fmt in (0, 1, 2, 3):
        for (cin, cout, h) in ((3, 32, 256), (64, 64, 128), (512, 256, 64), (512, 512, 16), (256, 512, 4)):
            if fmt != 0 and cin == 3:
                continue
            r8 = lib.aide_conv3x3_stat_rows(fmt, cin, cout, 8, h, h)
            r32 = lib.aide_conv3x3_stat_rows(fmt, cin, cout, 32, h, h)
            assert r8 > 0 and r32 == 4 * r8, (fmt, cin, cout, h, r8, r32)
