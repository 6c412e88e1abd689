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
        from models_twomodalinputs import fuseunet, fuseunetsa, fuseunetsaseparate
        from models_singlemodalinput import UNet, UNetsa, UNet2, UNet4, UNet8, UNet16, UNet32, UNet128
        from utils import (CrossEntropyLoss2d, DiceLoss, MulticlassDiceLoss, CEMDiceLoss, CEMDiceLossImage, PolyLR,
                           MulticlassDice_fn, MulticlassAccuracy_fn, Dice_fn, MulticlassMSELoss)
        import aide_b200
        assert fuseunet is aide_b200.fuseunet and UNet is aide_b200.UNet and UNetsa is aide_b200.UNetsa
        assert fuseunetsa is aide_b200.fuseunetsa and fuseunetsaseparate is aide_b200.fuseunetsaseparate
        with pytest.raises(NotImplementedError):
            UNet2()                       # 2-channel level: below the 4-channel access granularity of the NHWC kernels
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
    checked = pairs = 0
    for plan in (E.plan_fuseunet(2), E.plan_unet(2)):
        for S in (256, 320):
            for B in (1, 4, 8, 32, 40):
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
                            bn, mb, nacc, nbuf, rb, a_st, b_st, smem, flags = list(out)
                            stack, resident, occ = flags & 1, (flags >> 1) & 1, 2 if flags & 4 else 1
                            if flags & 8:        # CTA pair (cta_group::2): one CTA per SM, stacked planes, kind::f16 only
                                assert fmt != 1 and stack == (fmt == 3) and not resident and bn * (1 + stack) <= 256
                            assert cout % bn == 0 and bn in (32, 64, 128, 256) and mb in (1, 2, 4)
                            assert nbuf * mb * nacc * bn * (1 + stack) <= 512 // occ        # two CTAs per SM share TMEM ...
                            assert smem <= 227 * 1024 // occ and a_st >= 1 and b_st >= 2 and rb in (64, 128)   # ... and smem
                            assert not stack or (fmt != 2 and 2 * bn <= 256)
                            if resident:      # the whole weight matrix stays in shared memory: one cout tile, >= 2 halo stages
                                es = 4 if fmt == 1 else 2
                                assert bn == cout and a_st >= 2 and b_st == 9 * (cin * es // rb)
                            # same answer when asked again (the launch and the workspace query must agree)
                            again = (C.c_int * 9)()
                            lib.aide_conv3x3_plan_info(fmt, cin, cout, B, h, h, again)
                            assert list(again) == list(out)
                            checked += 1
                            pairs += bool(flags & 8)
    assert checked > 1000
    assert pairs > checked // 4          # the measured plan table is wired in: most >= 64-channel layers run as CTA pairs


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


def test_transposed_conv_is_conv3x3_on_zero_inserted_input():
    """Host-side algebra behind learned_bilinear=True (netblocks.py:11-14): ConvTranspose2d(k=2,s=2)(x) equals
    conv3x3(pad 1) of the zero-inserted input with K[co,ci,1-a,1-b] = W[ci,co,a,b] (engine.conv_weight_oihw), and the
    gradient of W is the re-indexed gradient of K (engine.transposed_weight_grad)."""
    import torch.nn.functional as F
    from aide_b200 import engine as E
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 5, 6, generator=g)
    W = torch.randn(8, 4, 2, 2, generator=g)
    b = torch.randn(4, generator=g)
    u = E.Unit("t", "c", "bn", 8, 4, 0, ("a", 0), None, [], False, True)
    K = E.conv_weight_oihw(u, {"c.weight": W})
    xz = torch.zeros(2, 8, 10, 12)
    xz[:, :, ::2, ::2] = x
    assert torch.equal(F.conv2d(xz, K, b, padding=1), F.conv_transpose2d(x, W, b, stride=2))
    gK = torch.randn(4, 8, 3, 3, generator=g)
    W2 = W.clone().requires_grad_(True)
    K2 = torch.zeros(4, 8, 3, 3)
    K2[:, :, 0:2, 0:2] = W2.permute(1, 0, 2, 3).flip(2, 3)
    (K2 * gK).sum().backward()
    assert torch.equal(W2.grad, E.transposed_weight_grad(gK))
    # plans: the ConvTranspose parameter names / indices of the reference's Sequential
    p = E.plan_fuseunet(2, learned_bilinear=True)
    ups = [v for v in p.units if v.transposed]
    assert [v.conv for v in ups] == [f"up_block{i}.bilinear_up.0" for i in range(1, 5)]
    assert [v.bn for v in ups] == [f"up_block{i}.bilinear_up.1" for i in range(1, 5)]
    assert all(op.mode == "zero" for op in p.ops if isinstance(op, E.Upsample))


def test_trainer_reverse_aug_parameters_follow_pil(oracle):
    """AideTrainer._aug_host builds, per (view, sample), the inverse affine matrix / fast-path mode / flip flag the
    reverse-augmentation kernel consumes; they must be PIL's (oracle.rotate_matrix restates PIL.Image.rotate), with
    views beyond a sample's `augno` left as the identity."""
    from aide_b200.trainer import AideTrainer
    augset = {"augno": [4, 2, 4], "degree1": [10.0, -33.5, 0.0], "degree2": [90.0, 45.0, 180.0],
              "degree3": [-60.0, 12.0, 7.25], "degree4": [0.0, 5.0, -90.0],
              "hflip1": [1, 0, 1], "hflip2": [0, 1, 0], "hflip3": [1, 1, 0], "hflip4": [0, 1, 1]}
    mats, modes, flips = AideTrainer._aug_host(augset, 4, 3, 64, 64)
    for v in range(4):
        for b in range(3):
            if augset["augno"][b] <= v:
                assert int(modes[v, b]) == 1 and int(flips[v, b]) == 0
                continue
            mode, m = oracle.rotate_matrix(0 - augset[f"degree{v + 1}"][b], 64, 64)
            assert int(modes[v, b]) == mode and mats[v, b].tolist() == m
            assert int(flips[v, b]) == augset[f"hflip{v + 1}"][b]


def test_gradient_buckets_cover_the_flat_buffer_in_completion_order():
    """Bucketed all-reduce: the float ranges handed to NCCL while the backward is still running must
    tile the flat gradient buffer exactly once, and every range must be final when its trigger unit is done -- i.e. all
    parameters inside it belong to units the backward pass (reverse forward order) has already visited."""
    from aide_b200 import engine as E
    for plan in (E.plan_fuseunet(2), E.plan_unet(2), E.plan_fuseunet(2, attention=True), E.plan_unet(2, attention=True)):
        gl = E.GradLayout(plan)
        buckets = E.gradient_buckets(plan, gl, 4)
        cov = sorted(r for _, rs in buckets for r in rs)
        assert cov[0][0] == 0 and cov[-1][1] == gl.total
        assert all(cov[i][1] == cov[i + 1][0] for i in range(len(cov) - 1))
        # the peer-memory all-reduce kernel moves 128-bit vectors: every cut is a multiple of 4 floats (the tail is padded)
        assert all(lo % 4 == 0 for lo, _ in cov) and all(hi % 4 == 0 for _, hi in cov[:-1])
        order = [u.name for u in reversed(plan.units)]                     # backward visiting order of the units
        done_at = {name: i for i, name in enumerate(order)}
        owner = {}                                                          # parameter -> unit / gate that finishes it
        for u in plan.units:
            for k in (u.conv + ".weight", u.conv + ".bias", u.bn + ".weight", u.bn + ".bias"):
                owner[k] = u.name
        last_trigger = buckets[-1][0]
        for trig, ranges in buckets:
            for lo, hi in ranges:
                for name, (off, shape) in gl.off.items():
                    if lo <= off < hi and name in owner:
                        assert done_at[owner[name]] <= done_at[trig], (plan.kind, trig, name)
                    if lo <= off < hi and name not in owner:               # gates / head: only in the last bucket
                        assert trig == last_trigger, (plan.kind, trig, name)
        assert [t for t, _ in buckets] == sorted([t for t, _ in buckets], key=lambda t: done_at[t])


def test_flat_gradient_slots_have_the_parameter_shapes():
    """The fused trainer keeps parameters, Adam state and gradients in ONE flat layout (engine.GradLayout): every slot must
    have exactly the shape of the module parameter it belongs to -- including the ConvTranspose2d weights [Cin,Cout,2,2] of
    learned_bilinear=True (netblocks.py:11-14), whose 3x3 stand-in gradient is re-indexed into the slot -- and the slots
    must tile the buffer without overlap (4-float alignment gaps only)."""
    import aide_b200
    from aide_b200 import engine as E
    for ctor, plan in ((aide_b200.fuseunet, E.plan_fuseunet), (aide_b200.UNet, E.plan_unet)):
        for lb in (False, True):
            net = ctor(num_classes=2, learned_bilinear=lb)
            gl = E.GradLayout(plan(2, learned_bilinear=lb))
            named = dict(net.named_parameters())
            assert set(named) == set(gl.off)
            end = 0
            for name, (off, shape) in sorted(gl.off.items(), key=lambda kv: kv[1][0]):
                assert tuple(named[name].shape) == shape, (name, lb)
                assert off >= end and off - end < 4 and (off % 4 == 0 or name.endswith((".bias", "bn.weight")))
                end = off + named[name].numel()
            assert gl.total - end < 4
