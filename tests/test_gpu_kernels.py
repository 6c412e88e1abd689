"""GPU: every kernel of libaide_b200 against the same operator on the CPU (torch fp32 / the oracle),
called through the C ABI (aide_b200.ops / aide_b200.losses are thin ctypes wrappers).

Tolerances: 'exact' (fp32 CUDA cores) and 'parity' (3xTF32 tcgen05) must agree with fp32 to ~1e-5 of the
tensor's max (fp32 summation-order noise); 'fast' (single-pass BF16) is only sanity-checked."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {0: 2e-5, 1: 3e-5, 2: 3e-2, 3: 3e-5}          # FMT_F32, FMT_TF32X2, FMT_BF16, FMT_F16X2
NAMES = {0: "exact", 1: "parity(tf32x2)", 2: "fast", 3: "parity(f16x2)"}


def relmax(a, b):
    return ((a.cpu().double() - b.cpu().double()).abs().max() / b.cpu().double().abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


CONV_SHAPES = [  # N, H, W, cin, cout
    (2, 16, 16, 32, 32), (1, 32, 32, 64, 64), (2, 8, 8, 128, 256), (1, 20, 12, 64, 32),
    (1, 2, 2, 512, 128), (1, 16, 128, 32, 64), (3, 24, 40, 96, 96), (1, 8, 8, 1024, 512),
]


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", CONV_SHAPES + [(2, 32, 32, 256, 512), (4, 64, 64, 64, 256)])
def test_conv3x3_forward_and_stats(dev, fmt, shape):
    from aide_b200 import ops
    N, H, W, cin, cout = shape
    x, w, b = rnd(N, cin, H, W, seed=1), rnd(cout, cin, 3, 3, seed=2, scale=(9 * cin) ** -0.5), rnd(cout, seed=3)
    ref = F.conv2d(x, w, b, padding=1)
    a = ops.from_nchw(x.to(dev), fmt)
    z, part = ops.conv3x3(a, w.to(dev), b.to(dev), stats=True)
    torch.cuda.synchronize()
    got = ops.nhwc_to_nchw(z)
    assert relmax(got, ref) < TOL[fmt], NAMES[fmt]
    s = part.sum(0).cpu()
    assert relmax(s[0], got.cpu().sum((0, 2, 3))) < 1e-4
    assert relmax(s[1], (got.cpu() ** 2).sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("fmt", [2, 3])
@pytest.mark.parametrize("shape", [(2, 16, 16, 32, 32), (1, 32, 32, 64, 64), (2, 8, 8, 128, 256), (1, 20, 12, 64, 32),
                                   (3, 24, 40, 96, 96), (1, 8, 8, 1024, 512), (1, 16, 8, 64, 128), (5, 8, 16, 32, 64)])
def test_conv3x3_cta_pair_kernel_equals_single_cta_kernel(dev, fmt, shape, monkeypatch):
    """The cta_group::2 kernel (two CTAs, one MMA of 256 pixels, half of the weight rows in each CTA) forced on small and
    ragged maps -- odd tile counts exercise the zero-filled padding tile of the last pair: same tolerance against fp32
    as every conv kernel, and the SAME BITS as the single-CTA kernel on the same tiling (identical accumulation chains:
    the pair only changes which SM holds which rows)."""
    from aide_b200 import ops
    N, H, W, cin, cout = shape
    x, w, b = rnd(N, cin, H, W, seed=1), rnd(cout, cin, 3, 3, seed=2, scale=(9 * cin) ** -0.5), rnd(cout, seed=3)
    ref = F.conv2d(x, w, b, padding=1)
    a = ops.from_nchw(x.to(dev), fmt)
    out = {}
    monkeypatch.setenv("AIDE_CONV_TABLE", "0")
    monkeypatch.setenv("AIDE_CONV_BN", "32" if cout % 64 else "64")
    monkeypatch.setenv("AIDE_CONV_MB", "1")
    monkeypatch.setenv("AIDE_CONV_STACK", "1")
    for occ in (4, 1):
        monkeypatch.setenv("AIDE_CONV_OCC", str(occ))
        z, part = ops.conv3x3(a, w.to(dev), b.to(dev), stats=True)
        torch.cuda.synchronize()
        out[occ] = (ops.nhwc_to_nchw(z).cpu(), part.cpu())
    assert relmax(out[4][0], ref) < TOL[fmt]
    assert torch.equal(out[4][0], out[1][0]) and torch.equal(out[4][1], out[1][1])


def test_conv3x3_first_layer_cin3(dev):
    from aide_b200 import ops
    x, w, b = rnd(2, 3, 32, 48, seed=1), rnd(32, 3, 3, 3, seed=2, scale=0.2), rnd(32, seed=3)
    a = ops.from_nchw(x.to(dev), 0)
    z, _ = ops.conv3x3(a, w.to(dev), b.to(dev))
    assert relmax(ops.nhwc_to_nchw(z), F.conv2d(x, w, b, padding=1)) < 2e-5
    dz = rnd(2, 32, 32, 48, seed=5)
    dw = ops.conv3x3_wgrad(a, ops.from_nchw(dz.to(dev), 0))
    ref = torch.nn.grad.conv2d_weight(x, w.shape, dz, padding=1)
    assert relmax(dw, ref) < 3e-5


@pytest.mark.parametrize("cout", [32, 64])
@pytest.mark.parametrize("shape", [(2, 20, 44), (1, 16, 16), (3, 33, 7), (2, 64, 64)])
def test_conv3x3_first_layer_dedicated_kernels(dev, cout, shape):
    """cin == 3 -> conv_first.cu (forward + BN partial statistics + weight gradient), ragged tiles included."""
    from aide_b200 import ops
    N, H, W = shape
    x, w, b = rnd(N, 3, H, W, seed=1), rnd(cout, 3, 3, 3, seed=2, scale=0.2), rnd(cout, seed=3)
    ref = F.conv2d(x, w, b, padding=1)
    a = ops.from_nchw(x.to(dev), 0)
    z, part = ops.conv3x3(a, w.to(dev), b.to(dev), stats=True)
    got = ops.nhwc_to_nchw(z)
    assert relmax(got, ref) < 2e-6
    s = part.sum(0).cpu()
    assert relmax(s[0], ref.sum((0, 2, 3))) < 1e-5 and relmax(s[1], (ref ** 2).sum((0, 2, 3))) < 1e-5
    dz = rnd(N, cout, H, W, seed=5)
    dw = ops.conv3x3_wgrad(a, ops.from_nchw(dz.to(dev), 0))
    assert relmax(dw, torch.nn.grad.conv2d_weight(x, w.shape, dz, padding=1)) < 1e-5


def test_bn_finalize_folds_many_rows(dev):
    """> 128 partial rows take the two-stage (fold in place, then finalize) path; same statistics either way."""
    from aide_b200 import ops
    C, rows = 96, 1000
    part = torch.randn(rows, 2, C, generator=torch.Generator().manual_seed(3)).abs() + 0.1
    part[:, 1] += 50.0                                   # keep the variance positive
    count = 4096.0
    gamma, beta = rnd(C, seed=1).abs() + 0.5, rnd(C, seed=2)
    s = part.double().sum(0)
    mean = s[0] / count
    var = s[1] / count - mean ** 2
    ss, mr = ops.bn_finalize(part.to(dev), count, gamma.to(dev), beta.to(dev), None, None, True)
    assert relmax(mr[0], mean) < 1e-6 and relmax(mr[1], 1.0 / torch.sqrt(var + 1e-5)) < 1e-6
    assert relmax(ss[0], gamma.double() / torch.sqrt(var + 1e-5)) < 1e-6
    # one-launch path (fold + finalize by the last block of each channel group): bit-identical, tickets left at zero;
    # with running statistics and several statistics groups (the stacked pseudo-label forward)
    G = 3
    parts = torch.randn(G, rows, 2, C, generator=torch.Generator().manual_seed(4)).abs() + 0.1
    parts[:, :, 1] += 50.0
    outs = []
    for fused in (False, True):
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        tk = torch.zeros(ops.lib.aide_bn_ticket_slots(C), dtype=torch.int32, device=dev)
        ss2, mr2 = ops.bn_finalize_grouped(parts.to(dev), count, gamma.to(dev), beta.to(dev), rm, rv,
                                           tickets=tk if fused else None)
        assert int(tk.abs().sum()) == 0
        outs.append((ss2, mr2, rm, rv))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    m2 = parts[2].double().sum(0)[0] / count
    assert relmax(outs[1][1][2, 0], m2) < 1e-6


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3x3_dgrad_wgrad(dev, fmt, shape):
    from aide_b200 import ops
    N, H, W, cin, cout = shape
    x, w, dz = rnd(N, cin, H, W, seed=4), rnd(cout, cin, 3, 3, seed=5, scale=(9 * cin) ** -0.5), rnd(N, cout, H, W, seed=6)
    ref_dx = torch.nn.grad.conv2d_input(x.shape, w, dz, padding=1)
    ref_dw = torch.nn.grad.conv2d_weight(x, w.shape, dz, padding=1)
    dza = ops.from_nchw(dz.to(dev), fmt)
    dx = ops.nhwc_to_nchw(ops.conv3x3_dgrad(dza, w.to(dev)))
    dw = ops.conv3x3_wgrad(ops.from_nchw(x.to(dev), fmt), dza)
    torch.cuda.synchronize()
    assert relmax(dx, ref_dx) < TOL[fmt], "dgrad " + NAMES[fmt]
    assert relmax(dw, ref_dw) < TOL[fmt], "wgrad " + NAMES[fmt]


@pytest.mark.parametrize("mag", [1e-6, 30.0])
def test_conv3x3_f16x2_backward_with_gradient_scale(dev, mag):
    """F16X2 gradient planes hold dz*s (s a device-side power of two); dgrad/wgrad multiply 1/s back in."""
    from aide_b200 import ops
    N, H, W, cin, cout = 2, 16, 24, 64, 128
    x, w = rnd(N, cin, H, W, seed=4), rnd(cout, cin, 3, 3, seed=5, scale=(9 * cin) ** -0.5)
    dz = rnd(N, cout, H, W, seed=6) * mag
    s_ = 2.0 ** round(float(torch.log2(torch.tensor(1024.0 / (4.5 * mag)))))      # max|dz|*s ~ 2^10
    dza = ops.from_nchw((dz * (s_ / 256.0)).to(dev), 3)                             # from_nchw multiplies by 2^8
    inv = torch.full((1,), 1.0 / s_, device=dev)
    dx = ops.nhwc_to_nchw(ops.conv3x3_dgrad(dza, w.to(dev), inv_scale=inv))
    dw = ops.conv3x3_wgrad(ops.from_nchw(x.to(dev), 3), dza, inv_scale=inv)
    assert relmax(dx, torch.nn.grad.conv2d_input(x.shape, w, dz, padding=1)) < TOL[3]
    assert relmax(dw, torch.nn.grad.conv2d_weight(x, w.shape, dz, padding=1)) < TOL[3]


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_conv3x3_channel_views(dev, fmt):
    """conv reading a channel slice of a wider buffer and writing... (zero-copy concat, fuseunet.py:49-55)."""
    from aide_b200 import ops
    x, w = rnd(2, 96, 16, 16, seed=7), rnd(64, 32, 3, 3, seed=8, scale=0.06)
    a = ops.from_nchw(x.to(dev), fmt)
    z, _ = ops.conv3x3(a, w.to(dev), None, cin=32, coff=64)
    assert relmax(ops.nhwc_to_nchw(z), F.conv2d(x[:, 64:96], w, None, padding=1)) < TOL[fmt]
    dz = rnd(2, 64, 16, 16, seed=9)
    dw = ops.conv3x3_wgrad(a, ops.from_nchw(dz.to(dev), fmt), cin=32, coff=64)
    assert relmax(dw, torch.nn.grad.conv2d_weight(x[:, 64:96], w.shape, dz, padding=1)) < TOL[fmt]


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_conv_adjoint_identity_full_size(dev, fmt):
    """Size-independent property at a BASELINE-sized layer (B=4, 128x128, 128->64):
    <conv(x,w), dz> == <x, dgrad(dz,w)> == <w, wgrad(x,dz)>."""
    from aide_b200 import ops
    N, H, W, cin, cout = 4, 128, 128, 128, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(N, cin, H, W, device=dev, generator=g)
    w = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * (9 * cin) ** -0.5
    dz = torch.randn(N, cout, H, W, device=dev, generator=g)
    xa, dza = ops.from_nchw(x, fmt), ops.from_nchw(dz, fmt)
    z, _ = ops.conv3x3(xa, w, None)
    a = (z.double() * dza.float().double()).sum().item()
    b = (ops.conv3x3_dgrad(dza, w).double() * xa.float().double()).sum().item()
    c = (ops.conv3x3_wgrad(xa, dza).double() * w.double()).sum().item()
    tol = 1e-5 if fmt != 2 else 2e-2
    scale = (z.double().norm() * dza.float().double().norm()).item()
    assert abs(a - b) < tol * scale and abs(a - c) < tol * scale, (a, b, c)


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("training", [True, False])
def test_bn_relu_pool_forward(dev, fmt, training):
    from aide_b200 import ops
    N, C, H, W = 3, 64, 12, 20
    x, w = rnd(N, 32, H, W, seed=1), rnd(C, 32, 3, 3, seed=2, scale=0.1)
    gamma, beta = rnd(C, seed=3).abs() + 0.5, rnd(C, seed=4)
    rm, rv = rnd(C, seed=5) * 0.1, rnd(C, seed=6).abs() + 0.5
    zr = F.conv2d(x, w, None, padding=1)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    yr = F.relu(F.batch_norm(zr, rm_ref, rv_ref, gamma, beta, training, 0.1, 1e-5))
    a = ops.from_nchw(x.to(dev), 0)
    z, part = ops.conv3x3(a, w.to(dev), None, stats=True)
    rmd, rvd = rm.to(dev), rv.to(dev)
    ss, mr = ops.bn_finalize(part, N * H * W, gamma.to(dev), beta.to(dev), rmd, rvd, training)
    y, p = ops.bn_relu_apply(z, ss, fmt, pool=True)
    tol = 5e-5 if fmt != 2 else 1e-2
    assert relmax(ops.to_nchw(y), yr) < tol
    assert relmax(ops.to_nchw(p), F.max_pool2d(yr, 2, 2)) < tol
    assert relmax(rmd, rm_ref) < 1e-5 and relmax(rvd, rv_ref) < 1e-5


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("gscale", [1.0, 3e-7, 2e3])
@pytest.mark.parametrize("fmt", [0, 1, 3])
def test_bn_relu_pool_backward(dev, fmt, gscale, fused):
    """g routing: one same-resolution source + two pooled sources (the modal-2 level-1 pattern).  F16X2 stores dZ
    with a power-of-two scale chosen on the device: exercised with tiny and huge upstream gradients."""
    if fmt != 3 and gscale != 1.0:
        pytest.skip("gradient magnitude only matters for the dynamically scaled format")
    from aide_b200 import ops
    from aide_b200._lib import call, lib
    import ctypes as C
    N, Cc, H, W = 2, 32, 8, 12
    z = rnd(N, Cc, H, W, seed=1).requires_grad_()
    gamma, beta = (rnd(Cc, seed=2).abs() + 0.5).requires_grad_(), rnd(Cc, seed=3).requires_grad_()
    y = F.relu(F.batch_norm(z, None, None, gamma, beta, True, 0.1, 1e-5))
    p = F.max_pool2d(y, 2, 2)
    gd, gp1, gp2 = (rnd(N, 64, H, W, seed=4) * gscale, rnd(N, 64, H // 2, W // 2, seed=5) * gscale,
                    rnd(N, Cc, H // 2, W // 2, seed=6) * gscale)
    loss = (y * gd[:, 16:48]).sum() + (p * gp1[:, 32:64]).sum() + (p * gp2).sum()
    dz_ref, dg_ref, db_ref = torch.autograd.grad(loss, [z, gamma, beta])
    st = torch.cuda.current_stream().cuda_stream
    zd = z.detach().permute(0, 2, 3, 1).contiguous().to(dev)
    # statistics exactly as the forward would produce them
    part = torch.stack([zd.sum((0, 1, 2)), (zd ** 2).sum((0, 1, 2))])[None].contiguous()
    ss, mr = ops.bn_finalize(part, N * H * W, gamma.detach().to(dev), beta.detach().to(dev), None, None, True)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(dev)
    d0, p0, p1 = nhwc(gd), nhwc(gp1), nhwc(gp2)
    rows = lib.aide_bn_bwd_rows(N, H, W, Cc)
    g = torch.empty(N, H, W, Cc, device=dev)
    part1 = torch.empty(rows, 2, Cc, device=dev)
    part2 = torch.empty(rows, Cc, device=dev)
    dptr, dct, dco = (C.c_void_p * 3)(d0.data_ptr()), (C.c_int * 3)(64), (C.c_int * 3)(16)
    pptr, pct, pco = (C.c_void_p * 3)(p0.data_ptr(), p1.data_ptr()), (C.c_int * 3)(64, Cc), (C.c_int * 3)(32, 0)
    gs = torch.zeros(4, device=dev)             # [0] max|g| bits (zero on entry), [1] s, [2] 1/s
    tk = torch.zeros(16, dtype=torch.int32, device=dev)     # ticket block of the one-launch paths (zero on entry)
    dyn = fmt == 3
    call("aide_bn_relu_bwd_reduce", zd.data_ptr(), ss.data_ptr(), mr.data_ptr(), N, H, W, Cc, dptr, dct, dco, 1,
         pptr, pct, pco, 2, g.data_ptr(), part1.data_ptr(), gs[0:].data_ptr() if dyn else None, st)
    dz = ops.Act(N, H, W, Cc, fmt, dev)
    small = torch.empty(3, Cc, device=dev)
    call("aide_bn_relu_bwd_apply", fmt, g.data_ptr(), zd.data_ptr(), mr.data_ptr(), gamma.detach().to(dev).data_ptr(),
         part1.data_ptr(), rows, N, H, W, Cc, dz.p0, dz.p1, small[1].data_ptr(), small[0].data_ptr(),
         small[2].data_ptr(), part2.data_ptr(), gs[0:].data_ptr() if dyn else None, gs[1:].data_ptr() if dyn else None,
         tk.data_ptr() if fused else None, st)
    got = ops.to_nchw(dz)
    assert int(tk.abs().sum()) == 0             # tickets are left at zero
    if dyn:
        s_, inv = gs[1].item(), gs[2].item()
        assert s_ > 0 and abs(s_ * inv - 1.0) < 1e-6 and abs(torch.tensor(s_).log2().item() % 1.0) < 1e-6   # power of two
        assert gs[0].item() == 0.0                  # max|g| was consumed and reset for the next unit
        got = got * 256.0 * inv                     # Act.to_nchw divides by the static activation scale 2^8
        assert (got.abs().max() * s_).item() < 65504 / 4      # head-room of the bound
    assert relmax(got, dz_ref) < 5e-5
    assert relmax(small[1], dg_ref) < 5e-5 and relmax(small[0], db_ref) < 5e-5
    assert small[2].abs().max().item() < 1e-4 * dz_ref.abs().max().item() * N * H * W   # sum dz == 0 analytically


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("hw", [(16, 16), (1, 1), (6, 10), (2, 2), (3, 5), (5, 2), (64, 48), (1, 7)])
def test_upsample_bilinear(dev, fmt, hw):
    from aide_b200 import ops
    h, w = hw
    x = rnd(2, 32, h, w, seed=1)
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    y = ops.upsample2x(ops.from_nchw(x.to(dev), fmt))
    assert relmax(ops.to_nchw(y), ref) < (1e-5 if fmt != 2 else 1e-2)
    if fmt == 0:
        g = rnd(2, 32, 2 * h, 2 * w, seed=2)
        xr = x.clone().requires_grad_()
        (F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True) * g).sum().backward()
        dlo = ops.upsample2x_bwd(g.permute(0, 2, 3, 1).contiguous().to(dev))
        assert relmax(ops.nhwc_to_nchw(dlo), xr.grad) < 1e-5


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
def test_conv1x1_head(dev, fmt):
    from aide_b200 import ops
    x, w, b = rnd(2, 64, 20, 24, seed=1), rnd(2, 64, 1, 1, seed=2, scale=0.1), rnd(2, seed=3)
    ref = F.conv2d(x, w, b)
    a = ops.from_nchw(x.to(dev), fmt)
    out = ops.conv1x1(a, w.view(2, 64).to(dev), b.to(dev))
    tol = 1e-5 if fmt != 2 else 1e-2
    assert relmax(out, ref) < tol
    dl = rnd(2, 2, 20, 24, seed=4)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    (F.conv2d(xr, wr, br) * dl).sum().backward()
    dx, dw, db = ops.conv1x1_bwd(a, w.view(2, 64).to(dev), dl.to(dev))
    assert relmax(ops.nhwc_to_nchw(dx), xr.grad) < 1e-5
    assert relmax(dw, wr.grad.view(2, 64)) < tol and relmax(db, br.grad) < 1e-5


def test_losses_against_oracle(dev, golden, oracle):
    import aide_b200 as A
    g = golden["loss"]
    lg, lg2, tg = g["logits"].to(dev), g["logits2"].to(dev), g["targets"].to(dev)
    assert torch.allclose(A.CEMDiceLossImage([1., 1.], torch.tensor([1., 1.]), [1., 1.])(lg, tg).cpu(), g["cedice_img"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(A.CEMDiceLossImage([0.5, 2.0], [0.3, 1.7], None)(lg, tg).cpu(), g["cedice_img_w"], rtol=1e-5, atol=1e-6)
    assert abs(A.CEMDiceLoss([1., 1.], [1., 1.], [1., 1.])(lg, tg).item() - g["cedice_mean"]) < 1e-5
    assert abs(A.DiceLoss()(lg, tg).item() - g["dice_loss"]) < 1e-6
    assert abs(A.Dice_fn(lg, tg).item() - g["dice_fn"]) < 1e-6
    assert torch.allclose(A.CrossEntropyLoss2d(reduction="none")(lg, tg).cpu(), g["ce_none"], rtol=1e-5, atol=1e-6)
    q, wm = A.pseudo_label([lg, lg2], 1.0)
    assert torch.allclose(q.cpu(), g["q"], atol=1e-6) and torch.allclose(wm.cpu(), g["wmap"], atol=1e-6)
    assert torch.allclose(A.pseudo_label([lg, lg2], 2.0, "chaos")[0].cpu(), g["q_T2_chaos"], atol=1e-6)
    assert torch.allclose(A.pseudo_label([lg, lg2], 2.0, "kidney")[0].cpu(), g["q_T2_kidney"], atol=1e-6)
    # empty / ragged: a 1-pixel image and an all-ignore target
    one = A.CEMDiceLossImage()(lg[:1, :, :1, :1].contiguous(), tg[:1, :1, :1].contiguous())
    assert torch.allclose(one.cpu(), oracle.ce_dice_per_image(g["logits"][:1, :, :1, :1], g["targets"][:1, :1, :1]), atol=1e-6)


def test_loss_backward_against_autograd(dev, golden, oracle):
    import aide_b200 as A
    g = golden["loss"]
    lg_c = g["logits"].clone().requires_grad_()
    tg = g["targets"]
    up = torch.tensor([0.3, -1.2, 2.0, 0.7, 1.1])
    (oracle.ce_dice_per_image(lg_c, tg, (0.5, 2.0), (0.3, 1.7)) * up).sum().backward()
    lg = g["logits"].to(dev).requires_grad_()
    (A.CEMDiceLossImage([0.5, 2.0], [0.3, 1.7], None)(lg, tg.to(dev)) * up.to(dev)).sum().backward()
    assert relmax(lg.grad, lg_c.grad) < 1e-5
    for red in ("mean", "sum"):
        lg_c.grad = None
        lg.grad = None
        import torch.nn.functional as F
        w = torch.tensor([0.3, 1.7])
        ce = F.cross_entropy(lg_c, tg, weight=w, reduction=red)
        dl = oracle.dice_per_image(lg_c, tg)
        ((ce * 0.5) + (dl.sum() / 5 if red == "mean" else dl.sum()) * 2.0).backward()
        A.CEMDiceLoss([0.5, 2.0], [0.3, 1.7], None, reduction=red)(lg, tg.to(dev)).backward()
        assert relmax(lg.grad, lg_c.grad) < 1e-5, red


def test_coteach_step_against_oracle(dev, golden, oracle):
    import aide_b200 as A
    g = golden["loss"]
    o1c, o2c = g["logits"].clone().requires_grad_(), g["logits2"].clone().requires_grad_()
    gen = torch.Generator().manual_seed(5)
    t1 = g["targets"]
    t2 = (torch.rand(t1.shape, generator=gen) < 0.3).long()
    a1, a2 = [torch.randn(o1c.shape, generator=gen) for _ in range(3)], [torch.randn(o1c.shape, generator=gen) for _ in range(3)]
    q1c, w1c = oracle.pseudo_label(a1)
    q2c, w2c = oracle.pseudo_label(a2)
    for n_clean in (2, 3):
        r = oracle.coteach_losses(o1c, o2c, t1, t2, q1c, w1c, q2c, w2c, 0.25, (1.0, 10.0), n_clean)
        g1c, = torch.autograd.grad(r["loss1"], o1c, retain_graph=True)
        g2c, = torch.autograd.grad(r["loss2"], o2c)
        o1, o2 = g["logits"].to(dev).requires_grad_(), g["logits2"].to(dev).requires_grad_()
        q1, w1 = A.pseudo_label([t.to(dev) for t in a1])
        q2, w2 = A.pseudo_label([t.to(dev) for t in a2])
        m = A.coteach_step(o1, o2, t1.to(dev), t2.to(dev), q1, w1, q2, w2, 0.25, (1.0, 10.0), n_clean)
        assert torch.equal(m["indx1"].cpu(), r["indx1"]) and torch.equal(m["indx2"].cpu(), r["indx2"])
        assert abs(m["loss1"].item() - r["loss1"].item()) < 2e-6 * max(1, abs(r["loss1"].item()))
        assert abs(m["loss2"].item() - r["loss2"].item()) < 2e-6 * max(1, abs(r["loss2"].item()))
        m["loss1"].backward(retain_graph=True)
        m["loss2"].backward()
        assert relmax(o1.grad, g1c) < 1e-5 and relmax(o2.grad, g2c) < 1e-5
        assert abs(m["dice1"].item() - oracle.dice_fn(o1c.detach(), t2).item()) < 1e-6


def test_coteach_classes(dev, golden):
    import aide_b200 as A
    g = golden["loss"]
    lg, lg2, tg = g["logits"].to(dev), g["logits2"].to(dev), g["targets"].to(dev)
    for name, (a, b) in g["coteach_classes"].items():
        cls, _, variant = name.partition("@")
        crit = getattr(A, cls)(reduction="none")
        if variant:
            o = crit(lg, lg2, tg, 0.4)
        else:
            o = crit(lg[:4].contiguous(), lg2[:4].contiguous(), tg[:4].contiguous(), 0.5)
        assert abs(float(o[0]) - a) < 1e-5 * max(1, abs(a)) and abs(float(o[1]) - b) < 1e-5 * max(1, abs(b)), name


def test_adam_amsgrad_matches_torch(dev):
    import aide_b200 as A
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s)) for s in [(7, 3, 3, 3), (7,), (130,)]]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt_ref = torch.optim.Adam(ref, lr=1e-3, amsgrad=True)
    dps = [torch.nn.Parameter(p.detach().clone().to(dev)) for p in ps]
    opt = A.FlatAdamAMSGrad(dps, lr=1e-3)
    for step in range(5):
        for p, r in zip(dps, ref):
            gr = torch.randn(r.shape) * (0.1 if step % 2 else 1.0)
            r.grad = gr.clone()
            p.grad = gr.to(dev)
        opt_ref.step()
        opt.step()
    for p, r in zip(dps, ref):
        assert torch.allclose(p.detach().cpu(), r.detach(), rtol=1e-5, atol=1e-7)


def test_reverse_augmentation_matches_pil_bit_for_bit(dev, oracle):
    """GPU un-flip / un-rotate of the augmented views' logits vs the reference's PIL round trip
    (trainchaos_proposed_30cases1labeled.py:81-95): bit-exact, including PIL's fast paths and samples with fewer views."""
    import random
    import aide_b200 as A
    rnd = random.Random(3)
    for (B, K, H, W) in [(4, 2, 64, 64), (3, 2, 48, 80), (2, 3, 33, 17)]:
        n_views = 4
        augset = {"augno": [rnd.choice([4, 4, 2]) for _ in range(B)]}
        for v in range(1, n_views + 1):
            augset[f"hflip{v}"] = [int(rnd.random() < 0.5) for _ in range(B)]
            augset[f"degree{v}"] = [rnd.choice([0, 0.0, 90, 180, -90, 30]) if rnd.random() < 0.3 else rnd.uniform(-60, 60)
                                    for _ in range(B)]
        outs = [rnd_t for rnd_t in (rnd_tensor(B, K, H, W, seed=10 + v) for v in range(n_views))]
        ref = oracle.reverseaug_pil(augset, [t.clone() for t in outs], K)
        got = A.reverseaug(augset, [t.clone().to(dev) for t in outs], K)
        torch.cuda.synchronize()
        for v in range(n_views):
            assert torch.equal(got[v].cpu(), ref[v]), (B, K, H, W, v)


def rnd_tensor(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def test_predict_mask_equals_reference_argmax(dev, golden):
    """argmax(softmax(logits,1),1) as uint8 in one kernel vs the reference's two torch calls on the CPU
    (trainchaos_proposed_30cases1labeled.py:407-409), on golden logits, on multi-class logits and on exact ties."""
    import aide_b200 as A
    lg = golden["loss"]["logits"] if "loss" in golden and "logits" in golden["loss"] else rnd(4, 2, 24, 40, seed=3)
    for t in (lg, rnd(2, 5, 17, 9, seed=4), torch.zeros(1, 3, 4, 4), rnd(3, 2, 32, 32, seed=5).round()):
        ref = torch.argmax(F.softmax(t, dim=1), dim=1).to(torch.uint8)
        got = A.predict_mask(t.to(dev))
        assert got.dtype == torch.uint8 and torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("fmt", [0, 3])
@pytest.mark.parametrize("C,pooled", [(32, True), (64, False), (256, True)])
def test_spatial_attention_gate_forward_backward(dev, fmt, C, pooled):
    """Spatial_Attention (netblocks.py:68-89) + the gating product t = gate * y as the engine runs it: forward
    (aide_sa_fwd, aide_bn_finalize, aide_sa_gate_apply incl. the 2x2 max-pool of t) and backward (aide_sa_bwd_gate,
    aide_sa_bwd_chain) against torch autograd on the CPU, one same-resolution and one max-pool routed upstream gradient."""
    import ctypes as Ct
    from aide_b200 import ops
    from aide_b200._lib import call, lib
    N, H, W, r, dil = 2, 16, 24, C // 16, 4
    g = torch.Generator().manual_seed(C)
    y = torch.randn(N, C, H, W, generator=g).relu()
    if fmt == 3:
        y = ops.to_nchw(ops.from_nchw(y.to(dev), 3)).cpu()          # the value the operand planes actually hold
    y.requires_grad_()
    mk = lambda *s_, sc=0.3: (torch.randn(*s_, generator=g) * sc).requires_grad_()
    w1, b1, w2, b2 = mk(r, C, 1, 1), mk(r), mk(r, r, 3, 3), mk(r)
    w3, b3, w4, b4 = mk(r, r, 3, 3), mk(r), mk(1, r, 1, 1), mk(1)
    gamma, beta = (torch.rand(1, generator=g) + 0.5).requires_grad_(), mk(1)
    rm, rv = torch.zeros(1), torch.ones(1)
    a = F.conv2d(y, w1, b1)
    a = F.conv2d(a, w2, b2, padding=dil, dilation=dil)
    a = F.conv2d(a, w3, b3, padding=dil, dilation=dil)
    a = F.conv2d(a, w4, b4)
    gate = torch.sigmoid(F.batch_norm(a, rm, rv, gamma, beta, True, 0.1, 1e-5))
    t = gate * y
    gd = torch.randn(N, C, H, W, generator=g)
    gp = torch.randn(N, C, H // 2, W // 2, generator=g)
    loss = (t * gd).sum() + ((F.max_pool2d(t, 2, 2) * gp).sum() if pooled else 0.0)
    leaves = [y, w1, b1, w2, b2, w3, b3, w4, b4, gamma, beta]
    ref = torch.autograd.grad(loss, leaves)
    # ---- engine
    st = torch.cuda.current_stream().cuda_stream
    D = lambda x: x.detach().to(dev).contiguous()
    ya = ops.from_nchw(D(y), fmt)
    dW = [D(x) for x in (w1, b1, w2, b2, w3, b3, w4, b4, gamma, beta)]
    f32 = lambda *s_: torch.empty(*s_, dtype=torch.float32, device=dev)
    a1, a2, a3, aa, gt = f32(N, H, W, r), f32(N, H, W, r), f32(N, H, W, r), f32(N, H, W), f32(N, H, W)
    rows = lib.aide_sa_stat_rows(N, H, W)
    stat = f32(rows, 2)
    call("aide_sa_fwd", fmt, *ya.view(), C, r, dil, *[x.data_ptr() for x in dW[:8]], a1.data_ptr(), a2.data_ptr(),
         a3.data_ptr(), aa.data_ptr(), stat.data_ptr(), N, H, W, st)
    drm, drv = torch.zeros(1, device=dev), torch.ones(1, device=dev)
    ss, mr = ops.bn_finalize(stat.view(rows, 2, 1), N * H * W, dW[8], dW[9], drm, drv, True)
    out = ops.Act(N, H, W, C, fmt, dev)
    pl = ops.Act(N, H // 2, W // 2, C, fmt, dev)
    none = (None, None, 0, 0)
    call("aide_sa_gate_apply", fmt, *ya.view(), aa.data_ptr(), ss.data_ptr(), N, N, H, W, C, gt.data_ptr(), *out.view(),
         *(pl.view() if pooled else none), *none, st)
    tol = 2e-5 if fmt == 0 else 3e-5
    assert relmax(ops.to_nchw(out), t.detach()) < tol
    assert relmax(gt, gate.detach()[:, 0]) < 1e-5
    assert relmax(drm, rm) < 1e-5 and relmax(drv, rv) < 1e-5            # F.batch_norm updated rm / rv in place
    if pooled:
        assert relmax(ops.to_nchw(pl), F.max_pool2d(t.detach(), 2, 2)) < tol
    nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous().to(dev)
    d0, p0 = nhwc(gd), nhwc(gp)
    dptr, dct, dco = (Ct.c_void_p * 3)(d0.data_ptr()), (Ct.c_int * 3)(C), (Ct.c_int * 3)(0)
    pptr, pct, pco = (Ct.c_void_p * 3)(p0.data_ptr()), (Ct.c_int * 3)(C), (Ct.c_int * 3)(0)
    dy, dah = f32(N, H, W, C), f32(N, H, W)
    prow = lib.aide_sa_bwd_rows(N, H, W, 1 if pooled else 0)
    part = f32(prow, 2)
    call("aide_sa_bwd_gate", fmt, *ya.view(), C, gt.data_ptr(), aa.data_ptr(), mr.data_ptr(), N, H, W, dptr, dct, dco, 1,
         pptr, pct, pco, 1 if pooled else 0, dy.data_ptr(), dah.data_ptr(), part.data_ptr(), st)
    nws = lib.aide_sa_bwd_workspace_floats(C, r, N, H, W)
    ws = f32(nws)
    gsz = [r * C, r, r * r * 9, r, r * r * 9, r, r, 1, 2]
    gbuf = f32(sum(gsz))
    offs = [sum(gsz[:i]) for i in range(len(gsz))]
    gp_ = [gbuf.data_ptr() + 4 * o for o in offs]
    call("aide_sa_bwd_chain", fmt, *ya.view(), C, r, dil, dW[0].data_ptr(), dW[2].data_ptr(), dW[4].data_ptr(),
         dW[6].data_ptr(), dW[8].data_ptr(), a1.data_ptr(), a2.data_ptr(), a3.data_ptr(), aa.data_ptr(), mr.data_ptr(),
         dah.data_ptr(), part.data_ptr(), prow, N, H, W, ws.data_ptr(), nws, dy.data_ptr(), *gp_, st)
    torch.cuda.synchronize()
    got = [gbuf[o:o + n] for o, n in zip(offs, gsz)]
    gtol = 2e-4
    assert relmax(ops.nhwc_to_nchw(dy), ref[0]) < gtol, "dy"
    for name, gg, rr in zip(("w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4"), got[:8], ref[1:9]):
        if name in ("b3", "b4"):
            # a constant added before BatchNorm2d(1) cancels: the true gradient is zero (rounding noise on both sides)
            assert (gg.cpu() - rr.flatten()).abs().max().item() < gtol * ref[5].abs().max().item(), name
        else:
            assert relmax(gg, rr.flatten()) < gtol, name
    assert relmax(got[8][1:2], ref[9]) < gtol and relmax(got[8][0:1], ref[10]) < gtol      # {dbeta, dgamma}


def test_forward_augmentation_matches_pil_bit_for_bit(dev, oracle):
    """GPU forward augmentation of the input images (aide_forward_aug) against the reference loader's chain through PIL
    and torch (datasetchaos_proposed/transform.py: rotate BILINEAR on uint8 RGB, FLIP_LEFT_RIGHT, ToTensor, Normalize)."""
    import random
    import numpy as np
    import aide_b200 as A
    rng = np.random.default_rng(3)
    for (H, W) in ((64, 64), (48, 80)):
        B = 6
        imgs = rng.integers(0, 256, size=(B, H, W, 3), dtype=np.uint8)
        degs = [random.Random(i).random() * 120 - 60 for i in range(B - 2)] + [0.0, 180.0]
        flips = [i % 2 for i in range(B)]
        base = torch.from_numpy(imgs.transpose(0, 3, 1, 2)).float() / 255.0
        mean, std = base.mean(dim=(2, 3)), base.std(dim=(2, 3))                 # transform.py:140-147, per image
        ref = torch.stack([oracle.forward_aug_pil(imgs[b], degs[b], bool(flips[b]), mean[b], std[b]) for b in range(B)])
        got = A.forward_aug(torch.from_numpy(imgs).to(dev), degs, flips, mean, std)
        assert torch.equal(got.cpu(), ref), (H, W, (got.cpu() - ref).abs().max().item())
    # the batch helper draws like the loader: 4 rotations, then 4 flips, per sample
    u1 = torch.from_numpy(rng.integers(0, 256, size=(2, 32, 32, 3), dtype=np.uint8)).to(dev)
    m, s_ = torch.full((2, 3), 0.4), torch.full((2, 3), 0.2)
    augset, views = A.augmented_views(u1, u1, m, s_, m, s_, rng=random.Random(7))
    assert len(views) == 4 and views[0][0].shape == (2, 3, 32, 32) and augset["augno"] == [4, 4]
    assert all(-60.0 <= d < 60.0 for k in range(1, 5) for d in augset[f"degree{k}"])


def test_pixel_loss_maps_against_autograd(dev):
    """The per-pixel maps behind CrossEntropyLoss2d('none'), MulticlassMSELoss('none'), the KL + CE drop map of
    Coteachingloss_dropimagedroppixel and the ceil-mode max-pool of Coteachingloss_dropregionce: values and gradients
    against the reference formulas evaluated by torch on the CPU (utils/loss2d.py:5-13,109-117, coteach_loss.py:85-92)."""
    import aide_b200 as A
    from aide_b200.losses import maxpool_nchw, pixel_loss_map
    g = torch.Generator().manual_seed(9)
    N, H, W = 3, 20, 28
    a = (torch.randn(N, 2, H, W, generator=g) * 2).requires_grad_()
    b = (torch.randn(N, 2, H, W, generator=g) * 2).requires_grad_()
    t = (torch.rand(N, H, W, generator=g) < 0.3).long()
    t[0, :3] = 255                                                    # ignored pixels
    up = torch.randn(N, H, W, generator=g)
    # cross-entropy map with class weights and ignore_index
    wcl = torch.tensor([0.3, 1.7])
    ref = F.nll_loss(F.log_softmax(a, dim=1), t, weight=wcl, reduction="none", ignore_index=255)
    ga_ref, = torch.autograd.grad((ref * up).sum(), a)
    ad = a.detach().to(dev).requires_grad_()
    got = A.CrossEntropyLoss2d(weight=wcl, reduction="none")(ad, t.to(dev))
    (got * up.to(dev)).sum().backward()
    assert relmax(got, ref.detach()) < 1e-5 and relmax(ad.grad, ga_ref) < 1e-5
    # bidirectional KL + CE (no ignore: coteach_loss.py uses plain nll_loss)
    t2 = (torch.rand(N, H, W, generator=g) < 0.3).long()
    pa, pb = F.softmax(a, dim=1), F.softmax(b, dim=1)
    kl = (pa * torch.log(pa / pb)).sum(1) + (pb * torch.log(pb / pa)).sum(1)
    ref = kl + F.nll_loss(F.log_softmax(a, dim=1), t2, reduction="none")
    ga_ref, gb_ref = torch.autograd.grad((ref * up).sum(), [a, b])
    ad, bd = a.detach().to(dev).requires_grad_(), b.detach().to(dev).requires_grad_()
    got = pixel_loss_map(ad, t2.to(dev), logits2=bd, kl=True)
    (got * up.to(dev)).sum().backward()
    assert relmax(got, ref.detach()) < 2e-5 and relmax(ad.grad, ga_ref) < 2e-5 and relmax(bd.grad, gb_ref) < 2e-5
    # MulticlassMSELoss
    q = F.softmax(torch.randn(N, 2, H, W, generator=g), dim=1)
    up4 = torch.randn(N, 2, H, W, generator=g)
    ref = F.mse_loss(F.softmax(a, dim=1), q, reduction="none")
    ga_ref, = torch.autograd.grad((ref * up4).sum(), a)
    ad = a.detach().to(dev).requires_grad_()
    got = A.MulticlassMSELoss(reduction="none")(ad, q.to(dev))
    (got * up4.to(dev)).sum().backward()
    assert relmax(got, ref.detach()) < 1e-5 and relmax(ad.grad, ga_ref) < 1e-5
    assert abs(A.MulticlassMSELoss()(a.detach().to(dev), q.to(dev)).item() - ref.mean().item()) < 1e-6
    # ceil-mode max-pool with stride = kernel (odd sizes: partial windows at the border)
    x = torch.randn(2, 2, 21, 30, generator=g).requires_grad_()
    ref = F.max_pool2d(x, kernel_size=(2, 4), stride=(2, 4), padding=0, ceil_mode=True)
    upp = torch.randn(ref.shape, generator=g)
    gx_ref, = torch.autograd.grad((ref * upp).sum(), x)
    xd = x.detach().to(dev).requires_grad_()
    got = maxpool_nchw(xd, 2, 4)
    (got * upp.to(dev)).sum().backward()
    assert torch.equal(got.detach().cpu(), ref.detach()) and torch.equal(xd.grad.cpu(), gx_ref)


def test_f16_saturation_flag_is_sticky_and_resettable(dev):
    """F16X2 planes hold x * 2^8 in fp16: |x| >= 255.87 is clipped.  The conversion kernels raise a sticky device flag
    (aide_f16_saturated) so that a silent loss of fp32 parity cannot go unnoticed."""
    from aide_b200 import ops
    from aide_b200._lib import lib
    lib.aide_f16_saturated(1)
    x = torch.randn(1, 32, 8, 8).to(dev)
    a = ops.from_nchw(x, 3)
    torch.cuda.synchronize()
    assert lib.aide_f16_saturated(0) == 0
    x[0, 3, 2, 2] = 300.0
    a = ops.from_nchw(x, 3)
    torch.cuda.synchronize()
    assert lib.aide_f16_saturated(0) == 1 and lib.aide_f16_saturated(1) == 1        # sticky until reset
    assert lib.aide_f16_saturated(0) == 0
    assert abs(ops.to_nchw(a)[0, 3, 2, 2].item() - 65504.0 / 256.0) < 1e-3           # clipped, not inf
