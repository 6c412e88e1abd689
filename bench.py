#!/usr/bin/env python
"""bench.py -- throughput of the AIDE hot path on B200, with the roofline of its dominant kernel and the
reference's CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode parity|fast|exact]

A *step* is one iteration of the AIDE proposed loop on one synthetic batch (BASELINE.json config 3,
train_files/trainchaos_proposed_30cases1labeled.py:263-325): two fuseunets, per-GPU batch 8 of 2-modality
256x256 slices, 4 augmented train-mode forwards per net -> pseudo labels, 2 train forwards, per-image CE+Dice,
cross small-loss selection + weighted-MSE consistency, 2 backwards, (N>1: one gradient all-reduce per net over
NCCL), 2 Adam(amsgrad) updates.  Nothing is skipped or cached inside the timed region.

    value  = dual-network train slices/s, WHOLE job (all ranks), inputs resident in HBM
    e2e    = the same through AideTrainer.step_from_host: pinned host buffers, H2D copies and the D2H read of
             the losses inside the timed region
    roofline     = the tcgen05 conv3x3 kernel (fwd/dgrad), every fuseunet layer shape timed alone with CUDA events
    cpu_baseline = the oracle (CPU restatement of the reference step, torch CPU ops, all host threads) on a
                   bounded sample of the same workload
Under torchrun every rank runs its own batch (weak scaling); timing = max over ranks between barriers.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

FUSEUNET_FWD_GFLOP_256 = 116.207          # per slice per net, conv FLOPs only
METRIC = "dual-FuseUNet train slices/sec @256x256x2 (AIDE proposed step)"


L2_NOTE = ("per-step working set (activations + weights, several GB) exceeds the 126 MB L2; distinct synthetic batches "
           "alternate between timed steps")


def workload_config(B, S, world, model="fuseunet", **extra):
    """The `config` object BOTH arms print -- identical for the same workload; arm-specific settings (precision mode,
    CUDA graph, ...) go to the top-level "engine" key."""
    unet = model == "unet"
    fwd = {256: 130.703, 320: 204.223, 512: 522.812}.get(S) if unet else FUSEUNET_FWD_GFLOP_256 * (S / 256.0) ** 2
    cfg = {"workload": ("AIDE proposed step, kidney flavour: 2x UNet (single modality), 4 eval-mode pseudo-label forwards + "
                        "train forward + backward per net, co-teaching selection, Adam-amsgrad (BASELINE.json configs[4] shape)"
                        if unet else
                        "AIDE proposed step: 2x fuseunet, 4 augmented forwards + train forward + backward per net, "
                        "co-teaching selection, Adam-amsgrad (BASELINE.json configs[2]; configs[3] for N>1)"),
           "model": model, "per_gpu_batch": B, "global_batch": B * world, "img_size": S, "modalities": 1 if unet else 2,
           "aug_views": 4, "rate": 0.25, "parallelism": f"dp{world}",
           "algorithmic_gflop_per_slice": round((3 + 4) * 2 * fwd, 1) if fwd else None, "l2": L2_NOTE}
    cfg.update(extra)
    return cfg


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--mode", default=os.environ.get("AIDE_B200_MODE", "parity"),
                    choices=["parity", "parity_tf32", "parity_mixed", "fast", "exact"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch (slices per network per step)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--model", default="fuseunet", choices=["fuseunet", "unet"],
                    help="unet = BASELINE.json configs[4] shape (single-modal UNet pair, kidney flavour: eval-mode "
                         "pseudo-label forwards); the default is the metric's configuration")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch slices per GPU (default); strong: --global-batch slices split over the ranks "
                         "(SURVEY.md 8d: fixed global 64)")
    ap.add_argument("--global-batch", type=int, default=64)
    ap.add_argument("--global-select", action="store_true",
                    help="nn.DataParallel selection semantics: all-gather the per-image losses, rank the global batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the fast-mode / train-only side measurements")
    ap.add_argument("--roofline-json", default="", help="also write the per-layer conv table to this file")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# synthetic workload: randn images, Bernoulli(0.08) masks, 4 augmented views, rate 0.25
# ---------------------------------------------------------------------------------------------------------
def make_batch(B, S, seed, device=None, pin=False, modalities=2):
    g = torch.Generator().manual_seed(seed)
    img = lambda: torch.randn(B, 3, S, S, generator=g)
    if modalities == 1:
        x = img()
        t1 = (torch.rand(B, S, S, generator=g) < 0.08).long()
        t2 = (torch.rand(B, S, S, generator=g) < 0.08).long()
        augs = [img() for _ in range(4)]
        mv = (lambda t: t.pin_memory()) if pin else ((lambda t: t.to(device)) if device is not None else (lambda t: t))
        return dict(x=mv(x), t1=mv(t1), t2=mv(t2), augs=[mv(a) for a in augs])
    x = (img(), img())
    t1 = (torch.rand(B, S, S, generator=g) < 0.08).long()
    t2 = (torch.rand(B, S, S, generator=g) < 0.08).long()
    augs = [(img(), img()) for _ in range(4)]
    mv = (lambda t: t.pin_memory()) if pin else ((lambda t: t.to(device)) if device is not None else (lambda t: t))
    return dict(x=tuple(mv(t) for t in x), t1=mv(t1), t2=mv(t2), augs=[tuple(mv(t) for t in a) for a in augs])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.th.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and args.impl == "engine":
        torch.cuda.set_device(local)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def timed_steps(fn, steps, world, device):
    """barrier + sync, K steps between CUDA events, sync + barrier; returns max-over-ranks milliseconds."""
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    timed_steps.host_ms = (time.perf_counter() - t0) * 1e3 / max(steps, 1)     # CPU time to ENQUEUE one step
    e1.record()
    torch.cuda.synchronize(device)
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    timed_steps.rank_ms = [round(ms.item() / max(steps, 1), 3)]
    if world > 1:
        allms = torch.empty(world, device=device)
        torch.distributed.all_gather_into_tensor(allms, ms)
        timed_steps.rank_ms = [round(v / max(steps, 1), 3) for v in allms.tolist()]      # per-rank skew, for the record
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.barrier()
    return ms.item()


# ---------------------------------------------------------------------------------------------------------
# roofline of the dominant kernel: conv3x3 on tcgen05, every fuseunet layer shape, timed alone
# ---------------------------------------------------------------------------------------------------------
FMT_NAMES = {0: "f32", 1: "tf32x2", 2: "bf16", 3: "f16x2"}


def conv_roofline(fmt, B, S, device, peaks):
    import aide_b200 as A
    from aide_b200 import engine as E, ops
    plan = E.plan_fuseunet(2)
    shapes = {}
    for u in plan.units:
        if u.first:
            continue
        key = (u.cin, u.cout, S >> u.level)
        shapes[key] = shapes.get(key, 0) + 1
    rows, tot_flop, tot_ms = [], 0.0, 0.0
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)     # > 126 MB L2
    for (cin, cout, hw), count in sorted(shapes.items(), key=lambda kv: -kv[0][2]):
        x = ops.Act(B, hw, hw, cin, fmt, device)
        x.planes.normal_()
        w = torch.randn(cout, cin, 3, 3, device=device) * (9 * cin) ** -0.5
        bias = torch.zeros(cout, device=device)
        w0, w1, keep = ops.weight_prep(w, fmt)
        z = torch.empty((B, hw, hw, cout), dtype=torch.float32, device=device)
        nrows = A.lib.aide_conv3x3_stat_rows(fmt, cin, cout, B, hw, hw)
        part = torch.empty((nrows, 2, cout), dtype=torch.float32, device=device)
        st = torch.cuda.current_stream().cuda_stream

        def launch():
            ops.call("aide_conv3x3_fwd", fmt, x.p0, x.p1, x.C, 0, cin, w0, w1, bias.data_ptr(), z.data_ptr(), cout, 0,
                     cout, B, hw, hw, part.data_ptr(), st)
        for _ in range(3):
            launch()
        reps, ms = 5, 0.0
        for _ in range(reps):
            flush.zero_()                                   # cold L2 for every timed launch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); launch(); e1.record()
            torch.cuda.synchronize(device)
            ms += e0.elapsed_time(e1)
        ms /= reps
        flop = 2.0 * B * hw * hw * cout * cin * 9
        rows.append(dict(cin=cin, cout=cout, hw=hw, launches_per_fwd=count, ms=round(ms, 4),
                         tflops=round(flop / ms / 1e9, 1)))
        tot_flop += flop * count
        tot_ms += ms * count
        del x, z, part
    achieved = tot_flop / tot_ms / 1e9
    traffic, traffic_note = conv_traffic(fmt, B, S, rows)
    peak = peaks.get("bf16_tflops")
    src = "measured (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)"
    if peak is None:
        peak, src = 1590.0, "fallback (B200_PROFILING.md)"
    passes = {1: 3, 2: 1, 3: 3}.get(fmt, 0)
    ceiling = {1: peak / 6.0, 2: peak, 3: peak / 3.0}.get(fmt, peak)
    sustained = peaks.get("bf16_tflops_sustained")
    return dict(bound="tensor", kernel=f"conv3x3_halo2_tc_kernel / conv3x3_halo_tc_kernel<{FMT_NAMES[fmt]}> (persistent tcgen05 "
                                       "implicit GEMM on TMA halo tiles: CTA pairs (cta_group::2, M = 256) or single CTAs per the "
                                       "measured plan table; dgrad is the same kernel; maps below 8x8 run conv3x3_fwd_tc_kernel)",
                achieved=round(achieved, 1), peak=peak, unit="TFLOP/s", frac=round(achieved / peak, 4),
                frac_sustained=round(achieved / sustained, 4) if sustained else None, traffic=traffic,
                traffic_note=traffic_note, launches=int(sum(r["launches_per_fwd"] for r in rows)),
                algorithmic_gflop_per_launch=round(tot_flop / 1e9 / sum(r["launches_per_fwd"] for r in rows), 2),
                peak_source=src, operand_format=FMT_NAMES[fmt], mma_passes=passes,
                frac_of_format_ceiling=round(achieved / ceiling, 4),
                note=("algorithmic conv FLOPs (2*B*H*W*Cout*Cin*9) of one fuseunet forward's tensor-core layers / summed "
                      "CUDA-event time of one launch per layer, L2 flushed before each launch.  Split-precision formats "
                      "issue 3 MMAs per algorithmic product: f16x2 at the bf16 rate (ceiling peak/3), tf32x2 at half of "
                      "it (ceiling peak/6)"),
                layers=rows)


def conv_traffic(fmt, B, S, rows):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, averaged over the launches of one forward,
    from the committed ncu capture of tools/profile_conv.py (profiles/conv_traffic.json); None if the capture does
    not cover this format / batch / size."""
    path = os.path.join(ROOT, "profiles", "conv_traffic.json")
    try:
        cap = json.load(open(path))
    except (OSError, ValueError):
        return None, "no ncu capture committed"
    ent = cap.get(f"{FMT_NAMES[fmt]}:B{B}:S{S}")
    if not ent:
        return None, f"profiles/conv_traffic.json has no capture for {FMT_NAMES[fmt]} at batch {B}, {S}x{S}"
    tot, n = 0.0, 0
    for r in rows:
        b = ent.get(f"{r['cin']},{r['cout']},{r['hw']}")
        if b is None:
            return None, "capture incomplete"
        tot += b * r["launches_per_fwd"]
        n += r["launches_per_fwd"]
    return round(tot / n), ("average DRAM bytes per conv launch over the %d launches of one forward, ncu "
                            "dram__bytes_read.sum + dram__bytes_write.sum (profiles/conv_traffic.json)" % n)


def _time_launch(fn, flush, device, reps=5):
    for _ in range(3):
        fn()
    ms = 0.0
    for _ in range(reps):
        flush.zero_()                                       # cold L2 for every timed launch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(device)
        ms += e0.elapsed_time(e1)
    return ms / reps


def wgrad_roofline(fmt, B, S, device, peaks):
    """conv3x3 weight-gradient kernel (conv_wgrad_halo.cu), every fuseunet layer shape
    at the train batch, incl. the split-K reduce.  FLOPs = 2*B*H*W*Cout*Cin*9 per launch."""
    import aide_b200 as A
    from aide_b200 import engine as E, ops
    plan = E.plan_fuseunet(2)
    shapes = {}
    for u in plan.units:
        if not u.first:
            key = (u.cin, u.cout, S >> u.level)
            shapes[key] = shapes.get(key, 0) + 1
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    rows, tot_flop, tot_ms = [], 0.0, 0.0
    inv = torch.full((1,), 1.0 / 256.0, dtype=torch.float32, device=device)
    for (cin, cout, hw), count in sorted(shapes.items(), key=lambda kv: -kv[0][2]):
        x = ops.Act(B, hw, hw, cin, fmt, device); x.planes.normal_()
        dz = ops.Act(B, hw, hw, cout, fmt, device); dz.planes.normal_()
        nbytes = A.lib.aide_conv3x3_wgrad_workspace_bytes(fmt, cin, cout, B, hw, hw)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
        dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=device)
        st = torch.cuda.current_stream().cuda_stream
        ms = _time_launch(lambda: ops.call("aide_conv3x3_wgrad", fmt, x.p0, x.p1, x.C, 0, cin, dz.p0, dz.p1,
                                           inv.data_ptr() if fmt == 3 else None, cout, B, hw, hw, ws.data_ptr(), nbytes,
                                           dw.data_ptr(), st), flush, device)
        flop = 2.0 * B * hw * hw * cout * cin * 9
        rows.append(dict(cin=cin, cout=cout, hw=hw, launches=count, ms=round(ms, 4), tflops=round(flop / ms / 1e9, 1)))
        tot_flop += flop * count
        tot_ms += ms * count
        del x, dz, ws
    achieved = tot_flop / tot_ms / 1e9
    peak = peaks.get("bf16_tflops") or 1590.0
    ceiling = {1: peak / 6.0, 2: peak, 3: peak / 3.0}.get(fmt, peak)
    return dict(bound="tensor", kernel="wgrad_halo_kernel (three dx taps per CTA on one halo box; Cin = 64: hi/lo planes stacked along M; 32-channel layers on zero-padded 64-channel boxes) + split-K reduce",
                achieved=round(achieved, 1), peak=peak, unit="TFLOP/s", frac=round(achieved / peak, 4),
                frac_of_format_ceiling=round(achieved / ceiling, 4), batch_per_launch=B, operand_format=FMT_NAMES[fmt],
                layers=rows)


def hbm_roofline(fmt, B, S, device, peaks):
    """The BatchNorm-apply kernel (bn_relu_apply: reads fp32 z, writes the operand planes + the 2x2 max-pooled copy) at
    the largest level of the stacked forward -- the heaviest HBM-bound kernel of the step.  Algorithmic bytes per
    element: 4 read + planes*esize written (+ 1/4 of that for the pooled copy)."""
    from aide_b200 import ops
    C_, hw = 64, S
    z = torch.randn(B, hw, hw, C_, device=device)
    ss = torch.randn(2, C_, device=device)
    y = ops.Act(B, hw, hw, C_, fmt, device)
    pl = ops.Act(B, hw // 2, hw // 2, C_, fmt, device)
    none = (None, None, 0, 0)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    ms = _time_launch(lambda: ops.call("aide_bn_relu_apply", fmt, z.data_ptr(), B, hw, hw, C_, ss.data_ptr(), *y.view(),
                                       *pl.view(), *none, st), flush, device)
    out_b = {0: 4, 1: 8, 2: 2, 3: 4}[fmt]
    nbytes = B * hw * hw * C_ * (4 + out_b * 1.25)
    peak = peaks.get("hbm_gbs") or 6500.0
    gbs = nbytes / ms / 1e6
    return dict(bound="hbm", kernel=f"bn_relu_apply_kernel<{FMT_NAMES[fmt]}> (+ fused 2x2 max-pool copy), {B}x{hw}x{hw}x{C_}",
                achieved=round(gbs, 1), peak=peak, unit="GB/s", frac=round(gbs / peak, 4), bytes_per_launch=int(nbytes),
                ms=round(ms, 4))


def eval_single_slice(args, S, device, n_slices=200):
    """The per-epoch evaluation / pseudo-label rewrite loops (trainchaos_proposed_30cases1labeled.py:373-496,
    evalchaos_comparison_1cases.py:203-214): one slice at a time -- H2D of the two modalities, eval-mode forward
    (net.graphed_eval: BatchNorm folded into the conv epilogues, one CUDA graph), softmax -> argmax mask
    (aide_argmax_mask), D2H of the uint8 mask."""
    import aide_b200 as A
    torch.manual_seed(2)
    net = A.fuseunet(num_classes=2, mode=args.mode).to(device).eval()
    g = torch.Generator().manual_seed(99)
    host = [tuple(torch.randn(1, 3, S, S, generator=g).pin_memory() for _ in range(2)) for _ in range(4)]
    masks = [torch.empty((1, S, S), dtype=torch.uint8).pin_memory() for _ in range(4)]
    dev_in = tuple(torch.empty(1, 3, S, S, device=device) for _ in range(2))
    f = net.graphed_eval(*dev_in)
    l0 = A.lib.aide_launch_count()

    def one(i):
        for d, h in zip(dev_in, host[i % 4]):
            d.copy_(h, non_blocking=True)
        with torch.no_grad():
            masks[i % 4].copy_(A.predict_mask(f(*dev_in)), non_blocking=True)

    for i in range(5):
        one(i)
    per_fwd = f.n_kernels
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_slices):
        one(i)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / n_slices
    return dict(value=round(1e3 / ms, 1), unit="slices/s", ms_per_slice=round(ms, 4), slices=n_slices,
                launches_per_slice=per_fwd + 1, h2d_bytes_per_slice=2 * 3 * S * S * 4, d2h_bytes_per_slice=S * S,
                api="net.graphed_eval(modal1, modal2) + predict_mask, one 256x256 slice per call (pinned host in / out)")


def eval_module_single_slice(args, S, device, n_slices=200):
    """The same loop written like the unmodified script (trainchaos_proposed_30cases1labeled.py:403-411): net.eval(),
    `with torch.no_grad(): output = net(inphase, outphase)`, torch.argmax(softmax(output), 1), .cpu() -- no engine-specific
    call.  The module captures its no-grad forward into a CUDA graph on the second call (nets._nograd_forward_graphed)."""
    import aide_b200 as A
    import torch.nn.functional as F
    torch.manual_seed(2)
    net = A.fuseunet(num_classes=2, mode=args.mode).to(device).eval()
    g = torch.Generator().manual_seed(99)
    host = [tuple(torch.randn(1, 3, S, S, generator=g) for _ in range(2)) for _ in range(4)]
    keep = {}

    def one(i):
        a, b = (t.to(device) for t in host[i % 4])
        with torch.no_grad():
            out = net(a, b)
        keep["mask"] = torch.argmax(F.softmax(out, dim=1), dim=1).cpu()

    for i in range(5):
        one(i)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_slices):
        one(i)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / n_slices
    return dict(value=round(1e3 / ms, 1), unit="slices/s", ms_per_slice=round(ms, 4), slices=n_slices,
                api="unmodified-script pattern: net.eval(); with torch.no_grad(): net(modal1, modal2); argmax(softmax).cpu() "
                    "(pageable host tensors, one 256x256 slice per call)")


def module_e2e(args, B, S, device, K):
    """The call pattern of an UNMODIFIED training script (train_files/trainchaos_proposed_30cases1labeled.py:263-325)
    on the drop-in modules: 8 detached augmented forwards, F.softmax / sharpen in torch, 2 train forwards through
    autograd, CEMDiceLossImage on index-gathered copies, weighted MulticlassMSELoss, loss.backward() twice,
    torch.optim.Adam(amsgrad=True).step() twice, loss.item() / Dice_fn per step -- eager launches, no CUDA graph."""
    import aide_b200 as A
    import torch.nn.functional as F
    torch.manual_seed(2)
    net1 = A.fuseunet(num_classes=2, mode=args.mode).to(device).train()
    net2 = A.fuseunet(num_classes=2, mode=args.mode).to(device).train()
    opt1 = torch.optim.Adam(net1.parameters(), lr=1e-4, amsgrad=True)
    opt2 = torch.optim.Adam(net2.parameters(), lr=1e-4, amsgrad=True)
    crit = A.CEMDiceLossImage(cediceweight=[1.0, 1.0], ceclassweight=torch.tensor([1.0, 1.0]), diceclassweight=[1.0, 1.0]).to(device)
    mse = A.MulticlassMSELoss(reduction="none").to(device)
    host = [make_batch(B, S, 777 + i, pin=True) for i in range(2)]
    rate, io = 0.25, {}

    def sharpen(m, T):
        m = torch.pow(m, T)
        return m / m.sum(dim=1).unsqueeze(dim=1)

    def step(i):
        hb = host[i % 2]
        d = lambda t: t.to(device, non_blocking=True)
        a1, a2 = [], []
        for a in hb["augs"]:
            xa, xb = d(a[0]), d(a[1])
            a1.append(net1(xa, xb).detach())
            a2.append(net2(xa, xb).detach())
        for k in range(len(a1)):
            m1, m2 = F.softmax(a1[k], dim=1), F.softmax(a2[k], dim=1)
            if k == 0:
                pl1, pl2 = m1, m2
            else:
                pl1 += m1; pl2 += m2
        pl1, pl2 = sharpen(pl1 / 4.0, 1.0), sharpen(pl2 / 4.0, 1.0)
        w1 = (1.0 - 4.0 * pl1[:, 0] * pl1[:, 1]).unsqueeze(dim=1)
        w2 = (1.0 - 4.0 * pl2[:, 0] * pl2[:, 1]).unsqueeze(dim=1)
        x1, x2, t1, t2 = d(hb["x"][0]), d(hb["x"][1]), d(hb["t1"]), d(hb["t2"])
        opt1.zero_grad(); opt2.zero_grad()
        o1, o2 = net1(x1, x2), net2(x1, x2)
        _, i1 = crit(o1, t2).sort()
        _, i2 = crit(o2, t1).sort()
        l1 = (crit(o1[i2[0:2]], t2[i2[0:2]]).mean() + (1.0 - rate) * crit(o1[i2[2:]], t2[i2[2:]]).mean()) \
            + 10.0 * rate * (w2[i2[2:]] * mse(o1[i2[2:]], pl2[i2[2:]])).mean()
        l2 = (crit(o2[i1[0:2]], t1[i1[0:2]]).mean() + (1.0 - rate) * crit(o2[i1[2:]], t1[i1[2:]]).mean()) \
            + 10.0 * rate * (w1[i1[2:]] * mse(o2[i1[2:]], pl1[i1[2:]])).mean()
        l1.backward(retain_graph=True)
        opt1.step()
        l2.backward()
        opt2.step()
        io["v"] = l1.item() + l2.item() + A.Dice_fn(o1, t2).item() + A.Dice_fn(o2, t1).item()

    for i in range(2):
        step(i)
    ms = timed_steps(step, K, 1, device)
    h2d = sum(t.numel() * t.element_size() for t in list(host[0]["x"]) + [host[0]["t1"], host[0]["t2"]]
              + [t for a in host[0]["augs"] for t in a])
    return dict(value=round(B * K / (ms / 1e3), 3), unit="slices/s", ms_per_step=round(ms / K, 3),
                host_enqueue_ms_per_step=round(timed_steps.host_ms, 2), h2d_bytes_per_step=h2d, d2h_bytes_per_step=16,
                api="drop-in modules driven like the unmodified script: fuseunet.forward x10, CEMDiceLossImage / "
                    "MulticlassMSELoss, loss.backward() x2, torch.optim.Adam(amsgrad=True).step() x2 (eager, no graph)")


# ---------------------------------------------------------------------------------------------------------
# CPU legs.  kind "reference": the UNMODIFIED reference modules (baseline/_ref/, copied there from /root/reference by
# baseline/fetch_reference.py in the build container; git-ignored, travels to the GPU box) driven by the inline step of
# train_files/trainchaos_proposed_30cases1labeled.py:263-325 with torch.optim.Adam(amsgrad=True).  kind "port": the
# oracle (CPU restatement, bit-identical to the reference -- tests/golden/make_golden.py) when baseline/_ref is absent.
# ---------------------------------------------------------------------------------------------------------
def load_reference():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "models_twomodalinputs")):
        return None
    import types
    for name in ("matplotlib", "matplotlib.pyplot"):          # utils/metrics2d.py:6 imports matplotlib (not installed)
        sys.modules.setdefault(name, types.ModuleType(name))
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import models_twomodalinputs as m2, models_singlemodalinput as m1, utils as ru
    assert os.path.abspath(m2.__file__).startswith(ref) and os.path.abspath(ru.__file__).startswith(ref)
    return dict(fuseunet=m2.fuseunet, UNet=m1.UNet, utils=ru)


def reference_step_fn(ref, B, S, model="fuseunet"):
    """The reference's own modules, losses and optimiser; the loop body follows the script line by line (the script
    itself cannot be imported: pydicom / SimpleITK / skimage are not installed and it reads DICOM folders).  The
    augmented forwards run with autograd enabled and are then detached, as in :267-268; reverseaug (:271-272) is the
    identity for the benchmark's views (degree 0, no flip) and is skipped in BOTH arms."""
    import torch.nn.functional as F
    from torch.optim import Adam
    ru = ref["utils"]
    unet = model == "unet"
    torch.manual_seed(2)
    ctor = ref["UNet"] if unet else ref["fuseunet"]
    net1, net2 = ctor(num_classes=2), ctor(num_classes=2)
    opt1 = Adam(net1.parameters(), lr=1e-4, amsgrad=True)
    opt2 = Adam(net2.parameters(), lr=1e-4, amsgrad=True)
    crit = ru.CEMDiceLossImage(cediceweight=[1.0, 1.0], ceclassweight=torch.tensor([1.0, 1.0]), diceclassweight=[1.0, 1.0])
    mse = ru.MulticlassMSELoss(reduction="none")
    rate, segcor = 0.25, (1.0, 10.0)

    def sharpen(mask, T):                                     # :97-101
        m = torch.pow(mask, T)
        return m / m.sum(dim=1).unsqueeze(dim=1)

    def run(i):
        b = make_batch(B, S, 1234 + i, modalities=1 if unet else 2)
        ins = (lambda t: (t,)) if unet else (lambda t: t)
        net1.train(); net2.train()
        if unet:                                              # kidney flavour: eval-mode views (trainkidney...:267-268)
            net1.eval(); net2.eval()
        a1 = [net1(*ins(a)).detach() for a in b["augs"]]
        a2 = [net2(*ins(a)).detach() for a in b["augs"]]
        net1.train(); net2.train()
        for k in range(len(a1)):
            m1, m2 = F.softmax(a1[k], dim=1), F.softmax(a2[k], dim=1)
            if k == 0:
                pl1, pl2 = m1, m2
            else:
                pl1 += m1; pl2 += m2
        pl1, pl2 = sharpen(pl1 / float(len(a1)), 1.0), sharpen(pl2 / float(len(a2)), 1.0)
        w1 = (1.0 - 4.0 * pl1[:, 0] * pl1[:, 1]).unsqueeze(dim=1)
        w2 = (1.0 - 4.0 * pl2[:, 0] * pl2[:, 1]).unsqueeze(dim=1)
        t1, t2 = b["t1"], b["t2"]
        opt1.zero_grad(); opt2.zero_grad()
        o1, o2 = net1(*ins(b["x"])), net2(*ins(b["x"]))
        _, i1 = crit(o1, t2).sort()
        _, i2 = crit(o2, t1).sort()
        l1 = segcor[0] * (crit(o1[i2[0:2]], t2[i2[0:2]]).mean() + (1.0 - rate) * crit(o1[i2[2:]], t2[i2[2:]]).mean()) \
            + segcor[1] * rate * (w2[i2[2:]] * mse(o1[i2[2:]], pl2[i2[2:]])).mean()
        l2 = segcor[0] * (crit(o2[i1[0:2]], t1[i1[0:2]]).mean() + (1.0 - rate) * crit(o2[i1[2:]], t1[i1[2:]]).mean()) \
            + segcor[1] * rate * (w1[i1[2:]] * mse(o2[i1[2:]], pl1[i1[2:]])).mean()
        l1.backward(retain_graph=True)
        opt1.step()
        l2.backward()
        opt2.step()
        return l1.item() + ru.Dice_fn(o1, t2).item()          # the script reads both back every step (:327-331)
    return run


def cpu_step_fn(B, S, model="fuseunet"):
    """(step function, kind).  The reference itself when baseline/_ref is present, else the oracle port."""
    ref = load_reference()
    if ref is not None:
        return reference_step_fn(ref, B, S, model), "reference"
    from oracle import aide_oracle as O
    torch.manual_seed(2)
    unet = model == "unet"
    init, fwd = (O.init_unet, O.unet_forward) if unet else (O.init_fuseunet, O.fuseunet_forward)
    p1 = O.clone_params(init(2), requires_grad=True)
    p2 = O.clone_params(init(2), requires_grad=True)
    st1, st2 = {}, {}
    state = dict(step=0)

    def run(i):
        b = make_batch(B, S, 1234 + i, modalities=1 if unet else 2)
        ins = (lambda t: (t,)) if unet else (lambda t: t)
        r = O.aide_step(fwd, p1, p2, ins(b["x"]), [ins(a) for a in b["augs"]], b["t1"], b["t2"], 0.25,
                        flavour="kidney" if unet else "chaos")
        state["step"] += 1
        O.adam_amsgrad_step(p1, r["grads1"], st1, state["step"])
        O.adam_amsgrad_step(p2, r["grads2"], st2, state["step"])
        return float(r["loss1"].detach())
    return run, "port"


KIND_TEXT = {"reference": "the unmodified reference modules (baseline/_ref) + torch.optim.Adam(amsgrad=True)",
             "port": "oracle/aide_oracle.py (CPU restatement, bit-identical to the reference)"}


def cpu_baseline(B, S, model="fuseunet", budget_s=24.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    probe, kind = cpu_step_fn(3, S, model)     # 3 = smallest batch the step accepts (2 "clean" + >= 1 "rest")
    t0 = time.time(); probe(0); t_one = (time.time() - t0) / 3.0   # also the warm-up (oneDNN primitive cache)
    b = max(3, min(B, int(budget_s / 2.0 / max(t_one, 1e-3))))
    run, kind = cpu_step_fn(b, S, model)
    n, t0 = 0, time.time()
    while n < 1 or (time.time() - t0 < budget_s / 2.0 and n < 3):
        run(n); n += 1
    dt = (time.time() - t0) / n
    return dict(value=round(b / dt, 4), unit="slices/s", cores=cores, kind=kind,
                sample=f"{n} AIDE step(s) at batch {b} (of {B}), {S}x{S}, after one batch-3 warm-up step; "
                       f"{KIND_TEXT[kind]}, torch CPU fp32, {cores} threads, {dt:.2f} s/step")


def run_reference(args, world, rank):
    """--impl reference: the reference's CPU path on all host threads, rank 0 only.  Runs the CONFIGURED per-GPU batch
    whenever the whole --steps/--warmup run fits ~5 minutes (it does for the default configuration on a 16-core box);
    otherwise the batch is reduced and `sample_batch` says so."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, S = args.batch, args.size
    probe, kind = cpu_step_fn(3, S, args.model)
    t0 = time.time(); probe(0)
    t0 = time.time(); probe(1); t_one = (time.time() - t0) / 3.0          # second step: primitive caches are warm
    total = args.steps + args.warmup
    b = max(3, min(B, int(330.0 / total / max(t_one, 1e-3))))
    del probe
    run, kind = cpu_step_fn(b, S, args.model)
    for i in range(args.warmup):
        run(i)
    t0 = time.time()
    for i in range(args.steps):
        run(args.warmup + i)
    dt = (time.time() - t0) / max(args.steps, 1)
    v = b / dt
    sample = (f"{args.steps} timed AIDE steps at batch {b} (per-GPU batch of the configuration: {B}), {S}x{S}; "
              f"{KIND_TEXT[kind]} on {cores} host threads")
    cfg = workload_config(B, S, max(args.gpus, 1), args.model, **({} if b == B else {"sample_batch": b}))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "slices/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "engine": {"mode": f"cpu-{kind}", "threads": cores},
        "cpu_baseline": {"value": round(v, 4), "unit": "slices/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(v, 4), "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    world, rank, local = dist_setup(args)
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback (use --impl reference for the CPU path)")
    import aide_b200 as A
    from aide_b200.trainer import AideTrainer
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    B, S, K, W = args.batch, args.size, args.steps, max(args.warmup, 0)
    if args.scaling == "strong":
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} does not divide over {world} ranks")
        B = args.global_batch // world
        if 4 * B > 64:                   # a stacked 4-view forward of > 64 images needs tens of GB of scratch: run the views in turn
            os.environ["AIDE_B200_GROUP_AUGS"] = "0"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass

    unet = args.model == "unet"
    nmod = 1 if unet else 2

    def build(mode):
        tr = AideTrainer(args.model, mode=mode, device=device, seed=2 + 0, flavour="kidney" if unet else "chaos",
                         global_select=args.global_select)
        tr.broadcast_parameters(0)
        return tr

    tr = build(args.mode)
    n_pool = 2
    dev_batches = [make_batch(B, S, 1234 + 97 * rank + i, device=device, modalities=nmod) for i in range(n_pool)]

    def dev_step(i):
        b = dev_batches[i % n_pool]
        tr.step(b["x"], b["t1"], b["t2"], b["augs"], 0.25)

    for i in range(max(W, 3)):
        dev_step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = A.lib.aide_launch_count() + tr.graph_launches
    ms = timed_steps(dev_step, K, world, device)
    launches = A.lib.aide_launch_count() + tr.graph_launches - l0
    host_enqueue_ms = timed_steps.host_ms
    rank_ms = timed_steps.rank_ms
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * K / (ms / 1e3)

    # ---- end to end: pinned host buffers -> H2D -> step -> D2H of losses / Dice
    host_batches = [make_batch(B, S, 4321 + 97 * rank + i, pin=True, modalities=nmod) for i in range(n_pool)]
    io = {}

    def host_step(i):
        # like a training loop with a prefetching loader: the H2D copy of the NEXT batch is started right after this
        # step's launch and overlaps its compute (copy stream -> staging buffers -> D2D into the graph's inputs); every
        # step still performs one full H2D of a step's inputs and the synchronising D2H read of its own results
        _, io["h2d"], io["d2h"] = tr.step_from_host(host_batches[i % n_pool], 0.25, prefetch=host_batches[(i + 1) % n_pool])

    for i in range(max(2, n_pool)):
        host_step(i)                                 # the last warm-up step prefetches batch 0 = the first timed step's
    ms_e2e = timed_steps(host_step, K, world, device)
    e2e = dict(value=round(world * B * K / (ms_e2e / 1e3), 3), unit="slices/s", ms_per_step=round(ms_e2e / K, 3),
               h2d_bytes_per_step=io["h2d"], d2h_bytes_per_step=io["d2h"],
               api="aide_b200.trainer.AideTrainer.step_from_host(batch, rate, prefetch=next_batch): pinned host tensors in, python "
                   "floats out; the next batch's H2D runs on a copy stream during this step's compute")

    out = {
        "metric": METRIC, "value": round(value, 3), "unit": "slices/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": {"parity": "f16x2: split-precision tcgen05 (two fp16 planes, 3 kind::f16 MMAs per product = fp32-equivalent "
                            "22-bit products), fp32 accumulate in TMEM; gradients with a device-chosen power-of-two scale",
                  "parity_tf32": "tf32x2: split-precision tcgen05 (3 kind::tf32 MMAs per product), fp32 accumulate",
                  "parity_mixed": "tf32x2 for the train forward/backward, f16x2 for the pseudo-label forwards",
                  "fast": "bf16 operands, fp32 accumulate (NOT a parity mode)", "exact": "f32 CUDA cores"}[args.mode],
        "data": "synthetic",
        "config": workload_config(B, S, world, args.model),
        "engine": {"mode": args.mode, "cuda_graph": bool(tr.cuda_graph), "stacked_aug_forward": bool(tr.group_augs), "stacked_train_forward": bool(tr.stack_train),
                   "resident_batches": n_pool, "global_select": bool(tr.global_select),
                   "gradient_all_reduce": {"p2p": "aide_allreduce_p2p (NVLink peer memory, csrc/comm.cu)", "nccl": "torch.distributed.all_reduce (NCCL)",
                                           "none": None}[tr.comm_kind] if tr._comm else "disabled (AIDE_B200_NO_COMM diagnostic)"},
        "gpu_launches": int(launches), "gpu_launches_per_step": round(launches / K, 1),
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 2), "rank_ms_per_step": rank_ms,
        "clocks": clocks, "e2e": e2e,
    }
    out["algorithmic_tflops"] = round(value * out["config"]["algorithmic_gflop_per_slice"] / 1e3, 1)

    if unet:
        # side configuration: the step rate (+ the CPU reference beside it at N = 1); the roofline legs describe the
        # metric's fuseunet workload
        if rank == 0:
            if world == 1 and not args.no_cpu_baseline:
                out["cpu_baseline"] = cpu_baseline(B, S, args.model)
            print(json.dumps(out), flush=True)
    elif rank == 0:
        from aide_b200 import engine as E
        fmt_inf, fmt_train = E.mode_format(args.mode, False), E.mode_format(args.mode, True)
        # The dominant kernel launches of the step are the convolutions of the stacked pseudo-label forward (4 views x B
        # images per launch, 4/7 of the step's conv FLOPs); the train forward / dgrad run the same kernel at batch B.
        # (with the train forward riding as a fifth statistics group: 5 x B images per launch)
        Ba = (5 * B if (tr.stack_train and tr.flavour == "chaos") else 4 * B) if tr.group_augs else B
        rl = conv_roofline(fmt_inf, Ba, S, device, peaks)
        layers = {f"{FMT_NAMES[fmt_inf]}:B{Ba}": rl.pop("layers")}
        rl["batch_per_launch"] = Ba
        out["roofline"] = rl
        rl2 = conv_roofline(fmt_train, B, S, device, peaks)       # train forward + dgrad
        layers[f"{FMT_NAMES[fmt_train]}:B{B}"] = rl2.pop("layers")
        out["roofline_train_batch"] = {k: rl2[k] for k in ("kernel", "achieved", "frac", "frac_of_format_ceiling",
                                                           "operand_format", "mma_passes", "traffic")}
        out["roofline_train_batch"]["batch_per_launch"] = B
        if not args.no_extras:
            rw = wgrad_roofline(fmt_train, B, S, device, peaks)
            layers["wgrad"] = rw.pop("layers")
            out["roofline_wgrad"] = rw
            out["roofline_hbm"] = hbm_roofline(fmt_inf, Ba, S, device, peaks)
        if args.roofline_json:
            with open(args.roofline_json, "w") as f:
                json.dump(dict(mode=args.mode, batch=B, size=S, summary=rl, layers=layers), f, indent=1)
        if world == 1 and not args.no_extras:
            # R_train: forward + backward + Adam of both nets only (no pseudo-label forwards, no
            # consistency term) = 697.2 GFLOP per slice; context for BASELINE.json's 1000 slices/s target
            def train_only(i):
                b = dev_batches[i % n_pool]
                tr.step(b["x"], b["t1"], b["t2"], [], 0.25)
            for i in range(3):
                train_only(i)
            ms_t = timed_steps(train_only, K, world, device)
            out["train_only"] = dict(value=round(B * K / (ms_t / 1e3), 3), unit="slices/s", ms_per_step=round(ms_t / K, 3),
                                     algorithmic_tflops=round(B * K / (ms_t / 1e3) * 6 * FUSEUNET_FWD_GFLOP_256 * (S / 256.0) ** 2 / 1e3, 1),
                                     note="train forward + backward + Adam of both nets, no augmented forwards (R_train)")
        if world == 1 and not args.no_extras:
            try:
                out["eval_single_slice"] = eval_single_slice(args, S, device)
            except Exception as e:  # noqa: BLE001
                out["eval_single_slice"] = dict(value=None, error=f"{type(e).__name__}: {e}")
            try:
                out["eval_module_single_slice"] = eval_module_single_slice(args, S, device)
            except Exception as e:  # noqa: BLE001
                out["eval_module_single_slice"] = dict(value=None, error=f"{type(e).__name__}: {e}")
            try:
                out["e2e_module"] = module_e2e(args, B, S, device, max(2, min(K, 5)))
            except Exception as e:  # noqa: BLE001
                out["e2e_module"] = dict(value=None, error=f"{type(e).__name__}: {e}")
        if world == 1 and not args.no_extras and args.mode != "fast":
            del tr
            torch.cuda.empty_cache()
            tr = build("fast")
            for i in range(3):
                dev_step(i)
            ms_f = timed_steps(dev_step, K, world, device)
            for i in range(3):
                train_only(i)
            ms_ft = timed_steps(train_only, K, world, device)
            out["fast_mode"] = dict(value=round(B * K / (ms_f / 1e3), 3), unit="slices/s", ms_per_step=round(ms_f / K, 3),
                                    train_only=round(B * K / (ms_ft / 1e3), 3),
                                    note="single-pass bf16 operands; fails the 1e-3 logit parity bar (SURVEY.md 8d), "
                                         "reported for context only")
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(B, S, args.model)
            except Exception as e:  # noqa: BLE001 -- the GPU numbers above must still be reported
                out["cpu_baseline"] = dict(value=None, unit="slices/s", cores=os.cpu_count(), kind="port",
                                           sample=f"failed: {type(e).__name__}: {e}")
        print(json.dumps(out), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: the captured CUDA graphs of the step still hold the communicator's kernels,
        # and destroying the process group under them can block forever (seen on 2 x B200).  Every rank has finished
        # its work at the barrier; flush and exit with status 0.
        torch.distributed.barrier()
        torch.cuda.synchronize(device)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
