"""Turn the raw ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_r1.csv profiles/r1_launches
    python tools/summarize_profiles.py full gpurun_out/conv_full.ncu-rep profiles/r1_conv_f16x2_full

`launches`: the `ncu --metrics gpu__time_duration.sum` launch list of `bench.py`.  Steps are delimited by the two
Adam kernels that end each AIDE step; the LAST complete step (a timed CUDA-graph replay) is written out launch by
launch (<out>_step.csv) and aggregated per kernel (<out>_summary.md).
`full`: selected metrics of an `ncu --set full` report (<out>.csv, one column per captured launch).
"""
import collections
import csv
import re
import subprocess
import sys


def short(name: str) -> str:
    n = re.sub(r"^void ", "", name)
    m = re.match(r"([\w:]+)(<[^(]*>)?", n)
    return ((m.group(1) + (m.group(2) or "")) if m else n)[:100]


def launches(src: str, out: str) -> None:
    rows = []
    with open(src) as f:
        for line in f:
            if line.startswith('"ID"'):
                break
        for r in csv.reader(f):
            if len(r) >= 15:
                rows.append((int(r[0]), r[4], r[6], r[7], r[8], float(r[14].replace(",", ""))))
    ends = [i for i, r in enumerate(rows) if "adam_amsgrad_kernel" in r[1]]
    # two Adam kernels per step (one per net); a step = (previous step's 2nd Adam, this step's 2nd Adam]
    ends = ends[1::2]
    if len(ends) < 2:
        raise SystemExit("fewer than two complete steps in the launch list")
    # choose the last step whose length equals the most common step length (graph replays are identical)
    lens = [ends[i] - ends[i - 1] for i in range(1, len(ends))]
    common = collections.Counter(lens).most_common(1)[0][0]
    k = max(i for i in range(1, len(ends)) if ends[i] - ends[i - 1] == common)
    seg = rows[ends[k - 1] + 1: ends[k] + 1]
    with open(out + "_step.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ncu_id", "kernel", "stream", "block", "grid", "gpu__time_duration_ns"])
        for r in seg:
            w.writerow([r[0], short(r[1]), r[2], r[3], r[4], int(r[5])])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in seg:
        a = agg[short(r[1])]
        a[0] += 1
        a[1] += r[5]
    tot = sum(v[1] for v in agg.values())
    with open(out + "_summary.md", "w") as f:
        f.write(f"# ncu launch list of one AIDE step ({len(seg)} launches, serialised total {tot / 1e6:.3f} ms)\n\n")
        f.write(f"Source: `{src}` ({len(rows)} launches in the whole `bench.py` run, {len(ends)} complete steps; "
                f"step {k} shown).\nPer-launch times are cold-cache and serialised by ncu: compare SHARES, not "
                "absolutes (the two networks overlap on two streams in a real step).\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {t / 1e6:.3f} | {100 * t / tot:.1f} % |\n")
    print(f"wrote {out}_step.csv and {out}_summary.md ({len(seg)} launches, {tot / 1e6:.2f} ms)")


FULL_PAT = re.compile(
    r"^(Kernel Name|Grid Size|Block Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?"
    r"|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|launch__registers_per_thread"
    r"|launch__shared_mem_per_block_dynamic|launch__occupancy_limit_\w+|sm__warps_active\.avg\.pct_of_peak_sustained_active"
    r"|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__cycles_elapsed\.(avg|max)(\.per_second)?"
    r"|sm__cycles_active\.avg|sm__inst_executed_pipe_tensor\w*\.sum|sm__pipe_tensor[\w.]*"
    r"|\w+\.TriageCompute\.sm__pipe_tensor[\w.]*"
    r"|l1tex__m_xbar2l1tex_read_bytes(_mem_global_op_tma_ld)?\.sum(\.per_second)?"
    r"|lts__t_bytes\.sum(\.per_second)?|lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed"
    r"|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__cycles_active\.avg)$")


def full(src: str, out: str) -> None:
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for i, h in enumerate(hdr):
            if FULL_PAT.match(h):
                w.writerow([h, units[i]] + [r[i] for r in data])
    print(f"wrote {out}.csv ({len(data)} launches)")


def traffic(src: str, out: str) -> None:
    """`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` log of `tools/profile_conv.py --reps 1
    --layers ...` -> profiles/conv_traffic.json entry.  Extra argv: key (e.g. f16x2:B8:S256) then the layer specs in
    launch order."""
    import json
    import os
    key, layers = sys.argv[4], sys.argv[5:]
    per = []
    with open(src) as f:
        for line in f:
            if line.startswith('"ID"'):
                break
        cur = {}
        for r in csv.reader(f):
            if len(r) < 15 or "conv3x3" not in r[4]:
                continue
            v = float(r[14].replace(",", ""))
            unit = r[13].lower()
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
            cur.setdefault(int(r[0]), 0.0)
            cur[int(r[0])] += v
        per = [cur[k] for k in sorted(cur)]
    if len(per) != len(layers):
        raise SystemExit(f"{len(per)} conv launches in the log, {len(layers)} layer specs")
    data = {}
    if os.path.exists(out):
        data = json.load(open(out))
    data[key] = {",".join(l.split(",")[:3]): round(b) for l, b in zip(layers, per)}
    with open(out, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print(f"wrote {out}: {key} ({len(per)} layers, {sum(per) / 1e6:.1f} MB total)")


def hbm(src: str, out: str) -> None:
    """`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log of
    tools/profile_step.py -> <out>.md: per kernel launches, time, DRAM bytes, GB/s against the measured copy peak."""
    import json
    import os
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        peak = 6544.0
    per = {}
    with open(src) as f:
        for line in f:
            if line.startswith('"ID"'):
                break
        for r in csv.reader(f):
            if len(r) < 15:
                continue
            e = per.setdefault(int(r[0]), dict(name=short(r[4]), ns=0.0, bytes=0.0))
            v = float(r[14].replace(",", ""))
            unit = r[13].lower()
            if "time_duration" in r[12]:
                e["ns"] += v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
            else:
                e["bytes"] += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for e in per.values():
        a = agg[e["name"]]
        a[0] += 1
        a[1] += e["ns"]
        a[2] += e["bytes"]
    with open(out + ".md", "w") as f:
        f.write("# Memory-side kernels of one stacked pseudo-label forward (4 views x 8) + one train forward / backward "
                "(fuseunet, 256x256, F16X2)\n\n")
        f.write("ncu `gpu__time_duration.sum`, `dram__bytes_read.sum + dram__bytes_write.sum` per kernel (`tools/profile_step.py`, "
                f"eager launches, cold caches, serialised); peak = measured copy bandwidth {peak:.0f} GB/s.  Kernels that move "
                "little data (statistics folds, column reductions) are listed for their time.\n\n")
        f.write("| kernel | launches | total ms | DRAM GB | GB/s | of measured peak |\n|---|---:|---:|---:|---:|---:|\n")
        for name, (n, ns, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            gbs = b / ns if ns else 0.0
            f.write(f"| `{name}` | {n} | {ns / 1e6:.3f} | {b / 1e9:.3f} | {gbs:.0f} | {gbs / peak:.2f} |\n")
    print(f"wrote {out}.md ({len(per)} launches)")


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic, "hbm": hbm}[sys.argv[1]](sys.argv[2], sys.argv[3])
