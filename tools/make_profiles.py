#!/usr/bin/env python
"""Turn gpurun_out/ ncu captures into the small, committed summaries under profiles/.

    python tools/make_profiles.py r01 target gpurun_out/r01_launches_target.csv gpurun_out/r01_full_target.ncu-rep
"""
import collections
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "profiles")


def short(name):
    n = name.replace("void ", "").replace("fyx::", "")
    return n.split("(")[0]


def launches(tag, workload, path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
    hdr = rows[0]
    i = {h: k for k, h in enumerate(hdr)}
    seq = []
    for r in rows[1:]:
        if r[i["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[i["Metric Value"]])
        unit = r[i["Metric Unit"]]
        us = v / 1e3 if unit.startswith("ns") else (v if unit.startswith("us") else v * 1e3)
        seq.append((short(r[i["Kernel Name"]]), us, r[i["Grid Size"]] if "Grid Size" in i else ""))
    agg = collections.OrderedDict()
    for n, us, _ in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    # the last frame = everything after the last k_skin but one
    skins = [k for k, (n, _, _) in enumerate(seq) if n.startswith("k_skin")]
    frame = seq[skins[-2] + 1: skins[-1] + 1] if len(skins) >= 2 else seq
    # prefer the last frame that ran the default FYX_UPDATE_ALL kernels (the run ends with incremental-update frames)
    for k in range(len(skins) - 1, 0, -1):
        cand = seq[skins[k - 1] + 1: skins[k] + 1]
        if any(", 20>" in n for n, _, _ in cand):
            frame = cand
            break
    out = os.path.join(OUT, f"{tag}_launches_{workload}.md")
    with open(out, "w") as f:
        f.write(f"# ncu launch list — workload `{workload}` ({tag})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py` (cold-cache, serialised: compare shares, not absolutes).\n\n")
        f.write("## One frame (last frame of the run, launch order)\n\n| # | kernel | grid | µs | share |\n|---|---|---|---|---|\n")
        tot = sum(us for _, us, _ in frame)
        for k, (n, us, g) in enumerate(frame):
            f.write(f"| {k} | `{n}` | {g} | {us:.1f} | {100 * us / tot:.1f} % |\n")
        f.write(f"| | **frame total** | | **{tot:.1f}** | |\n\n")
        f.write("## All launches of the process (load + warm-up + timed), by kernel\n\n| kernel | launches | total ms | mean µs |\n|---|---|---|---|\n")
        for n, (c, t) in agg.items():
            f.write(f"| `{n}` | {c} | {t / 1e3:.3f} | {t / c:.2f} |\n")
    print("wrote", out)


WANT = [
    ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory pipe %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank-conflict wavefronts"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]


def full(tag, workload, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: k for k, h in enumerate(hdr)}
    out = os.path.join(OUT, f"{tag}_ncu_full_{workload}.md")
    traffic = {}
    with open(out, "w") as f:
        f.write(f"# ncu --set full — workload `{workload}` ({tag})\n\n`ncu --set full --clock-control none --import-source on`; one row per profiled launch.\n\n")
        for r in rows[2:]:
            name = short(r[idx["Kernel Name"]])
            f.write(f"## `{name}` (grid {r[idx['launch__grid_size']]})\n\n| metric | value |\n|---|---|\n")
            for m, label in WANT:
                if m in idx:
                    f.write(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |\n")
            stalls = []
            for h, k in idx.items():
                if "issue_stalled" in h and "average" in h and "not_issued" not in h:
                    try:
                        v = float(r[k])
                    except ValueError:
                        continue
                    if v > 0.1:
                        stalls.append((v, h.split("issue_stalled_")[1].split("_per_")[0]))
            f.write("| stall cycles per issued instruction | " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)) + " |\n\n")

            def gb(metric):
                v = float(r[idx[metric]])
                u = units[idx[metric]]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)

            t = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
            prev = traffic.get(name)
            if prev is None or t > prev["dram_bytes_per_launch"]:
                traffic[name] = {"dram_bytes_per_launch": t, "grid": r[idx["launch__grid_size"]]}
    print("wrote", out)
    return traffic


if __name__ == "__main__":
    tag, workload, lpath, rep = sys.argv[1:5]
    os.makedirs(OUT, exist_ok=True)
    launches(tag, workload, lpath)
    tr = full(tag, workload, rep)
    tpath = os.path.join(OUT, "ncu_traffic.json")
    allt = json.load(open(tpath)) if os.path.exists(tpath) else {}
    ent = {}
    for name, v in tr.items():
        key = "k_skin" if name.startswith("k_skin") else ("k_update_level+cull" if name.startswith("k_update_level") else name)
        ent[key] = v
    allt[workload] = ent
    json.dump(allt, open(tpath, "w"), indent=1, sort_keys=True)
    print("wrote", tpath)
