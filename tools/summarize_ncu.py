"""Turn the scratch ncu outputs of a gpurun session into the committed summaries under profiles/.

  python tools/summarize_ncu.py <tag>     e.g. r01b
reads  gpurun_out/launches.csv            (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/prof_<tag>.ncu-rep      (ncu --set full capture)
writes profiles/<tag>_launch_list_summary.json, profiles/<tag>_launches.csv,
       profiles/<tag>_ncu_full_summary.md, profiles/ncu_summary.json (read by bench.py for `traffic`)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SCR = os.path.join(ROOT, "gpurun_out")


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")


def launch_list(tag, command):
    path = os.path.join(SCR, "launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        us = v / 1000.0 if r[iu].startswith("n") else (v if r[iu].startswith("u") else v * 1000.0)
        k = short(r[ik])
        tot[k] = tot.get(k, 0.0) + us
        cnt[k] += 1
    total = sum(tot.values())
    out = [{"kernel": k, "launches": cnt[k], "total_us": round(v, 1), "share_pct": round(100 * v / total, 2)}
           for k, v in sorted(tot.items(), key=lambda kv: -kv[1])]
    json.dump({"command": command, "note": "per-launch times are cold-cache and serialised under ncu: compare shares",
               "kernels": out}, open(os.path.join(OUT, tag + "_launch_list_summary.json"), "w"), indent=1)
    shutil.copy(path, os.path.join(OUT, tag + "_launches.csv"))
    for o in out:
        print("%-42s launches=%5d total_us=%10.1f share=%5.2f%%" % (o["kernel"], o["launches"], o["total_us"],
                                                                    o["share_pct"]))


def tobytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def full(tag, command):
    rep = os.path.join(SCR, "prof_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    M = {"dur": "gpu__time_duration.sum", "grid": "launch__grid_size", "regs": "launch__registers_per_thread",
         "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
         "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
         "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
         "fp64": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
         "warps": "sm__warps_active.avg.pct_of_peak_sustained_active"}
    lines = ["# ncu --set full summary (%s)" % tag, "", "Command: `%s`" % command,
             "Report: gpurun_out/prof_%s.ncu-rep (scratch).  Values are per launch." % tag, "",
             "| kernel | grid | regs | duration | DRAM read | DRAM write | DRAM % | tensor (DMMA) pipe % | fp64 pipe % "
             "| warps active % |", "|---|---|---|---|---|---|---|---|---|---|"]
    summary, seen = {}, collections.Counter()
    for r in rows[2:]:
        name = short(r[idx["Kernel Name"]])
        grid = int(r[idx[M["grid"]]])
        seen[name] += 1
        if seen[name] <= 3:
            def g(m):
                return r[idx[M[m]]] + " " + units[idx[M[m]]]
            lines.append("| %s | %d | %s | %s | %s | %s | %.1f | %.1f | %.1f | %.1f |" % (
                name, grid, r[idx[M["regs"]]], g("dur"), g("rd"), g("wr"), float(r[idx[M["dram"]]]),
                float(r[idx[M["tensor"]]]), float(r[idx[M["fp64"]]]), float(r[idx[M["warps"]]])))
        key = name.split("<")[0]
        traffic = tobytes(r[idx[M["rd"]]], units[idx[M["rd"]]]) + tobytes(r[idx[M["wr"]]], units[idx[M["wr"]]])
        entry = {"grid": grid, "dram_bytes_per_launch": traffic, "duration_us": float(r[idx[M["dur"]]]),
                 "tensor_pipe_pct": float(r[idx[M["tensor"]]]), "fp64_pipe_pct": float(r[idx[M["fp64"]]])}
        if key not in summary or grid > summary[key]["grid"]:
            summary[key] = entry
    open(os.path.join(OUT, tag + "_ncu_full_summary.md"), "w").write("\n".join(lines) + "\n")
    json.dump(summary, open(os.path.join(OUT, "ncu_summary.json"), "w"), indent=1)
    print("\n".join(lines[5:]))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launch_list(tag, "ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 python tools/ncu_target.py 8192 2048")
    full(tag, "ncu --set full --clock-control none --import-source on -k regex:... python tools/ncu_target.py 8192 2048")
