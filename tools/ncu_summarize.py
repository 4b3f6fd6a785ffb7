"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summarize.py launches gpurun_out/launches.csv profiles/out.csv "<command line>"
  python tools/ncu_summarize.py full gpurun_out/prof.ncu-rep profiles/out.txt
  python tools/ncu_summarize.py traffic gpurun_out/prof.ncu-rep profiles/ncu_traffic.json [kernel-substring] [algorithmic bytes]
      -> the json bench.py reads for roofline.traffic: DRAM bytes of the longest captured launch of the dominant kernel
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "sm__cycles_active.avg",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def launches(src, dst, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        d = tot.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += v
    allt = sum(v[1] for v in tot.values())
    out = [f"# ncu launch list summary: {cmd} (all launches of the process incl. warm-up and weight calibration; cold-cache, serialised: compare SHARES)",
           f"# total {allt / 1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches", "kernel,launches,total_us,share"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k},{v[0]},{v[1]:.1f},{v[1] / allt:.4f}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:25]))


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in KEYS or h == "Kernel Name"]
    out = [f"# ncu --set full --clock-control none summary of {src} (per launch; durations are cold-cache replays)"]
    for r in rows[2:]:
        out.append("---")
        for i in idx:
            out.append(f"{hdr[i]} = {r[i]} {units[i]}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def traffic(src, dst, kernel="gemm_tcgen05_2cta_kernel", algorithmic=None):
    import json
    import os

    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def num(r, key):
        v = float(r[col[key]].replace(",", ""))
        u = units[col[key]]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3,
                 "msecond": 1e3}.get(u, 1.0)
        return v * scale

    best = None
    for r in rows[2:]:
        if kernel not in r[col["Kernel Name"]]:
            continue
        dur = num(r, "gpu__time_duration.sum")
        if best is None or dur > best[0]:
            best = (dur, r)
    if best is None:
        raise SystemExit(f"no launch of {kernel} in {src}")
    dur, r = best
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = {"kernel": re.sub(r"\(.*", "", r[col["Kernel Name"]])[:120], "bytes_per_launch": int(num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")),
           "dram_read_bytes": int(num(r, "dram__bytes_read.sum")), "dram_write_bytes": int(num(r, "dram__bytes_write.sum")),
           "duration_us_under_ncu": round(dur, 1), "grid": r[col["launch__grid_size"]] if "launch__grid_size" in col else None,
           "tensor_pipe_active_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
           if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else None,
           "algorithmic_bytes": int(algorithmic) if algorithmic else None,
           "source": os.path.basename(src), "commit": commit,
           "note": "longest captured launch of the kernel, ncu --set full --clock-control none (cold-cache replay; a number under ncu is never a bench value)"}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], *(sys.argv[4:6]))
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
