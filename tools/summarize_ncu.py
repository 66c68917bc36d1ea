"""Summarise ncu outputs into the small text/csv files kept under profiles/ (run here, on the CPU box).

  python summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launch_list_summary.csv "<command>"
  python summarize_ncu.py full gpurun_out/round_prof.ncu-rep profiles/r01_ncu_kernels_summary.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import csv
import io
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(CUtensorMap_st.*", "", name)
    name = re.sub(r"\((int|bool|long|float|const|__nv|unsigned|double|StreamArgs|at::).*", "", name)
    name = name.replace("(int)", "").replace("(bool)", "")
    return name.strip()[:110]


def launches(src, dst, command):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        k = short(r[ki])
        t, c = agg.get(k, (0.0, 0))
        agg[k] = (t + v, c + 1)
    tot = sum(t for t, _ in agg.values())
    n = sum(c for _, c in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list of a bf16 train-step window (%d launches, eager mode, gpu__time_duration.sum, --clock-control none)\n" % n)
        f.write("# command: %s\n# per-launch times are cold-cache and serialised: compare SHARES\n" % command)
        f.write("kernel,launches,total_us,share\n")
        for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
            f.write('"%s",%d,%.1f,%.4f\n' % (k, c, t, t / tot))
    print(open(dst).read()[:3000])


def full(src, dst):
    if src.endswith(".csv"):      # `ncu -i x.ncu-rep --page raw --csv` already run on the GPU box (a large .ncu-rep cannot travel)
        out = open(src).read()
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg.per_second"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    ki = hdr.index("Kernel Name")
    col = {w: i for w, i in idx}
    peak = 6543.1
    try:
        import json
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def num(r, w):
        try:
            return float(r[col[w]].replace(",", ""))
        except Exception:
            return None

    def to_us(r):
        v, u = num(r, "gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]]
        return None if v is None else (v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v))

    def to_bytes(r, w):
        v, u = num(r, w), units[col[w]].lower() if w in col else ""
        if v is None:
            return None
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    # every launch of a kernel (the scripts launch each shape twice: the second one is warm)
    launches = {}
    for r in rows[2:]:
        launches.setdefault(short(r[ki]), []).append(r)
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none (profiler replay, serialised): %s\n" % src)
        f.write("# HBM column: (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum vs the measured %.0f GB/s\n" % peak)
        f.write("# kernels that move a few hundred KB are launch-latency bound (~2-5 us): their GB/s is informational\n\n")
        f.write("%-58s %4s %9s %10s %9s %7s %7s %6s\n" % ("kernel (last launch of each)", "n", "us", "dram MB", "GB/s", "%HBM", "tensor%", "regs"))
        for k, rs in launches.items():
            r = rs[-1]
            us = to_us(r)
            rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
            mb = (rd + wr) / 1e6 if rd is not None and wr is not None else None
            gbs = (rd + wr) / (us * 1e-6) / 1e9 if (mb is not None and us) else None
            tp = num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
            rg = num(r, "launch__registers_per_thread")
            f.write("%-58s %4d %9.1f %10.2f %9.0f %7.1f %7s %6s\n" % (
                k[:58], len(rs), us or 0, mb or 0, gbs or 0, 100.0 * (gbs or 0) / peak, "%.1f" % tp if tp else "-", "%d" % rg if rg else "-"))
        f.write("\n# full metric rows (first launch of each kernel)\n")
        seen = {}
        for r in rows[2:]:
            k = short(r[ki])
            seen[k] = seen.get(k, 0) + 1
            if seen[k] > 1:
                continue
            f.write("\n%s\n" % k)
            for w, i in idx:
                f.write("    %-70s %s %s\n" % (w, r[i], units[i]))
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
