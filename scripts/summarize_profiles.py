"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (run in the build container).
   python scripts/summarize_profiles.py <tag>      e.g. r1a
"""
import collections
import csv
import gzip
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def launches(tag, name="launches.csv"):
    name = name if os.path.exists(os.path.join(OUT, name)) else f"{tag}_launches.csv"
    path = os.path.join(OUT, name)
    if not os.path.exists(path):
        return
    lines = open(path).read().splitlines(True)
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines[start:]):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        k = re.sub(r"\(.*", "", row["Kernel Name"]).strip()
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold-cache): compare SHARES\n")
        f.write(f"# source: gpurun_out/{name} ({sum(v[0] for v in agg.values())} launches, {tot/1e6:.1f} ms of kernel time)\n")
        f.write(f"{'share':>8} {'launches':>9} {'avg_us':>10}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{v[1]/tot*100:7.2f}% {v[0]:9d} {v[1]/v[0]/1e3:10.2f}  {k}\n")
    with open(path, "rb") as fi, gzip.open(os.path.join(PROF, f"{tag}_launches.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)


def full(tag, rep):
    path = os.path.join(OUT, rep + ".ncu-rep")
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(PROF, f"{tag}_{rep}_ncu_full.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source gpurun_out/{rep}.ncu-rep\n")
        for n, row in enumerate(rows[2:]):
            f.write(f"\n## launch {n}: {row[hdr.index('Kernel Name')][:110]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:80s} {row[i]} {units[i]}\n")
            for i, h in enumerate(hdr):
                if ("issue_stalled" in h and h.endswith("_per_warp_active.pct")) or \
                        ("pipe_tensor" in h and "pct_of_peak" in h) or "mem_tensor_cycles_active.avg.pct" in h:
                    if h not in KEYS:
                        f.write(f"{h:80s} {row[i]} {units[i]}\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    if len(sys.argv) > 2:   # on the GPU box: write the text summaries next to the reports (only gpurun_out/ travels back,
        PROF = sys.argv[2]  # and it is capped at 64 MiB: the .ncu-rep files themselves are deleted by the calling script)
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    import glob
    for path in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
        full(tag, os.path.basename(path)[:-len(".ncu-rep")])
    for f in ("bench_n1.json",):
        if os.path.exists(os.path.join(OUT, f)):
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f"{tag}_{f}"))
    print(os.listdir(PROF))
