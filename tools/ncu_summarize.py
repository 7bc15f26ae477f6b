"""Summaries of the ncu artefacts of a round for profiles/ (B200_PROFILING.md: launch list = shares of the step, cold-cache and
serialised; `--set full` capture = the counters of the top kernels).  Usage:
  python tools/ncu_summarize.py launches <csv> > profiles/rNN_ncu_launches_X.md
  python tools/ncu_summarize.py full <a.ncu-rep> [<b.ncu-rep> ...] > profiles/rNN_ncu_full.md"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum"]


def base_name(k):
    k = re.sub(r"^void\s+", "", k)
    k = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", k)
    m = re.match(r"([A-Za-z_][\w:]*)", k)
    return m.group(1) if m else k[:40]


def launches(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in data:
        if len(r) > iv:
            per.setdefault(r[iid], {"name": base_name(r[ik])})[r[im]] = float(r[iv].replace(",", ""))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for v in per.values():
        a = agg[v["name"]]
        a[0] += 1; a[1] += v.get("gpu__time_duration.sum", 0.0); a[2] += v.get("dram__bytes_read.sum", 0.0); a[3] += v.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    out = [f"ncu launch list `{path}`: {len(per)} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised: compare SHARES).", "",
           "| kernel | launches | total us | share | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|"]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / tot:.3f} | {a[2] / 1e6:.1f} | {a[3] / 1e6:.1f} |")
    return "\n".join(out) + "\n"


def full(paths):
    out = []
    for p in paths:
        txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = base_name(vals[hdr.index("Kernel Name")])
            out += [f"### `{name}` ({p})", "", "| metric | value | unit |", "|---|---|---|"]
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    out.append(f"| {m} | {vals[i]} | {units[i]} |")
            out.append("")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    print(launches(sys.argv[2]) if sys.argv[1] == "launches" else full(sys.argv[2:]))
