"""Summarise ncu artefacts into profiles/ (tracked).  Usage:
    python tools/ncu_summary.py launches <launches.csv> <out.md> [title]
    python tools/ncu_summary.py full <report.ncu-rep> <out.md> [title]
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
]


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    cols, data = rows[h], rows[h + 1:]
    ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    agg, tot, seq = collections.OrderedDict(), 0.0, []
    for r in data:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("fs2::<unnamed>::", "fs2::")
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
        seq.append((name, v))
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off` around one step "
                f"(`tools/profile_step.py`).  Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
                f"{len(data)} launches, {tot:.3f} ms total.\n\n| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k[:100]}` | {n} | {t:.3f} | {100 * t / tot:.1f}% |\n")
        f.write("\n## launch sequence (ms)\n\n```\n")
        for i, (n, t) in enumerate(seq):
            f.write(f"{i:3d} {t:8.4f}  {n[:100]}\n")
        f.write("```\n")


def full(path, out, title):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` (`{path.split('/')[-1]}`, kept in gpurun_out/), "
                "read with `ncu -i ... --page raw --csv`.\n\n")
        for v in rows[2:]:
            name = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
            f.write(f"## `{name[:120]}`\n\n| metric | unit | value |\n|---|---|---:|\n")
            for i, n in enumerate(h):
                if n in KEYS:
                    f.write(f"| {n} | {u[i]} | {v[i]} |\n")
            f.write("\n")


if __name__ == "__main__":
    mode, src, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, out, title)
