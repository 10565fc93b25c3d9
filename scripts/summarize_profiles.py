"""Turn the ncu artefacts in gpurun_out/ into the small text summaries kept under profiles/.

    python scripts/summarize_profiles.py r1      # writes profiles/r1_*.txt / .json
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402  (the hash bench.py checks before quoting these numbers)
os.makedirs(PROF, exist_ok=True)

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
]


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def unit_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def summarize(rep, name):
    if not os.path.exists(rep):
        return None
    hdr, units, rows = ncu_raw(rep)
    lines = []
    traffic = []
    for r in rows:
        d = dict(zip(hdr, r))
        lines.append(f"--- {d.get('Kernel Name', '?')[:90]}")
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                lines.append(f"  {k} = {d[k]} {units[hdr.index(k)]}")
        try:
            traffic.append(unit_bytes(d["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")]) +
                           unit_bytes(d["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")]))
        except Exception:
            pass
    path = os.path.join(PROF, f"{tag}_{name}_ncu_full.txt")
    open(path, "w").write(f"# ncu --set full --clock-control none, {os.path.basename(rep)}\n" + "\n".join(lines) + "\n")
    return traffic


def launch_list():
    p = os.path.join(OUT, "launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, f"{tag}_launch_list.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 1\n")
        f.write("# (cold-cache, serialised: compare SHARES, not absolutes; first 600 launches)\n")
        f.write(f"{'kernel':44s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:44]:44s} {v[0]:8d} {v[1] / 1e3:10.3f} {v[1] / v[0]:10.2f} {v[1] / tot:7.3f}\n")


launch_list()
t = summarize(os.path.join(OUT, "prof_gemv.ncu-rep"), "gemv")
if t:
    json.dump({"bytes_per_launch": sum(t) / len(t), "launches_captured": len(t),
               "kernel_source_sha1_16": kernel_source_hash(),
               "source": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full"},
              open(os.path.join(PROF, "gemv_dram_bytes_per_launch.json"), "w"))
t = summarize(os.path.join(OUT, "prof_assemble.ncu-rep"), "assemble")
if t:
    # the capture holds the launch(es) of ONE assembly (the stream kernel: one): the step's traffic is their sum
    json.dump({"bytes_per_step": sum(t), "launches_captured": len(t),
               "kernel_source_sha1_16": kernel_source_hash(),
               "source": "sum over the k_assemble_rows launches of one assembly of dram__bytes_read.sum + "
                         "dram__bytes_write.sum, ncu --set full"},
              open(os.path.join(PROF, "assemble_dram_bytes_per_step.json"), "w"))
for f in ("bench.json", "gpu_info.csv", "assemble_dram_vs_group.txt", "bench_cfg3_g1.json", "bench_cfg5_20k.json", "bench_cfg5_20k_batched.json", "bench_cfg5_4k.json", "bench_reference.json"):
    src = os.path.join(OUT, f)
    if os.path.exists(src):
        open(os.path.join(PROF, f"{tag}_{f}"), "w").write(open(src).read())
subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_excerpts.py"), tag])
print("profiles written to", PROF)
