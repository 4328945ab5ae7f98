"""Turns ncu outputs brought back in gpurun_out/ into the small text/CSV summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python profiles/summarize.py full gpurun_out/prof_tc_r1b.ncu-rep profiles/r1_gmm_tc_full.md
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__sass_inst_executed_op_tmem_ldt.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = defaultdict(list)
    for r in rows[rows.index(hdr) + 1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        d[name].append(v)
    tot = sum(sum(v) for v in d.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}); per-launch device time, cold cache + serialised -> compare SHARES\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {k} | {len(v)} | {sum(v):.3f} | {sum(v) / tot:.3f} |\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\n## {name}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}\n\n| metric | unit | value |\n|---|---|---:|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {units[i]} | {r[i]} |\n")
            st = sorted(((int(float(r[i])), hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
                         if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued")), reverse=True)[:6]
            f.write("\nTop warp-stall reasons (pc samples): " + ", ".join(f"{n} {c}" for c, n in st) + "\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
