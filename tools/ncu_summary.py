"""Condense ncu artefacts (gpurun_out/, scratch) into small tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1a.csv profiles/r1_launches.csv
    python tools/ncu_summary.py full gpurun_out/prof_r1_tc3n4.ncu-rep profiles/r1_predict_tc3n4_ncu.txt
"""
import csv
import io
import subprocess
import sys

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_bytes.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
    "sm__cycles_elapsed.max.per_second", "smsp__cycles_active.avg", "launch__occupancy_limit_shared_mem",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
    "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
    "smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_selected_per_warp_active.pct",
)


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, bi, gi, vi = hdr.index("Kernel Name"), hdr.index("Block Size"), hdr.index("Grid Size"), hdr.index("Metric Value")
    out = [("id", "kernel", "block", "grid", "gpu__time_duration_ns")]
    tot = {}
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        out.append((r[0], name, r[bi], r[gi], r[vi]))
        tot[name] = tot.get(name, 0.0) + float(r[vi])
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerows(out)
        w.writerow([])
        w.writerow(("# share of summed device time per kernel (cold-cache, serialised: compare shares, not absolutes)",))
        s = sum(tot.values())
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            w.writerow((f"# {k}", f"{v / 1e6:.3f} ms", f"{100 * v / s:.2f} %"))
    print(f"wrote {dst}: {len(out) - 1} launches")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; source report: {src} (scratch, not tracked)\n")
        for r in rows[2:]:
            d = dict(zip(hdr, zip(units, r)))
            f.write(f"\n## {d['Kernel Name'][1]}  grid={d['Grid Size'][1]} block={d['Block Size'][1]}\n")
            for k in hdr:
                base = k.split(".Triage")[0]
                if k in KEEP or base in KEEP:
                    f.write(f"{k:95s} {d[k][1]:>22s} {d[k][0]}\n")
    print(f"wrote {dst}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
