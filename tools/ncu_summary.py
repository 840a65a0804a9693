"""Developer tool: condense `ncu -i X.ncu-rep --page raw --csv` into the few lines kept under profiles/.
    python -m tools.ncu_summary gpurun_out/prof.ncu-rep > profiles/r1_<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    unit_of = dict(zip(hdr, units))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for h in hdr:
            if h in KEYS or h.startswith("launch__occupancy_limit"):
                print(f"  {h:85s} {d[h]:>16s} {unit_of[h]}")
        stalls = [(float(d[h]), h.split("stalled_")[1].split("_per_issue")[0]) for h in hdr
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and d[h]]
        print("  top stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:6]))
        rd, wr = float(d.get("dram__bytes_read.sum", 0) or 0), float(d.get("dram__bytes_write.sum", 0) or 0)
        print(f"  dram traffic (read+write): {rd + wr:.3f} {unit_of.get('dram__bytes_read.sum')}")


if __name__ == "__main__":
    main(sys.argv[1])
