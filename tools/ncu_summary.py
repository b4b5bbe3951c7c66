"""Key metrics of every kernel in an ncu report (development tool): python tools/ncu_summary.py report.ncu-rep [out.csv]"""
import csv
import subprocess
import sys

WANT = [
    ("time_ms", "gpu__time_duration.sum", 1), ("inst", "smsp__inst_executed.sum", 1), ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("regs", "launch__registers_per_thread", 1),
    ("lsu_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
    ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1),
    ("smem_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1),
    ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1), ("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    ("pipe_fmaheavy_pct", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", 1),
    ("pipe_xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1), ("pipe_lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
    ("pipe_tensor_pct", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", 1),
    ("dram_read", "dram__bytes_read.sum", 1), ("dram_write", "dram__bytes_write.sum", 1),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct", 1),
] + [("stall_" + k, f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio", 1) for k in
     ["wait", "short_scoreboard", "long_scoreboard", "barrier", "math_pipe_throttle", "not_selected", "branch_resolving", "mio_throttle",
      "dispatch_stall", "no_instruction", "lg_throttle", "membar", "sleeping", "tex_throttle", "drain", "imc_miss", "selected"]]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[h.index("Kernel Name")][:70]}
        for name, metric, scale in WANT:
            if metric in h:
                v = r[h.index(metric)].replace(",", "")
                try:
                    d[name] = round(float(v) * scale, 4)
                except ValueError:
                    d[name] = v
        out.append(d)
    keys = ["kernel"] + [n for n, _, _ in WANT]
    w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
    w.writerow(["metric"] + [d["kernel"] for d in out])
    for k in keys[1:]:
        w.writerow([k] + [d.get(k, "") for d in out])


if __name__ == "__main__":
    main()
