"""Summarise an `ncu --page raw --csv` export: one line per profiled launch with the metrics the roofline needs."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "lim_reg"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "st_long"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "st_bar"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "st_short"),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "st_lg"),
    ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "st_mio"),
    ("smsp__warp_issue_stalled_membar_per_warp_active.pct", "st_membar"),
    ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "st_wait"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "st_math"),
    ("smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "st_br"),
    ("smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "st_noinst"),
    ("smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "st_disp"),
    ("smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "st_sleep"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx.get("Kernel Name")
    for r in rows[2:]:
        out = [r[name_i][:48]]
        for k, short in KEYS:
            if k in idx:
                v = r[idx[k]]
                u = units[idx[k]]
                try:
                    f = float(v.replace(",", ""))
                    if short in ("dram_rd", "dram_wr", "l2_bytes"):
                        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                        v = f"{f * mult / 1e6:.1f}MB"
                    elif short == "dur":
                        mult = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                        v = f"{f * mult:.1f}us"
                    else:
                        v = f"{f:.1f}" if f != int(f) else str(int(f))
                except ValueError:
                    pass
                out.append(f"{short}={v}")
        print("  ".join(out))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("#", p)
        main(p)
