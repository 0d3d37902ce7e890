#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one row per captured launch with the metrics the
roofline discussion needs (duration, DRAM bytes, DRAM/tensor/issue utilisation, occupancy, registers)."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("dram__bytes_read.sum", "rd_MB", None),
    ("dram__bytes_write.sum", "wr_MB", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%", 1),
    ("sm__inst_executed_pipe_tensor.sum", "tc_inst", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf", 1),
    ("lts__t_sector_hit_rate.pct", "l2hit%", 1),
]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * m.get(unit, 1) / 1e6


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [k for k, _, _ in KEYS if k in col]
    print("%-44s %-16s " % ("kernel", "grid") + " ".join("%9s" % lab for k, lab, _ in KEYS if k in col))
    for r in data:
        kn = r[col["Kernel Name"]].split("(")[0].split("::")[-1][:44]
        out = []
        for k, lab, scale in KEYS:
            if k not in col:
                continue
            try:
                v = float(r[col[k]].replace(",", ""))
            except ValueError:
                out.append("%9s" % "-"); continue
            u = units[col[k]]
            if scale is None:
                v = to_bytes(v, u)
            elif lab == "dur_us":
                v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
            out.append("%9.2f" % v)
        print("%-44s %-16s " % (kn, r[col["Grid Size"]].replace(" ", "")) + " ".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
