#!/usr/bin/env python3
"""ncu raw-page CSV (one launch) -> the compact `metric,unit,value` form kept under profiles/.
usage: tools/ncu_compact.py <raw.csv> <out.csv>"""
import csv
import sys

KEEP = ("dram__bytes", "dram__throughput", "gpu__time_duration", "l1tex__t_requests_pipe_lsu_mem", "l1tex__t_sectors_pipe_lsu_mem",
        "l1tex__t_sector_hit_rate", "l1tex__data_bank_conflicts", "lts__t_sector_hit_rate", "lts__t_sectors_srcunit_tex_op",
        "launch__", "sm__warps_active", "sm__throughput", "sm__icc_request", "sm__cycles_elapsed", "smsp__inst_executed",
        "smsp__thread_inst_executed", "smsp__issue_active", "smsp__average_warps_issue_stalled", "smsp__cycles_active",
        "sm__inst_executed_pipe_lsu", "smsp__sass_inst_executed_op", "sm__sass_inst_executed_op_global_red",
        "lts__t_sectors_op_red", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red", "smsp__inst_executed_op")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    head, units, vals = rows[0], rows[1], rows[2]
    out = csv.writer(open(sys.argv[2], "w"), lineterminator="\n")
    out.writerow(["metric", "unit", "value"])
    for h, u, v in zip(head, units, vals):
        if h == "Kernel Name":
            out.writerow([h, "", v])
        elif h.startswith(KEEP):
            try:
                if "issue_stalled" in h and float(v.replace(",", "")) < 0.01:
                    continue
            except ValueError:
                pass
            out.writerow([h, u, v])


if __name__ == "__main__":
    main()
