#!/usr/bin/env python3
"""Rebuilds the "kernels" section of profiles/traffic.json from the ncu raw-page CSVs kept under
profiles/ (one vlcu_k_run launch each): bench.py's roofline reads DRAM bytes, warp instructions and
active lanes per launch from there and divides by the launch duration it measures itself."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# workload -> (csv, instances, cycles of the profiled launch, bound)
PROFILES = {
    "rv32i": ("r02_ncu_full_rv32i_vlcu_k_run.csv", 65536, 2048, "issue"),  # (r01_* is used while no r02 capture exists)
    "x_initial": ("r02_ncu_full_x_initial_vlcu_k_run.csv", 16384, 4096, "issue"),
    "make_tracing_c": ("r02_ncu_full_make_tracing_c_vlcu_k_run.csv", 65536, 8192, "issue"),
    "wide_bus": ("r02_ncu_full_wide_bus_vlcu_k_run.csv", 4096, 8192, "issue"),
    "cov_campaign": ("r02_ncu_full_cov_campaign_vlcu_k_run.csv", 16384, 1024, "issue"),
    "veer_class": ("r02_ncu_full_veer_class_vlcu_k_run.csv", 16384, 256, "hbm"),
    "veer_class_64k": ("r02_ncu_full_veer_class_64k_vlcu_k_run.csv", 65536, 128, "hbm"),
    "cov_campaign_veer": ("r02_ncu_full_cov_campaign_veer_vlcu_k_run.csv", 16384, 64, "hbm"),
    "x_initial_il": ("r02_ncu_full_x_initial_il_vlcu_k_run.csv", 16384, 4096, "issue"),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(path))
    kernels = data.setdefault("kernels", {})
    for workload, (name, instances, cycles, bound) in PROFILES.items():
        f = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(f):
            name = name.replace("r02_", "r01_")  # the previous round's capture of the same kernel
            f = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(f):
            continue
        rows = list(csv.reader(open(f)))
        if rows[0][:3] == ["metric", "unit", "value"]:  # compact form (tools/ncu_compact.py)
            col = {r[0]: (r[1], r[2]) for r in rows[1:] if len(r) >= 3}
        else:  # ncu --page raw --csv
            col = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}

        def num(key):
            u, v = col[key]
            return float(v.replace(",", "")) * UNIT.get(u, 1.0)
        kernels[workload] = {
            "instances": instances, "cycles": cycles, "bound": bound,
            "dram_bytes": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
            "warp_instructions": num("smsp__inst_executed.sum"),
            "lanes_active": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "launch_ms_under_ncu": num("gpu__time_duration.sum") / (1e6 if col["gpu__time_duration.sum"][0] == "ns" else
                                                                     1e3 if col["gpu__time_duration.sum"][0] == "us" else 1.0),
            "source": f"profiles/{name} (ncu --set full --clock-control none, one launch of vlcu_k_run, "
                      f"{instances} instances x {cycles} cycles)",
        }
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
