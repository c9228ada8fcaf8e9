"""debug: dump every member of a veer_class_cov variant after k evals -> gpurun_out/members<v>_<k>.npz"""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from verilator_b200 import abi
VC = dict(tb_flags=abi.FUZZ_RESET_LOW, tb_arg=(4, 0b11100000, 0, 0))
v = sys.argv[1]
abi.load_design_library(f"/root/repo/tests/veercov/libvlcu_veercov{v}.so")
n = 40
for k in [int(a) for a in sys.argv[2:]]:
    with abi.Model("veer_class_cov" + v, n, inst_base=9, seed=3, rand_reset=2) as m:
        m.run("fuzz", k, tb_seed=12, **VC)
        out = {name.replace("/", "_"): m.get_member(name) for name in m.members if "__Vtogcov" not in name and "__Vcoverage" not in name}
        np.savez_compressed(f"gpurun_out/members{v}_{k}.npz", **out)
        print(v, k, len(out))
