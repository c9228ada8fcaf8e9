import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from verilator_b200 import abi
import oracle_util as O
VC = dict(tb_flags=abi.FUZZ_RESET_LOW, tb_arg=(4, 0b11100000, 0, 0))
VC_O = dict(tb_flags=1, reset_periods=4, sparse_mask=0b11100000)
n, evals = 40, 260
want_obs, want = O.generic_fuzz("veer_class_cov", n, evals, inst_base=9, seed=3, rand_reset=2, tb_seed=12, obs_evals=evals, threads=8, **VC_O)
for v in sys.argv[1:]:
    lib = f"/root/repo/tests/veercov/libvlcu_veercov{v}.so"
    if not os.path.exists(lib):
        print(v, "missing"); continue
    abi.load_design_library(lib)
    name = "veer_class_cov" + v
    with abi.Model(name, n, inst_base=9, seed=3, rand_reset=2) as m:
        obs, _ = m.run("fuzz", evals, obs_evals=evals, tb_seed=12, **VC)
        sig_ok = (m.signatures() == want).all()
        bad = np.argwhere(obs != want_obs)
        print(v or "main", "sigs equal:", sig_ok, "first diff (eval, word, inst):", bad[0] if len(bad) else None, "n diffs", len(bad))
