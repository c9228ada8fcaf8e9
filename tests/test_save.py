"""device state <-> verilated_save images.

The reference side is the genuine --savable model of rv32_soc.v and the genuine
include/verilated_save.cpp (oracle/_ref/libvlref.so): `os << *top` / `os >> *top`."""
import ctypes

import numpy as np
import pytest

import oracle_util as O

SCALARS = ["rv32_soc__DOT__pc", "rv32_soc__DOT__insn", "rv32_soc__DOT__ld_addr",
           "rv32_soc__DOT__instret", "rv32_soc__DOT__state", "trap", "retired", "clk"]


def test_context_image_layout_matches_the_reference_runtime():
    """The library restates VerilatedContext::Serialized (verilated.h:402-441); its image of a
    freshly constructed context must be the genuine runtime's, byte for byte."""
    from verilator_b200 import abi
    want = O.context_image()
    buf = np.zeros(512, dtype=np.uint8)
    n = abi.lib.vlcu_save_context_image(buf.ctypes.data, buf.size)
    assert n == len(want) == 160
    assert (buf[:n] == want).all()


def test_reference_save_restore_round_trip_is_transparent(tmp_path):
    """Sanity of the oracle itself: run 300 evals == run 120, save, restore into a new model, run 180."""
    path = str(tmp_path / "cpu.vltsave")
    regs, scal, mem, t = O.rv32_saverun(7, 300, seed=3, tb_seed=11)
    O.rv32_saverun(7, 120, seed=3, tb_seed=11, save_path=path)
    regs2, scal2, mem2, t2 = O.rv32_saverun(7, 180, seed=3, tb_seed=11, restore_path=path)
    assert t == t2 == 300 and (regs == regs2).all() and (scal == scal2).all() and (mem == mem2).all()
    raw = open(path, "rb").read()
    assert raw.startswith(b"verilatorsave02\n") and raw.endswith(b"vltsaved")


@pytest.fixture(scope="module")
def abi():
    from verilator_b200 import abi
    assert abi.lib.vlcu_device_count() > 0, "no CUDA device: GPU tests need a B200"
    return abi


def state_of(m, inst):
    regs = m.get_member("rv32_soc__DOT__regs")[inst]
    scal = np.array([int(m.get_member(name)[inst]) for name in SCALARS], dtype=np.uint32)
    mem = m.get_member("rv32_soc__DOT__mem")[inst]
    return regs, scal, mem


@pytest.mark.gpu
def test_device_instance_replays_in_the_cpp_model(abi, tmp_path):
    """GPU -> CPU: the image of a device instance is byte-identical to what the genuine model
    saves at the same point, restores into the genuine model, and the replay ends where the
    device instance ends."""
    n, k1, k2 = 70, 220, 333
    with abi.Model("rv32i", n, inst_base=1000, seed=5, rand_reset=2) as m:
        m.run("rv32_rand", k1, flags=abi.RUNF_TB_BEGIN, tb_seed=21)
        for inst in (0, 33, 69):
            ours, ref = str(tmp_path / f"gpu{inst}.vltsave"), str(tmp_path / f"cpu{inst}.vltsave")
            m.save_instance(inst, ours)
            O.rv32_saverun(1000 + inst, k1, seed=5, tb_seed=21, save_path=ref)
            assert open(ours, "rb").read() == open(ref, "rb").read(), inst
        m.run("rv32_rand", k2, tb_seed=21)
        for inst in (0, 33, 69):
            regs, scal, mem, t = O.rv32_saverun(1000 + inst, k2, seed=5, tb_seed=21,
                                                restore_path=str(tmp_path / f"gpu{inst}.vltsave"))
            g_regs, g_scal, g_mem = state_of(m, inst)
            assert t == k1 + k2 == int(m.status()["time"][inst])
            assert (regs == g_regs).all() and (scal == g_scal).all() and (mem == g_mem).all(), inst


@pytest.mark.gpu
def test_cpp_model_state_continues_on_the_device(abi, tmp_path):
    """CPU -> GPU: states saved by the genuine model at different times are loaded into instances
    of one batch (each continues from its own time), and end where uninterrupted CPU runs end."""
    n, k2 = 40, 250
    picks = {3: 0, 17: 90, 39: 201}  # instance -> evals already simulated on the CPU
    with abi.Model("rv32i", n, inst_base=64, seed=8, rand_reset=2) as m:
        m.run("rv32_rand", 0, flags=abi.RUNF_TB_BEGIN, tb_seed=2)
        before = {i: state_of(m, i) for i in (4, 18)}
        for inst, k1 in picks.items():
            path = str(tmp_path / f"cpu{inst}.vltsave")
            O.rv32_saverun(64 + inst, k1, seed=8, tb_seed=2, save_path=path)
            m.restore_instance(inst, path)
        for i in (4, 18):  # neighbours untouched
            assert all((a == b).all() for a, b in zip(before[i], state_of(m, i)))
        m.run("rv32_rand", k2, tb_seed=2)
        for inst, k1 in picks.items():
            regs, scal, mem, t = O.rv32_saverun(64 + inst, k1 + k2, seed=8, tb_seed=2)
            g_regs, g_scal, g_mem = state_of(m, inst)
            assert int(m.status()["time"][inst]) == t == k1 + k2
            assert (regs == g_regs).all() and (scal == g_scal).all() and (mem == g_mem).all(), inst
        # an image of another instance (different seed) or of another model is refused
        with pytest.raises(RuntimeError, match="randSeed"):
            m.restore_instance(5, str(tmp_path / "cpu3.vltsave"))
    bad = str(tmp_path / "bad.vltsave")
    open(bad, "wb").write(b"verilatorsave02\n" + b"\0" * 64 + b"vltsaved")
    with abi.Model("rv32i", 4) as m, pytest.raises(RuntimeError, match="different model|deserialize"):
        m.restore_instance(0, bad)
    with abi.Model("rv32i_cov", 4) as m, pytest.raises(RuntimeError, match="coverage"):
        m.save_instance(0, bad)
