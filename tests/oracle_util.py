"""Bindings to the ORACLE (oracle/_ref/libvlref.so) -- TEST INFRASTRUCTURE.

libvlref.so = the reference's own runtime (include/verilated*.cpp compiled from
/root/reference) + the GENUINE Verilated C++ models the reference's own compiler
(oracle/_ref/verilator_lite) writes for each design + the harnesses under oracle/models/.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libvlref.so")

u64, u32, i32, i64, vp, cp = (ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_int64,
                              ctypes.c_void_p, ctypes.c_char_p)

STATUS_DTYPE = np.dtype([
    ("time", "<u8"), ("first_finish_time", "<u8"), ("first_stop_time", "<u8"),
    ("first_fatal_time", "<u8"), ("status", "<u4"), ("error_count", "<u4"),
    ("finish_count", "<u4"), ("stop_count", "<u4"), ("fatal_count", "<u4"),
    ("first_stop_line", "<u4"), ("first_fatal_line", "<u4"), ("fatal_code", "<u4"),
    ("msg_count", "<u4"), ("pad_", "<u4")])

_lib = None


def ref():
    global _lib
    if _lib is None:
        if not os.path.exists(REF_PATH):
            raise RuntimeError(f"{REF_PATH} missing: run `make -C oracle` where /root/reference exists")
        _lib = ctypes.CDLL(REF_PATH)
        _lib.vlref_murmur64.restype = u64
        _lib.vlref_murmur64.argtypes = [cp]
        _lib.vlref_rand_reset_i.restype = u32
        _lib.vlref_rand_reset_i.argtypes = [i32, u64, u64, u64, i32]
        _lib.vlref_rand_reset_q.restype = u64
        _lib.vlref_rand_reset_q.argtypes = [i32, u64, u64, u64, i32]
        _lib.vlref_rand_reset_w.restype = None
        _lib.vlref_rand_reset_w.argtypes = [i32, vp, u64, u64, u64, i32]
        _lib.vlref_rng_rand64.restype = None
        _lib.vlref_rng_rand64.argtypes = [u64, vp, i32]
        _lib.vlref_mtc_run.restype = i32
        _lib.vlref_mtc_run.argtypes = [u64, u64, u64, i32, u64, u32, u64, u32, u32, i32, vp, u64, vp,
                                       vp, vp, vp, cp, i64, cp, i64, vp, u64, i32]
    return _lib


def murmur64(s):
    return ref().vlref_murmur64(s.encode())


def rand_reset_w(obits, scope_hash, salt, seed, rand_reset=2):
    words = (obits + 31) // 32
    out = np.zeros(words, dtype=np.uint32)
    ref().vlref_rand_reset_w(obits, out.ctypes.data, scope_hash, salt, seed, rand_reset)
    return out


def rng_rand64(seed, n):
    out = np.zeros(n, dtype=np.uint64)
    ref().vlref_rng_rand64(seed, out.ctypes.data, n)
    return out


class MtcResult:
    pass


def mtc_run(n, n_evals, inst_base=0, seed=1, rand_reset=2, stop_on_finish=False, tb_seed=0,
            per_instance=False, reset_period=0, trace_plusarg=False, obs_evals=0, vcd_path=None,
            vcd_instance=-1, cov_path=None, cov_instance=-1, threads=1):
    """Run the genuine Verilated model of examples/make_tracing_c for n instances."""
    r = MtcResult()
    r.obs = np.zeros((max(obs_evals, 1), 6, n), dtype=np.uint32)
    r.sigs = np.zeros(n, dtype=np.uint64)
    r.status = np.zeros(n, dtype=STATUS_DTYPE)
    r.final = np.zeros((n, 4), dtype=np.uint32)
    r.cov = np.zeros(ref().vlref_mtc_cov_bins(), dtype=np.uint64)
    con = ctypes.create_string_buffer(1 << 16)
    rc = ref().vlref_mtc_run(
        n, inst_base, seed, rand_reset, n_evals, 2 if stop_on_finish else 0, tb_seed,
        1 if per_instance else 0, reset_period, 1 if trace_plusarg else 0,
        r.obs.ctypes.data if obs_evals else None, obs_evals, r.sigs.ctypes.data,
        r.status.ctypes.data, r.final.ctypes.data, r.cov.ctypes.data,
        vcd_path.encode() if vcd_path else None, vcd_instance,
        cov_path.encode() if cov_path else None, cov_instance, con, len(con), threads)
    assert rc == 0
    r.console = con.value.decode()
    if not obs_evals:
        r.obs = None
    return r


# ---- rv32_soc (oracle/models/rv32i) ------------------------------------------------------
def _rv32():
    lib = ref()
    if not getattr(lib, "_rv32_ready", False):
        lib.vlref_rv32_run.restype = i32
        sig = [u64, u64, u64, i32, u64, u64, u32, vp, vp, u64, vp, vp, vp, vp, vp, i32, u32, cp, i64, vp, cp, i64, u32]
        lib.vlref_rv32_run.argtypes = sig
        lib.vlref_rv32cov_run.restype = i32
        lib.vlref_rv32cov_run.argtypes = sig
        lib.vlref_rv32thr_run.restype = i32
        lib.vlref_rv32thr_run.argtypes = sig
        lib.vlref_rv32cov_bins.restype = i32
        lib.vlref_rv32_gen_program.restype = None
        lib.vlref_rv32_gen_program.argtypes = [u64, vp]
        lib.vlref_rv32_saverun.restype = i32
        lib.vlref_rv32_saverun.argtypes = [u64, u64, i32, u64, u32, cp, u64, cp, vp, vp, vp, vp]
        lib.vlref_context_image.restype = i32
        lib.vlref_context_image.argtypes = [vp, ctypes.c_size_t]
        lib._rv32_ready = True
    return lib


def rv32_gen_program(seed):
    out = np.zeros(1024, dtype=np.uint32)
    _rv32().vlref_rv32_gen_program(seed, out.ctypes.data)
    return out


class Rv32Result:
    pass


def rv32_run(n, n_evals, inst_base=0, seed=1, rand_reset=2, tb_seed=0, reset_time=0, mem_images=None,
             obs_evals=0, want_mem=False, threads=1, rounds=0, vcd_path=None, vcd_instance=-1, cov=False,
             cov_path=None, cov_instance=-1, same_program=False, mtasks=False):
    """rounds=0: one run, resetn released when time() > reset_time (default 16).
    rounds=k>0: campaign mode, k runs of n_evals evals on the same models, each with a fresh
    program from (tb_seed + round) ^ index and resetn low for the run's first reset_time evals
    (the device testbench flag RV32_TBF_RESET_PER_RUN)."""
    r = Rv32Result()
    r.obs = np.zeros((max(obs_evals, 1), 8, n), dtype=np.uint32)
    r.sigs = np.zeros(n, dtype=np.uint64)
    r.times = np.zeros(n, dtype=np.uint64)
    r.regs = np.zeros((n, 32), dtype=np.uint32)
    r.mem = np.zeros((n, 1024), dtype=np.uint32) if want_mem else None
    r.scalars = np.zeros((n, 8), dtype=np.uint32)  # pc insn ld_addr instret state trap retired clk
    if mem_images is not None:
        mem_images = np.ascontiguousarray(mem_images, dtype=np.uint32).reshape(n, 1024)
    lib = _rv32()
    r.cov = np.zeros(lib.vlref_rv32cov_bins(), dtype=np.uint64) if cov else None
    rc = (lib.vlref_rv32cov_run if cov else lib.vlref_rv32thr_run if mtasks else lib.vlref_rv32_run)(
        n, inst_base, seed, rand_reset, n_evals, tb_seed, reset_time,
        mem_images.ctypes.data if mem_images is not None else None,
        r.obs.ctypes.data if obs_evals else None, obs_evals, r.sigs.ctypes.data, r.times.ctypes.data,
        r.regs.ctypes.data, r.mem.ctypes.data if want_mem else None, r.scalars.ctypes.data, threads,
        rounds, vcd_path.encode() if vcd_path else None, vcd_instance,
        r.cov.ctypes.data if cov else None, cov_path.encode() if cov_path else None, cov_instance,
        1 if same_program else 0)
    assert rc == 0
    if not obs_evals:
        r.obs = None
    return r


def rv32_saverun(index, n_evals, seed=1, rand_reset=2, tb_seed=0, reset_time=0, restore_path=None,
                 save_path=None):
    """One genuine (--savable) model: constructed like global instance `index` or restored from a
    verilated_save image, run for n_evals evals, optionally saved.  Returns regs, scalars, mem, time."""
    regs = np.zeros(32, dtype=np.uint32)
    scalars = np.zeros(8, dtype=np.uint32)
    mem = np.zeros(1024, dtype=np.uint32)
    t = np.zeros(1, dtype=np.uint64)
    rc = _rv32().vlref_rv32_saverun(index, seed, rand_reset, tb_seed, reset_time,
                                    restore_path.encode() if restore_path else None, n_evals,
                                    save_path.encode() if save_path else None, regs.ctypes.data,
                                    scalars.ctypes.data, mem.ctypes.data, t.ctypes.data)
    assert rc == 0
    return regs, scalars, mem, int(t[0])


def context_image():
    buf = np.zeros(512, dtype=np.uint8)
    n = _rv32().vlref_context_image(buf.ctypes.data, buf.size)
    assert n > 0
    return buf[:n].copy()


# ---- wide_bus (oracle/models/wide_bus) ---------------------------------------------------
class WideResult:
    pass


def wide_run(n, n_evals, inst_base=0, seed=1, rand_reset=2, tb_seed=0, reset_time=0, obs_evals=0,
             threads=1, vcd_path=None, vcd_instance=-1):
    """Genuine Verilated wide_bus model, stimulus of the device testbench TbWideRand."""
    lib = ref()
    lib.vlref_wide_run.restype = i32
    lib.vlref_wide_run.argtypes = [u64, u64, u64, i32, u64, u64, u32, vp, u64, vp, vp, cp, i64, i32]
    r = WideResult()
    r.obs = np.zeros((max(obs_evals, 1), 19, n), dtype=np.uint32)
    r.sigs = np.zeros(n, dtype=np.uint64)
    r.acc = np.zeros((n, 8), dtype=np.uint32)
    rc = lib.vlref_wide_run(n, inst_base, seed, rand_reset, n_evals, tb_seed, reset_time,
                            r.obs.ctypes.data if obs_evals else None, obs_evals, r.sigs.ctypes.data,
                            r.acc.ctypes.data, vcd_path.encode() if vcd_path else None, vcd_instance,
                            threads)
    assert rc == 0
    if not obs_evals:
        r.obs = None
    return r


# ---- generic harness (oracle/models/generic): any design, "trace" testbench ------------------
def generic_words(design):
    """(input row words, output row words) of a design built with the generic harness."""
    fn = getattr(ref(), f"vlref_{design}_run")
    fn.restype = i32
    fn.argtypes = [u64, u64, u64, i32, u64, vp, vp, vp, ctypes.POINTER(i32), ctypes.POINTER(i32), i32]
    iw, ow = i32(0), i32(0)
    fn(0, 0, 0, 0, 0, None, None, None, ctypes.byref(iw), ctypes.byref(ow), 1)
    return iw.value, ow.value


def generic_run(design, stim, inst_base=0, seed=1, rand_reset=2, threads=1):
    """stim: uint32 [evals][in_words][n] (exactly n columns).  Returns (obs [evals][out_words][n],
    sigs [n]) from the genuine Verilated model driven like the CUDA runtime's "trace" testbench."""
    iw, ow = generic_words(design)
    stim = np.ascontiguousarray(stim, dtype=np.uint32)
    evals, words, n = stim.shape
    assert words == iw, (words, iw)
    obs = np.zeros((evals, ow, n), dtype=np.uint32)
    sigs = np.zeros(n, dtype=np.uint64)
    fn = getattr(ref(), f"vlref_{design}_run")
    rc = fn(n, inst_base, seed, rand_reset, evals, stim.ctypes.data, obs.ctypes.data, sigs.ctypes.data,
            None, None, threads)
    assert rc == 0
    return obs, sigs


# ---- "fuzz" testbench (include/verilated_cuda_launch.h VlcuTbFuzz) on the CPU side -------------
FUZZ_RESET_LOW, FUZZ_RESET_HIGH, FUZZ_UNIFORM = 1, 2, 4
_M64 = (1 << 64) - 1


def _fuzz_mix(x):
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
    return x ^ (x >> np.uint64(31))


def fuzz_rows(n, n_evals, in_words, inst_base=0, seed=1, tb_seed=0, tb_flags=0, reset_periods=0,
              sparse_mask=0, eval_base=0):
    """The stimulus rows VlcuTbFuzz makes in the kernel, restated with numpy:
    uint32 [evals][in_words][n]."""
    with np.errstate(over="ignore"):
        inst = np.arange(n, dtype=np.uint64) + np.uint64(inst_base)
        rand_seed = np.uint64(seed) ^ inst  # VlcuContext::randSeed()
        key = np.uint64(tb_seed) ^ (np.zeros(n, dtype=np.uint64) if tb_flags & FUZZ_UNIFORM
                                    else _fuzz_mix(rand_seed + np.uint64(1)))
        e = np.arange(n_evals, dtype=np.uint64) + np.uint64(eval_base)
        period = e >> np.uint64(1)
        rows = np.zeros((n_evals, in_words, n), dtype=np.uint32)
        for w in range(in_words):
            draw = _fuzz_mix(key[None, :] + period[:, None] * np.uint64(0x9e3779b97f4a7c15)
                             + np.uint64(((w + 1) * 0xd1342543de82ef95) & _M64))
            val = (draw >> np.uint64(32)).astype(np.uint32)
            if w < 32 and (sparse_mask >> w) & 1:
                val = np.where((draw & np.uint64(0x1f)) != 0, np.uint32(0), val)
            rows[:, w, :] = val
        rows[:, 0, :] = ((e & np.uint64(1)) ^ np.uint64(1)).astype(np.uint32)[:, None]
        if in_words > 1 and tb_flags & FUZZ_RESET_LOW:
            rows[:, 1, :] = (period >= np.uint64(reset_periods)).astype(np.uint32)[:, None]
        if in_words > 1 and tb_flags & FUZZ_RESET_HIGH:
            rows[:, 1, :] = (period < np.uint64(reset_periods)).astype(np.uint32)[:, None]
    return rows


def generic_fuzz(design, n, n_evals, inst_base=0, seed=1, rand_reset=2, tb_seed=0, tb_flags=0,
                 reset_periods=0, sparse_mask=0, obs_evals=0, threads=1):
    """The genuine Verilated model driven by the harness's own restatement of the fuzz testbench
    (no stimulus rows in memory).  Returns (obs [obs_evals][out_words][n] or None, sigs [n])."""
    _, ow = generic_words(design)
    fn = getattr(ref(), f"vlref_{design}_fuzz")
    fn.restype = i32
    fn.argtypes = [u64, u64, u64, i32, u64, u64, u32, u32, u32, vp, u64, vp, i32]
    obs = np.zeros((obs_evals, ow, n), dtype=np.uint32) if obs_evals else None
    sigs = np.zeros(n, dtype=np.uint64)
    rc = fn(n, inst_base, seed, rand_reset, n_evals, tb_seed, tb_flags, reset_periods, sparse_mask,
            obs.ctypes.data if obs_evals else None, obs_evals, sigs.ctypes.data, threads)
    assert rc == 0
    return obs, sigs


def last_loop_seconds():
    fn = ref().vlref_last_loop_seconds
    fn.restype = ctypes.c_double
    return fn()


def generic_cov(design, n, n_evals, cov_path, inst_base=0, seed=1, rand_reset=2, tb_seed=0, tb_flags=0,
                reset_periods=0, sparse_mask=0):
    """--coverage build of a generic-harness design: a fuzz campaign of n instances whose coverage
    the reference's own VerilatedCovContext::write merges into `cov_path` (all models in one
    context; verilated_cov.cpp:368-438 sums the counters of points with equal keys).  Returns sigs."""
    fn = getattr(ref(), f"vlref_{design}_cov")
    fn.restype = i32
    fn.argtypes = [u64, u64, u64, i32, u64, u64, u32, u32, u32, vp, cp]
    sigs = np.zeros(n, dtype=np.uint64)
    rc = fn(n, inst_base, seed, rand_reset, n_evals, tb_seed, tb_flags, reset_periods, sparse_mask,
            sigs.ctypes.data, cov_path.encode())
    assert rc == 0
    return sigs
