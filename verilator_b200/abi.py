"""ctypes binding of the C ABI in include/verilated_cuda_abi.h.

Python is test/CI glue only (BASELINE.json north_star): the product is
``libverilated_cuda.so`` (CUDA kernels + C++ host runtime).  This module just
loads it and mirrors, for N instances at once, the calls a testbench makes on
the reference's generated model class (examples/make_tracing_c/sim_main.cpp:
construct, poke ports, eval(), read ports, final()).

There is no fallback: if the shared library is missing this import fails.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# VLCU_LIB: experiments only (e.g. a build with another block size); default is the in-tree build
LIB_PATH = os.environ.get("VLCU_LIB") or os.path.join(_HERE, "libverilated_cuda.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(make -C verilator_b200/csrc). There is no CPU/Python fallback for the evaluator.")

lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)

c_u64, c_u32, c_i32, c_vp, c_cp = (ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_char_p)


class PortInfo(ctypes.Structure):
    _fields_ = [("name", c_cp), ("dir", c_i32), ("width", c_i32), ("elem_bytes", c_i32),
                ("elems", c_i32)]


class RunDesc(ctypes.Structure):
    _fields_ = [("tb", c_cp), ("n_evals", c_u64), ("flags", c_u32), ("pad_", c_u32),
                ("stim", c_vp), ("obs", c_vp), ("obs_evals", c_u64),
                ("sample_instances", c_vp), ("n_samples", c_u32), ("trace_stream_words", c_u32),
                ("trace_out", c_vp), ("trace_evals", c_u64),
                ("tb_seed", c_u64), ("tb_flags", c_u32), ("tb_arg", c_u32 * 4), ("pad3_", c_u32),
                ("tb_data", c_vp), ("tb_data_bytes", c_u64)]


STATUS_DTYPE = np.dtype([
    ("time", "<u8"), ("first_finish_time", "<u8"), ("first_stop_time", "<u8"),
    ("first_fatal_time", "<u8"), ("status", "<u4"), ("error_count", "<u4"),
    ("finish_count", "<u4"), ("stop_count", "<u4"), ("fatal_count", "<u4"),
    ("first_stop_line", "<u4"), ("first_fatal_line", "<u4"), ("fatal_code", "<u4"),
    ("msg_count", "<u4"), ("pad_", "<u4")])
MSG_DTYPE = np.dtype([("time", "<u8"), ("id", "<u4"), ("nargs", "<u4"), ("args", "<u8", 8)])

DIR_IN, DIR_OUT, DIR_INOUT = 0, 1, 2
RUNF_NO_STREAM = 0x20
RUNF_BULK_1D = 0x40
RUNF_TRACE_CHANGES = 0x80
RUNF_TBDATA_DEVICE = 0x100
RUNF_TB_BEGIN, RUNF_STOP_ON_FINISH, RUNF_STIM_DEVICE, RUNF_OBS_DEVICE, RUNF_STIM_BROADCAST = (
    1, 2, 4, 8, 16)
FUZZ_RESET_LOW, FUZZ_RESET_HIGH, FUZZ_UNIFORM = 1, 2, 4  # tb_flags of the "fuzz" testbench
ST_DIDINIT, ST_FINISH, ST_ERROR, ST_STOPPED, ST_FATAL, ST_MSGDROP = 1, 2, 4, 8, 16, 32
TRACE_HDR = 16

# Every symbol include/verilated_cuda_abi.h declares: (restype, argtypes)
_H = c_vp
SIGNATURES = {
    "vlcu_abi_version": (c_i32, []),
    "vlcu_last_error": (c_cp, []),
    "vlcu_device_count": (c_i32, []),
    "vlcu_design_count": (c_i32, []),
    "vlcu_design_name": (c_cp, [c_i32]),
    "vlcu_host_alloc": (c_vp, [ctypes.c_size_t]),
    "vlcu_host_free": (None, [c_vp]),
    "vlcu_device_alloc": (c_vp, [_H, ctypes.c_size_t]),
    "vlcu_device_free": (None, [_H, c_vp]),
    "vlcu_device_copy": (c_i32, [_H, c_vp, c_vp, ctypes.c_size_t, c_i32]),
    "vlcu_set_pipeline_chunk": (c_i32, [_H, ctypes.c_size_t]),
    "vlcu_create": (_H, [c_cp, c_u64, c_i32, c_u64, c_u64, c_i32, c_u32, c_cp]),
    "vlcu_destroy": (None, [_H]),
    "vlcu_reset": (c_i32, [_H]),
    "vlcu_instances": (c_u64, [_H]),
    "vlcu_padded_instances": (c_u64, [_H]),
    "vlcu_state_bytes": (c_u32, [_H]),
    "vlcu_in_words": (c_u32, [_H]),
    "vlcu_out_words": (c_u32, [_H]),
    "vlcu_in_bytes": (c_u32, [_H]),
    "vlcu_out_bytes": (c_u32, [_H]),
    "vlcu_trace_record_bytes": (c_u32, [_H]),
    "vlcu_touched_read_bytes": (c_u32, [_H]),
    "vlcu_touched_write_bytes": (c_u32, [_H]),
    "vlcu_port_count": (c_i32, [_H]),
    "vlcu_port_info": (c_i32, [_H, c_i32, ctypes.POINTER(PortInfo)]),
    "vlcu_port_index": (c_i32, [_H, c_cp]),
    "vlcu_set_input": (c_i32, [_H, c_i32, c_vp, ctypes.c_size_t]),
    "vlcu_get_output": (c_i32, [_H, c_i32, c_vp, ctypes.c_size_t]),
    "vlcu_member_count": (c_i32, [_H]),
    "vlcu_member_name": (c_cp, [_H, c_i32]),
    "vlcu_member_index": (c_i32, [_H, c_cp]),
    "vlcu_member_info": (c_i32, [_H, c_i32, ctypes.POINTER(PortInfo)]),
    "vlcu_set_member": (c_i32, [_H, c_i32, c_vp, ctypes.c_size_t]),
    "vlcu_get_member": (c_i32, [_H, c_i32, c_vp, ctypes.c_size_t]),
    "vlcu_stage_member": (c_i32, [_H, c_i32, c_vp, ctypes.c_size_t, c_i32]),
    "vlcu_commit_staged": (c_i32, [_H]),
    "vlcu_member_device_ptr": (c_i32, [_H, c_i32, ctypes.POINTER(c_vp), ctypes.POINTER(c_u64)]),
    "vlcu_member_offset": (c_u32, [_H, c_i32]),
    "vlcu_tile": (c_u32, [_H]),
    "vlcu_time_inc": (c_i32, [_H, c_u64]),
    "vlcu_eval": (c_i32, [_H]),
    "vlcu_run": (c_i32, [_H, ctypes.POINTER(RunDesc)]),
    "vlcu_final": (c_i32, [_H]),
    "vlcu_sync": (c_i32, [_H]),
    "vlcu_flush_l2": (c_i32, [_H]),
    "vlcu_last_kernel_ms": (c_i32, [_H, ctypes.POINTER(ctypes.c_float)]),
    "vlcu_launch_count": (c_u64, [_H]),
    "vlcu_status": (c_i32, [_H, c_vp]),
    "vlcu_signatures": (c_i32, [_H, c_vp]),
    "vlcu_reset_signatures": (c_i32, [_H]),
    "vlcu_results_begin": (c_i32, [_H, c_vp, c_vp, c_i32, ctypes.POINTER(c_i32), ctypes.POINTER(c_vp)]),
    "vlcu_results_wait": (c_i32, [_H]),
    "vlcu_messages": (c_i32, [_H, c_u64, c_vp, c_u32, ctypes.POINTER(c_u32)]),
    "vlcu_message_format": (c_cp, [_H, c_u32]),
    "vlcu_snapshot": (c_i32, [_H, c_u64, c_vp, ctypes.c_size_t]),
    "vlcu_save_instance": (c_i32, [_H, c_u64, c_cp]),
    "vlcu_restore_instance": (c_i32, [_H, c_u64, c_cp]),
    "vlcu_save_context_image": (c_i32, [c_vp, ctypes.c_size_t]),
    "vlcu_coverage_bins": (c_u32, [_H]),
    "vlcu_coverage_points": (c_u32, [_H]),
    "vlcu_coverage_point": (c_i32, [_H, c_u32, ctypes.POINTER(c_u32), ctypes.POINTER(c_cp), ctypes.POINTER(c_i32),
                                    ctypes.POINTER(c_i32), ctypes.POINTER(c_cp),
                                    ctypes.POINTER(c_cp), ctypes.POINTER(c_cp),
                                    ctypes.POINTER(c_cp)]),
    "vlcu_coverage_reduce": (c_i32, [_H, c_vp]),
    "vlcu_coverage_write": (c_i32, [_H, c_vp, c_cp]),
    "vlcu_vcd_write": (c_i32, [_H, c_vp, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, c_cp]),
    "vlcu_vcd_write_changes": (c_i32, [_H, c_vp, c_u32, c_u32, c_u32, c_cp]),
    "vlcu_comm_unique_id": (c_i32, [c_vp]),
    "vlcu_comm_init": (c_i32, [_H, c_vp, c_i32, c_i32]),
    "vlcu_comm_destroy": (c_i32, [_H]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library lacks a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


def load_design_library(path):
    """Load a library of further emitted designs (it registers them with the runtime at load
    time): tests/reftests/libvlcu_reftests.so, tests/veercov/libvlcu_veercov.so ..."""
    return ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)


class VlcuError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise VlcuError(f"vlcu error {rc}: {lib.vlcu_last_error().decode()}")


def designs():
    return [lib.vlcu_design_name(i).decode() for i in range(lib.vlcu_design_count())]


_ELEM_DTYPE = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


class Model:
    """N instances of one emitted design on one GPU (one shard)."""

    def __init__(self, design, n_instances, device=0, inst_base=0, seed=1, rand_reset=0,
                 msg_cap=4, plusargs=""):
        self.h = lib.vlcu_create(design.encode(), n_instances, device, inst_base, seed, rand_reset,
                                 msg_cap, plusargs.encode())
        if not self.h:
            raise VlcuError(f"vlcu_create failed: {lib.vlcu_last_error().decode()}")
        self.design = design
        self.n = n_instances
        self.npad = lib.vlcu_padded_instances(self.h)
        self.state_bytes = lib.vlcu_state_bytes(self.h)
        self.in_words = lib.vlcu_in_words(self.h)
        self.out_words = lib.vlcu_out_words(self.h)
        self.in_bytes = lib.vlcu_in_bytes(self.h)
        self.out_bytes = lib.vlcu_out_bytes(self.h)
        self.touched_read_bytes = lib.vlcu_touched_read_bytes(self.h)
        self.touched_write_bytes = lib.vlcu_touched_write_bytes(self.h)
        self.ports = {}
        for p in range(lib.vlcu_port_count(self.h)):
            info = PortInfo()
            _check(lib.vlcu_port_info(self.h, p, ctypes.byref(info)))
            self.ports[info.name.decode()] = (p, info.dir, info.width, info.elem_bytes, info.elems)
        self.members = {}
        for m in range(lib.vlcu_member_count(self.h)):
            info = PortInfo()
            _check(lib.vlcu_member_info(self.h, m, ctypes.byref(info)))
            self.members[info.name.decode()] = (m, info.dir, info.width, info.elem_bytes,
                                                info.elems, lib.vlcu_member_offset(self.h, m))

    def close(self):
        if self.h:
            lib.vlcu_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        """Re-construct every instance (time 0, _ctor_var_reset with the same seeds)."""
        _check(lib.vlcu_reset(self.h))

    # -- ports ---------------------------------------------------------------
    def _as_port_array(self, info, values):
        _, _, _, eb, elems = info[:5]
        arr = np.asarray(values, dtype=_ELEM_DTYPE[eb])
        if arr.size == elems and self.n != 1:
            return np.ascontiguousarray(arr.reshape(1, elems)), 0  # one value for all instances
        arr = np.ascontiguousarray(arr.reshape(self.n, elems))
        return arr, eb * elems

    def set_input(self, name, values):
        info = self.ports[name]
        arr, stride = self._as_port_array(info, values)
        _check(lib.vlcu_set_input(self.h, info[0], arr.ctypes.data, stride))

    def get_output(self, name):
        p, _, _, eb, elems = self.ports[name]
        out = np.zeros((self.n, elems), dtype=_ELEM_DTYPE[eb])
        _check(lib.vlcu_get_output(self.h, p, out.ctypes.data, eb * elems))
        return out[:, 0] if elems == 1 else out

    def set_member(self, name, values):
        info = self.members[name]
        arr, stride = self._as_port_array(info, values)
        _check(lib.vlcu_set_member(self.h, info[0], arr.ctypes.data, stride))

    def stage_member(self, name, arr, mask_width=False):
        """Start the asynchronous host->device copy of a member's next value; `arr` is an
        (n, elems) array in page-locked memory that stays untouched until after commit_staged()
        and the next synchronising call."""
        info = self.members[name]
        assert arr.shape[0] == self.n and arr.flags["C_CONTIGUOUS"]
        _check(lib.vlcu_stage_member(self.h, info[0], arr.ctypes.data, arr.strides[0], int(mask_width)))

    def commit_staged(self):
        _check(lib.vlcu_commit_staged(self.h))

    def get_member(self, name):
        m, _, _, eb, elems, _ = self.members[name]
        out = np.zeros((self.n, elems), dtype=_ELEM_DTYPE[eb])
        _check(lib.vlcu_get_member(self.h, m, out.ctypes.data, eb * elems))
        return out[:, 0] if elems == 1 else out

    # -- evaluation ------------------------------------------------------------
    def time_inc(self, add=1):
        _check(lib.vlcu_time_inc(self.h, add))

    def eval(self):
        _check(lib.vlcu_eval(self.h))

    def run(self, tb, n_evals, flags=0, stim=None, obs_evals=0, samples=None, trace_evals=0,
            tb_seed=0, tb_flags=0, tb_arg=(0, 0, 0, 0), trace_stream_words=0, tb_data=None):
        """Returns (obs, trace): obs is uint32 [obs_evals][out_words][n] or None; trace is a
        uint8 array [trace_evals][n_samples][record_bytes] or None -- or, with
        trace_stream_words, the uint32 change streams [n_samples][trace_stream_words]."""
        d = RunDesc()
        d.tb = tb.encode()
        d.n_evals = n_evals
        d.flags = flags
        keep = []
        if stim is not None:
            stim = np.ascontiguousarray(stim, dtype=np.uint32)
            # the runtime reads in_words * npad words per eval (in_words with STIM_BROADCAST): a
            # buffer of any other shape would skew every row, so it is rejected -- except rows of
            # exactly n columns, which are padded here
            if flags & RUNF_STIM_BROADCAST:
                assert stim.shape == (n_evals, self.in_words), (stim.shape, n_evals, self.in_words)
            else:
                assert stim.shape[:2] == (n_evals, self.in_words) and stim.shape[2] in (self.n, self.npad), \
                    (stim.shape, n_evals, self.in_words, self.n, self.npad)
                if stim.shape[2] != self.npad:
                    padded = np.zeros((n_evals, self.in_words, self.npad), dtype=np.uint32)
                    padded[:, :, :self.n] = stim
                    stim = padded
            keep.append(stim)
            d.stim = stim.ctypes.data
        obs = None
        if obs_evals:
            obs = np.zeros((obs_evals, self.out_words, self.npad), dtype=np.uint32)
            d.obs = obs.ctypes.data
            d.obs_evals = obs_evals
        trace = None
        if samples is not None and len(samples) and trace_evals:
            samples = np.ascontiguousarray(samples, dtype=np.uint64)
            keep.append(samples)
            rec = lib.vlcu_trace_record_bytes(self.h)
            if trace_stream_words:
                trace = np.zeros((len(samples), trace_stream_words), dtype=np.uint32)
                d.flags |= RUNF_TRACE_CHANGES
                d.trace_stream_words = trace_stream_words
            else:
                trace = np.zeros((trace_evals, len(samples), rec), dtype=np.uint8)
            d.sample_instances = samples.ctypes.data
            d.n_samples = len(samples)
            d.trace_out = trace.ctypes.data
            d.trace_evals = trace_evals
        if tb_data is not None:  # per-run data of the design's testbench (e.g. a seed per instance)
            tb_data = np.ascontiguousarray(tb_data)
            keep.append(tb_data)
            d.tb_data = tb_data.ctypes.data
            d.tb_data_bytes = tb_data.nbytes
        d.tb_seed = tb_seed
        d.tb_flags = tb_flags
        for k in range(4):
            d.tb_arg[k] = tb_arg[k]
        _check(lib.vlcu_run(self.h, ctypes.byref(d)))
        if obs is not None:
            obs = obs[:, :, :self.n]
        return obs, trace

    def run_rows(self, n_evals, stim_ptr, obs_ptr=None, flags=0):
        """'trace' run over caller-owned row buffers (host pointers, or device pointers with
        RUNF_STIM_DEVICE / RUNF_OBS_DEVICE): stim uint32 [n_evals][in_words][npad], obs uint32
        [n_evals][out_words][npad]."""
        d = RunDesc()
        d.tb = b"trace"
        d.n_evals = n_evals
        d.flags = flags
        d.stim = stim_ptr
        if obs_ptr:
            d.obs = obs_ptr
            d.obs_evals = n_evals
        _check(lib.vlcu_run(self.h, ctypes.byref(d)))

    def set_pipeline_chunk(self, nbytes):
        _check(lib.vlcu_set_pipeline_chunk(self.h, nbytes))

    def final(self):
        _check(lib.vlcu_final(self.h))

    def sync(self):
        _check(lib.vlcu_sync(self.h))

    def flush_l2(self):
        _check(lib.vlcu_flush_l2(self.h))

    def last_kernel_ms(self):
        ms = ctypes.c_float()
        _check(lib.vlcu_last_kernel_ms(self.h, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        return lib.vlcu_launch_count(self.h)

    # -- observation -------------------------------------------------------------
    def status(self):
        out = np.zeros(self.n, dtype=STATUS_DTYPE)
        _check(lib.vlcu_status(self.h, out.ctypes.data))
        return out

    def signatures(self):
        out = np.zeros(self.n, dtype=np.uint64)
        _check(lib.vlcu_signatures(self.h, out.ctypes.data))
        return out

    def results_begin(self, sigs=None, status=None, members=()):
        """Start the asynchronous read-back of signatures (uint64 [n]), status (STATUS_DTYPE [n]) and
        `members` = [(name, array [n][elems])...] into page-locked arrays; results_wait() completes it."""
        k = len(members)
        idx = (c_i32 * max(k, 1))(*[self.members[name][0] for name, _ in members])
        dsts = (c_vp * max(k, 1))(*[arr.ctypes.data for _, arr in members])
        _check(lib.vlcu_results_begin(self.h, sigs.ctypes.data if sigs is not None else None,
                                      status.ctypes.data if status is not None else None, k, idx, dsts))

    def results_wait(self):
        _check(lib.vlcu_results_wait(self.h))

    def messages(self, instance, cap=64):
        out = np.zeros(cap, dtype=MSG_DTYPE)
        count = c_u32()
        _check(lib.vlcu_messages(self.h, instance, out.ctypes.data, cap, ctypes.byref(count)))
        return out[:min(count.value, cap)], count.value

    def message_format(self, msg_id):
        s = lib.vlcu_message_format(self.h, msg_id)
        return s.decode() if s is not None else None

    def format_message(self, msg):
        """Host-side formatting of a logged $display/$write (subset: %%, %0t/%t, %d, %x, %b)."""
        fmt = self.message_format(int(msg["id"]))
        out, args, i = [], list(msg["args"][:int(msg["nargs"])]), 0
        while i < len(fmt):
            c = fmt[i]
            if c != "%":
                out.append(c)
                i += 1
                continue
            j = i + 1
            while j < len(fmt) and (fmt[j].isdigit() or fmt[j] in "-."):
                j += 1
            spec = fmt[j].lower()
            if spec == "%":
                out.append("%")
            else:
                v = int(args.pop(0)) if args else 0
                out.append({"t": str(v), "d": str(v), "x": format(v, "x"), "h": format(v, "x"),
                            "b": format(v, "b")}.get(spec, str(v)))
            i = j + 1
        return "".join(out)

    def snapshot(self, instance):
        out = np.zeros(self.state_bytes, dtype=np.uint8)
        _check(lib.vlcu_snapshot(self.h, instance, out.ctypes.data, out.nbytes))
        return out

    def save_instance(self, instance, filename):
        """Write instance `instance` as a verilated_save image a --savable C++ model can restore."""
        _check(lib.vlcu_save_instance(self.h, instance, filename.encode()))

    def restore_instance(self, instance, filename):
        _check(lib.vlcu_restore_instance(self.h, instance, filename.encode()))

    def record_member(self, record, name):
        """Extract one member from a packed state record (snapshot or trace sample body)."""
        _, _, _, eb, elems, off = self.members[name]
        vals = np.frombuffer(record.tobytes()[off:off + eb * elems], dtype=_ELEM_DTYPE[eb])
        return vals[0] if elems == 1 else vals

    # -- coverage -----------------------------------------------------------------
    def coverage_bins(self):
        return lib.vlcu_coverage_bins(self.h)

    def coverage_reduce(self):
        out = np.zeros(self.coverage_bins(), dtype=np.uint64)
        _check(lib.vlcu_coverage_reduce(self.h, out.ctypes.data))
        return out

    def vcd_write(self, trace, sample, filename):
        """trace: the uint8 [trace_evals][n_samples][record] array run() returned."""
        trace = np.ascontiguousarray(trace, dtype=np.uint8)
        _check(lib.vlcu_vcd_write(self.h, trace.ctypes.data, trace.shape[0], trace.shape[1], sample,
                                  filename.encode()))

    def vcd_write_changes(self, streams, sample, filename):
        """streams: the uint32 [n_samples][trace_stream_words] array run(trace_stream_words=) returned."""
        streams = np.ascontiguousarray(streams, dtype=np.uint32)
        _check(lib.vlcu_vcd_write_changes(self.h, streams.ctypes.data, streams.shape[1], streams.shape[0],
                                          sample, filename.encode()))

    def coverage_write(self, counts, filename):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        _check(lib.vlcu_coverage_write(self.h, counts.ctypes.data, filename.encode()))

    # -- multi-GPU -------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = ctypes.create_string_buffer(128)
        _check(lib.vlcu_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid, rank, world):
        _check(lib.vlcu_comm_init(self.h, uid, rank, world))
