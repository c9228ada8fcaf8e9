"""Operator differential-test support (TEST INFRASTRUCTURE).

Generates seeded random test vectors for every VL_* operator listed in
oracle/optest_ops.inc, honouring each operator's documented preconditions
(clean inputs, width classes of the I/Q/W type letters, in-range bit
positions -- include/verilated_funcs.h EMIT_RULE comments), runs them through

  * the ORACLE: oracle/_ref/libvlref.so = the reference's own
    include/verilated_funcs.h + verilated.cpp compiled from /root/reference
  * the device library include/verilated_cuda_funcs.h, either on a B200
    (tests/optest/libvlcu_optest.so) or host-compiled for CPU-side debugging
    (tests/optest/libvlcu_optest_host.so)

and compares the complete result limbs bit for bit (dirty bits included).
"""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NP, NIN, NOUT = 8, 16, 160
OPVEC = np.dtype([("p", "<i4", NP), ("a", "<u4", NIN), ("b", "<u4", NIN), ("c", "<u4", NIN),
                  ("out", "<u4", NOUT)])
assert OPVEC.itemsize == 4 * (NP + 3 * NIN + NOUT)


def op_names():
    names = []
    with open(os.path.join(ROOT, "oracle", "optest_ops.inc")) as f:
        for line in f:
            m = re.match(r"^OP\((\w+),", line)
            if m:
                names.append(m.group(1))
    return names


OPS = op_names()
OPID = {n: i for i, n in enumerate(OPS)}


def _load(path, fn):
    lib = ctypes.CDLL(path)
    f = getattr(lib, fn)
    f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64]
    f.restype = ctypes.c_int
    return lib, f


def load_ref():
    return _load(os.path.join(ROOT, "oracle", "_ref", "libvlref.so"), "vlref_optest_run")


def load_host():
    return _load(os.path.join(ROOT, "tests", "optest", "libvlcu_optest_host.so"),
                 "vlcu_optest_run_host")


def load_cuda():
    return _load(os.path.join(ROOT, "tests", "optest", "libvlcu_optest.so"),
                 "vlcu_optest_run_cuda")


def run(fn, op, vecs):
    out = np.ascontiguousarray(vecs.copy())
    rc = fn(OPID[op], out.ctypes.data, len(out))
    assert rc == 0, f"{op}: dispatcher returned {rc}"
    return out


# ---------------------------------------------------------------------------
# value generation

SPECIAL = np.array([0, 1, 2, 0x7fffffff, 0x80000000, 0x80000001, 0xfffffffe, 0xffffffff,
                    0x0000ffff, 0xffff0000, 0x55555555, 0xaaaaaaaa], dtype=np.uint32)


def mask_to_bits(words, bits):
    """Zero every bit at position >= bits[i] of row i (words: (n, W) uint32)."""
    n, w = words.shape
    bits = np.asarray(bits, dtype=np.int64).reshape(n, 1)
    lo = np.arange(w, dtype=np.int64).reshape(1, w) * 32
    nb = np.clip(bits - lo, 0, 32)  # valid bits in each limb
    m = np.where(nb >= 32, 0xffffffff, (np.uint64(1) << nb.astype(np.uint64)) - np.uint64(1))
    return (words & m.astype(np.uint32)).astype(np.uint32)


def rand_val(rng, n, bits=None, w=NIN):
    """n random values of up to w limbs, drawn from a mix of distributions that
    exercises carries, borrows, sign bits and sparse/dense patterns."""
    u = lambda: rng.integers(0, 1 << 32, size=(n, w), dtype=np.uint64).astype(np.uint32)
    mode = rng.integers(0, 8, size=n)
    v = u()
    sparse = u() & u() & u()
    dense = u() | u() | u()
    special = SPECIAL[rng.integers(0, len(SPECIAL), size=(n, w))]
    small = np.zeros((n, w), dtype=np.uint32)
    small[:, 0] = rng.integers(0, 4, size=n)
    onehot = np.zeros((n, w), dtype=np.uint32)
    pos = rng.integers(0, w * 32, size=n)
    onehot[np.arange(n), pos // 32] = (np.uint32(1) << (pos % 32).astype(np.uint32))
    out = np.select([(mode == 0)[:, None], (mode == 1)[:, None], (mode == 2)[:, None],
                     (mode == 3)[:, None], (mode == 4)[:, None], (mode == 5)[:, None],
                     (mode == 6)[:, None]],
                    [v, sparse, dense, special, small, np.full((n, w), 0xffffffff, np.uint32),
                     onehot], default=v).astype(np.uint32)
    zero = rng.integers(0, 24, size=n) == 0
    out[zero] = 0
    if bits is not None:
        out = mask_to_bits(out, bits)
        # a one-hot beyond the width was masked to zero: that is fine (tests zero)
    return out


def new_vecs(n):
    return np.zeros(n, dtype=OPVEC)


def ri(rng, lo, hi, n):
    """Random ints in [lo, hi] inclusive; lo/hi may be arrays."""
    lo = np.broadcast_to(np.asarray(lo, dtype=np.int64), (n,))
    hi = np.broadcast_to(np.asarray(hi, dtype=np.int64), (n,))
    return (lo + (rng.random(n) * (hi - lo + 1)).astype(np.int64)).clip(lo, hi)


def shift_amounts(rng, n, obits):
    """Mostly in-range shift counts, plus boundary and huge ones."""
    kind = rng.integers(0, 8, size=n)
    inr = ri(rng, 0, np.maximum(np.asarray(obits) - 1, 0), n)
    near = np.asarray(obits) + rng.integers(-2, 3, size=n)
    huge = rng.integers(0, 1 << 32, size=n, dtype=np.int64)
    words = (inr // 32) * 32
    return np.select([kind <= 3, kind == 4, kind == 5, kind == 6], [inr, near, huge, words],
                     default=inr).clip(0, (1 << 32) - 1).astype(np.uint32)


def set_q(arr, q):
    q = np.asarray(q, dtype=np.uint64)
    arr[:, 0] = (q & np.uint64(0xffffffff)).astype(np.uint32)
    arr[:, 1] = (q >> np.uint64(32)).astype(np.uint32)


def gen(op, n, seed):
    """Build n vectors for operator 'op'."""
    rng = np.random.default_rng(seed ^ (OPID[op] * 7919))
    v = new_vecs(n)
    p, a, b, out = v["p"], v["a"], v["b"], v["out"]
    wI = lambda: ri(rng, 1, 32, n)
    wQ = lambda: ri(rng, 33, 64, n)
    wIQ = lambda: ri(rng, 1, 64, n)
    wW = lambda hi=512: ri(rng, 65, hi, n)
    full = lambda: rand_val(rng, n)
    preload = lambda: rng.integers(0, 1 << 32, size=(n, NOUT), dtype=np.uint64).astype(np.uint32)
    base = re.sub(r"_(C|S|I)$", "", op) if re.search(r"_(II|IO|IQ|IW)_(C|S|I)$", op) else op
    sfx = op[len(base):]
    narrow_bits = {"_C": 8, "_S": 16, "_I": 32}.get(sfx, 32)

    if op in ("VL_CLEAN_II",):
        p[:, 0] = wI(); a[:] = full()
    elif op == "VL_CLEAN_QQ":
        p[:, 0] = wIQ(); a[:] = full()
    elif op in ("VL_CLEAN_WW", "VL_ASSIGN_W"):
        p[:, 0] = wW(); a[:] = full(); out[:] = preload()
    elif op in ("VL_ZERO_W", "VL_ALLONES_W"):
        p[:, 0] = wW(); out[:] = preload()
    elif base in ("VL_ASSIGNBIT_II", "VL_ASSIGNBIT_IO"):
        p[:, 0] = ri(rng, 0, narrow_bits - 1, n); a[:, 0] = rng.integers(0, 2, size=n)
        out[:, 0] = rand_val(rng, n, np.full(n, narrow_bits))[:, 0]
    elif op in ("VL_ASSIGNBIT_QI", "VL_ASSIGNBIT_QO"):
        p[:, 0] = ri(rng, 0, 63, n); a[:, 0] = rng.integers(0, 2, size=n)
        out[:, :2] = full()[:, :2]
    elif op in ("VL_ASSIGNBIT_WI", "VL_ASSIGNBIT_WO"):
        bits = wW(); p[:, 0] = ri(rng, 0, bits - 1, n); a[:, 0] = rng.integers(0, 2, size=n)
        out[:, :NIN] = rand_val(rng, n, bits)
    elif op in ("VL_EXTEND_WI", "VL_EXTENDS_WI"):
        p[:, 0] = wW(); p[:, 1] = wI(); a[:] = rand_val(rng, n, p[:, 1]); out[:] = preload()
    elif op in ("VL_EXTEND_WQ", "VL_EXTENDS_WQ"):
        p[:, 0] = wW(); p[:, 1] = wQ(); a[:] = rand_val(rng, n, p[:, 1]); out[:] = preload()
    elif op in ("VL_EXTEND_WW", "VL_EXTENDS_WW"):
        p[:, 0] = wW(); p[:, 1] = ri(rng, 65, p[:, 0], n); a[:] = rand_val(rng, n, p[:, 1])
        out[:] = preload()
    elif op in ("VL_EXTENDS_II", "VL_EXTENDS_QI"):
        p[:, 0] = 32; p[:, 1] = wI(); a[:] = rand_val(rng, n, p[:, 1])
    elif op == "VL_EXTENDS_QQ":
        p[:, 0] = 64; p[:, 1] = wIQ(); a[:] = rand_val(rng, n, p[:, 1])
    elif op in ("VL_REDAND_II", "VL_REDAND_IQ", "VL_REDAND_IW"):
        p[:, 0] = {"VL_REDAND_II": wI, "VL_REDAND_IQ": wQ, "VL_REDAND_IW": wW}[op]()
        a[:] = rand_val(rng, n, p[:, 0])
        ones = rng.integers(0, 3, size=n) == 0  # make all-ones common
        a[ones] = mask_to_bits(np.full((n, NIN), 0xffffffff, np.uint32), p[:, 0])[ones]
    elif op in ("VL_REDOR_W", "VL_REDXOR_W", "VL_COUNTONES_W", "VL_ONEHOT_W", "VL_ONEHOT0_W",
                "VL_CLOG2_W", "VL_MOSTSETBITP1_W", "VL_NOT_W", "VL_NEGATE_W"):
        p[:, 0] = ri(rng, 1, NIN, n); a[:] = full()
        if op in ("VL_NOT_W", "VL_NEGATE_W"):
            out[:] = preload()
    elif op == "VL_NEGATE_INPLACE_W":
        p[:, 0] = ri(rng, 1, NIN, n); out[:, :NIN] = full()
    elif re.match(r"VL_(REDXOR_\d+|COUNTONES_[IQ]|ONEHOT0?_[IQ]|CLOG2_[IQ]|MOSTSETBITP1_[IQ])$",
                  op):
        a[:] = full()
    elif op in ("VL_COUNTBITS_I", "VL_COUNTBITS_Q", "VL_COUNTBITS_W"):
        p[:, 0] = {"I": wI, "Q": wQ, "W": wW}[op[-1]]()
        for k in (1, 2, 3):
            p[:, k] = rng.integers(0, 2, size=n)
        a[:] = rand_val(rng, n, p[:, 0])
    elif op in ("VL_AND_W", "VL_OR_W", "VL_XOR_W", "VL_CHANGEXOR_W", "VL_ADD_W", "VL_SUB_W",
                "VL_MUL_W", "VL_EQ_W", "VL_NEQ_W", "VL_LT_W", "VL_LTE_W", "VL_GT_W", "VL_GTE_W"):
        p[:, 0] = ri(rng, 1, NIN, n); a[:] = full(); b[:] = full(); out[:] = preload()
        same = rng.integers(0, 4, size=n) == 0
        b[same] = a[same]
        tweak = rng.integers(0, 4, size=n) == 1  # differ in exactly one limb
        idx = ri(rng, 0, p[:, 0] - 1, n)
        b[tweak] = a[tweak]
        b[np.where(tweak)[0], idx[tweak]] ^= np.uint32(1) << rng.integers(0, 32, size=int(tweak.sum())).astype(np.uint32)
        if op not in ("VL_AND_W", "VL_OR_W", "VL_XOR_W", "VL_ADD_W", "VL_SUB_W", "VL_MUL_W"):
            out[:] = 0
    elif op.endswith("S_IWW") or op in ("VL_MULS_WWW", "VL_DIVS_WWW", "VL_MODDIVS_WWW",
                                        "VL_DIV_WWW", "VL_MODDIV_WWW"):
        p[:, 0] = ri(rng, 33, 512, n) if op.endswith("S_IWW") else wW()
        a[:] = rand_val(rng, n, p[:, 0]); b[:] = rand_val(rng, n, p[:, 0])
        same = rng.integers(0, 6, size=n) == 0
        b[same] = a[same]
        if "DIV" in op:
            # divisors of assorted magnitudes: short, one limb, same size as dividend
            kind = rng.integers(0, 6, size=n)
            nb = np.select([kind == 0, kind == 1, kind == 2],
                           [ri(rng, 1, 32, n), ri(rng, 33, 64, n), ri(rng, 1, p[:, 0], n)],
                           default=p[:, 0])
            b[:] = mask_to_bits(b, nb)
            out[:] = preload()
        elif "MULS" in op:
            out[:] = preload()
    elif re.match(r"VL_(GTS|GTES|LTS|LTES)_III$", op) or op in ("VL_MULS_III", "VL_DIVS_III",
                                                              "VL_MODDIVS_III"):
        p[:, 0] = wI(); a[:] = rand_val(rng, n, p[:, 0]); b[:] = rand_val(rng, n, p[:, 0])
        same = rng.integers(0, 6, size=n) == 0
        b[same] = a[same]
        if "DIV" in op:  # the MIN / -1 corner
            k = rng.integers(0, 16, size=n) == 0
            p[k, 0] = 32; a[k, 0] = 0x80000000; b[k, 0] = 0xffffffff
    elif re.match(r"VL_(GTS|GTES|LTS|LTES)_IQQ$", op) or op in ("VL_MULS_QQQ", "VL_DIVS_QQQ",
                                                              "VL_MODDIVS_QQQ"):
        p[:, 0] = wQ(); a[:] = rand_val(rng, n, p[:, 0]); b[:] = rand_val(rng, n, p[:, 0])
        same = rng.integers(0, 6, size=n) == 0
        b[same] = a[same]
        if "DIV" in op:
            k = rng.integers(0, 16, size=n) == 0
            p[k, 0] = 64; a[k, 0] = 0; a[k, 1] = 0x80000000; b[k, 0] = 0xffffffff; b[k, 1] = 0xffffffff
    elif op in ("VL_DIV_III", "VL_MODDIV_III", "VL_DIV_QQQ", "VL_MODDIV_QQQ"):
        p[:, 0] = 32 if op.endswith("III") else 64
        a[:] = full(); b[:] = full()
        small = rng.integers(0, 3, size=n) == 0
        b[small, 1] = 0
    elif op in ("VL_POW_III", "VL_POWSS_III"):
        p[:, 0] = wI(); p[:, 1] = p[:, 0]; p[:, 2] = wI(); p[:, 3] = rng.integers(0, 4, size=n)
        a[:] = rand_val(rng, n, p[:, 0]); b[:] = rand_val(rng, n, p[:, 2])
    elif op in ("VL_POW_QQQ", "VL_POWSS_QQQ"):
        p[:, 0] = wIQ(); p[:, 1] = p[:, 0]; p[:, 2] = wIQ(); p[:, 3] = rng.integers(0, 4, size=n)
        a[:] = rand_val(rng, n, p[:, 0]); b[:] = rand_val(rng, n, p[:, 2])
    elif op in ("VL_POW_WWW", "VL_POWSS_WWW"):
        p[:, 0] = wW(256); p[:, 1] = p[:, 0]; p[:, 2] = ri(rng, 65, 128, n)
        p[:, 3] = rng.integers(0, 4, size=n)
        a[:] = rand_val(rng, n, p[:, 0]); b[:] = rand_val(rng, n, p[:, 2]); out[:] = preload()
    elif op in ("VL_POW_WWQ", "VL_POWSS_WWQ"):
        p[:, 0] = wW(256); p[:, 1] = p[:, 0]; p[:, 2] = wIQ(); p[:, 3] = rng.integers(0, 4, size=n)
        a[:] = rand_val(rng, n, p[:, 0]); b[:] = rand_val(rng, n, p[:, 2]); out[:] = preload()
    elif op in ("VL_POW_QQW", "VL_POWSS_QQW"):
        p[:, 0] = wIQ(); p[:, 1] = p[:, 0]; p[:, 2] = ri(rng, 65, 128, n)
        p[:, 3] = rng.integers(0, 4, size=n)
        a[:] = rand_val(rng, n, p[:, 0]); b[:] = rand_val(rng, n, p[:, 2])
    elif base == "_vl_insert_II":
        rb = ri(rng, 1, narrow_bits, n); hb = ri(rng, 0, rb - 1, n); lb = ri(rng, 0, hb, n)
        p[:, 0] = hb; p[:, 1] = lb; p[:, 2] = rb; a[:] = full()
        out[:, 0] = rand_val(rng, n, rb)[:, 0]
    elif op == "_vl_insert_QQ":
        rb = wIQ(); hb = ri(rng, 0, rb - 1, n); lb = ri(rng, 0, hb, n)
        p[:, 0] = hb; p[:, 1] = lb; p[:, 2] = rb; a[:] = full()
        out[:, :2] = rand_val(rng, n, rb)[:, :2]
    elif op in ("_vl_insert_WI", "_vl_insert_WQ", "_vl_insert_WW"):
        rb = wW(); lim = {"_vl_insert_WI": 32, "_vl_insert_WQ": 64, "_vl_insert_WW": 512}[op]
        lb = ri(rng, 0, rb - 1, n)
        aligned = rng.integers(0, 4, size=n) == 0
        lb = np.where(aligned, (lb // 32) * 32, lb)
        width = ri(rng, 1, np.minimum(lim, rb - lb), n)
        fullword = rng.integers(0, 5, size=n) == 0  # whole-limb fields
        width = np.where(fullword & aligned, np.minimum(((width + 31) // 32) * 32, np.minimum(lim, ((rb - lb) // 32) * 32)), width)
        width = np.maximum(width, 1)
        p[:, 0] = lb + width - 1; p[:, 1] = lb; p[:, 2] = rb
        norb = rng.integers(0, 5, size=n) == 0  # callers that pass rbits=0 (concat, replicate, shift)
        p[norb, 2] = 0
        a[:] = full()
        out[:] = preload(); out[:, :NIN] = rand_val(rng, n, rb)
    elif op == "_vl_insert_IW":
        src = wW(); rb = wI(); width = ri(rng, 1, rb, n); lb = ri(rng, 0, src - width, n)
        p[:, 0] = lb + width - 1; p[:, 1] = lb; p[:, 2] = rb
        a[:] = rand_val(rng, n, src); out[:, 0] = rand_val(rng, n, rb)[:, 0]
    elif op in ("VL_REPLICATE_IOI", "VL_REPLICATE_QOI"):
        p[:, 0] = 1; p[:, 1] = ri(rng, 1, 32 if "IOI" in op else 64, n)
        a[:, 0] = rng.integers(0, 2, size=n)
    elif op == "VL_REPLICATE_III":
        lb = ri(rng, 1, 16, n); p[:, 0] = lb; p[:, 1] = ri(rng, 1, 32 // lb, n)
        a[:] = rand_val(rng, n, lb)
    elif op == "VL_REPLICATE_QII":
        lb = ri(rng, 2, 32, n); p[:, 0] = lb; p[:, 1] = ri(rng, (32 // lb) + 1, 64 // lb, n)
        a[:] = rand_val(rng, n, lb)
    elif op == "VL_REPLICATE_WII":
        lb = wI(); p[:, 0] = lb; p[:, 1] = ri(rng, (64 // lb) + 1, 512 // lb, n)
        a[:] = rand_val(rng, n, lb); out[:] = preload()
    elif op == "VL_REPLICATE_WQI":
        lb = wQ(); p[:, 0] = lb; p[:, 1] = ri(rng, 2, 512 // lb, n)
        a[:] = rand_val(rng, n, lb); out[:] = preload()
    elif op == "VL_REPLICATE_WWI":
        lb = wW(256); p[:, 0] = lb; p[:, 1] = ri(rng, 1, 512 // lb, n)
        a[:] = rand_val(rng, n, lb); out[:] = preload()
    elif op in ("VL_STREAML_FAST_III", "VL_STREAML_FAST_QQI"):
        q = op.endswith("QQI")
        p[:, 0] = wQ() if q else wI(); p[:, 1] = ri(rng, 0, 6 if q else 5, n)
        a[:] = rand_val(rng, n, p[:, 0])
    elif op in ("VL_STREAML_III", "VL_STREAML_QQI", "VL_STREAML_WWI"):
        p[:, 0] = {"VL_STREAML_III": wI, "VL_STREAML_QQI": wQ, "VL_STREAML_WWI": wW}[op]()
        p[:, 1] = ri(rng, 1, p[:, 0], n)
        small = rng.integers(0, 2, size=n) == 0
        p[small, 1] = np.minimum(ri(rng, 1, 9, n), p[:, 0])[small]
        a[:] = rand_val(rng, n, p[:, 0]); out[:] = preload() if op.endswith("WWI") else 0
    elif op.startswith("VL_CONCAT_"):
        cls = {"I": wI, "Q": wQ, "W": lambda: wW(256)}
        lt, rt = op[-2], op[-1]
        lbits = cls[lt](); rbits = cls[rt]()
        if op[-3] == "I":  # narrow outputs: keep the sum inside the type
            lbits = ri(rng, 1, 31, n); rbits = ri(rng, 1, 32 - lbits, n)
        elif op[-3] == "Q":
            if lt == "I" and rt == "I":
                lbits = wI(); rbits = ri(rng, np.maximum(33 - lbits, 1), 32, n)
            elif lt == "I":
                rbits = ri(rng, 33, 63, n); lbits = ri(rng, 1, np.minimum(32, 64 - rbits), n)
            elif rt == "I":
                lbits = ri(rng, 33, 63, n); rbits = ri(rng, 1, np.minimum(32, 64 - lbits), n)
            else:
                lbits = ri(rng, 1, 63, n); rbits = ri(rng, 1, 64 - lbits, n)
        p[:, 0] = lbits + rbits; p[:, 1] = lbits; p[:, 2] = rbits
        a[:] = rand_val(rng, n, lbits); b[:] = rand_val(rng, n, rbits)
        if op[-3] == "W":
            out[:] = preload()
    elif re.match(r"VL_SHIFT(L|R|RS)_", op):
        ot, lt, rt = op[-3], op[-2], op[-1]
        ob = {"I": wI, "Q": wQ, "W": wW}[ot]()
        lb = ob.copy()
        if "SHIFTRS" in op:
            if ot == "W":  # sign must sit in the top limb of the output
                lb = ri(rng, np.maximum(((ob - 1) // 32) * 32 + 1, 65), ob, n)
            elif ot == "I" and lt == "Q":
                lb = wQ()
            else:
                lb = ri(rng, 1, ob, n)
        rb = {"I": np.full(n, 32), "Q": wQ(), "W": ri(rng, 65, 128, n)}[rt]
        p[:, 0] = ob; p[:, 1] = lb; p[:, 2] = rb
        a[:] = rand_val(rng, n, lb if "SHIFTRS" in op else ob)
        amt = shift_amounts(rng, n, ob)
        b[:, 0] = amt
        if rt != "I":  # occasionally a shift count with high limbs set
            hi = rand_val(rng, n, rb)
            big = rng.integers(0, 6, size=n) == 0
            b[big, 1:] = hi[big, 1:]
            b[:] = mask_to_bits(b, rb)
        if ot == "W":
            out[:] = preload()
    elif op == "VL_BITSEL_IWII":
        lb = wW(448); p[:, 0] = lb; p[:, 1] = ri(rng, 0, lb + 3, n); a[:] = rand_val(rng, n, lb)
    elif op in ("VL_SEL_IWII", "VL_SEL_QWII"):
        lb = wW(440); wd = wI() if op == "VL_SEL_IWII" else wQ()
        lsb = ri(rng, 0, lb, n)
        inr = rng.integers(0, 5, size=n) != 0
        lsb = np.where(inr, ri(rng, 0, np.maximum(lb - wd, 0), n), lsb)
        p[:, 0] = lb; p[:, 1] = lsb; p[:, 2] = wd; a[:] = rand_val(rng, n, lb)
    elif op == "VL_SEL_WWII":
        lb = ri(rng, 140, 440, n); wd = ri(rng, 65, lb - 8, n)
        lsb = ri(rng, 0, lb - wd, n)
        aligned = rng.integers(0, 4, size=n) == 0
        lsb = np.where(aligned, (lsb // 32) * 32, lsb)
        oob = rng.integers(0, 8, size=n) == 0
        lsb = np.where(oob, lb - wd + ri(rng, 2, 40, n), lsb)
        p[:, 0] = wd; p[:, 1] = lb; p[:, 2] = lsb; p[:, 3] = wd
        a[:] = rand_val(rng, n, lb); out[:] = preload()
    elif base == "VL_ASSIGNSEL_II":
        rb = ri(rng, 1, narrow_bits, n); ob = ri(rng, 1, rb, n); lsb = ri(rng, 0, rb - ob, n)
        p[:, 0] = rb; p[:, 1] = ob; p[:, 2] = lsb; a[:] = full()
        out[:, 0] = rand_val(rng, n, rb)[:, 0]
    elif op in ("VL_ASSIGNSEL_QI", "VL_ASSIGNSEL_QQ"):
        rb = wQ(); ob = ri(rng, 1, np.minimum(rb, 32 if op.endswith("QI") else 64), n)
        lsb = ri(rng, 0, rb - ob, n)
        p[:, 0] = rb; p[:, 1] = ob; p[:, 2] = lsb; a[:] = full()
        out[:, :2] = rand_val(rng, n, rb)[:, :2]
    elif op in ("VL_ASSIGNSEL_WI", "VL_ASSIGNSEL_WQ", "VL_ASSIGNSEL_WW"):
        rb = wW(); lim = {"I": 32, "Q": 64, "W": 512}[op[-1]]
        ob = ri(rng, 1, np.minimum(rb, lim), n); lsb = ri(rng, 0, rb - ob, n)
        aligned = rng.integers(0, 4, size=n) == 0
        lsb = np.where(aligned, (lsb // 32) * 32, lsb)
        p[:, 0] = rb; p[:, 1] = ob; p[:, 2] = lsb; a[:] = full()
        out[:] = preload(); out[:, :NIN] = rand_val(rng, n, rb)
    elif base == "VL_SELASSIGN_II":
        rb = ri(rng, 1, narrow_bits, n); ob = ri(rng, 1, rb, n); ro = ri(rng, 0, 32 - ob, n)
        p[:, 0] = rb; p[:, 1] = ob; p[:, 2] = ro; a[:] = full()
        out[:, 0] = rand_val(rng, n, rb)[:, 0]
    elif base == "VL_SELASSIGN_IQ":
        rb = ri(rng, 1, narrow_bits, n); ob = ri(rng, 1, rb, n); ro = ri(rng, 0, 64 - ob, n)
        p[:, 0] = rb; p[:, 1] = ob; p[:, 2] = ro; a[:] = full()
        out[:, 0] = rand_val(rng, n, rb)[:, 0]
    elif op == "VL_SELASSIGN_QQ":
        rb = wQ(); ob = ri(rng, 1, rb, n); ro = ri(rng, 0, 64 - ob, n)
        p[:, 0] = rb; p[:, 1] = ob; p[:, 2] = ro; a[:] = full()
        out[:, :2] = rand_val(rng, n, rb)[:, :2]
    elif base == "VL_SELASSIGN_IW":
        src = wW(); rb = ri(rng, 1, narrow_bits, n); ob = ri(rng, 1, rb, n)
        ro = ri(rng, 0, src - ob, n)
        p[:, 0] = rb; p[:, 1] = ob; p[:, 2] = ro; a[:] = rand_val(rng, n, src)
        out[:, 0] = rand_val(rng, n, rb)[:, 0]
    elif op == "VL_SELASSIGN_QW":
        src = wW(); rb = wQ(); ob = ri(rng, 1, rb, n); ro = ri(rng, 0, src - ob, n)
        p[:, 0] = rb; p[:, 1] = ob; p[:, 2] = ro; a[:] = rand_val(rng, n, src)
        out[:, :2] = rand_val(rng, n, rb)[:, :2]
    elif op == "VL_SELASSIGN_WW":
        src = ri(rng, 130, 512, n); rb = ri(rng, 65, src, n); ob = ri(rng, 1, rb, n)
        ro = ri(rng, 0, src - ob, n)
        aligned = rng.integers(0, 4, size=n) == 0
        ro = np.where(aligned, (ro // 32) * 32, ro)
        p[:, 0] = rb; p[:, 1] = ob; p[:, 2] = ro; a[:] = rand_val(rng, n, src)
        out[:] = preload(); out[:, :NIN] = rand_val(rng, n, rb)
    elif op == "VL_COND_WIWW":
        p[:, 0] = wW(); p[:, 1] = rng.integers(0, 2, size=n); a[:] = full(); b[:] = full()
        out[:] = preload()
    elif re.match(r"VL_CONST_W_(\d)X$", op):
        k = int(op[-2]); p[:, 0] = ri(rng, 32 * (k - 1) + 1, 512, n); a[:] = full(); out[:] = preload()
    elif re.match(r"VL_CONSTHI_W_(\d)X$", op):
        k = int(op[-2]); p[:, 1] = 256; p[:, 0] = ri(rng, 256 + 32 * (k - 1) + 1, 512, n)
        a[:] = full(); out[:] = preload()
    elif op == "VL_CONSTLO_W_8X":
        p[:, 0] = rng.integers(0, 2, size=n) * 256; a[:] = full(); out[:] = preload()
    elif op.startswith("VL_COV_TOGGLE_CHG_ST_"):
        p[:, 0] = {"I": wI, "Q": wQ, "W": lambda: ri(rng, 65, 80, n)}[op[-1]]()
        a[:] = rand_val(rng, n, p[:, 0]); b[:] = rand_val(rng, n, p[:, 0])
        same = rng.integers(0, 4, size=n) == 0
        b[same] = a[same]
        out[:] = preload()
        wrap = rng.integers(0, 4, size=n) == 0  # counters about to wrap
        out[wrap] = 0xffffffff
    else:
        raise KeyError(f"no generator for {op}")
    return v


def first_mismatch(op, want, got):
    bad = np.nonzero((want["out"] != got["out"]).any(axis=1))[0]
    if not len(bad):
        return None
    i = int(bad[0])
    w = int(np.nonzero(want["out"][i] != got["out"][i])[0][0])
    hx = lambda arr: " ".join(f"{x:08x}" for x in arr[:8])
    return (f"{op}: {len(bad)}/{len(want)} vectors differ; first #{i} limb {w}: "
            f"p={list(want['p'][i][:4])} a=[{hx(want['a'][i])}] b=[{hx(want['b'][i])}] "
            f"want=[{hx(want['out'][i][max(0, w - 2):])}] got=[{hx(got['out'][i][max(0, w - 2):])}]")
