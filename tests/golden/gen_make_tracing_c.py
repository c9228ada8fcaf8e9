#!/usr/bin/env python3
"""Generates tests/golden/make_tracing_c_ports.json.

An INDEPENDENT behavioural model of examples/make_tracing_c/{top.v,sub.v} driven by the
stimulus of examples/make_tracing_c/sim_main.cpp:52-106, written straight from the Verilog
semantics (continuous assigns + one always_ff), not from the hand-scheduled C++/CUDA models.
It pins the restated oracle model (oracle/models/make_tracing_c) and, through it, the GPU path.
The expected console line "*-* All Finished *-*" at the 4th counted clock is the check the
reference's own example Makefile makes (examples/make_tracing_c/sub.v:25-27).
"""
import json
import os

M40 = (1 << 40) - 1
M70 = (1 << 70) - 1


def run(n_evals, in_small=1, in_quad=0x1234, in_wide=(0x3 << 64) | (0x22222222 << 32) | 0x11111111,
        count_c=0xdeadbeef, reset_period=0):
    clk, reset_l, t = 0, 1, 0
    prev_clk = None
    rows, finish_times = [], []
    for _ in range(n_evals):
        t += 1
        clk ^= 1
        if not clk:
            tt = t % reset_period if reset_period else t
            reset_l = 0 if 1 < tt < 10 else 1
            in_quad = (in_quad + 0x12) & M40
        # eval(): the first eval only initialises edge detection
        posedge = prev_clk is not None and clk == 1 and prev_clk == 0
        prev_clk = clk
        if posedge:
            old = count_c
            if not reset_l:
                count_c = 0
            else:
                count_c = (old + 1) & 0xffffffff
                if old >= 3:
                    finish_times.append(t)
        if reset_l:
            o_small, o_quad, o_wide = (in_small + 1) & 3, (in_quad + 1) & M40, (in_wide + 1) & M70
        else:
            o_small = o_quad = o_wide = 0
        rows.append([o_small, o_quad & 0xffffffff, o_quad >> 32, o_wide & 0xffffffff,
                     (o_wide >> 32) & 0xffffffff, o_wide >> 64, count_c])
    return rows, finish_times


if __name__ == "__main__":
    rows, fin = run(64)
    rows_p, fin_p = run(600, reset_period=128)
    out = {"default": {"n_evals": 64, "rows": rows, "finish_times": fin},
           "period128": {"n_evals": 600, "rows_last": rows_p[-8:], "finish_count": len(fin_p),
                         "first_finish": fin_p[0]}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "make_tracing_c_ports.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, "first finish at", fin[0])
