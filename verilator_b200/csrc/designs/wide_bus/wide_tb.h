// -*- mode: C++; c-file-style: "cc-mode" -*-
// Device testbench of wide_bus.v: synchronous clock, reset for the first tbArg[0] (default 10)
// time units, fresh random inputs on every falling edge.  Included into the emitted
// Vwide_bus__cuda.cu through VLCU_TESTBENCH_INCLUDE.
#ifndef WIDE_TB_H_
#define WIDE_TB_H_

#define WIDE_STIM_FN VLCU_D
#include "wide_stim.h"

struct TbWideRand final {
    static constexpr const char* name() { return "wide_rand"; }
    template <class T_Design>
    VLCU_D static void begin(typename T_Design::Root& top, VlcuContext&, const VlcuRunDev&, uint64_t,
                             uint64_t) {
        top.clk = 0;
        top.rst_n = 0;
        top.in_valid = 0;
        top.in_op = 0;
        top.in_shift = 0;
        for (int w = 0; w < 8; ++w) top.in_data[w] = 0;
    }
    template <class T_Design>
    VLCU_D static void drive(typename T_Design::Root& top, VlcuContext& ctx, const VlcuRunDev& r,
                             uint64_t, uint64_t e) {
        ctx.timeInc(1);
        top.clk = !top.clk;
        if (!top.clk) {
            const uint64_t resetTime = r.tbArg[0] ? r.tbArg[0] : 10;
            top.rst_n = ctx.time() > resetTime;
            // ctx.randSeed() = seed_base ^ global instance index: the per-instance stream id
            wide_drive_inputs(top, r.tbSeed ^ ctx.randSeed(), e);
        }
    }
};

// Registered with the design (VLCU_REGISTER_DESIGN at the end of the emitted model)
#define VLCU_TESTBENCHES , TbWideRand

#endif  // guard
