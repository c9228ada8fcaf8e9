// -*- mode: C++; c-file-style: "cc-mode" -*-
// Device-side testbench for examples/make_tracing_c: the body of the
// "while (!contextp->gotFinish())" loop of examples/make_tracing_c/sim_main.cpp:52-106,
// evaluated by every instance in lock-step.
//
// tbFlags bit 0: per-instance stimulus.  Instance g (global index) offsets the initial
//   in_small / in_quad / in_wide by draws from VlRNG(tbSeed ^ g)  (SURVEY 8d, config 1 batched)
// tbArg[0]: reset period P.  0 = the original "reset while 1 < time < 10"; otherwise reset is
//   asserted while 1 < (time mod P) < 10, which keeps count_c below the design's assertion
//   threshold on long runs (BASELINE config 1: "1M cycles").
#ifndef VLCU_SIM_MAIN_TB_H_
#define VLCU_SIM_MAIN_TB_H_

struct TbSimMain final {
    static constexpr const char* name() { return "sim_main"; }

    template <class T_Design>
    VLCU_D static void begin(typename T_Design::Root& top, VlcuContext&, const VlcuRunDev& r,
                             uint64_t, uint64_t globalInst) {
        // sim_main.cpp:53-61  initial inputs
        top.reset_l = !0;
        top.clk = 0;
        top.in_small = 1;
        top.in_quad = 0x1234;
        top.in_wide[0] = 0x11111111;
        top.in_wide[1] = 0x22222222;
        top.in_wide[2] = 0x3;
        if (r.tbFlags & 1U) {
            VlRNG rng{r.tbSeed ^ globalInst};
            const uint64_t r0 = rng.rand64();
            const uint64_t r1 = rng.rand64();
            const uint64_t r2 = rng.rand64();
            const uint64_t r3 = rng.rand64();
            top.in_small = static_cast<CData>((top.in_small + r0) & 3U);
            top.in_quad = (top.in_quad + r1) & 0x000000ffffffffffULL;
            VlWide<3> off;
            off[0] = static_cast<EData>(r2);
            off[1] = static_cast<EData>(r2 >> 32);
            off[2] = static_cast<EData>(r3) & 0x3fU;
            VlWide<3> sum;
            VL_ADD_W(3, sum, top.in_wide, off);
            top.in_wide[0] = sum[0];
            top.in_wide[1] = sum[1];
            top.in_wide[2] = sum[2] & 0x3fU;
        }
    }

    template <class T_Design>
    VLCU_D static void drive(typename T_Design::Root& top, VlcuContext& ctx, const VlcuRunDev& r,
                             uint64_t, uint64_t) {
        ctx.timeInc(1);  // sim_main.cpp:72
        top.clk = !top.clk;  // :78
        if (!top.clk) {  // :85-93
            const uint64_t period = r.tbArg[0];
            const uint64_t t = period ? (ctx.time() % period) : ctx.time();
            if (t > 1 && t < 10) {
                top.reset_l = !1;  // Assert reset
            } else {
                top.reset_l = !0;  // Deassert reset
            }
            top.in_quad = (top.in_quad + 0x12) & 0x000000ffffffffffULL;  // :92 (kept clean)
        }
    }
};

// Registered with the design (VLCU_REGISTER_DESIGN at the end of the emitted model)
#define VLCU_TESTBENCHES , TbSimMain

#endif  // guard
