// -*- mode: C++; c-file-style: "cc-mode" -*-
// Device testbenches of rv32_soc.v: the analogue of a user's C++ main loop (clock, reset,
// memory image), run inside the cycle-loop kernel.  Included into the emitted
// Vrv32_soc__cuda.cu through VLCU_TESTBENCH_INCLUDE.
#ifndef RV32_TB_H_
#define RV32_TB_H_

#define RV32_GEN_FN VLCU_D
#include "rv32_program.h"

// Random program image from the per-instance seed, synchronous clock, resetn low for the first
// tbArg[0] time units (default 16 = 8 clock cycles)
struct TbRv32Rand final {
    static constexpr const char* name() { return "rv32_rand"; }
    template <class T_Design>
    VLCU_D static void begin(typename T_Design::Root& top, VlcuContext&, const VlcuRunDev& r,
                             uint64_t, uint64_t globalInst) {
        top.clk = 0;
        top.resetn = 0;
        // RV32_TBF_SAME_PROGRAM (tb_flags bit 1): one program image for all instances, so that
        // they differ only in their --x-initial unique reset values (BASELINE config 3's setup)
        rv32_gen_program(top.rv32_soc__DOT__mem, (r.tbFlags & 2U) ? r.tbSeed : (r.tbSeed ^ globalInst));
    }
    template <class T_Design>
    VLCU_D static void drive(typename T_Design::Root& top, VlcuContext& ctx, const VlcuRunDev& r,
                             uint64_t, uint64_t e) {
        ctx.timeInc(1);
        top.clk = !top.clk;
        const uint64_t resetTime = r.tbArg[0] ? r.tbArg[0] : 16;
        // RV32_TBF_RESET_PER_RUN (tb_flags bit 0): reset is held for the first resetTime evals
        // of every vlcu_run (a fuzz campaign: new program, reset, run); otherwise once, by time
        top.resetn = (r.tbFlags & 1U) ? ((e - r.evalBase) >= resetTime) : (ctx.time() > resetTime);
    }
};
// Same clocking, memory image left as loaded by the host (vlcu_set_member "rv32_soc__DOT__mem")
struct TbRv32Host final {
    static constexpr const char* name() { return "rv32_host"; }
    template <class T_Design>
    VLCU_D static void begin(typename T_Design::Root& top, VlcuContext&, const VlcuRunDev&, uint64_t,
                             uint64_t) {
        top.clk = 0;
        top.resetn = 0;
    }
    template <class T_Design>
    VLCU_D static void drive(typename T_Design::Root& top, VlcuContext& ctx, const VlcuRunDev& r,
                             uint64_t inst, uint64_t e) {
        TbRv32Rand::drive<T_Design>(top, ctx, r, inst, e);
    }
};
// A fuzz campaign as the host drives it: ONE 64-bit program seed per instance per round, supplied
// by the host (vlcu_run_desc_t.tb_data: uint64 [instances of this shard]); the program image is
// made on the device ("stimulus generated on device"), so a round moves 8 bytes per
// instance over PCIe instead of a 4 KB image.  Clocking as rv32_rand.
struct TbRv32Seeds final {
    static constexpr const char* name() { return "rv32_seeds"; }
    template <class T_Design>
    VLCU_D static void begin(typename T_Design::Root& top, VlcuContext&, const VlcuRunDev& r,
                             uint64_t inst, uint64_t) {
        top.clk = 0;
        top.resetn = 0;
        rv32_gen_program(top.rv32_soc__DOT__mem, static_cast<const uint64_t*>(r.tbPtr[0])[inst]);
    }
    template <class T_Design>
    VLCU_D static void drive(typename T_Design::Root& top, VlcuContext& ctx, const VlcuRunDev& r,
                             uint64_t inst, uint64_t e) {
        TbRv32Rand::drive<T_Design>(top, ctx, r, inst, e);
    }
};
// Registered with the design (VLCU_REGISTER_DESIGN at the end of the emitted model)
#define VLCU_TESTBENCHES , TbRv32Rand, TbRv32Host, TbRv32Seeds

#endif  // guard
