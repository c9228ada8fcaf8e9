// -*- mode: C++; c-file-style: "cc-mode" -*-
// DESCRIPTION: rv32_soc.v (PicoRV32-class RV32I core + private 4 KB memory) as a batched
// CUDA model, in the form V3EmitCuda writes: root struct, SoA layout, load/store, the
// scheduled functions (Vrv32_soc__body.inc), host tables and registration.
// This is the workload of BASELINE.json config 2 ("RV32I core, 65,536 independent
// instances each running a random-seeded synthetic program").

#include "verilated_cuda_launch.h"

#include <vector>

struct Vrv32_soc___024root final {
    // DESIGN SPECIFIC STATE (ordered by storage size)
    IData /*31:0*/ mem_addr;
    IData /*31:0*/ mem_wdata;
    IData /*31:0*/ mem_rdata;
    IData /*31:0*/ retired;
    IData /*31:0*/ rv32_soc__DOT__pc;
    IData /*31:0*/ rv32_soc__DOT__insn;
    IData /*31:0*/ rv32_soc__DOT__ld_addr;
    IData /*31:0*/ rv32_soc__DOT__instret;
    VlcuSharedArray<IData /*31:0*/, 32> rv32_soc__DOT__regs;  // run-time indexed: storage in Shared
    // 4 KB per instance: stays in HBM, instance-major
    VlcuGlobalArray<IData /*31:0*/, 1024> rv32_soc__DOT__mem;
    CData /*0:0*/ clk;
    CData /*0:0*/ resetn;
    CData /*0:0*/ trap;
    CData /*0:0*/ mem_valid;
    CData /*0:0*/ mem_instr;
    CData /*3:0*/ mem_wstrb;
    CData /*1:0*/ rv32_soc__DOT__state;
    CData /*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    // Scheduling temporaries (never stored)
    VlUnpacked<QData /*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData /*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData /*63:0*/, 1> __VnbaTriggered;
    IData __VstlIterCount, __VactIterCount, __VnbaIterCount;
    CData __VstlFirstIteration, __VstlPhaseResult, __VstlExecute;
    CData __VactFirstIteration, __VactPhaseResult, __VactExecute;
    CData __VnbaFirstIteration, __VnbaPhaseResult, __VnbaExecute;
};

struct Vrv32_soc__cuda final {
    typedef Vrv32_soc___024root Root;
    enum : uint32_t {
        OFF_mem_addr = 0,
        OFF_mem_wdata = 4,
        OFF_mem_rdata = 8,
        OFF_retired = 12,
        OFF_pc = 16,
        OFF_insn = 20,
        OFF_ld_addr = 24,
        OFF_instret = 28,
        OFF_regs = 32,
        OFF_mem = 160,
        OFF_clk = 160 + 4096,
        OFF_resetn = OFF_clk + 1,
        OFF_trap = OFF_clk + 2,
        OFF_mem_valid = OFF_clk + 3,
        OFF_mem_instr = OFF_clk + 4,
        OFF_mem_wstrb = OFF_clk + 5,
        OFF_state = OFF_clk + 6,
        OFF_prev_clk = OFF_clk + 7,
        kStateBytes = OFF_clk + 8,  // S_persist = 4264 B (4096 of them the private memory)
        kInWords = 2,
        kOutWords = 8,
        kInBytes = 2,
        kOutBytes = 20,
        // Persistent bytes the triggered leaves touch per eval (SURVEY 8d "touched only"):
        // read pc,insn,ld_addr,instret,state,prev_clk,clk,resetn + 2 regs + 1 mem word;
        // write pc,insn|ld_addr,instret,state,prev_clk + 1 reg or 1 mem word + the 8 outputs
        kTouchedReadBytes = 4 * 4 + 4 + 2 * 4 + 4,
        kTouchedWriteBytes = 3 * 4 + 2 + 4 + 20,
    };

    struct Locals final {};
    struct Shared final {
        IData rv32_soc__DOT__regs[32][VLCU_BLOCK];  // 8 KB per CTA, bank = lane: conflict free
    };
    VLCU_D static void bind(Root& s, Locals&, Shared& sh, const VlcuState& st, uint64_t i) {
        s.rv32_soc__DOT__regs.bind(sh.rv32_soc__DOT__regs, threadIdx.x);
        s.rv32_soc__DOT__mem.bind(st, OFF_mem, i);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.mem_addr = st.ld<IData>(OFF_mem_addr, i);
        s.mem_wdata = st.ld<IData>(OFF_mem_wdata, i);
        s.mem_rdata = st.ld<IData>(OFF_mem_rdata, i);
        s.retired = st.ld<IData>(OFF_retired, i);
        s.rv32_soc__DOT__pc = st.ld<IData>(OFF_pc, i);
        s.rv32_soc__DOT__insn = st.ld<IData>(OFF_insn, i);
        s.rv32_soc__DOT__ld_addr = st.ld<IData>(OFF_ld_addr, i);
        s.rv32_soc__DOT__instret = st.ld<IData>(OFF_instret, i);
        for (uint32_t e = 0; e < 32; ++e) s.rv32_soc__DOT__regs[e] = st.ldE<IData>(OFF_regs, e, i);
        s.clk = st.ld<CData>(OFF_clk, i);
        s.resetn = st.ld<CData>(OFF_resetn, i);
        s.trap = st.ld<CData>(OFF_trap, i);
        s.mem_valid = st.ld<CData>(OFF_mem_valid, i);
        s.mem_instr = st.ld<CData>(OFF_mem_instr, i);
        s.mem_wstrb = st.ld<CData>(OFF_mem_wstrb, i);
        s.rv32_soc__DOT__state = st.ld<CData>(OFF_state, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<CData>(OFF_prev_clk, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<IData>(OFF_mem_addr, i, s.mem_addr);
        st.st<IData>(OFF_mem_wdata, i, s.mem_wdata);
        st.st<IData>(OFF_mem_rdata, i, s.mem_rdata);
        st.st<IData>(OFF_retired, i, s.retired);
        st.st<IData>(OFF_pc, i, s.rv32_soc__DOT__pc);
        st.st<IData>(OFF_insn, i, s.rv32_soc__DOT__insn);
        st.st<IData>(OFF_ld_addr, i, s.rv32_soc__DOT__ld_addr);
        st.st<IData>(OFF_instret, i, s.rv32_soc__DOT__instret);
        for (uint32_t e = 0; e < 32; ++e) st.stE<IData>(OFF_regs, e, i, s.rv32_soc__DOT__regs[e]);
        st.st<CData>(OFF_clk, i, s.clk);
        st.st<CData>(OFF_resetn, i, s.resetn);
        st.st<CData>(OFF_trap, i, s.trap);
        st.st<CData>(OFF_mem_valid, i, s.mem_valid);
        st.st<CData>(OFF_mem_instr, i, s.mem_instr);
        st.st<CData>(OFF_mem_wstrb, i, s.mem_wstrb);
        st.st<CData>(OFF_state, i, s.rv32_soc__DOT__state);
        st.st<CData>(OFF_prev_clk, i, s.__Vtrigprevexpr___TOP__clk__0);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<CData>(rowp[0 * stride + i] & 1U);
        s.resetn = static_cast<CData>(rowp[1 * stride + i] & 1U);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.trap;
        rowp[1 * stride + i] = s.mem_valid;
        rowp[2 * stride + i] = s.mem_instr;
        rowp[3 * stride + i] = s.mem_addr;
        rowp[4 * stride + i] = s.mem_wdata;
        rowp[5 * stride + i] = s.mem_wstrb;
        rowp[6 * stride + i] = s.mem_rdata;
        rowp[7 * stride + i] = s.retired;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        sig = vlcu_sig_mix(sig, (static_cast<uint32_t>(s.trap) << 6) | (static_cast<uint32_t>(s.mem_valid) << 5)
                                    | (static_cast<uint32_t>(s.mem_instr) << 4) | s.mem_wstrb);
        sig = vlcu_sig_mix(sig, s.mem_addr);
        sig = vlcu_sig_mix(sig, s.mem_wdata);
        sig = vlcu_sig_mix(sig, s.mem_rdata);
        sig = vlcu_sig_mix(sig, s.retired);
        return sig;
    }

    // _ctor_var_reset: --x-initial unique (V3EmitCFunc.cpp:600-660).  The 4 KB memory array is
    // reset word by word like any unpacked array (V3EmitCFunc.cpp:560-590: one salt, element
    // index folded into the scope hash is NOT done by the reference -- every element of an
    // array draws from the same (scope, salt) pair, so all elements start equal).
    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        auto& vlSelfRef = *vlSelf;
        const uint64_t __VscopeHash = 1586713752769493771ULL;  // VL_MURMUR64_HASH("TOP")
        vlSelfRef.clk = VL_SCOPED_RAND_RESET_I(vlCtx, 1, __VscopeHash, 16707436170211756652ull);
        vlSelfRef.resetn = VL_SCOPED_RAND_RESET_I(vlCtx, 1, __VscopeHash, 8624841754543469506ull);
        vlSelfRef.trap = VL_SCOPED_RAND_RESET_I(vlCtx, 1, __VscopeHash, 18214934560881419504ull);
        vlSelfRef.mem_valid = VL_SCOPED_RAND_RESET_I(vlCtx, 1, __VscopeHash, 1551974250180885553ull);
        vlSelfRef.mem_instr = VL_SCOPED_RAND_RESET_I(vlCtx, 1, __VscopeHash, 1406578376079151150ull);
        vlSelfRef.mem_addr = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 326597072690670135ull);
        vlSelfRef.mem_wdata = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 5431754401481461448ull);
        vlSelfRef.mem_wstrb = VL_SCOPED_RAND_RESET_I(vlCtx, 4, __VscopeHash, 8859681292774497410ull);
        vlSelfRef.mem_rdata = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 9659133473039683418ull);
        vlSelfRef.retired = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 9241213606999339685ull);
        for (int __Vi0 = 0; __Vi0 < 1024; ++__Vi0) {
            vlSelfRef.rv32_soc__DOT__mem[__Vi0]
                = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 157536494676281617ull);
        }
        for (int __Vi0 = 0; __Vi0 < 32; ++__Vi0) {
            vlSelfRef.rv32_soc__DOT__regs[__Vi0]
                = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 11523106369459488163ull);
        }
        vlSelfRef.rv32_soc__DOT__pc = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 11841601083557326370ull);
        vlSelfRef.rv32_soc__DOT__insn = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 6790827016312138654ull);
        vlSelfRef.rv32_soc__DOT__ld_addr = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 18411059245981549376ull);
        vlSelfRef.rv32_soc__DOT__instret = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 1322355925642281443ull);
        vlSelfRef.rv32_soc__DOT__state = VL_SCOPED_RAND_RESET_I(vlCtx, 2, __VscopeHash, 769792777424239494ull);
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = 0;
    }

#define VLB_FN VLCU_D static
#define VLB_ROOT Root
#define VLB_CTXPARAM , VlcuContext& vlCtx
#define VLB_FATAL_DIDNOTCONVERGE(region) VLCU_FATAL(vlCtx, 13, VLCU_FATAL_DIDNOTCONVERGE)
#include "Vrv32_soc__body.inc"
#undef VLB_FN
#undef VLB_ROOT
#undef VLB_CTXPARAM
#undef VLB_FATAL_DIDNOTCONVERGE

    VLCU_D static void eval_static(Root* vlSelf, VlcuContext&) { rv32_eval_static(*vlSelf); }
    VLCU_D static void eval_initial(Root*, VlcuContext&) {}
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { rv32_eval_settle(*vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { rv32_eval(*vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root*, VlcuContext&) {}

    static const VlcuDesignInfo& info();
};

// Device testbench: random program image from the per-instance seed, synchronous clock,
// resetn low for the first tbArg[0] time units (default 16 = 8 clock cycles)
struct TbRv32Rand final {
    static constexpr const char* name() { return "rv32_rand"; }
    template <class T_Design>
    VLCU_D static void begin(typename T_Design::Root& top, VlcuContext&, const VlcuRunDev& r,
                             uint64_t, uint64_t globalInst) {
        top.clk = 0;
        top.resetn = 0;
        T_Design::rv32_gen_program(top.rv32_soc__DOT__mem, r.tbSeed ^ globalInst);
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

const VlcuDesignInfo& Vrv32_soc__cuda::info() {
    typedef Vrv32_soc__cuda T;
    static const VlcuMemberDesc members[] = {
        {"clk", "rv32_soc.clk", VLCU_MEMBER_IN, 1, 1, 1, 1, 0, 0, T::OFF_clk},
        {"resetn", "rv32_soc.resetn", VLCU_MEMBER_IN, 1, 2, 1, 1, 0, 0, T::OFF_resetn},
        {"trap", "rv32_soc.trap", VLCU_MEMBER_OUT, 1, 3, 1, 1, 0, 0, T::OFF_trap},
        {"mem_valid", "rv32_soc.mem_valid", VLCU_MEMBER_OUT, 1, 4, 1, 1, 0, 0, T::OFF_mem_valid},
        {"mem_instr", "rv32_soc.mem_instr", VLCU_MEMBER_OUT, 1, 5, 1, 1, 0, 0, T::OFF_mem_instr},
        {"mem_addr", "rv32_soc.mem_addr", VLCU_MEMBER_OUT, 4, 6, 1, 32, 31, 0, T::OFF_mem_addr},
        {"mem_wdata", "rv32_soc.mem_wdata", VLCU_MEMBER_OUT, 4, 7, 1, 32, 31, 0, T::OFF_mem_wdata},
        {"mem_wstrb", "rv32_soc.mem_wstrb", VLCU_MEMBER_OUT, 1, 8, 1, 4, 3, 0, T::OFF_mem_wstrb},
        {"mem_rdata", "rv32_soc.mem_rdata", VLCU_MEMBER_OUT, 4, 9, 1, 32, 31, 0, T::OFF_mem_rdata},
        {"retired", "rv32_soc.retired", VLCU_MEMBER_OUT, 4, 10, 1, 32, 31, 0, T::OFF_retired},
        {"rv32_soc__DOT__pc", "rv32_soc.pc", VLCU_MEMBER_STATE, 4, 11, 1, 32, 31, 0, T::OFF_pc},
        {"rv32_soc__DOT__insn", "rv32_soc.insn", VLCU_MEMBER_STATE, 4, 12, 1, 32, 31, 0, T::OFF_insn},
        {"rv32_soc__DOT__ld_addr", "rv32_soc.ld_addr", VLCU_MEMBER_STATE, 4, 13, 1, 32, 31, 0, T::OFF_ld_addr},
        {"rv32_soc__DOT__instret", "rv32_soc.instret", VLCU_MEMBER_STATE, 4, 14, 1, 32, 31, 0, T::OFF_instret},
        {"rv32_soc__DOT__state", "rv32_soc.state", VLCU_MEMBER_STATE, 1, 15, 1, 2, 1, 0, T::OFF_state},
        {"rv32_soc__DOT__regs", "rv32_soc.regs", VLCU_MEMBER_STATE, 4, 0, 32, 32, 31, 0, T::OFF_regs},
        {"rv32_soc__DOT__mem", "rv32_soc.mem", VLCU_MEMBER_MEMORY, 4, 0, 1024, 32, 31, 0, T::OFF_mem},
        {"__Vtrigprevexpr___TOP__clk__0", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_prev_clk},
    };
    // trace_init: ports at the root scope, the same signals plus the internal registers inside
    // `rv32_soc`; regs[] (32 entries <= --trace-max-array) element by element, mem[] not traced
    static std::vector<VlcuTraceDecl> decls;
    if (decls.empty()) {
        const auto ports = [&]() {
            uint32_t code = 1;
            for (int mi = 0; mi < 10; ++mi) {
                const VlcuMemberDesc& md = members[mi];
                decls.push_back({VLCU_TRACE_SIG, static_cast<uint8_t>(md.width > 1), md.name, code, mi,
                                 -1, md.msb, md.lsb});
                code += (md.width + 31) / 32;
            }
            return code;
        };
        ports();
        decls.push_back({VLCU_TRACE_PUSH, 0, "rv32_soc", 0, -1, -1, 0, 0});
        uint32_t code = ports();
        const char* const names[] = {"pc", "insn", "ld_addr", "instret", "state"};
        for (int k = 0; k < 5; ++k) {
            const VlcuMemberDesc& md = members[10 + k];
            decls.push_back({VLCU_TRACE_SIG, 1, names[k], code++, 10 + k, -1, md.msb, md.lsb});
        }
        for (int e = 0; e < 32; ++e) decls.push_back({VLCU_TRACE_SIG, 1, "regs", code++, 15, e, 31, 0});
        decls.push_back({VLCU_TRACE_POP, 0, nullptr, 0, -1, -1, 0, 0});
    }
    static const VlcuDesignInfo s = {
        "rv32i", "rv32_soc", members, sizeof(members) / sizeof(members[0]),
        kStateBytes, kInWords, kOutWords, kInBytes, kOutBytes, kTouchedReadBytes, kTouchedWriteBytes,
        nullptr, 0, 0, nullptr, 0, nullptr, 0, decls.data(), static_cast<uint32_t>(decls.size())};
    return s;
}

VLCU_REGISTER_DESIGN(Vrv32_soc__cuda, TbRv32Rand, TbRv32Host);
