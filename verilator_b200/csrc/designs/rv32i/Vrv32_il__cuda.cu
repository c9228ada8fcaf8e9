// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vrv32_il"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vrv32_il___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ resetn;
    CData/*0:0*/ trap;
    CData/*0:0*/ mem_valid;
    CData/*0:0*/ mem_instr;
    IData/*31:0*/ mem_addr;
    IData/*31:0*/ mem_wdata;
    CData/*3:0*/ mem_wstrb;
    IData/*31:0*/ mem_rdata;
    IData/*31:0*/ retired;
    CData/*1:0*/ rv32_soc__DOT__state;
    CData/*0:0*/ rv32_soc__DOT__is_load;
    CData/*0:0*/ rv32_soc__DOT__is_store;
    CData/*0:0*/ rv32_soc__DOT__is_alur;
    CData/*0:0*/ rv32_soc__DOT__ls_misaligned;
    CData/*3:0*/ rv32_soc__DOT__st_strb;
    CData/*0:0*/ rv32_soc__DOT__pc_bad;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_0;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ rv32_soc__DOT__pc;
    IData/*31:0*/ rv32_soc__DOT__insn;
    IData/*31:0*/ rv32_soc__DOT__ld_addr;
    IData/*31:0*/ rv32_soc__DOT__instret;
    IData/*31:0*/ rv32_soc__DOT__imm_i;
    IData/*31:0*/ rv32_soc__DOT__imm_u;
    IData/*31:0*/ rv32_soc__DOT__rs1_val;
    IData/*31:0*/ rv32_soc__DOT__rs2_val;
    IData/*31:0*/ rv32_soc__DOT__alu_b;
    IData/*31:0*/ rv32_soc__DOT__ls_addr;
    IData/*31:0*/ rv32_soc__DOT__st_data;
    IData/*31:0*/ rv32_soc__DOT__ld_shift;
    IData/*31:0*/ __VactIterCount;
    VlcuStateArray<IData/*31:0*/, 1024> rv32_soc__DOT__mem;  // stays in HBM, [element][instance]: instance-uniform indices coalesce
    VlcuSharedArray<IData/*31:0*/, 32> rv32_soc__DOT__regs;  // run-time indexed: shared memory [index][lane]
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vrv32_il__cuda final {
    typedef Vrv32_il___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 4316,
        OFF_resetn = 4317,
        OFF_trap = 4318,
        OFF_mem_valid = 4319,
        OFF_mem_instr = 4320,
        OFF_mem_addr = 24,
        OFF_mem_wdata = 28,
        OFF_mem_wstrb = 4321,
        OFF_mem_rdata = 32,
        OFF_retired = 36,
        OFF_rv32_soc__DOT__state = 4322,
        OFF_rv32_soc__DOT__is_load = 4323,
        OFF_rv32_soc__DOT__is_store = 4324,
        OFF_rv32_soc__DOT__is_alur = 4325,
        OFF_rv32_soc__DOT__ls_misaligned = 4326,
        OFF_rv32_soc__DOT__st_strb = 4327,
        OFF_rv32_soc__DOT__pc_bad = 4328,
        OFF___VdfgRegularize_h6e95ff9d_0_0 = 4329,
        OFF___VstlFirstIteration = 4330,
        OFF___VstlPhaseResult = 4331,
        OFF___Vtrigprevexpr___TOP__clk__0 = 4332,
        OFF___VactPhaseResult = 4333,
        OFF___VnbaPhaseResult = 4334,
        OFF_rv32_soc__DOT__pc = 40,
        OFF_rv32_soc__DOT__insn = 44,
        OFF_rv32_soc__DOT__ld_addr = 48,
        OFF_rv32_soc__DOT__instret = 52,
        OFF_rv32_soc__DOT__imm_i = 56,
        OFF_rv32_soc__DOT__imm_u = 60,
        OFF_rv32_soc__DOT__rs1_val = 64,
        OFF_rv32_soc__DOT__rs2_val = 68,
        OFF_rv32_soc__DOT__alu_b = 72,
        OFF_rv32_soc__DOT__ls_addr = 76,
        OFF_rv32_soc__DOT__st_data = 80,
        OFF_rv32_soc__DOT__ld_shift = 84,
        OFF___VactIterCount = 88,
        OFF_rv32_soc__DOT__mem = 92,
        OFF_rv32_soc__DOT__regs = 4188,
        OFF___VstlTriggered = 0,
        OFF___VactTriggered = 8,
        OFF___VnbaTriggered = 16,
        kStateBytes = 4336,  // S_persist
        kTouchedReadBytes = 126,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 139,  // S_state_wr
    };
    struct Locals final {
    };
    struct Shared final {
        IData/*31:0*/ rv32_soc__DOT__regs[32][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.rv32_soc__DOT__mem.bind(st, OFF_rv32_soc__DOT__mem, i);
        s.rv32_soc__DOT__regs.bind(sh.rv32_soc__DOT__regs, threadIdx.x);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.resetn = st.ld<VlcuMem8>(OFF_resetn, i);
        s.trap = st.ld<VlcuMem8>(OFF_trap, i);
        s.mem_valid = st.ld<VlcuMem8>(OFF_mem_valid, i);
        s.mem_instr = st.ld<VlcuMem8>(OFF_mem_instr, i);
        s.mem_addr = st.ld<IData>(OFF_mem_addr, i);
        s.mem_wdata = st.ld<IData>(OFF_mem_wdata, i);
        s.mem_wstrb = st.ld<VlcuMem8>(OFF_mem_wstrb, i);
        s.mem_rdata = st.ld<IData>(OFF_mem_rdata, i);
        s.retired = st.ld<IData>(OFF_retired, i);
        s.rv32_soc__DOT__state = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__state, i);
        s.rv32_soc__DOT__is_load = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__is_load, i);
        s.rv32_soc__DOT__is_store = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__is_store, i);
        s.rv32_soc__DOT__is_alur = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__is_alur, i);
        s.rv32_soc__DOT__ls_misaligned = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__ls_misaligned, i);
        s.rv32_soc__DOT__st_strb = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__st_strb, i);
        s.rv32_soc__DOT__pc_bad = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__pc_bad, i);
        s.__VdfgRegularize_h6e95ff9d_0_0 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_0, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.rv32_soc__DOT__pc = st.ld<IData>(OFF_rv32_soc__DOT__pc, i);
        s.rv32_soc__DOT__insn = st.ld<IData>(OFF_rv32_soc__DOT__insn, i);
        s.rv32_soc__DOT__ld_addr = st.ld<IData>(OFF_rv32_soc__DOT__ld_addr, i);
        s.rv32_soc__DOT__instret = st.ld<IData>(OFF_rv32_soc__DOT__instret, i);
        s.rv32_soc__DOT__imm_i = st.ld<IData>(OFF_rv32_soc__DOT__imm_i, i);
        s.rv32_soc__DOT__imm_u = st.ld<IData>(OFF_rv32_soc__DOT__imm_u, i);
        s.rv32_soc__DOT__rs1_val = st.ld<IData>(OFF_rv32_soc__DOT__rs1_val, i);
        s.rv32_soc__DOT__rs2_val = st.ld<IData>(OFF_rv32_soc__DOT__rs2_val, i);
        s.rv32_soc__DOT__alu_b = st.ld<IData>(OFF_rv32_soc__DOT__alu_b, i);
        s.rv32_soc__DOT__ls_addr = st.ld<IData>(OFF_rv32_soc__DOT__ls_addr, i);
        s.rv32_soc__DOT__st_data = st.ld<IData>(OFF_rv32_soc__DOT__st_data, i);
        s.rv32_soc__DOT__ld_shift = st.ld<IData>(OFF_rv32_soc__DOT__ld_shift, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 32; ++e) s.rv32_soc__DOT__regs[e] = st.ldE<IData>(OFF_rv32_soc__DOT__regs, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_resetn, i, s.resetn);
        st.st<VlcuMem8>(OFF_trap, i, s.trap);
        st.st<VlcuMem8>(OFF_mem_valid, i, s.mem_valid);
        st.st<VlcuMem8>(OFF_mem_instr, i, s.mem_instr);
        st.st<IData>(OFF_mem_addr, i, s.mem_addr);
        st.st<IData>(OFF_mem_wdata, i, s.mem_wdata);
        st.st<VlcuMem8>(OFF_mem_wstrb, i, s.mem_wstrb);
        st.st<IData>(OFF_mem_rdata, i, s.mem_rdata);
        st.st<IData>(OFF_retired, i, s.retired);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__state, i, s.rv32_soc__DOT__state);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__is_load, i, s.rv32_soc__DOT__is_load);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__is_store, i, s.rv32_soc__DOT__is_store);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__is_alur, i, s.rv32_soc__DOT__is_alur);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__ls_misaligned, i, s.rv32_soc__DOT__ls_misaligned);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__st_strb, i, s.rv32_soc__DOT__st_strb);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__pc_bad, i, s.rv32_soc__DOT__pc_bad);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_0, i, s.__VdfgRegularize_h6e95ff9d_0_0);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_rv32_soc__DOT__pc, i, s.rv32_soc__DOT__pc);
        st.st<IData>(OFF_rv32_soc__DOT__insn, i, s.rv32_soc__DOT__insn);
        st.st<IData>(OFF_rv32_soc__DOT__ld_addr, i, s.rv32_soc__DOT__ld_addr);
        st.st<IData>(OFF_rv32_soc__DOT__instret, i, s.rv32_soc__DOT__instret);
        st.st<IData>(OFF_rv32_soc__DOT__imm_i, i, s.rv32_soc__DOT__imm_i);
        st.st<IData>(OFF_rv32_soc__DOT__imm_u, i, s.rv32_soc__DOT__imm_u);
        st.st<IData>(OFF_rv32_soc__DOT__rs1_val, i, s.rv32_soc__DOT__rs1_val);
        st.st<IData>(OFF_rv32_soc__DOT__rs2_val, i, s.rv32_soc__DOT__rs2_val);
        st.st<IData>(OFF_rv32_soc__DOT__alu_b, i, s.rv32_soc__DOT__alu_b);
        st.st<IData>(OFF_rv32_soc__DOT__ls_addr, i, s.rv32_soc__DOT__ls_addr);
        st.st<IData>(OFF_rv32_soc__DOT__st_data, i, s.rv32_soc__DOT__st_data);
        st.st<IData>(OFF_rv32_soc__DOT__ld_shift, i, s.rv32_soc__DOT__ld_shift);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 32; ++e) st.stE<IData>(OFF_rv32_soc__DOT__regs, e, i, s.rv32_soc__DOT__regs[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.resetn = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
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
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.trap);
        acc = vlcu_sig_word(acc, s.mem_valid);
        acc = vlcu_sig_word(acc, s.mem_instr);
        acc = vlcu_sig_word(acc, s.mem_addr);
        acc = vlcu_sig_word(acc, s.mem_wdata);
        acc = vlcu_sig_word(acc, s.mem_wstrb);
        acc = vlcu_sig_word(acc, s.mem_rdata);
        acc = vlcu_sig_word(acc, s.retired);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 2, kOutWords = 8, kInBytes = 2, kOutBytes = 20, kTraceWords = 0 };

    VLCU_D static void Vrv32_il___024root___eval_static(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vrv32_il___024root___eval_initial(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrv32_il___024root___eval_final(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrv32_il___024root___eval_settle(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __VstlIterCount;
        // Body
        __VstlIterCount = 0U;
        vlSelfRef.__VstlFirstIteration = 1U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VstlIterCount)))) {
#ifdef VL_DEBUG
                Vrv32_il___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 16, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vrv32_il___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vrv32_il___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
        typedef bool VlcuFuncRet;
        // Locals
        IData/*31:0*/ n;
        // Body
        n = 0U;
        do {
            if (in[n]) {
                return (1U);
            }
            n = ((IData)(1U) + n);
        } while ((1U > n));
        return (0U);
    }

    VLCU_D static void Vrv32_il___024root___stl_sequent__TOP__0(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ rv32_soc__DOT__exec_ls;
        rv32_soc__DOT__exec_ls = 0;
        CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_1;
        __VdfgRegularize_h6e95ff9d_0_1 = 0;
        CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_2;
        __VdfgRegularize_h6e95ff9d_0_2 = 0;
        // Body
        vlSelfRef.trap = (3U == (IData)(vlSelfRef.rv32_soc__DOT__state));
        vlSelfRef.rv32_soc__DOT__imm_u = (0xfffff000U 
                                          & vlSelfRef.rv32_soc__DOT__insn);
        vlSelfRef.retired = vlSelfRef.rv32_soc__DOT__instret;
        vlSelfRef.rv32_soc__DOT__ld_shift = (vlSelfRef.rv32_soc__DOT__mem
                                             [(0x000003ffU 
                                               & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                                  >> 2U))] 
                                             >> (0x00000018U 
                                                 & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                                    << 3U)));
        vlSelfRef.rv32_soc__DOT__imm_i = (((- (IData)(
                                                      (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000001fU))) 
                                           << 0x0000000cU) 
                                          | (vlSelfRef.rv32_soc__DOT__insn 
                                             >> 0x00000014U));
        vlSelfRef.rv32_soc__DOT__is_alur = ((0x33U 
                                             == (0x0000007fU 
                                                 & vlSelfRef.rv32_soc__DOT__insn)) 
                                            & ((0U 
                                                == 
                                                (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x00000019U)) 
                                               | ((0x20U 
                                                   == 
                                                   (vlSelfRef.rv32_soc__DOT__insn 
                                                    >> 0x00000019U)) 
                                                  & ((0U 
                                                      == 
                                                      (7U 
                                                       & (vlSelfRef.rv32_soc__DOT__insn 
                                                          >> 0x0000000cU))) 
                                                     | (5U 
                                                        == 
                                                        (7U 
                                                         & (vlSelfRef.rv32_soc__DOT__insn 
                                                            >> 0x0000000cU)))))));
        vlSelfRef.rv32_soc__DOT__rs2_val = ((0U == 
                                             (0x0000001fU 
                                              & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x00000014U)))
                                             ? 0U : vlSelfRef.rv32_soc__DOT__regs
                                            [(0x0000001fU 
                                              & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x00000014U))]);
        vlSelfRef.rv32_soc__DOT__pc_bad = ((0U != (3U 
                                                   & vlSelfRef.rv32_soc__DOT__pc)) 
                                           | (0x00000400U 
                                              <= (vlSelfRef.rv32_soc__DOT__pc 
                                                  >> 2U)));
        vlSelfRef.rv32_soc__DOT__rs1_val = ((0U == 
                                             (0x0000001fU 
                                              & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000000fU)))
                                             ? 0U : vlSelfRef.rv32_soc__DOT__regs
                                            [(0x0000001fU 
                                              & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000000fU))]);
        __VdfgRegularize_h6e95ff9d_0_2 = ((0U == (7U 
                                                  & (vlSelfRef.rv32_soc__DOT__insn 
                                                     >> 0x0000000cU))) 
                                          | ((1U == 
                                              (7U & 
                                               (vlSelfRef.rv32_soc__DOT__insn 
                                                >> 0x0000000cU))) 
                                             | (2U 
                                                == 
                                                (7U 
                                                 & (vlSelfRef.rv32_soc__DOT__insn 
                                                    >> 0x0000000cU)))));
        vlSelfRef.rv32_soc__DOT__alu_b = ((IData)(vlSelfRef.rv32_soc__DOT__is_alur)
                                           ? vlSelfRef.rv32_soc__DOT__rs2_val
                                           : vlSelfRef.rv32_soc__DOT__imm_i);
        vlSelfRef.mem_instr = ((~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)) 
                               & (0U == (IData)(vlSelfRef.rv32_soc__DOT__state)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 = 
            (vlSelfRef.rv32_soc__DOT__rs1_val >= vlSelfRef.rv32_soc__DOT__rs2_val);
        vlSelfRef.rv32_soc__DOT__is_load = ((3U == 
                                             (0x0000007fU 
                                              & vlSelfRef.rv32_soc__DOT__insn)) 
                                            & ((IData)(__VdfgRegularize_h6e95ff9d_0_2) 
                                               | ((4U 
                                                   == 
                                                   (7U 
                                                    & (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000000cU))) 
                                                  | (5U 
                                                     == 
                                                     (7U 
                                                      & (vlSelfRef.rv32_soc__DOT__insn 
                                                         >> 0x0000000cU))))));
        vlSelfRef.rv32_soc__DOT__is_store = ((0x23U 
                                              == (0x0000007fU 
                                                  & vlSelfRef.rv32_soc__DOT__insn)) 
                                             & (IData)(__VdfgRegularize_h6e95ff9d_0_2));
        vlSelfRef.rv32_soc__DOT__ls_addr = (vlSelfRef.rv32_soc__DOT__rs1_val 
                                            + (((- (IData)(
                                                           (vlSelfRef.rv32_soc__DOT__insn 
                                                            >> 0x0000001fU))) 
                                                << 0x0000000cU) 
                                               | ((0x00000fe0U 
                                                   & (vlSelfRef.rv32_soc__DOT__insn 
                                                      >> 0x00000014U)) 
                                                  | (0x0000001fU 
                                                     & ((IData)(vlSelfRef.rv32_soc__DOT__is_store)
                                                         ? 
                                                        (vlSelfRef.rv32_soc__DOT__insn 
                                                         >> 7U)
                                                         : 
                                                        (vlSelfRef.rv32_soc__DOT__insn 
                                                         >> 0x00000014U))))));
        if ((0U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlSelfRef.rv32_soc__DOT__st_data = ((vlSelfRef.rv32_soc__DOT__rs2_val 
                                                 << 0x00000018U) 
                                                | ((0x00ff0000U 
                                                    & (vlSelfRef.rv32_soc__DOT__rs2_val 
                                                       << 0x00000010U)) 
                                                   | ((0x0000ff00U 
                                                       & (vlSelfRef.rv32_soc__DOT__rs2_val 
                                                          << 8U)) 
                                                      | (0x000000ffU 
                                                         & vlSelfRef.rv32_soc__DOT__rs2_val))));
            vlSelfRef.rv32_soc__DOT__st_strb = (0x0000000fU 
                                                & ((IData)(1U) 
                                                   << 
                                                   (3U 
                                                    & vlSelfRef.rv32_soc__DOT__ls_addr)));
        } else {
            vlSelfRef.rv32_soc__DOT__st_data = ((1U 
                                                 == 
                                                 (3U 
                                                  & (vlSelfRef.rv32_soc__DOT__insn 
                                                     >> 0x0000000cU)))
                                                 ? 
                                                ((vlSelfRef.rv32_soc__DOT__rs2_val 
                                                  << 0x00000010U) 
                                                 | (0x0000ffffU 
                                                    & vlSelfRef.rv32_soc__DOT__rs2_val))
                                                 : vlSelfRef.rv32_soc__DOT__rs2_val);
            vlSelfRef.rv32_soc__DOT__st_strb = (0x0000000fU 
                                                & (((IData)(3U) 
                                                    << 
                                                    (3U 
                                                     & vlSelfRef.rv32_soc__DOT__ls_addr)) 
                                                   | (- (IData)(
                                                                (1U 
                                                                 != 
                                                                 (3U 
                                                                  & (vlSelfRef.rv32_soc__DOT__insn 
                                                                     >> 0x0000000cU)))))));
        }
        vlSelfRef.rv32_soc__DOT__ls_misaligned = ((IData)(
                                                          ((0x00001000U 
                                                            == 
                                                            (0x00003000U 
                                                             & vlSelfRef.rv32_soc__DOT__insn)) 
                                                           & vlSelfRef.rv32_soc__DOT__ls_addr)) 
                                                  | (IData)(
                                                            ((0x00002000U 
                                                              == 
                                                              (0x00003000U 
                                                               & vlSelfRef.rv32_soc__DOT__insn)) 
                                                             & (0U 
                                                                != 
                                                                (3U 
                                                                 & vlSelfRef.rv32_soc__DOT__ls_addr)))));
        rv32_soc__DOT__exec_ls = ((1U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
                                  & ((~ ((0x00000400U 
                                          <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                              >> 2U)) 
                                         | (IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned))) 
                                     & ((IData)(vlSelfRef.rv32_soc__DOT__is_load) 
                                        | (IData)(vlSelfRef.rv32_soc__DOT__is_store))));
        vlSelfRef.mem_valid = ((IData)(vlSelfRef.mem_instr) 
                               | (IData)(rv32_soc__DOT__exec_ls));
        if (vlSelfRef.mem_instr) {
            vlSelfRef.mem_addr = vlSelfRef.rv32_soc__DOT__pc;
            vlSelfRef.mem_rdata = vlSelfRef.rv32_soc__DOT__mem
                [(0x000003ffU & (vlSelfRef.rv32_soc__DOT__pc 
                                 >> 2U))];
        } else {
            vlSelfRef.mem_addr = (0xfffffffcU & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                                 & (- (IData)((IData)(rv32_soc__DOT__exec_ls)))));
            vlSelfRef.mem_rdata = (((IData)(rv32_soc__DOT__exec_ls) 
                                    & (IData)(vlSelfRef.rv32_soc__DOT__is_load))
                                    ? vlSelfRef.rv32_soc__DOT__mem
                                   [(0x000003ffU & 
                                     (vlSelfRef.rv32_soc__DOT__ls_addr 
                                      >> 2U))] : 0U);
        }
        __VdfgRegularize_h6e95ff9d_0_1 = ((IData)(rv32_soc__DOT__exec_ls) 
                                          & (IData)(vlSelfRef.rv32_soc__DOT__is_store));
        vlSelfRef.mem_wdata = (vlSelfRef.rv32_soc__DOT__st_data 
                               & (- (IData)((IData)(__VdfgRegularize_h6e95ff9d_0_1))));
        vlSelfRef.mem_wstrb = ((IData)(vlSelfRef.rv32_soc__DOT__st_strb) 
                               & (- (IData)((IData)(__VdfgRegularize_h6e95ff9d_0_1))));
    }

    VLCU_D static bool Vrv32_il___024root___eval_phase__stl(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VstlExecute;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__stl
            vlSelfRef.__VstlTriggered[0U] = ((0xfffffffffffffffeULL 
                                              & vlSelfRef.__VstlTriggered[0U]) 
                                             | (IData)((IData)(vlSelfRef.__VstlFirstIteration)));
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrv32_il___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vrv32_il___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vrv32_il___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vrv32_il___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
        typedef bool VlcuFuncRet;
        // Locals
        IData/*31:0*/ n;
        // Body
        n = 0U;
        do {
            if (in[n]) {
                return (1U);
            }
            n = ((IData)(1U) + n);
        } while ((1U > n));
        return (0U);
    }

    VLCU_D static void Vrv32_il___024root___nba_sequent__TOP__0(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ rv32_soc__DOT__exec_ls;
        rv32_soc__DOT__exec_ls = 0;
        CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_1;
        __VdfgRegularize_h6e95ff9d_0_1 = 0;
        CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_2;
        __VdfgRegularize_h6e95ff9d_0_2 = 0;
        CData/*1:0*/ __Vdly__rv32_soc__DOT__state;
        __Vdly__rv32_soc__DOT__state = 0;
        IData/*31:0*/ __Vdly__rv32_soc__DOT__insn;
        __Vdly__rv32_soc__DOT__insn = 0;
        IData/*31:0*/ __Vdly__rv32_soc__DOT__pc;
        __Vdly__rv32_soc__DOT__pc = 0;
        IData/*31:0*/ __Vdly__rv32_soc__DOT__instret;
        __Vdly__rv32_soc__DOT__instret = 0;
        IData/*31:0*/ __Vdly__rv32_soc__DOT__ld_addr;
        __Vdly__rv32_soc__DOT__ld_addr = 0;
        IData/*31:0*/ __VdlyVal__rv32_soc__DOT__regs__v0;
        __VdlyVal__rv32_soc__DOT__regs__v0 = 0;
        CData/*4:0*/ __VdlyDim0__rv32_soc__DOT__regs__v0;
        __VdlyDim0__rv32_soc__DOT__regs__v0 = 0;
        CData/*0:0*/ __VdlySet__rv32_soc__DOT__regs__v0;
        __VdlySet__rv32_soc__DOT__regs__v0 = 0;
        IData/*31:0*/ __VdlyVal__rv32_soc__DOT__regs__v1;
        __VdlyVal__rv32_soc__DOT__regs__v1 = 0;
        CData/*4:0*/ __VdlyDim0__rv32_soc__DOT__regs__v1;
        __VdlyDim0__rv32_soc__DOT__regs__v1 = 0;
        CData/*0:0*/ __VdlySet__rv32_soc__DOT__regs__v1;
        __VdlySet__rv32_soc__DOT__regs__v1 = 0;
        IData/*31:0*/ __VdlyVal__rv32_soc__DOT__regs__v2;
        __VdlyVal__rv32_soc__DOT__regs__v2 = 0;
        CData/*4:0*/ __VdlyDim0__rv32_soc__DOT__regs__v2;
        __VdlyDim0__rv32_soc__DOT__regs__v2 = 0;
        CData/*0:0*/ __VdlySet__rv32_soc__DOT__regs__v2;
        __VdlySet__rv32_soc__DOT__regs__v2 = 0;
        IData/*31:0*/ __VdlyVal__rv32_soc__DOT__regs__v3;
        __VdlyVal__rv32_soc__DOT__regs__v3 = 0;
        CData/*4:0*/ __VdlyDim0__rv32_soc__DOT__regs__v3;
        __VdlyDim0__rv32_soc__DOT__regs__v3 = 0;
        CData/*0:0*/ __VdlySet__rv32_soc__DOT__regs__v3;
        __VdlySet__rv32_soc__DOT__regs__v3 = 0;
        IData/*31:0*/ __VdlyVal__rv32_soc__DOT__regs__v4;
        __VdlyVal__rv32_soc__DOT__regs__v4 = 0;
        CData/*4:0*/ __VdlyDim0__rv32_soc__DOT__regs__v4;
        __VdlyDim0__rv32_soc__DOT__regs__v4 = 0;
        CData/*0:0*/ __VdlySet__rv32_soc__DOT__regs__v4;
        __VdlySet__rv32_soc__DOT__regs__v4 = 0;
        CData/*7:0*/ __VdlyVal__rv32_soc__DOT__mem__v0;
        __VdlyVal__rv32_soc__DOT__mem__v0 = 0;
        SData/*9:0*/ __VdlyDim0__rv32_soc__DOT__mem__v0;
        __VdlyDim0__rv32_soc__DOT__mem__v0 = 0;
        CData/*0:0*/ __VdlySet__rv32_soc__DOT__mem__v0;
        __VdlySet__rv32_soc__DOT__mem__v0 = 0;
        CData/*7:0*/ __VdlyVal__rv32_soc__DOT__mem__v1;
        __VdlyVal__rv32_soc__DOT__mem__v1 = 0;
        SData/*9:0*/ __VdlyDim0__rv32_soc__DOT__mem__v1;
        __VdlyDim0__rv32_soc__DOT__mem__v1 = 0;
        CData/*0:0*/ __VdlySet__rv32_soc__DOT__mem__v1;
        __VdlySet__rv32_soc__DOT__mem__v1 = 0;
        CData/*7:0*/ __VdlyVal__rv32_soc__DOT__mem__v2;
        __VdlyVal__rv32_soc__DOT__mem__v2 = 0;
        SData/*9:0*/ __VdlyDim0__rv32_soc__DOT__mem__v2;
        __VdlyDim0__rv32_soc__DOT__mem__v2 = 0;
        CData/*0:0*/ __VdlySet__rv32_soc__DOT__mem__v2;
        __VdlySet__rv32_soc__DOT__mem__v2 = 0;
        CData/*7:0*/ __VdlyVal__rv32_soc__DOT__mem__v3;
        __VdlyVal__rv32_soc__DOT__mem__v3 = 0;
        SData/*9:0*/ __VdlyDim0__rv32_soc__DOT__mem__v3;
        __VdlyDim0__rv32_soc__DOT__mem__v3 = 0;
        CData/*0:0*/ __VdlySet__rv32_soc__DOT__mem__v3;
        __VdlySet__rv32_soc__DOT__mem__v3 = 0;
        IData/*31:0*/ __VdlyVal__rv32_soc__DOT__regs__v5;
        __VdlyVal__rv32_soc__DOT__regs__v5 = 0;
        CData/*4:0*/ __VdlyDim0__rv32_soc__DOT__regs__v5;
        __VdlyDim0__rv32_soc__DOT__regs__v5 = 0;
        CData/*0:0*/ __VdlySet__rv32_soc__DOT__regs__v5;
        __VdlySet__rv32_soc__DOT__regs__v5 = 0;
        // Body
        __Vdly__rv32_soc__DOT__instret = vlSelfRef.rv32_soc__DOT__instret;
        __Vdly__rv32_soc__DOT__ld_addr = vlSelfRef.rv32_soc__DOT__ld_addr;
        __VdlySet__rv32_soc__DOT__mem__v0 = 0U;
        __VdlySet__rv32_soc__DOT__mem__v1 = 0U;
        __VdlySet__rv32_soc__DOT__mem__v2 = 0U;
        __VdlySet__rv32_soc__DOT__mem__v3 = 0U;
        __Vdly__rv32_soc__DOT__pc = vlSelfRef.rv32_soc__DOT__pc;
        __Vdly__rv32_soc__DOT__state = vlSelfRef.rv32_soc__DOT__state;
        __VdlySet__rv32_soc__DOT__regs__v0 = 0U;
        __VdlySet__rv32_soc__DOT__regs__v1 = 0U;
        __VdlySet__rv32_soc__DOT__regs__v2 = 0U;
        __VdlySet__rv32_soc__DOT__regs__v3 = 0U;
        __VdlySet__rv32_soc__DOT__regs__v4 = 0U;
        __VdlySet__rv32_soc__DOT__regs__v5 = 0U;
        __Vdly__rv32_soc__DOT__insn = vlSelfRef.rv32_soc__DOT__insn;
        if (vlSelfRef.resetn) {
            if ((0U == (IData)(vlSelfRef.rv32_soc__DOT__state))) {
                if (vlSelfRef.rv32_soc__DOT__pc_bad) {
                    __Vdly__rv32_soc__DOT__state = 3U;
                } else {
                    __Vdly__rv32_soc__DOT__insn = vlSelfRef.rv32_soc__DOT__mem
                        [(0x000003ffU & (vlSelfRef.rv32_soc__DOT__pc 
                                         >> 2U))];
                    __Vdly__rv32_soc__DOT__state = 1U;
                }
            } else if ((1U == (IData)(vlSelfRef.rv32_soc__DOT__state))) {
                __Vdly__rv32_soc__DOT__state = 0U;
                __Vdly__rv32_soc__DOT__pc = ((IData)(4U) 
                                             + vlSelfRef.rv32_soc__DOT__pc);
                __Vdly__rv32_soc__DOT__instret = ((IData)(1U) 
                                                  + vlSelfRef.rv32_soc__DOT__instret);
                if ((0x37U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
                    if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                               >> 7U)))) {
                        __VdlyVal__rv32_soc__DOT__regs__v0 
                            = vlSelfRef.rv32_soc__DOT__imm_u;
                        __VdlyDim0__rv32_soc__DOT__regs__v0 
                            = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                              >> 7U));
                        __VdlySet__rv32_soc__DOT__regs__v0 = 1U;
                    }
                } else if ((0x17U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
                    if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                               >> 7U)))) {
                        __VdlyVal__rv32_soc__DOT__regs__v1 
                            = (vlSelfRef.rv32_soc__DOT__pc 
                               + vlSelfRef.rv32_soc__DOT__imm_u);
                        __VdlyDim0__rv32_soc__DOT__regs__v1 
                            = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                              >> 7U));
                        __VdlySet__rv32_soc__DOT__regs__v1 = 1U;
                    }
                } else if ((0x6fU == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
                    if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                               >> 7U)))) {
                        __VdlyVal__rv32_soc__DOT__regs__v2 
                            = ((IData)(4U) + vlSelfRef.rv32_soc__DOT__pc);
                        __VdlyDim0__rv32_soc__DOT__regs__v2 
                            = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                              >> 7U));
                        __VdlySet__rv32_soc__DOT__regs__v2 = 1U;
                    }
                    __Vdly__rv32_soc__DOT__pc = (vlSelfRef.rv32_soc__DOT__pc 
                                                 + 
                                                 ((((0x00000ffeU 
                                                     & ((- (IData)(
                                                                   (vlSelfRef.rv32_soc__DOT__insn 
                                                                    >> 0x0000001fU))) 
                                                        << 1U)) 
                                                    | (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000001fU)) 
                                                   << 0x00000014U) 
                                                  | ((((0x000001feU 
                                                        & (vlSelfRef.rv32_soc__DOT__insn 
                                                           >> 0x0000000bU)) 
                                                       | (1U 
                                                          & (vlSelfRef.rv32_soc__DOT__insn 
                                                             >> 0x00000014U))) 
                                                      << 0x0000000bU) 
                                                     | (0x000007feU 
                                                        & (vlSelfRef.rv32_soc__DOT__insn 
                                                           >> 0x00000014U)))));
                } else if ((IData)((0x00000067U == 
                                    (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
                    if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                               >> 7U)))) {
                        __VdlyVal__rv32_soc__DOT__regs__v3 
                            = ((IData)(4U) + vlSelfRef.rv32_soc__DOT__pc);
                        __VdlyDim0__rv32_soc__DOT__regs__v3 
                            = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                              >> 7U));
                        __VdlySet__rv32_soc__DOT__regs__v3 = 1U;
                    }
                    __Vdly__rv32_soc__DOT__pc = (0xfffffffeU 
                                                 & (vlSelfRef.rv32_soc__DOT__rs1_val 
                                                    + vlSelfRef.rv32_soc__DOT__imm_i));
                } else if (((0x63U == (0x0000007fU 
                                       & vlSelfRef.rv32_soc__DOT__insn)) 
                            & ((2U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                             >> 0x0000000cU))) 
                               & (3U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                               >> 0x0000000cU)))))) {
                    if (((0x00004000U & vlSelfRef.rv32_soc__DOT__insn)
                          ? ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)
                              ? ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)
                                  ? (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                  : (vlSelfRef.rv32_soc__DOT__rs1_val 
                                     < vlSelfRef.rv32_soc__DOT__rs2_val))
                              : ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)
                                  ? VL_GTES_III(32, vlSelfRef.rv32_soc__DOT__rs1_val, vlSelfRef.rv32_soc__DOT__rs2_val)
                                  : VL_LTS_III(32, vlSelfRef.rv32_soc__DOT__rs1_val, vlSelfRef.rv32_soc__DOT__rs2_val)))
                          : ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)
                              ? (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                              : ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)
                                  ? (vlSelfRef.rv32_soc__DOT__rs1_val 
                                     != vlSelfRef.rv32_soc__DOT__rs2_val)
                                  : (vlSelfRef.rv32_soc__DOT__rs1_val 
                                     == vlSelfRef.rv32_soc__DOT__rs2_val))))) {
                        __Vdly__rv32_soc__DOT__pc = 
                            (vlSelfRef.rv32_soc__DOT__pc 
                             + (((- (IData)((vlSelfRef.rv32_soc__DOT__insn 
                                             >> 0x0000001fU))) 
                                 << 0x0000000dU) | 
                                ((((2U & (vlSelfRef.rv32_soc__DOT__insn 
                                          >> 0x0000001eU)) 
                                   | (1U & (vlSelfRef.rv32_soc__DOT__insn 
                                            >> 7U))) 
                                  << 0x0000000bU) | 
                                 ((0x000007e0U & (vlSelfRef.rv32_soc__DOT__insn 
                                                  >> 0x00000014U)) 
                                  | (0x0000001eU & 
                                     (vlSelfRef.rv32_soc__DOT__insn 
                                      >> 7U))))));
                    }
                } else if ((((0x13U == (0x0000007fU 
                                        & vlSelfRef.rv32_soc__DOT__insn)) 
                             & ((0U == (vlSelfRef.rv32_soc__DOT__insn 
                                        >> 0x00000019U)) 
                                | ((1U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000000cU))) 
                                   & (IData)(((0x00005000U 
                                               != (0x00007000U 
                                                   & vlSelfRef.rv32_soc__DOT__insn)) 
                                              | (0x20U 
                                                 == 
                                                 (vlSelfRef.rv32_soc__DOT__insn 
                                                  >> 0x00000019U))))))) 
                            | (IData)(vlSelfRef.rv32_soc__DOT__is_alur))) {
                    if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                               >> 7U)))) {
                        __VdlyVal__rv32_soc__DOT__regs__v4 
                            = ((0x00004000U & vlSelfRef.rv32_soc__DOT__insn)
                                ? ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)
                                    ? ((0x00001000U 
                                        & vlSelfRef.rv32_soc__DOT__insn)
                                        ? (vlSelfRef.rv32_soc__DOT__rs1_val 
                                           & vlSelfRef.rv32_soc__DOT__alu_b)
                                        : (vlSelfRef.rv32_soc__DOT__rs1_val 
                                           | vlSelfRef.rv32_soc__DOT__alu_b))
                                    : ((0x00001000U 
                                        & vlSelfRef.rv32_soc__DOT__insn)
                                        ? ((0x40000000U 
                                            & vlSelfRef.rv32_soc__DOT__insn)
                                            ? VL_SHIFTRS_III(32,32,5, vlSelfRef.rv32_soc__DOT__rs1_val, 
                                                             (0x0000001fU 
                                                              & vlSelfRef.rv32_soc__DOT__alu_b))
                                            : (vlSelfRef.rv32_soc__DOT__rs1_val 
                                               >> (0x0000001fU 
                                                   & vlSelfRef.rv32_soc__DOT__alu_b)))
                                        : (vlSelfRef.rv32_soc__DOT__rs1_val 
                                           ^ vlSelfRef.rv32_soc__DOT__alu_b)))
                                : ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)
                                    ? ((0x00001000U 
                                        & vlSelfRef.rv32_soc__DOT__insn)
                                        ? (vlSelfRef.rv32_soc__DOT__rs1_val 
                                           < vlSelfRef.rv32_soc__DOT__alu_b)
                                        : VL_LTS_III(32, vlSelfRef.rv32_soc__DOT__rs1_val, vlSelfRef.rv32_soc__DOT__alu_b))
                                    : ((0x00001000U 
                                        & vlSelfRef.rv32_soc__DOT__insn)
                                        ? (vlSelfRef.rv32_soc__DOT__rs1_val 
                                           << (0x0000001fU 
                                               & vlSelfRef.rv32_soc__DOT__alu_b))
                                        : (((vlSelfRef.rv32_soc__DOT__insn 
                                             >> 0x0000001eU) 
                                            & (IData)(vlSelfRef.rv32_soc__DOT__is_alur))
                                            ? (vlSelfRef.rv32_soc__DOT__rs1_val 
                                               - vlSelfRef.rv32_soc__DOT__alu_b)
                                            : (vlSelfRef.rv32_soc__DOT__rs1_val 
                                               + vlSelfRef.rv32_soc__DOT__alu_b)))));
                        __VdlyDim0__rv32_soc__DOT__regs__v4 
                            = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                              >> 7U));
                        __VdlySet__rv32_soc__DOT__regs__v4 = 1U;
                    }
                } else if (vlSelfRef.rv32_soc__DOT__is_load) {
                    __Vdly__rv32_soc__DOT__pc = vlSelfRef.rv32_soc__DOT__pc;
                    __Vdly__rv32_soc__DOT__instret 
                        = vlSelfRef.rv32_soc__DOT__instret;
                    if (((IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned) 
                         | (0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                            >> 2U)))) {
                        __Vdly__rv32_soc__DOT__state = 3U;
                    } else {
                        __Vdly__rv32_soc__DOT__ld_addr 
                            = vlSelfRef.rv32_soc__DOT__ls_addr;
                        __Vdly__rv32_soc__DOT__state = 2U;
                    }
                } else if (vlSelfRef.rv32_soc__DOT__is_store) {
                    if (((IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned) 
                         | (0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                            >> 2U)))) {
                        __Vdly__rv32_soc__DOT__pc = vlSelfRef.rv32_soc__DOT__pc;
                        __Vdly__rv32_soc__DOT__instret 
                            = vlSelfRef.rv32_soc__DOT__instret;
                        __Vdly__rv32_soc__DOT__state = 3U;
                    } else {
                        if ((1U & (IData)(vlSelfRef.rv32_soc__DOT__st_strb))) {
                            __VdlyVal__rv32_soc__DOT__mem__v0 
                                = (0x000000ffU & vlSelfRef.rv32_soc__DOT__st_data);
                            __VdlyDim0__rv32_soc__DOT__mem__v0 
                                = (0x000003ffU & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                                  >> 2U));
                            __VdlySet__rv32_soc__DOT__mem__v0 = 1U;
                        }
                        if ((2U & (IData)(vlSelfRef.rv32_soc__DOT__st_strb))) {
                            __VdlyVal__rv32_soc__DOT__mem__v1 
                                = (0x000000ffU & (vlSelfRef.rv32_soc__DOT__st_data 
                                                  >> 8U));
                            __VdlyDim0__rv32_soc__DOT__mem__v1 
                                = (0x000003ffU & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                                  >> 2U));
                            __VdlySet__rv32_soc__DOT__mem__v1 = 1U;
                        }
                        if ((4U & (IData)(vlSelfRef.rv32_soc__DOT__st_strb))) {
                            __VdlyVal__rv32_soc__DOT__mem__v2 
                                = (0x000000ffU & (vlSelfRef.rv32_soc__DOT__st_data 
                                                  >> 0x10U));
                            __VdlyDim0__rv32_soc__DOT__mem__v2 
                                = (0x000003ffU & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                                  >> 2U));
                            __VdlySet__rv32_soc__DOT__mem__v2 = 1U;
                        }
                        if ((8U & (IData)(vlSelfRef.rv32_soc__DOT__st_strb))) {
                            __VdlyVal__rv32_soc__DOT__mem__v3 
                                = (vlSelfRef.rv32_soc__DOT__st_data 
                                   >> 0x18U);
                            __VdlyDim0__rv32_soc__DOT__mem__v3 
                                = (0x000003ffU & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                                  >> 2U));
                            __VdlySet__rv32_soc__DOT__mem__v3 = 1U;
                        }
                    }
                } else {
                    __Vdly__rv32_soc__DOT__pc = vlSelfRef.rv32_soc__DOT__pc;
                    __Vdly__rv32_soc__DOT__instret 
                        = vlSelfRef.rv32_soc__DOT__instret;
                    __Vdly__rv32_soc__DOT__state = 3U;
                }
            } else if ((2U == (IData)(vlSelfRef.rv32_soc__DOT__state))) {
                __Vdly__rv32_soc__DOT__pc = ((IData)(4U) 
                                             + vlSelfRef.rv32_soc__DOT__pc);
                __Vdly__rv32_soc__DOT__instret = ((IData)(1U) 
                                                  + vlSelfRef.rv32_soc__DOT__instret);
                if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                           >> 7U)))) {
                    __VdlyVal__rv32_soc__DOT__regs__v5 
                        = ((0U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                         >> 0x0000000cU)))
                            ? (((- (IData)((1U & (vlSelfRef.rv32_soc__DOT__ld_shift 
                                                  >> 7U)))) 
                                << 8U) | (0x000000ffU 
                                          & vlSelfRef.rv32_soc__DOT__ld_shift))
                            : ((1U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                             >> 0x0000000cU)))
                                ? (((- (IData)((1U 
                                                & (vlSelfRef.rv32_soc__DOT__ld_shift 
                                                   >> 0x0000000fU)))) 
                                    << 0x00000010U) 
                                   | (0x0000ffffU & vlSelfRef.rv32_soc__DOT__ld_shift))
                                : ((4U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000000cU)))
                                    ? (0x000000ffU 
                                       & vlSelfRef.rv32_soc__DOT__ld_shift)
                                    : ((5U == (7U & 
                                               (vlSelfRef.rv32_soc__DOT__insn 
                                                >> 0x0000000cU)))
                                        ? (0x0000ffffU 
                                           & vlSelfRef.rv32_soc__DOT__ld_shift)
                                        : vlSelfRef.rv32_soc__DOT__mem
                                       [(0x000003ffU 
                                         & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                            >> 2U))]))));
                    __VdlyDim0__rv32_soc__DOT__regs__v5 
                        = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                          >> 7U));
                    __VdlySet__rv32_soc__DOT__regs__v5 = 1U;
                }
                __Vdly__rv32_soc__DOT__state = 0U;
            } else {
                __Vdly__rv32_soc__DOT__state = 3U;
            }
        } else {
            __Vdly__rv32_soc__DOT__pc = 0U;
            __Vdly__rv32_soc__DOT__state = 0U;
            __Vdly__rv32_soc__DOT__instret = 0U;
        }
        vlSelfRef.rv32_soc__DOT__instret = __Vdly__rv32_soc__DOT__instret;
        vlSelfRef.rv32_soc__DOT__ld_addr = __Vdly__rv32_soc__DOT__ld_addr;
        if (__VdlySet__rv32_soc__DOT__mem__v0) {
            vlSelfRef.rv32_soc__DOT__mem[__VdlyDim0__rv32_soc__DOT__mem__v0] 
                = ((0xffffff00U & vlSelfRef.rv32_soc__DOT__mem
                    [__VdlyDim0__rv32_soc__DOT__mem__v0]) 
                   | (IData)(__VdlyVal__rv32_soc__DOT__mem__v0));
        }
        if (__VdlySet__rv32_soc__DOT__mem__v1) {
            vlSelfRef.rv32_soc__DOT__mem[__VdlyDim0__rv32_soc__DOT__mem__v1] 
                = ((0xffff00ffU & vlSelfRef.rv32_soc__DOT__mem
                    [__VdlyDim0__rv32_soc__DOT__mem__v1]) 
                   | ((IData)(__VdlyVal__rv32_soc__DOT__mem__v1) 
                      << 8U));
        }
        if (__VdlySet__rv32_soc__DOT__mem__v2) {
            vlSelfRef.rv32_soc__DOT__mem[__VdlyDim0__rv32_soc__DOT__mem__v2] 
                = ((0xff00ffffU & vlSelfRef.rv32_soc__DOT__mem
                    [__VdlyDim0__rv32_soc__DOT__mem__v2]) 
                   | ((IData)(__VdlyVal__rv32_soc__DOT__mem__v2) 
                      << 0x00000010U));
        }
        if (__VdlySet__rv32_soc__DOT__mem__v3) {
            vlSelfRef.rv32_soc__DOT__mem[__VdlyDim0__rv32_soc__DOT__mem__v3] 
                = ((0x00ffffffU & vlSelfRef.rv32_soc__DOT__mem
                    [__VdlyDim0__rv32_soc__DOT__mem__v3]) 
                   | ((IData)(__VdlyVal__rv32_soc__DOT__mem__v3) 
                      << 0x00000018U));
        }
        vlSelfRef.rv32_soc__DOT__pc = __Vdly__rv32_soc__DOT__pc;
        vlSelfRef.rv32_soc__DOT__state = __Vdly__rv32_soc__DOT__state;
        if (__VdlySet__rv32_soc__DOT__regs__v0) {
            vlSelfRef.rv32_soc__DOT__regs[__VdlyDim0__rv32_soc__DOT__regs__v0] 
                = __VdlyVal__rv32_soc__DOT__regs__v0;
        }
        if (__VdlySet__rv32_soc__DOT__regs__v1) {
            vlSelfRef.rv32_soc__DOT__regs[__VdlyDim0__rv32_soc__DOT__regs__v1] 
                = __VdlyVal__rv32_soc__DOT__regs__v1;
        }
        if (__VdlySet__rv32_soc__DOT__regs__v2) {
            vlSelfRef.rv32_soc__DOT__regs[__VdlyDim0__rv32_soc__DOT__regs__v2] 
                = __VdlyVal__rv32_soc__DOT__regs__v2;
        }
        if (__VdlySet__rv32_soc__DOT__regs__v3) {
            vlSelfRef.rv32_soc__DOT__regs[__VdlyDim0__rv32_soc__DOT__regs__v3] 
                = __VdlyVal__rv32_soc__DOT__regs__v3;
        }
        if (__VdlySet__rv32_soc__DOT__regs__v4) {
            vlSelfRef.rv32_soc__DOT__regs[__VdlyDim0__rv32_soc__DOT__regs__v4] 
                = __VdlyVal__rv32_soc__DOT__regs__v4;
        }
        if (__VdlySet__rv32_soc__DOT__regs__v5) {
            vlSelfRef.rv32_soc__DOT__regs[__VdlyDim0__rv32_soc__DOT__regs__v5] 
                = __VdlyVal__rv32_soc__DOT__regs__v5;
        }
        vlSelfRef.rv32_soc__DOT__insn = __Vdly__rv32_soc__DOT__insn;
        vlSelfRef.retired = vlSelfRef.rv32_soc__DOT__instret;
        vlSelfRef.rv32_soc__DOT__ld_shift = (vlSelfRef.rv32_soc__DOT__mem
                                             [(0x000003ffU 
                                               & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                                  >> 2U))] 
                                             >> (0x00000018U 
                                                 & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                                    << 3U)));
        vlSelfRef.rv32_soc__DOT__pc_bad = ((0U != (3U 
                                                   & vlSelfRef.rv32_soc__DOT__pc)) 
                                           | (0x00000400U 
                                              <= (vlSelfRef.rv32_soc__DOT__pc 
                                                  >> 2U)));
        vlSelfRef.trap = (3U == (IData)(vlSelfRef.rv32_soc__DOT__state));
        vlSelfRef.rv32_soc__DOT__imm_u = (0xfffff000U 
                                          & vlSelfRef.rv32_soc__DOT__insn);
        vlSelfRef.rv32_soc__DOT__imm_i = (((- (IData)(
                                                      (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000001fU))) 
                                           << 0x0000000cU) 
                                          | (vlSelfRef.rv32_soc__DOT__insn 
                                             >> 0x00000014U));
        vlSelfRef.rv32_soc__DOT__is_alur = ((0x33U 
                                             == (0x0000007fU 
                                                 & vlSelfRef.rv32_soc__DOT__insn)) 
                                            & ((0U 
                                                == 
                                                (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x00000019U)) 
                                               | ((0x20U 
                                                   == 
                                                   (vlSelfRef.rv32_soc__DOT__insn 
                                                    >> 0x00000019U)) 
                                                  & ((0U 
                                                      == 
                                                      (7U 
                                                       & (vlSelfRef.rv32_soc__DOT__insn 
                                                          >> 0x0000000cU))) 
                                                     | (5U 
                                                        == 
                                                        (7U 
                                                         & (vlSelfRef.rv32_soc__DOT__insn 
                                                            >> 0x0000000cU)))))));
        vlSelfRef.rv32_soc__DOT__rs2_val = ((0U == 
                                             (0x0000001fU 
                                              & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x00000014U)))
                                             ? 0U : vlSelfRef.rv32_soc__DOT__regs
                                            [(0x0000001fU 
                                              & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x00000014U))]);
        vlSelfRef.rv32_soc__DOT__rs1_val = ((0U == 
                                             (0x0000001fU 
                                              & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000000fU)))
                                             ? 0U : vlSelfRef.rv32_soc__DOT__regs
                                            [(0x0000001fU 
                                              & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000000fU))]);
        __VdfgRegularize_h6e95ff9d_0_2 = ((0U == (7U 
                                                  & (vlSelfRef.rv32_soc__DOT__insn 
                                                     >> 0x0000000cU))) 
                                          | ((1U == 
                                              (7U & 
                                               (vlSelfRef.rv32_soc__DOT__insn 
                                                >> 0x0000000cU))) 
                                             | (2U 
                                                == 
                                                (7U 
                                                 & (vlSelfRef.rv32_soc__DOT__insn 
                                                    >> 0x0000000cU)))));
        vlSelfRef.mem_instr = ((~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)) 
                               & (0U == (IData)(vlSelfRef.rv32_soc__DOT__state)));
        vlSelfRef.rv32_soc__DOT__alu_b = ((IData)(vlSelfRef.rv32_soc__DOT__is_alur)
                                           ? vlSelfRef.rv32_soc__DOT__rs2_val
                                           : vlSelfRef.rv32_soc__DOT__imm_i);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 = 
            (vlSelfRef.rv32_soc__DOT__rs1_val >= vlSelfRef.rv32_soc__DOT__rs2_val);
        vlSelfRef.rv32_soc__DOT__is_load = ((3U == 
                                             (0x0000007fU 
                                              & vlSelfRef.rv32_soc__DOT__insn)) 
                                            & ((IData)(__VdfgRegularize_h6e95ff9d_0_2) 
                                               | ((4U 
                                                   == 
                                                   (7U 
                                                    & (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000000cU))) 
                                                  | (5U 
                                                     == 
                                                     (7U 
                                                      & (vlSelfRef.rv32_soc__DOT__insn 
                                                         >> 0x0000000cU))))));
        vlSelfRef.rv32_soc__DOT__is_store = ((0x23U 
                                              == (0x0000007fU 
                                                  & vlSelfRef.rv32_soc__DOT__insn)) 
                                             & (IData)(__VdfgRegularize_h6e95ff9d_0_2));
        vlSelfRef.rv32_soc__DOT__ls_addr = (vlSelfRef.rv32_soc__DOT__rs1_val 
                                            + (((- (IData)(
                                                           (vlSelfRef.rv32_soc__DOT__insn 
                                                            >> 0x0000001fU))) 
                                                << 0x0000000cU) 
                                               | ((0x00000fe0U 
                                                   & (vlSelfRef.rv32_soc__DOT__insn 
                                                      >> 0x00000014U)) 
                                                  | (0x0000001fU 
                                                     & ((IData)(vlSelfRef.rv32_soc__DOT__is_store)
                                                         ? 
                                                        (vlSelfRef.rv32_soc__DOT__insn 
                                                         >> 7U)
                                                         : 
                                                        (vlSelfRef.rv32_soc__DOT__insn 
                                                         >> 0x00000014U))))));
        if ((0U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlSelfRef.rv32_soc__DOT__st_data = ((vlSelfRef.rv32_soc__DOT__rs2_val 
                                                 << 0x00000018U) 
                                                | ((0x00ff0000U 
                                                    & (vlSelfRef.rv32_soc__DOT__rs2_val 
                                                       << 0x00000010U)) 
                                                   | ((0x0000ff00U 
                                                       & (vlSelfRef.rv32_soc__DOT__rs2_val 
                                                          << 8U)) 
                                                      | (0x000000ffU 
                                                         & vlSelfRef.rv32_soc__DOT__rs2_val))));
            vlSelfRef.rv32_soc__DOT__st_strb = (0x0000000fU 
                                                & ((IData)(1U) 
                                                   << 
                                                   (3U 
                                                    & vlSelfRef.rv32_soc__DOT__ls_addr)));
        } else {
            vlSelfRef.rv32_soc__DOT__st_data = ((1U 
                                                 == 
                                                 (3U 
                                                  & (vlSelfRef.rv32_soc__DOT__insn 
                                                     >> 0x0000000cU)))
                                                 ? 
                                                ((vlSelfRef.rv32_soc__DOT__rs2_val 
                                                  << 0x00000010U) 
                                                 | (0x0000ffffU 
                                                    & vlSelfRef.rv32_soc__DOT__rs2_val))
                                                 : vlSelfRef.rv32_soc__DOT__rs2_val);
            vlSelfRef.rv32_soc__DOT__st_strb = (0x0000000fU 
                                                & (((IData)(3U) 
                                                    << 
                                                    (3U 
                                                     & vlSelfRef.rv32_soc__DOT__ls_addr)) 
                                                   | (- (IData)(
                                                                (1U 
                                                                 != 
                                                                 (3U 
                                                                  & (vlSelfRef.rv32_soc__DOT__insn 
                                                                     >> 0x0000000cU)))))));
        }
        vlSelfRef.rv32_soc__DOT__ls_misaligned = ((IData)(
                                                          ((0x00001000U 
                                                            == 
                                                            (0x00003000U 
                                                             & vlSelfRef.rv32_soc__DOT__insn)) 
                                                           & vlSelfRef.rv32_soc__DOT__ls_addr)) 
                                                  | (IData)(
                                                            ((0x00002000U 
                                                              == 
                                                              (0x00003000U 
                                                               & vlSelfRef.rv32_soc__DOT__insn)) 
                                                             & (0U 
                                                                != 
                                                                (3U 
                                                                 & vlSelfRef.rv32_soc__DOT__ls_addr)))));
        rv32_soc__DOT__exec_ls = ((1U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
                                  & ((~ ((0x00000400U 
                                          <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                              >> 2U)) 
                                         | (IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned))) 
                                     & ((IData)(vlSelfRef.rv32_soc__DOT__is_load) 
                                        | (IData)(vlSelfRef.rv32_soc__DOT__is_store))));
        vlSelfRef.mem_valid = ((IData)(vlSelfRef.mem_instr) 
                               | (IData)(rv32_soc__DOT__exec_ls));
        if (vlSelfRef.mem_instr) {
            vlSelfRef.mem_addr = vlSelfRef.rv32_soc__DOT__pc;
            vlSelfRef.mem_rdata = vlSelfRef.rv32_soc__DOT__mem
                [(0x000003ffU & (vlSelfRef.rv32_soc__DOT__pc 
                                 >> 2U))];
        } else {
            vlSelfRef.mem_addr = (0xfffffffcU & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                                 & (- (IData)((IData)(rv32_soc__DOT__exec_ls)))));
            vlSelfRef.mem_rdata = (((IData)(rv32_soc__DOT__exec_ls) 
                                    & (IData)(vlSelfRef.rv32_soc__DOT__is_load))
                                    ? vlSelfRef.rv32_soc__DOT__mem
                                   [(0x000003ffU & 
                                     (vlSelfRef.rv32_soc__DOT__ls_addr 
                                      >> 2U))] : 0U);
        }
        __VdfgRegularize_h6e95ff9d_0_1 = ((IData)(rv32_soc__DOT__exec_ls) 
                                          & (IData)(vlSelfRef.rv32_soc__DOT__is_store));
        vlSelfRef.mem_wdata = (vlSelfRef.rv32_soc__DOT__st_data 
                               & (- (IData)((IData)(__VdfgRegularize_h6e95ff9d_0_1))));
        vlSelfRef.mem_wstrb = ((IData)(vlSelfRef.rv32_soc__DOT__st_strb) 
                               & (- (IData)((IData)(__VdfgRegularize_h6e95ff9d_0_1))));
    }

    VLCU_D static void Vrv32_il___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
        typedef void VlcuFuncRet;
        // Locals
        IData/*31:0*/ n;
        // Body
        n = 0U;
        do {
            out[n] = (out[n] | in[n]);
            n = ((IData)(1U) + n);
        } while ((0U >= n));
    }

    VLCU_D static bool Vrv32_il___024root___eval_phase__act(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__act
            vlSelfRef.__VactTriggered[0U] = (QData)((IData)(
                                                            ((IData)(vlSelfRef.clk) 
                                                             & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrv32_il___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vrv32_il___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vrv32_il___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
        typedef void VlcuFuncRet;
        // Locals
        IData/*31:0*/ n;
        // Body
        n = 0U;
        do {
            out[n] = 0ULL;
            n = ((IData)(1U) + n);
        } while ((1U > n));
    }

    VLCU_D static bool Vrv32_il___024root___eval_phase__nba(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vrv32_il___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vrv32_il___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vrv32_il___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vrv32_il___024root___eval(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __VnbaIterCount;
        // Body
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vrv32_il___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 16, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vrv32_il___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 16, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vrv32_il___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vrv32_il___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vrv32_il___024root___ctor_var_reset(Vrv32_il___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->resetn = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 8624841754543469506ull);
        vlSelf->trap = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 18214934560881419504ull);
        vlSelf->mem_valid = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 1551974250180885553ull);
        vlSelf->mem_instr = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 1406578376079151150ull);
        vlSelf->mem_addr = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 326597072690670135ull);
        vlSelf->mem_wdata = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 5431754401481461448ull);
        vlSelf->mem_wstrb = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 8859681292774497410ull);
        vlSelf->mem_rdata = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 9659133473039683418ull);
        vlSelf->retired = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 9241213606999339685ull);
        for (int __Vi0 = 0; __Vi0 < 1024; ++__Vi0) {
            vlSelf->rv32_soc__DOT__mem[__Vi0] = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 157536494676281617ull);
        }
        for (int __Vi0 = 0; __Vi0 < 32; ++__Vi0) {
            vlSelf->rv32_soc__DOT__regs[__Vi0] = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 11523106369459488163ull);
        }
        vlSelf->rv32_soc__DOT__pc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 11841601083557326370ull);
        vlSelf->rv32_soc__DOT__insn = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 6790827016312138654ull);
        vlSelf->rv32_soc__DOT__ld_addr = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 18411059245981549376ull);
        vlSelf->rv32_soc__DOT__instret = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 1322355925642281443ull);
        vlSelf->rv32_soc__DOT__state = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 769792777424239494ull);
        vlSelf->rv32_soc__DOT__imm_i = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 1712762055299309798ull);
        vlSelf->rv32_soc__DOT__imm_u = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 9881345184067898959ull);
        vlSelf->rv32_soc__DOT__is_load = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 2994363835360905129ull);
        vlSelf->rv32_soc__DOT__is_store = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 7894171225882899832ull);
        vlSelf->rv32_soc__DOT__is_alur = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 11965279755098390934ull);
        vlSelf->rv32_soc__DOT__rs1_val = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 4619132075120175768ull);
        vlSelf->rv32_soc__DOT__rs2_val = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 6237873797417355087ull);
        vlSelf->rv32_soc__DOT__alu_b = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 2322885805071923878ull);
        vlSelf->rv32_soc__DOT__ls_addr = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 327928301318749977ull);
        vlSelf->rv32_soc__DOT__ls_misaligned = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 13860169767913532206ull);
        vlSelf->rv32_soc__DOT__st_data = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 620096875113001671ull);
        vlSelf->rv32_soc__DOT__st_strb = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 9732364407211009767ull);
        vlSelf->rv32_soc__DOT__ld_shift = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 1520602722809905503ull);
        vlSelf->rv32_soc__DOT__pc_bad = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16411224859674924017ull);
        vlSelf->__VdfgRegularize_h6e95ff9d_0_0 = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VnbaTriggered[__Vi0] = 0;
        }
    }

    // tfp->dump(): every traced value stored at its trace code.  Full functions, then the
    // constants, as VerilatedTrace::dump runs them the first time (verilated_trace_imp.h:348-358)
    VLCU_D static void trace_sample(Root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
        (void)vlSelf; (void)vlCtx; (void)valsp;
    }

    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        Vrv32_il___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vrv32_il___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vrv32_il___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vrv32_il___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vrv32_il___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vrv32_il___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vrv32_il__cuda::info() {
    typedef Vrv32_il__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"resetn", "resetn", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_resetn},
        {"trap", "trap", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_trap},
        {"mem_valid", "mem_valid", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_mem_valid},
        {"mem_instr", "mem_instr", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_mem_instr},
        {"mem_addr", "mem_addr", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_mem_addr},
        {"mem_wdata", "mem_wdata", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_mem_wdata},
        {"mem_wstrb", "mem_wstrb", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_mem_wstrb},
        {"mem_rdata", "mem_rdata", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_mem_rdata},
        {"retired", "retired", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_retired},
        {"rv32_soc__DOT__state", "rv32_soc.state", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_rv32_soc__DOT__state},
        {"rv32_soc__DOT__is_load", "rv32_soc.is_load", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__is_load},
        {"rv32_soc__DOT__is_store", "rv32_soc.is_store", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__is_store},
        {"rv32_soc__DOT__is_alur", "rv32_soc.is_alur", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__is_alur},
        {"rv32_soc__DOT__ls_misaligned", "rv32_soc.ls_misaligned", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__ls_misaligned},
        {"rv32_soc__DOT__st_strb", "rv32_soc.st_strb", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_rv32_soc__DOT__st_strb},
        {"rv32_soc__DOT__pc_bad", "rv32_soc.pc_bad", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__pc_bad},
        {"__VdfgRegularize_h6e95ff9d_0_0", "__VdfgRegularize_h6e95ff9d_0_0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_0},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"rv32_soc__DOT__pc", "rv32_soc.pc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__pc},
        {"rv32_soc__DOT__insn", "rv32_soc.insn", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__insn},
        {"rv32_soc__DOT__ld_addr", "rv32_soc.ld_addr", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__ld_addr},
        {"rv32_soc__DOT__instret", "rv32_soc.instret", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__instret},
        {"rv32_soc__DOT__imm_i", "rv32_soc.imm_i", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__imm_i},
        {"rv32_soc__DOT__imm_u", "rv32_soc.imm_u", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__imm_u},
        {"rv32_soc__DOT__rs1_val", "rv32_soc.rs1_val", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__rs1_val},
        {"rv32_soc__DOT__rs2_val", "rv32_soc.rs2_val", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__rs2_val},
        {"rv32_soc__DOT__alu_b", "rv32_soc.alu_b", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__alu_b},
        {"rv32_soc__DOT__ls_addr", "rv32_soc.ls_addr", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__ls_addr},
        {"rv32_soc__DOT__st_data", "rv32_soc.st_data", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__st_data},
        {"rv32_soc__DOT__ld_shift", "rv32_soc.ld_shift", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__ld_shift},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"rv32_soc__DOT__mem", "rv32_soc.mem", VLCU_MEMBER_STATE, 4, 0, 1024, 32, 31, 0, T::OFF_rv32_soc__DOT__mem},
        {"rv32_soc__DOT__regs", "rv32_soc.regs", VLCU_MEMBER_STATE, 4, 0, 32, 32, 31, 0, T::OFF_rv32_soc__DOT__regs},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 7, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 5, 6, 8, 9, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "rv32_soc", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 41, 0, 0x2f9cd36244ca914aULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vrv32_il__cuda VLCU_TESTBENCHES);
