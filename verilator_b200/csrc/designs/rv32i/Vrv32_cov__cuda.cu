// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vrv32_cov"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vrv32_cov___024root final {
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
    CData/*3:0*/ rv32_soc__DOT____VlemCond_10;
    CData/*3:0*/ rv32_soc__DOT____VlemCond_6;
    CData/*1:0*/ rv32_soc__DOT__state;
    CData/*0:0*/ rv32_soc__DOT__is_jalr;
    CData/*0:0*/ rv32_soc__DOT__is_branch;
    CData/*0:0*/ rv32_soc__DOT__is_load;
    CData/*0:0*/ rv32_soc__DOT__is_store;
    CData/*0:0*/ rv32_soc__DOT__is_alui;
    CData/*0:0*/ rv32_soc__DOT__is_alur;
    CData/*0:0*/ rv32_soc__DOT__take_branch;
    CData/*0:0*/ rv32_soc__DOT__ls_misaligned;
    CData/*3:0*/ rv32_soc__DOT__st_strb;
    CData/*0:0*/ rv32_soc__DOT__pc_bad;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__clk;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__resetn;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__trap;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__mem_valid;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__mem_instr;
    CData/*3:0*/ rv32_soc__DOT____Vtogcov__mem_wstrb;
    CData/*1:0*/ rv32_soc__DOT____Vtogcov__state;
    CData/*6:0*/ rv32_soc__DOT____Vtogcov__opcode;
    CData/*4:0*/ rv32_soc__DOT____Vtogcov__rd;
    CData/*2:0*/ rv32_soc__DOT____Vtogcov__funct3;
    CData/*4:0*/ rv32_soc__DOT____Vtogcov__rs1;
    CData/*4:0*/ rv32_soc__DOT____Vtogcov__rs2;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__f7b5;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__is_lui;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__is_auipc;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__is_jal;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__is_jalr;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__is_branch;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__is_load;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__is_store;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__is_alui;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__is_alur;
    CData/*4:0*/ rv32_soc__DOT____Vtogcov__shamt;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__alu_sub;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__take_branch;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__ls_misaligned;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__ls_oob;
    CData/*3:0*/ rv32_soc__DOT____Vtogcov__st_strb;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__pc_bad;
    CData/*0:0*/ rv32_soc__DOT____Vtogcov__exec_ls;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__resetn__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ rv32_soc__DOT____VlemCond_11;
    IData/*31:0*/ rv32_soc__DOT____VlemCond_9;
    IData/*31:0*/ rv32_soc__DOT____VlemCond_8;
    IData/*31:0*/ rv32_soc__DOT____VlemCond_7;
    IData/*31:0*/ rv32_soc__DOT____VlemCond_5;
    IData/*31:0*/ rv32_soc__DOT____VlemCond_4;
    IData/*31:0*/ rv32_soc__DOT____VlemCond_3;
    IData/*31:0*/ rv32_soc__DOT____VlemCond_2;
    IData/*31:0*/ rv32_soc__DOT____VlemCond_1;
    IData/*31:0*/ rv32_soc__DOT____VlemCond_0;
    IData/*31:0*/ rv32_soc__DOT__pc;
    IData/*31:0*/ rv32_soc__DOT__insn;
    IData/*31:0*/ rv32_soc__DOT__ld_addr;
    IData/*31:0*/ rv32_soc__DOT__instret;
    IData/*31:0*/ rv32_soc__DOT__imm_i;
    IData/*31:0*/ rv32_soc__DOT__imm_b;
    IData/*31:0*/ rv32_soc__DOT__imm_u;
    IData/*31:0*/ rv32_soc__DOT__imm_j;
    IData/*31:0*/ rv32_soc__DOT__rs1_val;
    IData/*31:0*/ rv32_soc__DOT__alu_out;
    IData/*31:0*/ rv32_soc__DOT__ls_addr;
    IData/*31:0*/ rv32_soc__DOT__st_data;
    IData/*31:0*/ rv32_soc__DOT__ld_val;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__mem_addr;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__mem_wdata;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__mem_rdata;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__retired;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__pc;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__insn;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__ld_addr;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__imm_i;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__imm_s;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__imm_b;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__imm_u;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__imm_j;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__rs1_val;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__rs2_val;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__alu_b;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__alu_out;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__ls_addr;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__st_data;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__ld_word;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__ld_shift;
    IData/*31:0*/ rv32_soc__DOT____Vtogcov__ld_val;
    IData/*31:0*/ __VactIterCount;
    VlcuGlobalArray<IData/*31:0*/, 1024> rv32_soc__DOT__mem;  // stays in HBM, instance-major
    VlcuSharedArray<IData/*31:0*/, 32> rv32_soc__DOT__regs;  // run-time indexed: shared memory [index][lane]
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    VlcuWarpCovArray<104 * 16> __Vcoverage;  // stays in HBM as cov[warp][bin], uint64 per warp of instances
    const char* vlNamep;
};

struct Vrv32_cov__cuda final {
    typedef Vrv32_cov___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 4876,
        OFF_resetn = 4877,
        OFF_trap = 4878,
        OFF_mem_valid = 4879,
        OFF_mem_instr = 4880,
        OFF_mem_addr = 40,
        OFF_mem_wdata = 44,
        OFF_mem_wstrb = 4881,
        OFF_mem_rdata = 48,
        OFF_retired = 52,
        OFF_rv32_soc__DOT____VlemCond_10 = 4882,
        OFF_rv32_soc__DOT____VlemCond_6 = 4883,
        OFF_rv32_soc__DOT__state = 4884,
        OFF_rv32_soc__DOT__is_jalr = 4885,
        OFF_rv32_soc__DOT__is_branch = 4886,
        OFF_rv32_soc__DOT__is_load = 4887,
        OFF_rv32_soc__DOT__is_store = 4888,
        OFF_rv32_soc__DOT__is_alui = 4889,
        OFF_rv32_soc__DOT__is_alur = 4890,
        OFF_rv32_soc__DOT__take_branch = 4891,
        OFF_rv32_soc__DOT__ls_misaligned = 4892,
        OFF_rv32_soc__DOT__st_strb = 4893,
        OFF_rv32_soc__DOT__pc_bad = 4894,
        OFF_rv32_soc__DOT____Vtogcov__clk = 4895,
        OFF_rv32_soc__DOT____Vtogcov__resetn = 4896,
        OFF_rv32_soc__DOT____Vtogcov__trap = 4897,
        OFF_rv32_soc__DOT____Vtogcov__mem_valid = 4898,
        OFF_rv32_soc__DOT____Vtogcov__mem_instr = 4899,
        OFF_rv32_soc__DOT____Vtogcov__mem_wstrb = 4900,
        OFF_rv32_soc__DOT____Vtogcov__state = 4901,
        OFF_rv32_soc__DOT____Vtogcov__opcode = 4902,
        OFF_rv32_soc__DOT____Vtogcov__rd = 4903,
        OFF_rv32_soc__DOT____Vtogcov__funct3 = 4904,
        OFF_rv32_soc__DOT____Vtogcov__rs1 = 4905,
        OFF_rv32_soc__DOT____Vtogcov__rs2 = 4906,
        OFF_rv32_soc__DOT____Vtogcov__f7b5 = 4907,
        OFF_rv32_soc__DOT____Vtogcov__is_lui = 4908,
        OFF_rv32_soc__DOT____Vtogcov__is_auipc = 4909,
        OFF_rv32_soc__DOT____Vtogcov__is_jal = 4910,
        OFF_rv32_soc__DOT____Vtogcov__is_jalr = 4911,
        OFF_rv32_soc__DOT____Vtogcov__is_branch = 4912,
        OFF_rv32_soc__DOT____Vtogcov__is_load = 4913,
        OFF_rv32_soc__DOT____Vtogcov__is_store = 4914,
        OFF_rv32_soc__DOT____Vtogcov__is_alui = 4915,
        OFF_rv32_soc__DOT____Vtogcov__is_alur = 4916,
        OFF_rv32_soc__DOT____Vtogcov__shamt = 4917,
        OFF_rv32_soc__DOT____Vtogcov__alu_sub = 4918,
        OFF_rv32_soc__DOT____Vtogcov__take_branch = 4919,
        OFF_rv32_soc__DOT____Vtogcov__ls_misaligned = 4920,
        OFF_rv32_soc__DOT____Vtogcov__ls_oob = 4921,
        OFF_rv32_soc__DOT____Vtogcov__st_strb = 4922,
        OFF_rv32_soc__DOT____Vtogcov__pc_bad = 4923,
        OFF_rv32_soc__DOT____Vtogcov__exec_ls = 4924,
        OFF___VstlFirstIteration = 4925,
        OFF___VstlPhaseResult = 4926,
        OFF___Vtrigprevexpr___TOP__clk__0 = 4927,
        OFF___Vtrigprevexpr___TOP__resetn__0 = 4928,
        OFF___VicoDidInit = 4929,
        OFF___VicoPhaseResult = 4930,
        OFF___Vtrigprevexpr___TOP__clk__1 = 4931,
        OFF___VactPhaseResult = 4932,
        OFF___VnbaPhaseResult = 4933,
        OFF_rv32_soc__DOT____VlemCond_11 = 56,
        OFF_rv32_soc__DOT____VlemCond_9 = 60,
        OFF_rv32_soc__DOT____VlemCond_8 = 64,
        OFF_rv32_soc__DOT____VlemCond_7 = 68,
        OFF_rv32_soc__DOT____VlemCond_5 = 72,
        OFF_rv32_soc__DOT____VlemCond_4 = 76,
        OFF_rv32_soc__DOT____VlemCond_3 = 80,
        OFF_rv32_soc__DOT____VlemCond_2 = 84,
        OFF_rv32_soc__DOT____VlemCond_1 = 88,
        OFF_rv32_soc__DOT____VlemCond_0 = 92,
        OFF_rv32_soc__DOT__pc = 96,
        OFF_rv32_soc__DOT__insn = 100,
        OFF_rv32_soc__DOT__ld_addr = 104,
        OFF_rv32_soc__DOT__instret = 108,
        OFF_rv32_soc__DOT__imm_i = 112,
        OFF_rv32_soc__DOT__imm_b = 116,
        OFF_rv32_soc__DOT__imm_u = 120,
        OFF_rv32_soc__DOT__imm_j = 124,
        OFF_rv32_soc__DOT__rs1_val = 128,
        OFF_rv32_soc__DOT__alu_out = 132,
        OFF_rv32_soc__DOT__ls_addr = 136,
        OFF_rv32_soc__DOT__st_data = 140,
        OFF_rv32_soc__DOT__ld_val = 144,
        OFF_rv32_soc__DOT____Vtogcov__mem_addr = 148,
        OFF_rv32_soc__DOT____Vtogcov__mem_wdata = 152,
        OFF_rv32_soc__DOT____Vtogcov__mem_rdata = 156,
        OFF_rv32_soc__DOT____Vtogcov__retired = 160,
        OFF_rv32_soc__DOT____Vtogcov__pc = 164,
        OFF_rv32_soc__DOT____Vtogcov__insn = 168,
        OFF_rv32_soc__DOT____Vtogcov__ld_addr = 172,
        OFF_rv32_soc__DOT____Vtogcov__imm_i = 176,
        OFF_rv32_soc__DOT____Vtogcov__imm_s = 180,
        OFF_rv32_soc__DOT____Vtogcov__imm_b = 184,
        OFF_rv32_soc__DOT____Vtogcov__imm_u = 188,
        OFF_rv32_soc__DOT____Vtogcov__imm_j = 192,
        OFF_rv32_soc__DOT____Vtogcov__rs1_val = 196,
        OFF_rv32_soc__DOT____Vtogcov__rs2_val = 200,
        OFF_rv32_soc__DOT____Vtogcov__alu_b = 204,
        OFF_rv32_soc__DOT____Vtogcov__alu_out = 208,
        OFF_rv32_soc__DOT____Vtogcov__ls_addr = 212,
        OFF_rv32_soc__DOT____Vtogcov__st_data = 216,
        OFF_rv32_soc__DOT____Vtogcov__ld_word = 220,
        OFF_rv32_soc__DOT____Vtogcov__ld_shift = 224,
        OFF_rv32_soc__DOT____Vtogcov__ld_val = 228,
        OFF___VactIterCount = 232,
        OFF_rv32_soc__DOT__mem = 236,
        OFF_rv32_soc__DOT__regs = 4332,
        OFF___VstlTriggered = 0,
        OFF___VicoTriggered = 8,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        OFF___Vcoverage = 4460,
        kStateBytes = 4936,  // S_persist
        kTouchedReadBytes = 336,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 322,  // S_state_wr
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
        s.__Vcoverage.bind(st, OFF___Vcoverage, i);
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
        s.rv32_soc__DOT____VlemCond_10 = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____VlemCond_10, i);
        s.rv32_soc__DOT____VlemCond_6 = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____VlemCond_6, i);
        s.rv32_soc__DOT__state = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__state, i);
        s.rv32_soc__DOT__is_jalr = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__is_jalr, i);
        s.rv32_soc__DOT__is_branch = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__is_branch, i);
        s.rv32_soc__DOT__is_load = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__is_load, i);
        s.rv32_soc__DOT__is_store = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__is_store, i);
        s.rv32_soc__DOT__is_alui = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__is_alui, i);
        s.rv32_soc__DOT__is_alur = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__is_alur, i);
        s.rv32_soc__DOT__take_branch = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__take_branch, i);
        s.rv32_soc__DOT__ls_misaligned = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__ls_misaligned, i);
        s.rv32_soc__DOT__st_strb = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__st_strb, i);
        s.rv32_soc__DOT__pc_bad = st.ld<VlcuMem8>(OFF_rv32_soc__DOT__pc_bad, i);
        s.rv32_soc__DOT____Vtogcov__clk = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__clk, i);
        s.rv32_soc__DOT____Vtogcov__resetn = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__resetn, i);
        s.rv32_soc__DOT____Vtogcov__trap = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__trap, i);
        s.rv32_soc__DOT____Vtogcov__mem_valid = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__mem_valid, i);
        s.rv32_soc__DOT____Vtogcov__mem_instr = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__mem_instr, i);
        s.rv32_soc__DOT____Vtogcov__mem_wstrb = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__mem_wstrb, i);
        s.rv32_soc__DOT____Vtogcov__state = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__state, i);
        s.rv32_soc__DOT____Vtogcov__opcode = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__opcode, i);
        s.rv32_soc__DOT____Vtogcov__rd = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__rd, i);
        s.rv32_soc__DOT____Vtogcov__funct3 = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__funct3, i);
        s.rv32_soc__DOT____Vtogcov__rs1 = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__rs1, i);
        s.rv32_soc__DOT____Vtogcov__rs2 = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__rs2, i);
        s.rv32_soc__DOT____Vtogcov__f7b5 = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__f7b5, i);
        s.rv32_soc__DOT____Vtogcov__is_lui = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_lui, i);
        s.rv32_soc__DOT____Vtogcov__is_auipc = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_auipc, i);
        s.rv32_soc__DOT____Vtogcov__is_jal = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_jal, i);
        s.rv32_soc__DOT____Vtogcov__is_jalr = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_jalr, i);
        s.rv32_soc__DOT____Vtogcov__is_branch = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_branch, i);
        s.rv32_soc__DOT____Vtogcov__is_load = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_load, i);
        s.rv32_soc__DOT____Vtogcov__is_store = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_store, i);
        s.rv32_soc__DOT____Vtogcov__is_alui = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_alui, i);
        s.rv32_soc__DOT____Vtogcov__is_alur = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_alur, i);
        s.rv32_soc__DOT____Vtogcov__shamt = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__shamt, i);
        s.rv32_soc__DOT____Vtogcov__alu_sub = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__alu_sub, i);
        s.rv32_soc__DOT____Vtogcov__take_branch = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__take_branch, i);
        s.rv32_soc__DOT____Vtogcov__ls_misaligned = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__ls_misaligned, i);
        s.rv32_soc__DOT____Vtogcov__ls_oob = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__ls_oob, i);
        s.rv32_soc__DOT____Vtogcov__st_strb = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__st_strb, i);
        s.rv32_soc__DOT____Vtogcov__pc_bad = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__pc_bad, i);
        s.rv32_soc__DOT____Vtogcov__exec_ls = st.ld<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__exec_ls, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__resetn__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__resetn__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.rv32_soc__DOT____VlemCond_11 = st.ld<IData>(OFF_rv32_soc__DOT____VlemCond_11, i);
        s.rv32_soc__DOT____VlemCond_9 = st.ld<IData>(OFF_rv32_soc__DOT____VlemCond_9, i);
        s.rv32_soc__DOT____VlemCond_8 = st.ld<IData>(OFF_rv32_soc__DOT____VlemCond_8, i);
        s.rv32_soc__DOT____VlemCond_7 = st.ld<IData>(OFF_rv32_soc__DOT____VlemCond_7, i);
        s.rv32_soc__DOT____VlemCond_5 = st.ld<IData>(OFF_rv32_soc__DOT____VlemCond_5, i);
        s.rv32_soc__DOT____VlemCond_4 = st.ld<IData>(OFF_rv32_soc__DOT____VlemCond_4, i);
        s.rv32_soc__DOT____VlemCond_3 = st.ld<IData>(OFF_rv32_soc__DOT____VlemCond_3, i);
        s.rv32_soc__DOT____VlemCond_2 = st.ld<IData>(OFF_rv32_soc__DOT____VlemCond_2, i);
        s.rv32_soc__DOT____VlemCond_1 = st.ld<IData>(OFF_rv32_soc__DOT____VlemCond_1, i);
        s.rv32_soc__DOT____VlemCond_0 = st.ld<IData>(OFF_rv32_soc__DOT____VlemCond_0, i);
        s.rv32_soc__DOT__pc = st.ld<IData>(OFF_rv32_soc__DOT__pc, i);
        s.rv32_soc__DOT__insn = st.ld<IData>(OFF_rv32_soc__DOT__insn, i);
        s.rv32_soc__DOT__ld_addr = st.ld<IData>(OFF_rv32_soc__DOT__ld_addr, i);
        s.rv32_soc__DOT__instret = st.ld<IData>(OFF_rv32_soc__DOT__instret, i);
        s.rv32_soc__DOT__imm_i = st.ld<IData>(OFF_rv32_soc__DOT__imm_i, i);
        s.rv32_soc__DOT__imm_b = st.ld<IData>(OFF_rv32_soc__DOT__imm_b, i);
        s.rv32_soc__DOT__imm_u = st.ld<IData>(OFF_rv32_soc__DOT__imm_u, i);
        s.rv32_soc__DOT__imm_j = st.ld<IData>(OFF_rv32_soc__DOT__imm_j, i);
        s.rv32_soc__DOT__rs1_val = st.ld<IData>(OFF_rv32_soc__DOT__rs1_val, i);
        s.rv32_soc__DOT__alu_out = st.ld<IData>(OFF_rv32_soc__DOT__alu_out, i);
        s.rv32_soc__DOT__ls_addr = st.ld<IData>(OFF_rv32_soc__DOT__ls_addr, i);
        s.rv32_soc__DOT__st_data = st.ld<IData>(OFF_rv32_soc__DOT__st_data, i);
        s.rv32_soc__DOT__ld_val = st.ld<IData>(OFF_rv32_soc__DOT__ld_val, i);
        s.rv32_soc__DOT____Vtogcov__mem_addr = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__mem_addr, i);
        s.rv32_soc__DOT____Vtogcov__mem_wdata = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__mem_wdata, i);
        s.rv32_soc__DOT____Vtogcov__mem_rdata = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__mem_rdata, i);
        s.rv32_soc__DOT____Vtogcov__retired = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__retired, i);
        s.rv32_soc__DOT____Vtogcov__pc = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__pc, i);
        s.rv32_soc__DOT____Vtogcov__insn = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__insn, i);
        s.rv32_soc__DOT____Vtogcov__ld_addr = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__ld_addr, i);
        s.rv32_soc__DOT____Vtogcov__imm_i = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__imm_i, i);
        s.rv32_soc__DOT____Vtogcov__imm_s = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__imm_s, i);
        s.rv32_soc__DOT____Vtogcov__imm_b = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__imm_b, i);
        s.rv32_soc__DOT____Vtogcov__imm_u = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__imm_u, i);
        s.rv32_soc__DOT____Vtogcov__imm_j = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__imm_j, i);
        s.rv32_soc__DOT____Vtogcov__rs1_val = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__rs1_val, i);
        s.rv32_soc__DOT____Vtogcov__rs2_val = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__rs2_val, i);
        s.rv32_soc__DOT____Vtogcov__alu_b = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__alu_b, i);
        s.rv32_soc__DOT____Vtogcov__alu_out = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__alu_out, i);
        s.rv32_soc__DOT____Vtogcov__ls_addr = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__ls_addr, i);
        s.rv32_soc__DOT____Vtogcov__st_data = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__st_data, i);
        s.rv32_soc__DOT____Vtogcov__ld_word = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__ld_word, i);
        s.rv32_soc__DOT____Vtogcov__ld_shift = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__ld_shift, i);
        s.rv32_soc__DOT____Vtogcov__ld_val = st.ld<IData>(OFF_rv32_soc__DOT____Vtogcov__ld_val, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 32; ++e) s.rv32_soc__DOT__regs[e] = st.ldE<IData>(OFF_rv32_soc__DOT__regs, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
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
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____VlemCond_10, i, s.rv32_soc__DOT____VlemCond_10);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____VlemCond_6, i, s.rv32_soc__DOT____VlemCond_6);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__state, i, s.rv32_soc__DOT__state);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__is_jalr, i, s.rv32_soc__DOT__is_jalr);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__is_branch, i, s.rv32_soc__DOT__is_branch);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__is_load, i, s.rv32_soc__DOT__is_load);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__is_store, i, s.rv32_soc__DOT__is_store);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__is_alui, i, s.rv32_soc__DOT__is_alui);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__is_alur, i, s.rv32_soc__DOT__is_alur);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__take_branch, i, s.rv32_soc__DOT__take_branch);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__ls_misaligned, i, s.rv32_soc__DOT__ls_misaligned);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__st_strb, i, s.rv32_soc__DOT__st_strb);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT__pc_bad, i, s.rv32_soc__DOT__pc_bad);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__clk, i, s.rv32_soc__DOT____Vtogcov__clk);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__resetn, i, s.rv32_soc__DOT____Vtogcov__resetn);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__trap, i, s.rv32_soc__DOT____Vtogcov__trap);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__mem_valid, i, s.rv32_soc__DOT____Vtogcov__mem_valid);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__mem_instr, i, s.rv32_soc__DOT____Vtogcov__mem_instr);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__mem_wstrb, i, s.rv32_soc__DOT____Vtogcov__mem_wstrb);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__state, i, s.rv32_soc__DOT____Vtogcov__state);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__opcode, i, s.rv32_soc__DOT____Vtogcov__opcode);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__rd, i, s.rv32_soc__DOT____Vtogcov__rd);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__funct3, i, s.rv32_soc__DOT____Vtogcov__funct3);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__rs1, i, s.rv32_soc__DOT____Vtogcov__rs1);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__rs2, i, s.rv32_soc__DOT____Vtogcov__rs2);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__f7b5, i, s.rv32_soc__DOT____Vtogcov__f7b5);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_lui, i, s.rv32_soc__DOT____Vtogcov__is_lui);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_auipc, i, s.rv32_soc__DOT____Vtogcov__is_auipc);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_jal, i, s.rv32_soc__DOT____Vtogcov__is_jal);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_jalr, i, s.rv32_soc__DOT____Vtogcov__is_jalr);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_branch, i, s.rv32_soc__DOT____Vtogcov__is_branch);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_load, i, s.rv32_soc__DOT____Vtogcov__is_load);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_store, i, s.rv32_soc__DOT____Vtogcov__is_store);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_alui, i, s.rv32_soc__DOT____Vtogcov__is_alui);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__is_alur, i, s.rv32_soc__DOT____Vtogcov__is_alur);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__shamt, i, s.rv32_soc__DOT____Vtogcov__shamt);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__alu_sub, i, s.rv32_soc__DOT____Vtogcov__alu_sub);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__take_branch, i, s.rv32_soc__DOT____Vtogcov__take_branch);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__ls_misaligned, i, s.rv32_soc__DOT____Vtogcov__ls_misaligned);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__ls_oob, i, s.rv32_soc__DOT____Vtogcov__ls_oob);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__st_strb, i, s.rv32_soc__DOT____Vtogcov__st_strb);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__pc_bad, i, s.rv32_soc__DOT____Vtogcov__pc_bad);
        st.st<VlcuMem8>(OFF_rv32_soc__DOT____Vtogcov__exec_ls, i, s.rv32_soc__DOT____Vtogcov__exec_ls);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__resetn__0, i, s.__Vtrigprevexpr___TOP__resetn__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_rv32_soc__DOT____VlemCond_11, i, s.rv32_soc__DOT____VlemCond_11);
        st.st<IData>(OFF_rv32_soc__DOT____VlemCond_9, i, s.rv32_soc__DOT____VlemCond_9);
        st.st<IData>(OFF_rv32_soc__DOT____VlemCond_8, i, s.rv32_soc__DOT____VlemCond_8);
        st.st<IData>(OFF_rv32_soc__DOT____VlemCond_7, i, s.rv32_soc__DOT____VlemCond_7);
        st.st<IData>(OFF_rv32_soc__DOT____VlemCond_5, i, s.rv32_soc__DOT____VlemCond_5);
        st.st<IData>(OFF_rv32_soc__DOT____VlemCond_4, i, s.rv32_soc__DOT____VlemCond_4);
        st.st<IData>(OFF_rv32_soc__DOT____VlemCond_3, i, s.rv32_soc__DOT____VlemCond_3);
        st.st<IData>(OFF_rv32_soc__DOT____VlemCond_2, i, s.rv32_soc__DOT____VlemCond_2);
        st.st<IData>(OFF_rv32_soc__DOT____VlemCond_1, i, s.rv32_soc__DOT____VlemCond_1);
        st.st<IData>(OFF_rv32_soc__DOT____VlemCond_0, i, s.rv32_soc__DOT____VlemCond_0);
        st.st<IData>(OFF_rv32_soc__DOT__pc, i, s.rv32_soc__DOT__pc);
        st.st<IData>(OFF_rv32_soc__DOT__insn, i, s.rv32_soc__DOT__insn);
        st.st<IData>(OFF_rv32_soc__DOT__ld_addr, i, s.rv32_soc__DOT__ld_addr);
        st.st<IData>(OFF_rv32_soc__DOT__instret, i, s.rv32_soc__DOT__instret);
        st.st<IData>(OFF_rv32_soc__DOT__imm_i, i, s.rv32_soc__DOT__imm_i);
        st.st<IData>(OFF_rv32_soc__DOT__imm_b, i, s.rv32_soc__DOT__imm_b);
        st.st<IData>(OFF_rv32_soc__DOT__imm_u, i, s.rv32_soc__DOT__imm_u);
        st.st<IData>(OFF_rv32_soc__DOT__imm_j, i, s.rv32_soc__DOT__imm_j);
        st.st<IData>(OFF_rv32_soc__DOT__rs1_val, i, s.rv32_soc__DOT__rs1_val);
        st.st<IData>(OFF_rv32_soc__DOT__alu_out, i, s.rv32_soc__DOT__alu_out);
        st.st<IData>(OFF_rv32_soc__DOT__ls_addr, i, s.rv32_soc__DOT__ls_addr);
        st.st<IData>(OFF_rv32_soc__DOT__st_data, i, s.rv32_soc__DOT__st_data);
        st.st<IData>(OFF_rv32_soc__DOT__ld_val, i, s.rv32_soc__DOT__ld_val);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__mem_addr, i, s.rv32_soc__DOT____Vtogcov__mem_addr);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__mem_wdata, i, s.rv32_soc__DOT____Vtogcov__mem_wdata);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__mem_rdata, i, s.rv32_soc__DOT____Vtogcov__mem_rdata);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__retired, i, s.rv32_soc__DOT____Vtogcov__retired);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__pc, i, s.rv32_soc__DOT____Vtogcov__pc);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__insn, i, s.rv32_soc__DOT____Vtogcov__insn);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__ld_addr, i, s.rv32_soc__DOT____Vtogcov__ld_addr);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__imm_i, i, s.rv32_soc__DOT____Vtogcov__imm_i);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__imm_s, i, s.rv32_soc__DOT____Vtogcov__imm_s);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__imm_b, i, s.rv32_soc__DOT____Vtogcov__imm_b);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__imm_u, i, s.rv32_soc__DOT____Vtogcov__imm_u);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__imm_j, i, s.rv32_soc__DOT____Vtogcov__imm_j);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__rs1_val, i, s.rv32_soc__DOT____Vtogcov__rs1_val);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__rs2_val, i, s.rv32_soc__DOT____Vtogcov__rs2_val);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__alu_b, i, s.rv32_soc__DOT____Vtogcov__alu_b);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__alu_out, i, s.rv32_soc__DOT____Vtogcov__alu_out);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__ls_addr, i, s.rv32_soc__DOT____Vtogcov__ls_addr);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__st_data, i, s.rv32_soc__DOT____Vtogcov__st_data);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__ld_word, i, s.rv32_soc__DOT____Vtogcov__ld_word);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__ld_shift, i, s.rv32_soc__DOT____Vtogcov__ld_shift);
        st.st<IData>(OFF_rv32_soc__DOT____Vtogcov__ld_val, i, s.rv32_soc__DOT____Vtogcov__ld_val);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 32; ++e) st.stE<IData>(OFF_rv32_soc__DOT__regs, e, i, s.rv32_soc__DOT__regs[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
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

    VLCU_D static void Vrv32_cov___024root___eval_static(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__resetn__0 
            = vlSelfRef.resetn;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vrv32_cov___024root___eval_initial(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrv32_cov___024root___eval_final(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrv32_cov___024root___eval_settle(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vrv32_cov___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 16, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vrv32_cov___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vrv32_cov___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vrv32_cov___024root___stl_sequent__TOP__0(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ rv32_soc__DOT__rs2_val;
        rv32_soc__DOT__rs2_val = 0;
        IData/*31:0*/ rv32_soc__DOT__alu_b;
        rv32_soc__DOT__alu_b = 0;
        CData/*0:0*/ rv32_soc__DOT__alu_sub;
        rv32_soc__DOT__alu_sub = 0;
        IData/*31:0*/ rv32_soc__DOT__ld_shift;
        rv32_soc__DOT__ld_shift = 0;
        CData/*0:0*/ rv32_soc__DOT__exec_ls;
        rv32_soc__DOT__exec_ls = 0;
        CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_0;
        __VdfgRegularize_h6e95ff9d_0_0 = 0;
        // Body
        if ((IData)((0x00000067U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[850]);
        }
        if ((0U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[851]);
        }
        if ((0x67U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            ++(vlSelf->__Vcoverage[852]);
        }
        if ((((0x63U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn)) 
              & (2U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU)))) & 
             (3U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                           >> 0x0000000cU))))) {
            ++(vlSelf->__Vcoverage[855]);
        }
        if ((3U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[856]);
        }
        if ((2U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[857]);
        }
        if ((0x63U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            ++(vlSelf->__Vcoverage[858]);
        }
        if ((IData)((0x00005003U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[861]);
        }
        if ((IData)((0x00004003U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[862]);
        }
        if ((IData)((0x00002003U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[863]);
        }
        if ((IData)((0x00001003U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[864]);
        }
        if ((IData)((3U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[865]);
        }
        if ((((((0U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU))) & (1U 
                                                   != 
                                                   (7U 
                                                    & (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000000cU)))) 
               & (2U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                               >> 0x0000000cU)))) & 
              (4U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                            >> 0x0000000cU)))) & (5U 
                                                  != 
                                                  (7U 
                                                   & (vlSelfRef.rv32_soc__DOT__insn 
                                                      >> 0x0000000cU))))) {
            ++(vlSelf->__Vcoverage[866]);
        }
        if ((3U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            ++(vlSelf->__Vcoverage[867]);
        }
        if ((IData)((0x00002023U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[870]);
        }
        if ((IData)((0x00001023U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[871]);
        }
        if ((IData)((0x00000023U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[872]);
        }
        if ((((0U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                            >> 0x0000000cU))) & (1U 
                                                 != 
                                                 (7U 
                                                  & (vlSelfRef.rv32_soc__DOT__insn 
                                                     >> 0x0000000cU)))) 
             & (2U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                             >> 0x0000000cU))))) {
            ++(vlSelf->__Vcoverage[873]);
        }
        if ((0x23U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            ++(vlSelf->__Vcoverage[874]);
        }
        if ((IData)((0x00000013U == (0xfe00007fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[878]);
        }
        if ((IData)(((0x00000013U == (0xfe00007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                     & (5U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                     >> 0x0000000cU)))))) {
            ++(vlSelf->__Vcoverage[879]);
        }
        if ((IData)(((0x40000013U == (0xfe00007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                     & (1U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                     >> 0x0000000cU)))))) {
            ++(vlSelf->__Vcoverage[880]);
        }
        if ((IData)(((0x00000013U == (0xfe00007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                     & (1U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                     >> 0x0000000cU)))))) {
            ++(vlSelf->__Vcoverage[881]);
        }
        if ((((0x13U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn)) 
              & (1U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU)))) & 
             (5U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                           >> 0x0000000cU))))) {
            ++(vlSelf->__Vcoverage[882]);
        }
        if ((IData)(((0x00005000U == (0x00007000U & vlSelfRef.rv32_soc__DOT__insn)) 
                     & ((0U != (vlSelfRef.rv32_soc__DOT__insn 
                                >> 0x19U)) & (0x20U 
                                              != (vlSelfRef.rv32_soc__DOT__insn 
                                                  >> 0x19U)))))) {
            ++(vlSelf->__Vcoverage[883]);
        }
        if ((IData)(((0x00001000U == (0x00007000U & vlSelfRef.rv32_soc__DOT__insn)) 
                     & (0U != (vlSelfRef.rv32_soc__DOT__insn 
                               >> 0x19U))))) {
            ++(vlSelf->__Vcoverage[884]);
        }
        if ((0x13U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            ++(vlSelf->__Vcoverage[885]);
        }
        if ((IData)((0x40005033U == (0xfe00707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[888]);
        }
        if ((IData)((0x40000033U == (0xfe00707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[889]);
        }
        if ((IData)((0x00000033U == (0xfe00007fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            ++(vlSelf->__Vcoverage[890]);
        }
        if ((((0U != (vlSelfRef.rv32_soc__DOT__insn 
                      >> 0x19U)) & (0U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000000cU)))) 
             & (5U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                             >> 0x0000000cU))))) {
            ++(vlSelf->__Vcoverage[891]);
        }
        if (((0U != (vlSelfRef.rv32_soc__DOT__insn 
                     >> 0x19U)) & (0x20U != (vlSelfRef.rv32_soc__DOT__insn 
                                             >> 0x19U)))) {
            ++(vlSelf->__Vcoverage[892]);
        }
        if ((0x33U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            ++(vlSelf->__Vcoverage[893]);
        }
        if ((0x00000400U <= (vlSelfRef.rv32_soc__DOT__pc 
                             >> 2U))) {
            ++(vlSelf->__Vcoverage[1567]);
        }
        if ((0U != (3U & vlSelfRef.rv32_soc__DOT__pc))) {
            ++(vlSelf->__Vcoverage[1568]);
        }
        if (((0U == (3U & vlSelfRef.rv32_soc__DOT__pc)) 
             & (0x00000400U > (vlSelfRef.rv32_soc__DOT__pc 
                               >> 2U)))) {
            ++(vlSelf->__Vcoverage[1569]);
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.clk) ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__clk)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 0, vlSelfRef.clk, vlSelfRef.rv32_soc__DOT____Vtogcov__clk);
            vlSelfRef.rv32_soc__DOT____Vtogcov__clk 
                = vlSelfRef.clk;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.resetn) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__resetn)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 2, vlSelfRef.resetn, vlSelfRef.rv32_soc__DOT____Vtogcov__resetn);
            vlSelfRef.rv32_soc__DOT____Vtogcov__resetn 
                = vlSelfRef.resetn;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__instret 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__retired)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 210, vlSelfRef.rv32_soc__DOT__instret, vlSelfRef.rv32_soc__DOT____Vtogcov__retired);
            vlSelfRef.rv32_soc__DOT____Vtogcov__retired 
                = vlSelfRef.rv32_soc__DOT__instret;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__pc 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__pc)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 274, vlSelfRef.rv32_soc__DOT__pc, vlSelfRef.rv32_soc__DOT____Vtogcov__pc);
            vlSelfRef.rv32_soc__DOT____Vtogcov__pc 
                = vlSelfRef.rv32_soc__DOT__pc;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__insn 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__insn)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 338, vlSelfRef.rv32_soc__DOT__insn, vlSelfRef.rv32_soc__DOT____Vtogcov__insn);
            vlSelfRef.rv32_soc__DOT____Vtogcov__insn 
                = vlSelfRef.rv32_soc__DOT__insn;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__ld_addr 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__ld_addr)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 402, vlSelfRef.rv32_soc__DOT__ld_addr, vlSelfRef.rv32_soc__DOT____Vtogcov__ld_addr);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ld_addr 
                = vlSelfRef.rv32_soc__DOT__ld_addr;
        }
        if (vlcu_cov_any((0U != ((IData)(vlSelfRef.rv32_soc__DOT__state) 
                                 ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__state))))) {
            VL_COV_TOGGLE_CHG_ST_I(2, vlSelf->__Vcoverage + 466, vlSelfRef.rv32_soc__DOT__state, vlSelfRef.rv32_soc__DOT____Vtogcov__state);
            vlSelfRef.rv32_soc__DOT____Vtogcov__state 
                = vlSelfRef.rv32_soc__DOT__state;
        }
        if (vlcu_cov_any((0U != (0x0000007fU & (vlSelfRef.rv32_soc__DOT__insn 
                                                ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__opcode)))))) {
            VL_COV_TOGGLE_CHG_ST_I(7, vlSelf->__Vcoverage + 470, vlSelfRef.rv32_soc__DOT__insn, vlSelfRef.rv32_soc__DOT____Vtogcov__opcode);
            vlSelfRef.rv32_soc__DOT____Vtogcov__opcode 
                = (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn);
        }
        if (vlcu_cov_any((0U != (0x0000001fU & ((vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 7U) 
                                                ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__rd)))))) {
            VL_COV_TOGGLE_CHG_ST_I(5, vlSelf->__Vcoverage + 484, 
                                   (vlSelfRef.rv32_soc__DOT__insn 
                                    >> 7U), vlSelfRef.rv32_soc__DOT____Vtogcov__rd);
            vlSelfRef.rv32_soc__DOT____Vtogcov__rd 
                = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                  >> 7U));
        }
        if (vlcu_cov_any((0U != (7U & ((vlSelfRef.rv32_soc__DOT__insn 
                                        >> 0x0000000cU) 
                                       ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__funct3)))))) {
            VL_COV_TOGGLE_CHG_ST_I(3, vlSelf->__Vcoverage + 494, 
                                   (vlSelfRef.rv32_soc__DOT__insn 
                                    >> 0x0000000cU), vlSelfRef.rv32_soc__DOT____Vtogcov__funct3);
            vlSelfRef.rv32_soc__DOT____Vtogcov__funct3 
                = (7U & (vlSelfRef.rv32_soc__DOT__insn 
                         >> 0x0000000cU));
        }
        if (vlcu_cov_any((0U != (0x0000001fU & ((vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000000fU) 
                                                ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__rs1)))))) {
            VL_COV_TOGGLE_CHG_ST_I(5, vlSelf->__Vcoverage + 500, 
                                   (vlSelfRef.rv32_soc__DOT__insn 
                                    >> 0x0000000fU), vlSelfRef.rv32_soc__DOT____Vtogcov__rs1);
            vlSelfRef.rv32_soc__DOT____Vtogcov__rs1 
                = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                  >> 0x0000000fU));
        }
        if (vlcu_cov_any((0U != (0x0000001fU & ((vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x00000014U) 
                                                ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__rs2)))))) {
            VL_COV_TOGGLE_CHG_ST_I(5, vlSelf->__Vcoverage + 510, 
                                   (vlSelfRef.rv32_soc__DOT__insn 
                                    >> 0x00000014U), vlSelfRef.rv32_soc__DOT____Vtogcov__rs2);
            vlSelfRef.rv32_soc__DOT____Vtogcov__rs2 
                = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                  >> 0x00000014U));
        }
        if (vlcu_cov_any((1U & ((vlSelfRef.rv32_soc__DOT__insn 
                                 >> 0x0000001eU) ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__f7b5))))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 520, 
                                   (vlSelfRef.rv32_soc__DOT__insn 
                                    >> 0x0000001eU), vlSelfRef.rv32_soc__DOT____Vtogcov__f7b5);
            vlSelfRef.rv32_soc__DOT____Vtogcov__f7b5 
                = (1U & (vlSelfRef.rv32_soc__DOT__insn 
                         >> 0x0000001eU));
        }
        if (vlcu_cov_any((0U != ((((- (IData)((vlSelfRef.rv32_soc__DOT__insn 
                                               >> 0x0000001fU))) 
                                   << 0x0000000cU) 
                                  | ((0x00000fe0U & 
                                      (vlSelfRef.rv32_soc__DOT__insn 
                                       >> 0x00000014U)) 
                                     | (0x0000001fU 
                                        & (vlSelfRef.rv32_soc__DOT__insn 
                                           >> 7U)))) 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__imm_s)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 586, 
                                   (((- (IData)((vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000001fU))) 
                                     << 0x0000000cU) 
                                    | ((0x00000fe0U 
                                        & (vlSelfRef.rv32_soc__DOT__insn 
                                           >> 0x00000014U)) 
                                       | (0x0000001fU 
                                          & (vlSelfRef.rv32_soc__DOT__insn 
                                             >> 7U)))), vlSelfRef.rv32_soc__DOT____Vtogcov__imm_s);
            vlSelfRef.rv32_soc__DOT____Vtogcov__imm_s 
                = (((- (IData)((vlSelfRef.rv32_soc__DOT__insn 
                                >> 0x0000001fU))) << 0x0000000cU) 
                   | ((0x00000fe0U & (vlSelfRef.rv32_soc__DOT__insn 
                                      >> 0x00000014U)) 
                      | (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                        >> 7U))));
        }
        if (vlcu_cov_any((0U != ((0xfffff000U & vlSelfRef.rv32_soc__DOT__insn) 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__imm_u)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 714, 
                                   (0xfffff000U & vlSelfRef.rv32_soc__DOT__insn), vlSelfRef.rv32_soc__DOT____Vtogcov__imm_u);
            vlSelfRef.rv32_soc__DOT____Vtogcov__imm_u 
                = (0xfffff000U & vlSelfRef.rv32_soc__DOT__insn);
        }
        if (vlcu_cov_any(((0x37U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_lui)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 842, 
                                   (0x37U == (0x0000007fU 
                                              & vlSelfRef.rv32_soc__DOT__insn)), vlSelfRef.rv32_soc__DOT____Vtogcov__is_lui);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_lui 
                = (0x37U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn));
        }
        if (vlcu_cov_any(((0x17U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_auipc)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 844, 
                                   (0x17U == (0x0000007fU 
                                              & vlSelfRef.rv32_soc__DOT__insn)), vlSelfRef.rv32_soc__DOT____Vtogcov__is_auipc);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_auipc 
                = (0x17U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn));
        }
        if (vlcu_cov_any(((0x6fU == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_jal)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 846, 
                                   (0x6fU == (0x0000007fU 
                                              & vlSelfRef.rv32_soc__DOT__insn)), vlSelfRef.rv32_soc__DOT____Vtogcov__is_jal);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_jal 
                = (0x6fU == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn));
        }
        vlSelfRef.rv32_soc__DOT__imm_u = (0xfffff000U 
                                          & vlSelfRef.rv32_soc__DOT__insn);
        vlSelfRef.retired = vlSelfRef.rv32_soc__DOT__instret;
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__mem
                                 [(0x000003ffU & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                                  >> 2U))] 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__ld_word)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1363, vlSelfRef.rv32_soc__DOT__mem
                                   [(0x000003ffU & 
                                     (vlSelfRef.rv32_soc__DOT__ld_addr 
                                      >> 2U))], vlSelfRef.rv32_soc__DOT____Vtogcov__ld_word);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ld_word 
                = vlSelfRef.rv32_soc__DOT__mem[(0x000003ffU 
                                                & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                                   >> 2U))];
        }
        vlSelfRef.trap = (3U == (IData)(vlSelfRef.rv32_soc__DOT__state));
        vlSelfRef.rv32_soc__DOT__imm_b = (((- (IData)(
                                                      (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000001fU))) 
                                           << 0x0000000dU) 
                                          | ((((2U 
                                                & (vlSelfRef.rv32_soc__DOT__insn 
                                                   >> 0x0000001eU)) 
                                               | (1U 
                                                  & (vlSelfRef.rv32_soc__DOT__insn 
                                                     >> 7U))) 
                                              << 0x0000000bU) 
                                             | ((0x000007e0U 
                                                 & (vlSelfRef.rv32_soc__DOT__insn 
                                                    >> 0x00000014U)) 
                                                | (0x0000001eU 
                                                   & (vlSelfRef.rv32_soc__DOT__insn 
                                                      >> 7U)))));
        vlSelfRef.rv32_soc__DOT__imm_j = ((((0x00000ffeU 
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
                                                   >> 0x00000014U))));
        vlSelfRef.rv32_soc__DOT__is_jalr = (IData)(
                                                   (0x00000067U 
                                                    == 
                                                    (0x0000707fU 
                                                     & vlSelfRef.rv32_soc__DOT__insn)));
        vlSelfRef.rv32_soc__DOT__is_branch = ((0x63U 
                                               == (0x0000007fU 
                                                   & vlSelfRef.rv32_soc__DOT__insn)) 
                                              & ((2U 
                                                  != 
                                                  (7U 
                                                   & (vlSelfRef.rv32_soc__DOT__insn 
                                                      >> 0x0000000cU))) 
                                                 & (3U 
                                                    != 
                                                    (7U 
                                                     & (vlSelfRef.rv32_soc__DOT__insn 
                                                        >> 0x0000000cU)))));
        vlSelfRef.rv32_soc__DOT__is_alui = ((0x13U 
                                             == (0x0000007fU 
                                                 & vlSelfRef.rv32_soc__DOT__insn)) 
                                            & ((0U 
                                                == 
                                                (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x00000019U)) 
                                               | ((1U 
                                                   != 
                                                   (7U 
                                                    & (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000000cU))) 
                                                  & (IData)(
                                                            ((0x00005000U 
                                                              != 
                                                              (0x00007000U 
                                                               & vlSelfRef.rv32_soc__DOT__insn)) 
                                                             | (0x20U 
                                                                == 
                                                                (vlSelfRef.rv32_soc__DOT__insn 
                                                                 >> 0x00000019U)))))));
        rv32_soc__DOT__ld_shift = (vlSelfRef.rv32_soc__DOT__mem
                                   [(0x000003ffU & 
                                     (vlSelfRef.rv32_soc__DOT__ld_addr 
                                      >> 2U))] >> (0x00000018U 
                                                   & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                                      << 3U)));
        vlSelfRef.rv32_soc__DOT__imm_i = (((- (IData)(
                                                      (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000001fU))) 
                                           << 0x0000000cU) 
                                          | (vlSelfRef.rv32_soc__DOT__insn 
                                             >> 0x00000014U));
        vlSelfRef.rv32_soc__DOT__pc_bad = ((0U != (3U 
                                                   & vlSelfRef.rv32_soc__DOT__pc)) 
                                           | (0x00000400U 
                                              <= (vlSelfRef.rv32_soc__DOT__pc 
                                                  >> 2U)));
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
        if ((0U == (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x00000014U)))) {
            ++(vlSelf->__Vcoverage[1028]);
            vlSelfRef.rv32_soc__DOT____VlemCond_1 = 0U;
        } else {
            ++(vlSelf->__Vcoverage[1029]);
            vlSelfRef.rv32_soc__DOT____VlemCond_1 = vlSelfRef.rv32_soc__DOT__regs
                [(0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                 >> 0x00000014U))];
        }
        rv32_soc__DOT__rs2_val = vlSelfRef.rv32_soc__DOT____VlemCond_1;
        if ((0U == (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x00000014U)))) {
            ++(vlSelf->__Vcoverage[1026]);
        }
        if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x00000014U)))) {
            ++(vlSelf->__Vcoverage[1027]);
        }
        if ((0U == (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x0000000fU)))) {
            ++(vlSelf->__Vcoverage[960]);
            vlSelfRef.rv32_soc__DOT____VlemCond_0 = 0U;
        } else {
            ++(vlSelf->__Vcoverage[961]);
            vlSelfRef.rv32_soc__DOT____VlemCond_0 = vlSelfRef.rv32_soc__DOT__regs
                [(0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                 >> 0x0000000fU))];
        }
        vlSelfRef.rv32_soc__DOT__rs1_val = vlSelfRef.rv32_soc__DOT____VlemCond_0;
        if ((0U == (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x0000000fU)))) {
            ++(vlSelf->__Vcoverage[958]);
        }
        if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x0000000fU)))) {
            ++(vlSelf->__Vcoverage[959]);
        }
        __VdfgRegularize_h6e95ff9d_0_0 = ((0U == (7U 
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
        if (vlcu_cov_any(((IData)(vlSelfRef.trap) ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__trap)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 4, vlSelfRef.trap, vlSelfRef.rv32_soc__DOT____Vtogcov__trap);
            vlSelfRef.rv32_soc__DOT____Vtogcov__trap 
                = vlSelfRef.trap;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__imm_b 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__imm_b)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 650, vlSelfRef.rv32_soc__DOT__imm_b, vlSelfRef.rv32_soc__DOT____Vtogcov__imm_b);
            vlSelfRef.rv32_soc__DOT____Vtogcov__imm_b 
                = vlSelfRef.rv32_soc__DOT__imm_b;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__imm_j 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__imm_j)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 778, vlSelfRef.rv32_soc__DOT__imm_j, vlSelfRef.rv32_soc__DOT____Vtogcov__imm_j);
            vlSelfRef.rv32_soc__DOT____Vtogcov__imm_j 
                = vlSelfRef.rv32_soc__DOT__imm_j;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_jalr) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_jalr)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 848, vlSelfRef.rv32_soc__DOT__is_jalr, vlSelfRef.rv32_soc__DOT____Vtogcov__is_jalr);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_jalr 
                = vlSelfRef.rv32_soc__DOT__is_jalr;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_branch) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_branch)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 853, vlSelfRef.rv32_soc__DOT__is_branch, vlSelfRef.rv32_soc__DOT____Vtogcov__is_branch);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_branch 
                = vlSelfRef.rv32_soc__DOT__is_branch;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_alui) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_alui)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 875, vlSelfRef.rv32_soc__DOT__is_alui, vlSelfRef.rv32_soc__DOT____Vtogcov__is_alui);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_alui 
                = vlSelfRef.rv32_soc__DOT__is_alui;
        }
        if (vlcu_cov_any((0U != (rv32_soc__DOT__ld_shift 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__ld_shift)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1427, rv32_soc__DOT__ld_shift, vlSelfRef.rv32_soc__DOT____Vtogcov__ld_shift);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ld_shift 
                = rv32_soc__DOT__ld_shift;
        }
        if ((0U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[1557]);
            vlSelfRef.rv32_soc__DOT____VlemCond_7 = 
                (((- (IData)((1U & (rv32_soc__DOT__ld_shift 
                                    >> 7U)))) << 8U) 
                 | (0x000000ffU & rv32_soc__DOT__ld_shift));
        } else {
            ++(vlSelf->__Vcoverage[1564]);
            if ((1U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU)))) {
                ++(vlSelf->__Vcoverage[1558]);
                vlSelfRef.rv32_soc__DOT____VlemCond_7 
                    = (((- (IData)((1U & (rv32_soc__DOT__ld_shift 
                                          >> 0x0fU)))) 
                        << 0x00000010U) | (0x0000ffffU 
                                           & rv32_soc__DOT__ld_shift));
            } else {
                ++(vlSelf->__Vcoverage[1563]);
                if ((4U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                  >> 0x0000000cU)))) {
                    ++(vlSelf->__Vcoverage[1559]);
                    vlSelfRef.rv32_soc__DOT____VlemCond_7 
                        = (0x000000ffU & rv32_soc__DOT__ld_shift);
                } else {
                    ++(vlSelf->__Vcoverage[1562]);
                    if ((5U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                      >> 0x0000000cU)))) {
                        ++(vlSelf->__Vcoverage[1560]);
                        vlSelfRef.rv32_soc__DOT____VlemCond_7 
                            = (0x0000ffffU & rv32_soc__DOT__ld_shift);
                    } else {
                        ++(vlSelf->__Vcoverage[1561]);
                        vlSelfRef.rv32_soc__DOT____VlemCond_7 
                            = vlSelfRef.rv32_soc__DOT__mem
                            [(0x000003ffU & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                             >> 2U))];
                    }
                }
            }
        }
        vlSelfRef.rv32_soc__DOT__ld_val = vlSelfRef.rv32_soc__DOT____VlemCond_7;
        if ((0U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[1555]);
        }
        if ((0U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[1556]);
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__imm_i 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__imm_i)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 522, vlSelfRef.rv32_soc__DOT__imm_i, vlSelfRef.rv32_soc__DOT____Vtogcov__imm_i);
            vlSelfRef.rv32_soc__DOT____Vtogcov__imm_i 
                = vlSelfRef.rv32_soc__DOT__imm_i;
        }
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            ++(vlSelf->__Vcoverage[1582]);
        }
        if (vlSelfRef.rv32_soc__DOT__pc_bad) {
            ++(vlSelf->__Vcoverage[1583]);
        }
        if ((0U != (IData)(vlSelfRef.rv32_soc__DOT__state))) {
            ++(vlSelf->__Vcoverage[1584]);
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__pc_bad) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__pc_bad)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1565, vlSelfRef.rv32_soc__DOT__pc_bad, vlSelfRef.rv32_soc__DOT____Vtogcov__pc_bad);
            vlSelfRef.rv32_soc__DOT____Vtogcov__pc_bad 
                = vlSelfRef.rv32_soc__DOT__pc_bad;
        }
        vlSelfRef.mem_instr = ((~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)) 
                               & (0U == (IData)(vlSelfRef.rv32_soc__DOT__state)));
        if (((IData)(vlSelfRef.rv32_soc__DOT__is_alur) 
             & (vlSelfRef.rv32_soc__DOT__insn >> 0x0000001eU))) {
            ++(vlSelf->__Vcoverage[1110]);
        }
        if ((1U & (~ (vlSelfRef.rv32_soc__DOT__insn 
                      >> 0x0000001eU)))) {
            ++(vlSelf->__Vcoverage[1111]);
        }
        if ((1U & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_alur)))) {
            ++(vlSelf->__Vcoverage[1112]);
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_alur) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_alur)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 886, vlSelfRef.rv32_soc__DOT__is_alur, vlSelfRef.rv32_soc__DOT____Vtogcov__is_alur);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_alur 
                = vlSelfRef.rv32_soc__DOT__is_alur;
        }
        rv32_soc__DOT__alu_sub = ((vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x0000001eU) 
                                  & (IData)(vlSelfRef.rv32_soc__DOT__is_alur));
        if (vlcu_cov_any((0U != (rv32_soc__DOT__rs2_val 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__rs2_val)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 962, rv32_soc__DOT__rs2_val, vlSelfRef.rv32_soc__DOT____Vtogcov__rs2_val);
            vlSelfRef.rv32_soc__DOT____Vtogcov__rs2_val 
                = rv32_soc__DOT__rs2_val;
        }
        if ((0U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[1345]);
            vlSelfRef.rv32_soc__DOT____VlemCond_5 = 
                ((rv32_soc__DOT__rs2_val << 0x00000018U) 
                 | ((0x00ff0000U & (rv32_soc__DOT__rs2_val 
                                    << 0x00000010U)) 
                    | ((0x0000ff00U & (rv32_soc__DOT__rs2_val 
                                       << 8U)) | (0x000000ffU 
                                                  & rv32_soc__DOT__rs2_val))));
        } else {
            ++(vlSelf->__Vcoverage[1348]);
            if ((1U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU)))) {
                ++(vlSelf->__Vcoverage[1346]);
                vlSelfRef.rv32_soc__DOT____VlemCond_5 
                    = ((rv32_soc__DOT__rs2_val << 0x00000010U) 
                       | (0x0000ffffU & rv32_soc__DOT__rs2_val));
            } else {
                ++(vlSelf->__Vcoverage[1347]);
                vlSelfRef.rv32_soc__DOT____VlemCond_5 
                    = rv32_soc__DOT__rs2_val;
            }
        }
        vlSelfRef.rv32_soc__DOT__st_data = vlSelfRef.rv32_soc__DOT____VlemCond_5;
        if ((0U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[1343]);
        }
        if ((0U != (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[1344]);
        }
        if (vlSelfRef.rv32_soc__DOT__is_alur) {
            ++(vlSelf->__Vcoverage[1096]);
            vlSelfRef.rv32_soc__DOT____VlemCond_2 = rv32_soc__DOT__rs2_val;
        } else {
            ++(vlSelf->__Vcoverage[1097]);
            vlSelfRef.rv32_soc__DOT____VlemCond_2 = vlSelfRef.rv32_soc__DOT__imm_i;
        }
        rv32_soc__DOT__alu_b = vlSelfRef.rv32_soc__DOT____VlemCond_2;
        if (vlSelfRef.rv32_soc__DOT__is_alur) {
            ++(vlSelf->__Vcoverage[1094]);
        }
        if ((1U & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_alur)))) {
            ++(vlSelf->__Vcoverage[1095]);
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__rs1_val 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__rs1_val)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 894, vlSelfRef.rv32_soc__DOT__rs1_val, vlSelfRef.rv32_soc__DOT____Vtogcov__rs1_val);
            vlSelfRef.rv32_soc__DOT____Vtogcov__rs1_val 
                = vlSelfRef.rv32_soc__DOT__rs1_val;
        }
        if ((0x00004000U & vlSelfRef.rv32_soc__DOT__insn)) {
            if ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)) {
                if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
                    vlSelfRef.rv32_soc__DOT__take_branch 
                        = (vlSelfRef.rv32_soc__DOT__rs1_val 
                           >= rv32_soc__DOT__rs2_val);
                    ++(vlSelf->__Vcoverage[1201]);
                } else {
                    vlSelfRef.rv32_soc__DOT__take_branch 
                        = (vlSelfRef.rv32_soc__DOT__rs1_val 
                           < rv32_soc__DOT__rs2_val);
                    ++(vlSelf->__Vcoverage[1200]);
                }
            } else if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
                vlSelfRef.rv32_soc__DOT__take_branch 
                    = VL_GTES_III(32, vlSelfRef.rv32_soc__DOT__rs1_val, rv32_soc__DOT__rs2_val);
                ++(vlSelf->__Vcoverage[1199]);
            } else {
                vlSelfRef.rv32_soc__DOT__take_branch 
                    = VL_LTS_III(32, vlSelfRef.rv32_soc__DOT__rs1_val, rv32_soc__DOT__rs2_val);
                ++(vlSelf->__Vcoverage[1198]);
            }
        } else if ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)) {
            vlSelfRef.rv32_soc__DOT__take_branch = 
                (vlSelfRef.rv32_soc__DOT__rs1_val >= rv32_soc__DOT__rs2_val);
            ++(vlSelf->__Vcoverage[1201]);
        } else if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
            vlSelfRef.rv32_soc__DOT__take_branch = 
                (vlSelfRef.rv32_soc__DOT__rs1_val != rv32_soc__DOT__rs2_val);
            ++(vlSelf->__Vcoverage[1197]);
        } else {
            vlSelfRef.rv32_soc__DOT__take_branch = 
                (vlSelfRef.rv32_soc__DOT__rs1_val == rv32_soc__DOT__rs2_val);
            ++(vlSelf->__Vcoverage[1196]);
        }
        ++(vlSelf->__Vcoverage[1202]);
        vlSelfRef.rv32_soc__DOT__is_load = ((3U == 
                                             (0x0000007fU 
                                              & vlSelfRef.rv32_soc__DOT__insn)) 
                                            & ((IData)(__VdfgRegularize_h6e95ff9d_0_0) 
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
                                             & (IData)(__VdfgRegularize_h6e95ff9d_0_0));
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__ld_val 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__ld_val)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1491, vlSelfRef.rv32_soc__DOT__ld_val, vlSelfRef.rv32_soc__DOT____Vtogcov__ld_val);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ld_val 
                = vlSelfRef.rv32_soc__DOT__ld_val;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.mem_instr) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__mem_instr)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 8, vlSelfRef.mem_instr, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_instr);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_instr 
                = vlSelfRef.mem_instr;
        }
        if (vlcu_cov_any(((IData)(rv32_soc__DOT__alu_sub) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__alu_sub)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1108, rv32_soc__DOT__alu_sub, vlSelfRef.rv32_soc__DOT____Vtogcov__alu_sub);
            vlSelfRef.rv32_soc__DOT____Vtogcov__alu_sub 
                = rv32_soc__DOT__alu_sub;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__st_data 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__st_data)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1279, vlSelfRef.rv32_soc__DOT__st_data, vlSelfRef.rv32_soc__DOT____Vtogcov__st_data);
            vlSelfRef.rv32_soc__DOT____Vtogcov__st_data 
                = vlSelfRef.rv32_soc__DOT__st_data;
        }
        if (vlcu_cov_any((0U != (rv32_soc__DOT__alu_b 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__alu_b)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1030, rv32_soc__DOT__alu_b, vlSelfRef.rv32_soc__DOT____Vtogcov__alu_b);
            vlSelfRef.rv32_soc__DOT____Vtogcov__alu_b 
                = rv32_soc__DOT__alu_b;
        }
        if (vlcu_cov_any((0U != (0x0000001fU & (rv32_soc__DOT__alu_b 
                                                ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__shamt)))))) {
            VL_COV_TOGGLE_CHG_ST_I(5, vlSelf->__Vcoverage + 1098, rv32_soc__DOT__alu_b, vlSelfRef.rv32_soc__DOT____Vtogcov__shamt);
            vlSelfRef.rv32_soc__DOT____Vtogcov__shamt 
                = (0x0000001fU & rv32_soc__DOT__alu_b);
        }
        if ((0x00004000U & vlSelfRef.rv32_soc__DOT__insn)) {
            if ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)) {
                if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
                    vlSelfRef.rv32_soc__DOT__alu_out 
                        = (vlSelfRef.rv32_soc__DOT__rs1_val 
                           & rv32_soc__DOT__alu_b);
                    ++(vlSelf->__Vcoverage[1192]);
                } else {
                    vlSelfRef.rv32_soc__DOT__alu_out 
                        = (vlSelfRef.rv32_soc__DOT__rs1_val 
                           | rv32_soc__DOT__alu_b);
                    ++(vlSelf->__Vcoverage[1191]);
                }
            } else if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
                if ((0x40000000U & vlSelfRef.rv32_soc__DOT__insn)) {
                    ++(vlSelf->__Vcoverage[1188]);
                    vlSelfRef.rv32_soc__DOT____VlemCond_4 
                        = VL_SHIFTRS_III(32,32,5, vlSelfRef.rv32_soc__DOT__rs1_val, 
                                         (0x0000001fU 
                                          & rv32_soc__DOT__alu_b));
                } else {
                    ++(vlSelf->__Vcoverage[1189]);
                    vlSelfRef.rv32_soc__DOT____VlemCond_4 
                        = (vlSelfRef.rv32_soc__DOT__rs1_val 
                           >> (0x0000001fU & rv32_soc__DOT__alu_b));
                }
                vlSelfRef.rv32_soc__DOT__alu_out = vlSelfRef.rv32_soc__DOT____VlemCond_4;
                ++(vlSelf->__Vcoverage[1190]);
            } else {
                vlSelfRef.rv32_soc__DOT__alu_out = 
                    (vlSelfRef.rv32_soc__DOT__rs1_val 
                     ^ rv32_soc__DOT__alu_b);
                ++(vlSelf->__Vcoverage[1185]);
            }
        } else if ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)) {
            if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
                vlSelfRef.rv32_soc__DOT__alu_out = 
                    (vlSelfRef.rv32_soc__DOT__rs1_val 
                     < rv32_soc__DOT__alu_b);
                ++(vlSelf->__Vcoverage[1184]);
            } else {
                vlSelfRef.rv32_soc__DOT__alu_out = 
                    VL_LTS_III(32, vlSelfRef.rv32_soc__DOT__rs1_val, rv32_soc__DOT__alu_b);
                ++(vlSelf->__Vcoverage[1183]);
            }
        } else if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
            vlSelfRef.rv32_soc__DOT__alu_out = (vlSelfRef.rv32_soc__DOT__rs1_val 
                                                << 
                                                (0x0000001fU 
                                                 & rv32_soc__DOT__alu_b));
            ++(vlSelf->__Vcoverage[1182]);
        } else {
            if (rv32_soc__DOT__alu_sub) {
                ++(vlSelf->__Vcoverage[1179]);
                vlSelfRef.rv32_soc__DOT____VlemCond_3 
                    = (vlSelfRef.rv32_soc__DOT__rs1_val 
                       - rv32_soc__DOT__alu_b);
            } else {
                ++(vlSelf->__Vcoverage[1180]);
                vlSelfRef.rv32_soc__DOT____VlemCond_3 
                    = (vlSelfRef.rv32_soc__DOT__rs1_val 
                       + rv32_soc__DOT__alu_b);
            }
            vlSelfRef.rv32_soc__DOT__alu_out = vlSelfRef.rv32_soc__DOT____VlemCond_3;
            ++(vlSelf->__Vcoverage[1181]);
        }
        if (rv32_soc__DOT__alu_sub) {
            ++(vlSelf->__Vcoverage[1177]);
        }
        if ((1U & (~ (IData)(rv32_soc__DOT__alu_sub)))) {
            ++(vlSelf->__Vcoverage[1178]);
        }
        if ((0x40000000U & vlSelfRef.rv32_soc__DOT__insn)) {
            ++(vlSelf->__Vcoverage[1186]);
        }
        if ((1U & (~ (vlSelfRef.rv32_soc__DOT__insn 
                      >> 0x0000001eU)))) {
            ++(vlSelf->__Vcoverage[1187]);
        }
        ++(vlSelf->__Vcoverage[1193]);
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__take_branch) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__take_branch)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1194, vlSelfRef.rv32_soc__DOT__take_branch, vlSelfRef.rv32_soc__DOT____Vtogcov__take_branch);
            vlSelfRef.rv32_soc__DOT____Vtogcov__take_branch 
                = vlSelfRef.rv32_soc__DOT__take_branch;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_load) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_load)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 859, vlSelfRef.rv32_soc__DOT__is_load, vlSelfRef.rv32_soc__DOT____Vtogcov__is_load);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_load 
                = vlSelfRef.rv32_soc__DOT__is_load;
        }
        if (vlSelfRef.rv32_soc__DOT__is_store) {
            ++(vlSelf->__Vcoverage[1267]);
        }
        if ((1U & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_store)))) {
            ++(vlSelf->__Vcoverage[1268]);
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_store) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_store)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 868, vlSelfRef.rv32_soc__DOT__is_store, vlSelfRef.rv32_soc__DOT____Vtogcov__is_store);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_store 
                = vlSelfRef.rv32_soc__DOT__is_store;
        }
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
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__alu_out 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__alu_out)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1113, vlSelfRef.rv32_soc__DOT__alu_out, vlSelfRef.rv32_soc__DOT____Vtogcov__alu_out);
            vlSelfRef.rv32_soc__DOT____Vtogcov__alu_out 
                = vlSelfRef.rv32_soc__DOT__alu_out;
        }
        if ((IData)(((0x00002000U == (0x00003000U & vlSelfRef.rv32_soc__DOT__insn)) 
                     & (0U != (3U & vlSelfRef.rv32_soc__DOT__ls_addr))))) {
            ++(vlSelf->__Vcoverage[1271]);
        }
        if ((IData)(((0x00001000U == (0x00003000U & vlSelfRef.rv32_soc__DOT__insn)) 
                     & vlSelfRef.rv32_soc__DOT__ls_addr))) {
            ++(vlSelf->__Vcoverage[1272]);
        }
        if ((IData)((0U == (3U & vlSelfRef.rv32_soc__DOT__ls_addr)))) {
            ++(vlSelf->__Vcoverage[1273]);
        }
        if (((~ vlSelfRef.rv32_soc__DOT__ls_addr) & 
             (2U != (3U & (vlSelfRef.rv32_soc__DOT__insn 
                           >> 0x0000000cU))))) {
            ++(vlSelf->__Vcoverage[1274]);
        }
        if (((1U != (3U & (vlSelfRef.rv32_soc__DOT__insn 
                           >> 0x0000000cU))) & (0U 
                                                == 
                                                (3U 
                                                 & vlSelfRef.rv32_soc__DOT__ls_addr)))) {
            ++(vlSelf->__Vcoverage[1275]);
        }
        if (((1U != (3U & (vlSelfRef.rv32_soc__DOT__insn 
                           >> 0x0000000cU))) & (2U 
                                                != 
                                                (3U 
                                                 & (vlSelfRef.rv32_soc__DOT__insn 
                                                    >> 0x0000000cU))))) {
            ++(vlSelf->__Vcoverage[1276]);
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__ls_addr 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__ls_addr)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1203, vlSelfRef.rv32_soc__DOT__ls_addr, vlSelfRef.rv32_soc__DOT____Vtogcov__ls_addr);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ls_addr 
                = vlSelfRef.rv32_soc__DOT__ls_addr;
        }
        if (vlcu_cov_any(((0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                           >> 2U)) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__ls_oob)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1277, 
                                   (0x00000400U <= 
                                    (vlSelfRef.rv32_soc__DOT__ls_addr 
                                     >> 2U)), vlSelfRef.rv32_soc__DOT____Vtogcov__ls_oob);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ls_oob 
                = (0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                   >> 2U));
        }
        if ((0U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[1359]);
            vlSelfRef.rv32_soc__DOT____VlemCond_6 = 
                (0x0000000fU & ((IData)(1U) << (3U 
                                                & vlSelfRef.rv32_soc__DOT__ls_addr)));
        } else {
            ++(vlSelf->__Vcoverage[1362]);
            if ((1U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU)))) {
                ++(vlSelf->__Vcoverage[1360]);
                vlSelfRef.rv32_soc__DOT____VlemCond_6 
                    = (0x0000000fU & ((IData)(3U) << 
                                      (3U & vlSelfRef.rv32_soc__DOT__ls_addr)));
            } else {
                ++(vlSelf->__Vcoverage[1361]);
                vlSelfRef.rv32_soc__DOT____VlemCond_6 = 0x0fU;
            }
        }
        vlSelfRef.rv32_soc__DOT__st_strb = vlSelfRef.rv32_soc__DOT____VlemCond_6;
        if ((0U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[1357]);
        }
        if ((0U != (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            ++(vlSelf->__Vcoverage[1358]);
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
        if (vlcu_cov_any((0U != ((IData)(vlSelfRef.rv32_soc__DOT__st_strb) 
                                 ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__st_strb))))) {
            VL_COV_TOGGLE_CHG_ST_I(4, vlSelf->__Vcoverage + 1349, vlSelfRef.rv32_soc__DOT__st_strb, vlSelfRef.rv32_soc__DOT____Vtogcov__st_strb);
            vlSelfRef.rv32_soc__DOT____Vtogcov__st_strb 
                = vlSelfRef.rv32_soc__DOT__st_strb;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__ls_misaligned)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1269, vlSelfRef.rv32_soc__DOT__ls_misaligned, vlSelfRef.rv32_soc__DOT____Vtogcov__ls_misaligned);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ls_misaligned 
                = vlSelfRef.rv32_soc__DOT__ls_misaligned;
        }
        if (((((1U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
               & (IData)(vlSelfRef.rv32_soc__DOT__is_store)) 
              & (~ (IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned))) 
             & (0x00000400U > (vlSelfRef.rv32_soc__DOT__ls_addr 
                               >> 2U)))) {
            ++(vlSelf->__Vcoverage[1572]);
        }
        if (((((1U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
               & (IData)(vlSelfRef.rv32_soc__DOT__is_load)) 
              & (~ (IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned))) 
             & (0x00000400U > (vlSelfRef.rv32_soc__DOT__ls_addr 
                               >> 2U)))) {
            ++(vlSelf->__Vcoverage[1573]);
        }
        if ((0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                             >> 2U))) {
            ++(vlSelf->__Vcoverage[1574]);
        }
        if (vlSelfRef.rv32_soc__DOT__ls_misaligned) {
            ++(vlSelf->__Vcoverage[1575]);
        }
        if ((1U & ((~ (IData)(vlSelfRef.rv32_soc__DOT__is_load)) 
                   & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_store))))) {
            ++(vlSelf->__Vcoverage[1576]);
        }
        if ((1U != (IData)(vlSelfRef.rv32_soc__DOT__state))) {
            ++(vlSelf->__Vcoverage[1577]);
        }
        rv32_soc__DOT__exec_ls = ((1U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
                                  & ((~ ((0x00000400U 
                                          <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                              >> 2U)) 
                                         | (IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned))) 
                                     & ((IData)(vlSelfRef.rv32_soc__DOT__is_load) 
                                        | (IData)(vlSelfRef.rv32_soc__DOT__is_store))));
        if (rv32_soc__DOT__exec_ls) {
            ++(vlSelf->__Vcoverage[1578]);
        }
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            ++(vlSelf->__Vcoverage[1579]);
        }
        if (((IData)(vlSelfRef.rv32_soc__DOT__pc_bad) 
             & (~ (IData)(rv32_soc__DOT__exec_ls)))) {
            ++(vlSelf->__Vcoverage[1580]);
        }
        if (((0U != (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(rv32_soc__DOT__exec_ls)))) {
            ++(vlSelf->__Vcoverage[1581]);
        }
        if (vlcu_cov_any(((IData)(rv32_soc__DOT__exec_ls) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__exec_ls)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1570, rv32_soc__DOT__exec_ls, vlSelfRef.rv32_soc__DOT____Vtogcov__exec_ls);
            vlSelfRef.rv32_soc__DOT____Vtogcov__exec_ls 
                = rv32_soc__DOT__exec_ls;
        }
        vlSelfRef.mem_valid = ((IData)(vlSelfRef.mem_instr) 
                               | (IData)(rv32_soc__DOT__exec_ls));
        if (((IData)(rv32_soc__DOT__exec_ls) & (IData)(vlSelfRef.rv32_soc__DOT__is_store))) {
            ++(vlSelf->__Vcoverage[1595]);
            vlSelfRef.rv32_soc__DOT____VlemCond_9 = vlSelfRef.rv32_soc__DOT__st_data;
        } else {
            ++(vlSelf->__Vcoverage[1596]);
            vlSelfRef.rv32_soc__DOT____VlemCond_9 = 0U;
        }
        vlSelfRef.mem_wdata = vlSelfRef.rv32_soc__DOT____VlemCond_9;
        if (((IData)(rv32_soc__DOT__exec_ls) & (IData)(vlSelfRef.rv32_soc__DOT__is_store))) {
            ++(vlSelf->__Vcoverage[1592]);
        }
        if ((1U & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_store)))) {
            ++(vlSelf->__Vcoverage[1593]);
        }
        if ((1U & (~ (IData)(rv32_soc__DOT__exec_ls)))) {
            ++(vlSelf->__Vcoverage[1594]);
        }
        if (((IData)(rv32_soc__DOT__exec_ls) & (IData)(vlSelfRef.rv32_soc__DOT__is_store))) {
            ++(vlSelf->__Vcoverage[1600]);
            vlSelfRef.rv32_soc__DOT____VlemCond_10 
                = vlSelfRef.rv32_soc__DOT__st_strb;
        } else {
            ++(vlSelf->__Vcoverage[1601]);
            vlSelfRef.rv32_soc__DOT____VlemCond_10 = 0U;
        }
        vlSelfRef.mem_wstrb = vlSelfRef.rv32_soc__DOT____VlemCond_10;
        if (((IData)(rv32_soc__DOT__exec_ls) & (IData)(vlSelfRef.rv32_soc__DOT__is_store))) {
            ++(vlSelf->__Vcoverage[1597]);
        }
        if ((1U & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_store)))) {
            ++(vlSelf->__Vcoverage[1598]);
        }
        if ((1U & (~ (IData)(rv32_soc__DOT__exec_ls)))) {
            ++(vlSelf->__Vcoverage[1599]);
        }
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            ++(vlSelf->__Vcoverage[1588]);
            vlSelfRef.rv32_soc__DOT____VlemCond_8 = vlSelfRef.rv32_soc__DOT__pc;
        } else {
            ++(vlSelf->__Vcoverage[1591]);
            if (rv32_soc__DOT__exec_ls) {
                ++(vlSelf->__Vcoverage[1589]);
                vlSelfRef.rv32_soc__DOT____VlemCond_8 
                    = (0xfffffffcU & vlSelfRef.rv32_soc__DOT__ls_addr);
            } else {
                ++(vlSelf->__Vcoverage[1590]);
                vlSelfRef.rv32_soc__DOT____VlemCond_8 = 0U;
            }
        }
        vlSelfRef.mem_addr = vlSelfRef.rv32_soc__DOT____VlemCond_8;
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            ++(vlSelf->__Vcoverage[1585]);
        }
        if (vlSelfRef.rv32_soc__DOT__pc_bad) {
            ++(vlSelf->__Vcoverage[1586]);
        }
        if ((0U != (IData)(vlSelfRef.rv32_soc__DOT__state))) {
            ++(vlSelf->__Vcoverage[1587]);
        }
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            ++(vlSelf->__Vcoverage[1605]);
            vlSelfRef.rv32_soc__DOT____VlemCond_11 
                = vlSelfRef.rv32_soc__DOT__mem[(0x000003ffU 
                                                & (vlSelfRef.rv32_soc__DOT__pc 
                                                   >> 2U))];
        } else {
            ++(vlSelf->__Vcoverage[1608]);
            if (((IData)(rv32_soc__DOT__exec_ls) & (IData)(vlSelfRef.rv32_soc__DOT__is_load))) {
                ++(vlSelf->__Vcoverage[1606]);
                vlSelfRef.rv32_soc__DOT____VlemCond_11 
                    = vlSelfRef.rv32_soc__DOT__mem[
                    (0x000003ffU & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                    >> 2U))];
            } else {
                ++(vlSelf->__Vcoverage[1607]);
                vlSelfRef.rv32_soc__DOT____VlemCond_11 = 0U;
            }
        }
        vlSelfRef.mem_rdata = vlSelfRef.rv32_soc__DOT____VlemCond_11;
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            ++(vlSelf->__Vcoverage[1602]);
        }
        if (vlSelfRef.rv32_soc__DOT__pc_bad) {
            ++(vlSelf->__Vcoverage[1603]);
        }
        if ((0U != (IData)(vlSelfRef.rv32_soc__DOT__state))) {
            ++(vlSelf->__Vcoverage[1604]);
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.mem_valid) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__mem_valid)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 6, vlSelfRef.mem_valid, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_valid);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_valid 
                = vlSelfRef.mem_valid;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.mem_wdata 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wdata)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 74, vlSelfRef.mem_wdata, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wdata);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wdata 
                = vlSelfRef.mem_wdata;
        }
        if (vlcu_cov_any((0U != ((IData)(vlSelfRef.mem_wstrb) 
                                 ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wstrb))))) {
            VL_COV_TOGGLE_CHG_ST_I(4, vlSelf->__Vcoverage + 138, vlSelfRef.mem_wstrb, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wstrb);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wstrb 
                = vlSelfRef.mem_wstrb;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.mem_addr 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__mem_addr)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 10, vlSelfRef.mem_addr, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_addr);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_addr 
                = vlSelfRef.mem_addr;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.mem_rdata 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__mem_rdata)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 146, vlSelfRef.mem_rdata, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_rdata);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_rdata 
                = vlSelfRef.mem_rdata;
        }
    }

    VLCU_D static bool Vrv32_cov___024root___eval_phase__stl(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vrv32_cov___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vrv32_cov___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vrv32_cov___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vrv32_cov___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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
        } while ((2U > n));
        return (0U);
    }

    VLCU_D static bool Vrv32_cov___024root___eval_phase__ico(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VicoExecute;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__ico
            vlSelfRef.__VicoTriggered[0U] = (QData)((IData)(
                                                            ((((IData)(vlSelfRef.resetn) 
                                                               != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__resetn__0)) 
                                                              << 1U) 
                                                             | ((IData)(vlSelfRef.clk) 
                                                                != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__resetn__0 
                = vlSelfRef.resetn;
            if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.__VicoDidInit)))))) {
                vlSelfRef.__VicoDidInit = 1U;
                vlSelfRef.__VicoTriggered[0U] = (1ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (2ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
            }
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrv32_cov___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vrv32_cov___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            {
                // Inlined CFunc: _eval_ico
                if ((1ULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_sequent__TOP__0
                        if (vlcu_cov_any(((IData)(vlSelfRef.clk) 
                                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__clk)))) {
                            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 0, vlSelfRef.clk, vlSelfRef.rv32_soc__DOT____Vtogcov__clk);
                            vlSelfRef.rv32_soc__DOT____Vtogcov__clk 
                                = vlSelfRef.clk;
                        }
                    }
                }
                if ((2ULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_sequent__TOP__1
                        if (vlcu_cov_any(((IData)(vlSelfRef.resetn) 
                                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__resetn)))) {
                            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 2, vlSelfRef.resetn, vlSelfRef.rv32_soc__DOT____Vtogcov__resetn);
                            vlSelfRef.rv32_soc__DOT____Vtogcov__resetn 
                                = vlSelfRef.resetn;
                        }
                    }
                }
            }
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vrv32_cov___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vrv32_cov___024root___nba_sequent__TOP__0(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ rv32_soc__DOT__rs2_val;
        rv32_soc__DOT__rs2_val = 0;
        IData/*31:0*/ rv32_soc__DOT__alu_b;
        rv32_soc__DOT__alu_b = 0;
        CData/*0:0*/ rv32_soc__DOT__alu_sub;
        rv32_soc__DOT__alu_sub = 0;
        IData/*31:0*/ rv32_soc__DOT__ld_shift;
        rv32_soc__DOT__ld_shift = 0;
        CData/*0:0*/ rv32_soc__DOT__exec_ls;
        rv32_soc__DOT__exec_ls = 0;
        CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_0;
        __VdfgRegularize_h6e95ff9d_0_0 = 0;
        CData/*1:0*/ __Vdly__rv32_soc__DOT__state;
        __Vdly__rv32_soc__DOT__state = 0;
        IData/*31:0*/ __Vdly__rv32_soc__DOT__insn;
        __Vdly__rv32_soc__DOT__insn = 0;
        IData/*31:0*/ __Vdly__rv32_soc__DOT__pc;
        __Vdly__rv32_soc__DOT__pc = 0;
        IData/*31:0*/ __Vdly__rv32_soc__DOT__instret;
        __Vdly__rv32_soc__DOT__instret = 0;
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
        uint32_t vlCovHit26 = 0U;
        uint32_t vlCovHit27 = 0U;
        uint32_t vlCovHit29 = 0U;
        uint32_t vlCovHit30 = 0U;
        uint32_t vlCovHit32 = 0U;
        uint32_t vlCovHit34 = 0U;
        uint32_t vlCovHit36 = 0U;
        uint32_t vlCovHit37 = 0U;
        uint32_t vlCovHit39 = 0U;
        uint32_t vlCovHit41 = 0U;
        uint32_t vlCovHit42 = 0U;
        uint32_t vlCovHit48 = 0U;
        uint32_t vlCovHit49 = 0U;
        uint32_t vlCovHit50 = 0U;
        uint32_t vlCovHit51 = 0U;
        // Body
        __Vdly__rv32_soc__DOT__instret = vlSelfRef.rv32_soc__DOT__instret;
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
                    vlCovHit50 |= 512U;
                    __Vdly__rv32_soc__DOT__state = 3U;
                } else {
                    vlCovHit50 |= 1024U;
                    __Vdly__rv32_soc__DOT__insn = vlSelfRef.rv32_soc__DOT__mem
                        [(0x000003ffU & (vlSelfRef.rv32_soc__DOT__pc 
                                         >> 2U))];
                    __Vdly__rv32_soc__DOT__state = 1U;
                }
                vlCovHit50 |= 2048U;
            } else if ((1U == (IData)(vlSelfRef.rv32_soc__DOT__state))) {
                __Vdly__rv32_soc__DOT__state = 0U;
                __Vdly__rv32_soc__DOT__pc = ((IData)(4U) 
                                             + vlSelfRef.rv32_soc__DOT__pc);
                __Vdly__rv32_soc__DOT__instret = ((IData)(1U) 
                                                  + vlSelfRef.rv32_soc__DOT__instret);
                if ((0x37U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
                    if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                               >> 7U)))) {
                        vlCovHit50 |= 4096U;
                        __VdlyVal__rv32_soc__DOT__regs__v0 
                            = vlSelfRef.rv32_soc__DOT__imm_u;
                        __VdlyDim0__rv32_soc__DOT__regs__v0 
                            = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                              >> 7U));
                        __VdlySet__rv32_soc__DOT__regs__v0 = 1U;
                    } else {
                        vlCovHit50 |= 8192U;
                    }
                    vlCovHit51 |= 2097152U;
                } else if ((0x17U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
                    if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                               >> 7U)))) {
                        vlCovHit50 |= 16384U;
                        __VdlyVal__rv32_soc__DOT__regs__v1 
                            = (vlSelfRef.rv32_soc__DOT__pc 
                               + vlSelfRef.rv32_soc__DOT__imm_u);
                        __VdlyDim0__rv32_soc__DOT__regs__v1 
                            = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                              >> 7U));
                        __VdlySet__rv32_soc__DOT__regs__v1 = 1U;
                    } else {
                        vlCovHit50 |= 32768U;
                    }
                    vlCovHit51 |= 1048576U;
                } else if ((0x6fU == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
                    if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                               >> 7U)))) {
                        vlCovHit50 |= 65536U;
                        __VdlyVal__rv32_soc__DOT__regs__v2 
                            = ((IData)(4U) + vlSelfRef.rv32_soc__DOT__pc);
                        __VdlyDim0__rv32_soc__DOT__regs__v2 
                            = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                              >> 7U));
                        __VdlySet__rv32_soc__DOT__regs__v2 = 1U;
                    } else {
                        vlCovHit50 |= 131072U;
                    }
                    __Vdly__rv32_soc__DOT__pc = (vlSelfRef.rv32_soc__DOT__pc 
                                                 + vlSelfRef.rv32_soc__DOT__imm_j);
                    vlCovHit51 |= 524288U;
                } else if (vlSelfRef.rv32_soc__DOT__is_jalr) {
                    if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                               >> 7U)))) {
                        vlCovHit50 |= 262144U;
                        __VdlyVal__rv32_soc__DOT__regs__v3 
                            = ((IData)(4U) + vlSelfRef.rv32_soc__DOT__pc);
                        __VdlyDim0__rv32_soc__DOT__regs__v3 
                            = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                              >> 7U));
                        __VdlySet__rv32_soc__DOT__regs__v3 = 1U;
                    } else {
                        vlCovHit50 |= 524288U;
                    }
                    vlCovHit51 |= 262144U;
                    __Vdly__rv32_soc__DOT__pc = (0xfffffffeU 
                                                 & (vlSelfRef.rv32_soc__DOT__rs1_val 
                                                    + vlSelfRef.rv32_soc__DOT__imm_i));
                } else if (vlSelfRef.rv32_soc__DOT__is_branch) {
                    if (vlSelfRef.rv32_soc__DOT__take_branch) {
                        __Vdly__rv32_soc__DOT__pc = 
                            (vlSelfRef.rv32_soc__DOT__pc 
                             + vlSelfRef.rv32_soc__DOT__imm_b);
                        vlCovHit50 |= 1048576U;
                    } else {
                        vlCovHit50 |= 2097152U;
                    }
                    vlCovHit51 |= 131072U;
                } else {
                    if (((IData)(vlSelfRef.rv32_soc__DOT__is_alui) 
                         | (IData)(vlSelfRef.rv32_soc__DOT__is_alur))) {
                        if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                                   >> 7U)))) {
                            vlCovHit50 |= 4194304U;
                            __VdlyVal__rv32_soc__DOT__regs__v4 
                                = vlSelfRef.rv32_soc__DOT__alu_out;
                            __VdlyDim0__rv32_soc__DOT__regs__v4 
                                = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                                  >> 7U));
                            __VdlySet__rv32_soc__DOT__regs__v4 = 1U;
                        } else {
                            vlCovHit50 |= 8388608U;
                        }
                        vlCovHit51 |= 8192U;
                    } else if (vlSelfRef.rv32_soc__DOT__is_load) {
                        if (((IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned) 
                             | (0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                                >> 2U)))) {
                            __Vdly__rv32_soc__DOT__pc 
                                = vlSelfRef.rv32_soc__DOT__pc;
                            __Vdly__rv32_soc__DOT__instret 
                                = vlSelfRef.rv32_soc__DOT__instret;
                            vlCovHit50 |= 16777216U;
                            __Vdly__rv32_soc__DOT__state = 3U;
                        } else {
                            __Vdly__rv32_soc__DOT__pc 
                                = vlSelfRef.rv32_soc__DOT__pc;
                            __Vdly__rv32_soc__DOT__instret 
                                = vlSelfRef.rv32_soc__DOT__instret;
                            vlCovHit50 |= 33554432U;
                            vlSelfRef.rv32_soc__DOT__ld_addr 
                                = vlSelfRef.rv32_soc__DOT__ls_addr;
                            __Vdly__rv32_soc__DOT__state = 2U;
                        }
                        if ((0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                             >> 2U))) {
                            vlCovHit50 |= 67108864U;
                        }
                        if (vlSelfRef.rv32_soc__DOT__ls_misaligned) {
                            vlCovHit50 |= 134217728U;
                        }
                        if (((~ (IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned)) 
                             & (0x00000400U > (vlSelfRef.rv32_soc__DOT__ls_addr 
                                               >> 2U)))) {
                            vlCovHit50 |= 268435456U;
                        }
                        vlCovHit51 |= 4096U;
                    } else if (vlSelfRef.rv32_soc__DOT__is_store) {
                        if (((IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned) 
                             | (0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                                >> 2U)))) {
                            __Vdly__rv32_soc__DOT__pc 
                                = vlSelfRef.rv32_soc__DOT__pc;
                            __Vdly__rv32_soc__DOT__instret 
                                = vlSelfRef.rv32_soc__DOT__instret;
                            vlCovHit51 |= 32U;
                            __Vdly__rv32_soc__DOT__state = 3U;
                        } else {
                            if ((1U & (IData)(vlSelfRef.rv32_soc__DOT__st_strb))) {
                                vlCovHit50 |= 536870912U;
                                __VdlyVal__rv32_soc__DOT__mem__v0 
                                    = (0x000000ffU 
                                       & vlSelfRef.rv32_soc__DOT__st_data);
                                __VdlyDim0__rv32_soc__DOT__mem__v0 
                                    = (0x000003ffU 
                                       & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                          >> 2U));
                                __VdlySet__rv32_soc__DOT__mem__v0 = 1U;
                            } else {
                                vlCovHit50 |= 1073741824U;
                            }
                            if ((2U & (IData)(vlSelfRef.rv32_soc__DOT__st_strb))) {
                                vlCovHit50 |= 2147483648U;
                                __VdlyVal__rv32_soc__DOT__mem__v1 
                                    = (0x000000ffU 
                                       & (vlSelfRef.rv32_soc__DOT__st_data 
                                          >> 8U));
                                __VdlyDim0__rv32_soc__DOT__mem__v1 
                                    = (0x000003ffU 
                                       & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                          >> 2U));
                                __VdlySet__rv32_soc__DOT__mem__v1 = 1U;
                            } else {
                                vlCovHit51 |= 1U;
                            }
                            if ((4U & (IData)(vlSelfRef.rv32_soc__DOT__st_strb))) {
                                vlCovHit51 |= 2U;
                                __VdlyVal__rv32_soc__DOT__mem__v2 
                                    = (0x000000ffU 
                                       & (vlSelfRef.rv32_soc__DOT__st_data 
                                          >> 0x10U));
                                __VdlyDim0__rv32_soc__DOT__mem__v2 
                                    = (0x000003ffU 
                                       & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                          >> 2U));
                                __VdlySet__rv32_soc__DOT__mem__v2 = 1U;
                            } else {
                                vlCovHit51 |= 4U;
                            }
                            if ((8U & (IData)(vlSelfRef.rv32_soc__DOT__st_strb))) {
                                vlCovHit51 |= 8U;
                                __VdlyVal__rv32_soc__DOT__mem__v3 
                                    = (vlSelfRef.rv32_soc__DOT__st_data 
                                       >> 0x18U);
                                __VdlyDim0__rv32_soc__DOT__mem__v3 
                                    = (0x000003ffU 
                                       & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                          >> 2U));
                                __VdlySet__rv32_soc__DOT__mem__v3 = 1U;
                            } else {
                                vlCovHit51 |= 16U;
                            }
                            vlCovHit51 |= 64U;
                        }
                        if ((0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                             >> 2U))) {
                            vlCovHit51 |= 128U;
                        }
                        if (vlSelfRef.rv32_soc__DOT__ls_misaligned) {
                            vlCovHit51 |= 256U;
                        }
                        if (((~ (IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned)) 
                             & (0x00000400U > (vlSelfRef.rv32_soc__DOT__ls_addr 
                                               >> 2U)))) {
                            vlCovHit51 |= 512U;
                        }
                        vlCovHit51 |= 1024U;
                    } else {
                        __Vdly__rv32_soc__DOT__pc = vlSelfRef.rv32_soc__DOT__pc;
                        __Vdly__rv32_soc__DOT__instret 
                            = vlSelfRef.rv32_soc__DOT__instret;
                        vlCovHit51 |= 2048U;
                        __Vdly__rv32_soc__DOT__state = 3U;
                    }
                    if (vlSelfRef.rv32_soc__DOT__is_alur) {
                        vlCovHit51 |= 16384U;
                    }
                    if (vlSelfRef.rv32_soc__DOT__is_alui) {
                        vlCovHit51 |= 32768U;
                    }
                    if ((1U & ((~ (IData)(vlSelfRef.rv32_soc__DOT__is_alui)) 
                               & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_alur))))) {
                        vlCovHit51 |= 65536U;
                    }
                }
                vlCovHit51 |= 4194304U;
            } else if ((2U == (IData)(vlSelfRef.rv32_soc__DOT__state))) {
                __Vdly__rv32_soc__DOT__pc = ((IData)(4U) 
                                             + vlSelfRef.rv32_soc__DOT__pc);
                __Vdly__rv32_soc__DOT__instret = ((IData)(1U) 
                                                  + vlSelfRef.rv32_soc__DOT__instret);
                if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                           >> 7U)))) {
                    vlCovHit51 |= 8388608U;
                    __VdlyVal__rv32_soc__DOT__regs__v5 
                        = vlSelfRef.rv32_soc__DOT__ld_val;
                    __VdlyDim0__rv32_soc__DOT__regs__v5 
                        = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                          >> 7U));
                    __VdlySet__rv32_soc__DOT__regs__v5 = 1U;
                } else {
                    vlCovHit51 |= 16777216U;
                }
                __Vdly__rv32_soc__DOT__state = 0U;
                vlCovHit51 |= 33554432U;
            } else {
                vlCovHit51 |= 67108864U;
                __Vdly__rv32_soc__DOT__state = 3U;
            }
            vlCovHit51 |= 268435456U;
        } else {
            vlCovHit51 |= 134217728U;
            __Vdly__rv32_soc__DOT__pc = 0U;
            __Vdly__rv32_soc__DOT__state = 0U;
            __Vdly__rv32_soc__DOT__instret = 0U;
        }
        if ((1U & (~ (IData)(vlSelfRef.resetn)))) {
            vlCovHit51 |= 536870912U;
        }
        if (vlSelfRef.resetn) {
            vlCovHit51 |= 1073741824U;
        }
        vlCovHit51 |= 2147483648U;
        vlSelfRef.rv32_soc__DOT__instret = __Vdly__rv32_soc__DOT__instret;
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
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__instret 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__retired)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 210, vlSelfRef.rv32_soc__DOT__instret, vlSelfRef.rv32_soc__DOT____Vtogcov__retired);
            vlSelfRef.rv32_soc__DOT____Vtogcov__retired 
                = vlSelfRef.rv32_soc__DOT__instret;
        }
        vlSelfRef.retired = vlSelfRef.rv32_soc__DOT__instret;
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__ld_addr 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__ld_addr)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 402, vlSelfRef.rv32_soc__DOT__ld_addr, vlSelfRef.rv32_soc__DOT____Vtogcov__ld_addr);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ld_addr 
                = vlSelfRef.rv32_soc__DOT__ld_addr;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__mem
                                 [(0x000003ffU & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                                  >> 2U))] 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__ld_word)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1363, vlSelfRef.rv32_soc__DOT__mem
                                   [(0x000003ffU & 
                                     (vlSelfRef.rv32_soc__DOT__ld_addr 
                                      >> 2U))], vlSelfRef.rv32_soc__DOT____Vtogcov__ld_word);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ld_word 
                = vlSelfRef.rv32_soc__DOT__mem[(0x000003ffU 
                                                & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                                   >> 2U))];
        }
        rv32_soc__DOT__ld_shift = (vlSelfRef.rv32_soc__DOT__mem
                                   [(0x000003ffU & 
                                     (vlSelfRef.rv32_soc__DOT__ld_addr 
                                      >> 2U))] >> (0x00000018U 
                                                   & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                                      << 3U)));
        if ((0x00000400U <= (vlSelfRef.rv32_soc__DOT__pc 
                             >> 2U))) {
            vlCovHit48 |= 2147483648U;
        }
        if ((0U != (3U & vlSelfRef.rv32_soc__DOT__pc))) {
            vlCovHit49 |= 1U;
        }
        if (((0U == (3U & vlSelfRef.rv32_soc__DOT__pc)) 
             & (0x00000400U > (vlSelfRef.rv32_soc__DOT__pc 
                               >> 2U)))) {
            vlCovHit49 |= 2U;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__pc 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__pc)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 274, vlSelfRef.rv32_soc__DOT__pc, vlSelfRef.rv32_soc__DOT____Vtogcov__pc);
            vlSelfRef.rv32_soc__DOT____Vtogcov__pc 
                = vlSelfRef.rv32_soc__DOT__pc;
        }
        vlSelfRef.rv32_soc__DOT__pc_bad = ((0U != (3U 
                                                   & vlSelfRef.rv32_soc__DOT__pc)) 
                                           | (0x00000400U 
                                              <= (vlSelfRef.rv32_soc__DOT__pc 
                                                  >> 2U)));
        if (vlcu_cov_any((0U != ((IData)(vlSelfRef.rv32_soc__DOT__state) 
                                 ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__state))))) {
            VL_COV_TOGGLE_CHG_ST_I(2, vlSelf->__Vcoverage + 466, vlSelfRef.rv32_soc__DOT__state, vlSelfRef.rv32_soc__DOT____Vtogcov__state);
            vlSelfRef.rv32_soc__DOT____Vtogcov__state 
                = vlSelfRef.rv32_soc__DOT__state;
        }
        vlSelfRef.trap = (3U == (IData)(vlSelfRef.rv32_soc__DOT__state));
        if ((IData)((0x00000067U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit26 |= 262144U;
        }
        if ((0U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit26 |= 524288U;
        }
        if ((0x67U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            vlCovHit26 |= 1048576U;
        }
        if ((((0x63U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn)) 
              & (2U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU)))) & 
             (3U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                           >> 0x0000000cU))))) {
            vlCovHit26 |= 8388608U;
        }
        if ((3U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit26 |= 16777216U;
        }
        if ((2U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit26 |= 33554432U;
        }
        if ((0x63U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            vlCovHit26 |= 67108864U;
        }
        if ((IData)((0x00005003U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit26 |= 536870912U;
        }
        if ((IData)((0x00004003U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit26 |= 1073741824U;
        }
        if ((IData)((0x00002003U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit26 |= 2147483648U;
        }
        if ((IData)((0x00001003U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit27 |= 1U;
        }
        if ((IData)((3U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit27 |= 2U;
        }
        if ((((((0U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU))) & (1U 
                                                   != 
                                                   (7U 
                                                    & (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000000cU)))) 
               & (2U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                               >> 0x0000000cU)))) & 
              (4U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                            >> 0x0000000cU)))) & (5U 
                                                  != 
                                                  (7U 
                                                   & (vlSelfRef.rv32_soc__DOT__insn 
                                                      >> 0x0000000cU))))) {
            vlCovHit27 |= 4U;
        }
        if ((3U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            vlCovHit27 |= 8U;
        }
        if ((IData)((0x00002023U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit27 |= 64U;
        }
        if ((IData)((0x00001023U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit27 |= 128U;
        }
        if ((IData)((0x00000023U == (0x0000707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit27 |= 256U;
        }
        if ((((0U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                            >> 0x0000000cU))) & (1U 
                                                 != 
                                                 (7U 
                                                  & (vlSelfRef.rv32_soc__DOT__insn 
                                                     >> 0x0000000cU)))) 
             & (2U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                             >> 0x0000000cU))))) {
            vlCovHit27 |= 512U;
        }
        if ((0x23U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            vlCovHit27 |= 1024U;
        }
        if ((IData)((0x00000013U == (0xfe00007fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit27 |= 16384U;
        }
        if ((IData)(((0x00000013U == (0xfe00007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                     & (5U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                     >> 0x0000000cU)))))) {
            vlCovHit27 |= 32768U;
        }
        if ((IData)(((0x40000013U == (0xfe00007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                     & (1U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                     >> 0x0000000cU)))))) {
            vlCovHit27 |= 65536U;
        }
        if ((IData)(((0x00000013U == (0xfe00007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                     & (1U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                     >> 0x0000000cU)))))) {
            vlCovHit27 |= 131072U;
        }
        if ((((0x13U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn)) 
              & (1U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU)))) & 
             (5U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                           >> 0x0000000cU))))) {
            vlCovHit27 |= 262144U;
        }
        if ((IData)(((0x00005000U == (0x00007000U & vlSelfRef.rv32_soc__DOT__insn)) 
                     & ((0U != (vlSelfRef.rv32_soc__DOT__insn 
                                >> 0x19U)) & (0x20U 
                                              != (vlSelfRef.rv32_soc__DOT__insn 
                                                  >> 0x19U)))))) {
            vlCovHit27 |= 524288U;
        }
        if ((IData)(((0x00001000U == (0x00007000U & vlSelfRef.rv32_soc__DOT__insn)) 
                     & (0U != (vlSelfRef.rv32_soc__DOT__insn 
                               >> 0x19U))))) {
            vlCovHit27 |= 1048576U;
        }
        if ((0x13U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            vlCovHit27 |= 2097152U;
        }
        if ((IData)((0x40005033U == (0xfe00707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit27 |= 16777216U;
        }
        if ((IData)((0x40000033U == (0xfe00707fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit27 |= 33554432U;
        }
        if ((IData)((0x00000033U == (0xfe00007fU & vlSelfRef.rv32_soc__DOT__insn)))) {
            vlCovHit27 |= 67108864U;
        }
        if ((((0U != (vlSelfRef.rv32_soc__DOT__insn 
                      >> 0x19U)) & (0U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000000cU)))) 
             & (5U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                             >> 0x0000000cU))))) {
            vlCovHit27 |= 134217728U;
        }
        if (((0U != (vlSelfRef.rv32_soc__DOT__insn 
                     >> 0x19U)) & (0x20U != (vlSelfRef.rv32_soc__DOT__insn 
                                             >> 0x19U)))) {
            vlCovHit27 |= 268435456U;
        }
        if ((0x33U != (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn))) {
            vlCovHit27 |= 536870912U;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__insn 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__insn)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 338, vlSelfRef.rv32_soc__DOT__insn, vlSelfRef.rv32_soc__DOT____Vtogcov__insn);
            vlSelfRef.rv32_soc__DOT____Vtogcov__insn 
                = vlSelfRef.rv32_soc__DOT__insn;
        }
        if (vlcu_cov_any((0U != (0x0000007fU & (vlSelfRef.rv32_soc__DOT__insn 
                                                ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__opcode)))))) {
            VL_COV_TOGGLE_CHG_ST_I(7, vlSelf->__Vcoverage + 470, vlSelfRef.rv32_soc__DOT__insn, vlSelfRef.rv32_soc__DOT____Vtogcov__opcode);
            vlSelfRef.rv32_soc__DOT____Vtogcov__opcode 
                = (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn);
        }
        if (vlcu_cov_any((0U != (0x0000001fU & ((vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 7U) 
                                                ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__rd)))))) {
            VL_COV_TOGGLE_CHG_ST_I(5, vlSelf->__Vcoverage + 484, 
                                   (vlSelfRef.rv32_soc__DOT__insn 
                                    >> 7U), vlSelfRef.rv32_soc__DOT____Vtogcov__rd);
            vlSelfRef.rv32_soc__DOT____Vtogcov__rd 
                = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                  >> 7U));
        }
        if (vlcu_cov_any((0U != (7U & ((vlSelfRef.rv32_soc__DOT__insn 
                                        >> 0x0000000cU) 
                                       ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__funct3)))))) {
            VL_COV_TOGGLE_CHG_ST_I(3, vlSelf->__Vcoverage + 494, 
                                   (vlSelfRef.rv32_soc__DOT__insn 
                                    >> 0x0000000cU), vlSelfRef.rv32_soc__DOT____Vtogcov__funct3);
            vlSelfRef.rv32_soc__DOT____Vtogcov__funct3 
                = (7U & (vlSelfRef.rv32_soc__DOT__insn 
                         >> 0x0000000cU));
        }
        if (vlcu_cov_any((0U != (0x0000001fU & ((vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000000fU) 
                                                ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__rs1)))))) {
            VL_COV_TOGGLE_CHG_ST_I(5, vlSelf->__Vcoverage + 500, 
                                   (vlSelfRef.rv32_soc__DOT__insn 
                                    >> 0x0000000fU), vlSelfRef.rv32_soc__DOT____Vtogcov__rs1);
            vlSelfRef.rv32_soc__DOT____Vtogcov__rs1 
                = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                  >> 0x0000000fU));
        }
        if (vlcu_cov_any((0U != (0x0000001fU & ((vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x00000014U) 
                                                ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__rs2)))))) {
            VL_COV_TOGGLE_CHG_ST_I(5, vlSelf->__Vcoverage + 510, 
                                   (vlSelfRef.rv32_soc__DOT__insn 
                                    >> 0x00000014U), vlSelfRef.rv32_soc__DOT____Vtogcov__rs2);
            vlSelfRef.rv32_soc__DOT____Vtogcov__rs2 
                = (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                  >> 0x00000014U));
        }
        if (vlcu_cov_any((1U & ((vlSelfRef.rv32_soc__DOT__insn 
                                 >> 0x0000001eU) ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__f7b5))))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 520, 
                                   (vlSelfRef.rv32_soc__DOT__insn 
                                    >> 0x0000001eU), vlSelfRef.rv32_soc__DOT____Vtogcov__f7b5);
            vlSelfRef.rv32_soc__DOT____Vtogcov__f7b5 
                = (1U & (vlSelfRef.rv32_soc__DOT__insn 
                         >> 0x0000001eU));
        }
        if (vlcu_cov_any((0U != ((((- (IData)((vlSelfRef.rv32_soc__DOT__insn 
                                               >> 0x0000001fU))) 
                                   << 0x0000000cU) 
                                  | ((0x00000fe0U & 
                                      (vlSelfRef.rv32_soc__DOT__insn 
                                       >> 0x00000014U)) 
                                     | (0x0000001fU 
                                        & (vlSelfRef.rv32_soc__DOT__insn 
                                           >> 7U)))) 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__imm_s)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 586, 
                                   (((- (IData)((vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x0000001fU))) 
                                     << 0x0000000cU) 
                                    | ((0x00000fe0U 
                                        & (vlSelfRef.rv32_soc__DOT__insn 
                                           >> 0x00000014U)) 
                                       | (0x0000001fU 
                                          & (vlSelfRef.rv32_soc__DOT__insn 
                                             >> 7U)))), vlSelfRef.rv32_soc__DOT____Vtogcov__imm_s);
            vlSelfRef.rv32_soc__DOT____Vtogcov__imm_s 
                = (((- (IData)((vlSelfRef.rv32_soc__DOT__insn 
                                >> 0x0000001fU))) << 0x0000000cU) 
                   | ((0x00000fe0U & (vlSelfRef.rv32_soc__DOT__insn 
                                      >> 0x00000014U)) 
                      | (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                        >> 7U))));
        }
        if (vlcu_cov_any((0U != ((0xfffff000U & vlSelfRef.rv32_soc__DOT__insn) 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__imm_u)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 714, 
                                   (0xfffff000U & vlSelfRef.rv32_soc__DOT__insn), vlSelfRef.rv32_soc__DOT____Vtogcov__imm_u);
            vlSelfRef.rv32_soc__DOT____Vtogcov__imm_u 
                = (0xfffff000U & vlSelfRef.rv32_soc__DOT__insn);
        }
        if (vlcu_cov_any(((0x37U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_lui)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 842, 
                                   (0x37U == (0x0000007fU 
                                              & vlSelfRef.rv32_soc__DOT__insn)), vlSelfRef.rv32_soc__DOT____Vtogcov__is_lui);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_lui 
                = (0x37U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn));
        }
        if (vlcu_cov_any(((0x17U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_auipc)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 844, 
                                   (0x17U == (0x0000007fU 
                                              & vlSelfRef.rv32_soc__DOT__insn)), vlSelfRef.rv32_soc__DOT____Vtogcov__is_auipc);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_auipc 
                = (0x17U == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn));
        }
        if (vlcu_cov_any(((0x6fU == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn)) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_jal)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 846, 
                                   (0x6fU == (0x0000007fU 
                                              & vlSelfRef.rv32_soc__DOT__insn)), vlSelfRef.rv32_soc__DOT____Vtogcov__is_jal);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_jal 
                = (0x6fU == (0x0000007fU & vlSelfRef.rv32_soc__DOT__insn));
        }
        vlSelfRef.rv32_soc__DOT__imm_u = (0xfffff000U 
                                          & vlSelfRef.rv32_soc__DOT__insn);
        vlSelfRef.rv32_soc__DOT__imm_b = (((- (IData)(
                                                      (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000001fU))) 
                                           << 0x0000000dU) 
                                          | ((((2U 
                                                & (vlSelfRef.rv32_soc__DOT__insn 
                                                   >> 0x0000001eU)) 
                                               | (1U 
                                                  & (vlSelfRef.rv32_soc__DOT__insn 
                                                     >> 7U))) 
                                              << 0x0000000bU) 
                                             | ((0x000007e0U 
                                                 & (vlSelfRef.rv32_soc__DOT__insn 
                                                    >> 0x00000014U)) 
                                                | (0x0000001eU 
                                                   & (vlSelfRef.rv32_soc__DOT__insn 
                                                      >> 7U)))));
        vlSelfRef.rv32_soc__DOT__imm_j = ((((0x00000ffeU 
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
                                                   >> 0x00000014U))));
        vlSelfRef.rv32_soc__DOT__is_jalr = (IData)(
                                                   (0x00000067U 
                                                    == 
                                                    (0x0000707fU 
                                                     & vlSelfRef.rv32_soc__DOT__insn)));
        vlSelfRef.rv32_soc__DOT__is_branch = ((0x63U 
                                               == (0x0000007fU 
                                                   & vlSelfRef.rv32_soc__DOT__insn)) 
                                              & ((2U 
                                                  != 
                                                  (7U 
                                                   & (vlSelfRef.rv32_soc__DOT__insn 
                                                      >> 0x0000000cU))) 
                                                 & (3U 
                                                    != 
                                                    (7U 
                                                     & (vlSelfRef.rv32_soc__DOT__insn 
                                                        >> 0x0000000cU)))));
        vlSelfRef.rv32_soc__DOT__is_alui = ((0x13U 
                                             == (0x0000007fU 
                                                 & vlSelfRef.rv32_soc__DOT__insn)) 
                                            & ((0U 
                                                == 
                                                (vlSelfRef.rv32_soc__DOT__insn 
                                                 >> 0x00000019U)) 
                                               | ((1U 
                                                   != 
                                                   (7U 
                                                    & (vlSelfRef.rv32_soc__DOT__insn 
                                                       >> 0x0000000cU))) 
                                                  & (IData)(
                                                            ((0x00005000U 
                                                              != 
                                                              (0x00007000U 
                                                               & vlSelfRef.rv32_soc__DOT__insn)) 
                                                             | (0x20U 
                                                                == 
                                                                (vlSelfRef.rv32_soc__DOT__insn 
                                                                 >> 0x00000019U)))))));
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
        if ((0U == (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x00000014U)))) {
            vlCovHit32 |= 16U;
            vlSelfRef.rv32_soc__DOT____VlemCond_1 = 0U;
        } else {
            vlCovHit32 |= 32U;
            vlSelfRef.rv32_soc__DOT____VlemCond_1 = vlSelfRef.rv32_soc__DOT__regs
                [(0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                 >> 0x00000014U))];
        }
        rv32_soc__DOT__rs2_val = vlSelfRef.rv32_soc__DOT____VlemCond_1;
        if ((0U == (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x00000014U)))) {
            vlCovHit32 |= 4U;
        }
        if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x00000014U)))) {
            vlCovHit32 |= 8U;
        }
        if ((0U == (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x0000000fU)))) {
            vlCovHit30 |= 1U;
            vlSelfRef.rv32_soc__DOT____VlemCond_0 = 0U;
        } else {
            vlCovHit30 |= 2U;
            vlSelfRef.rv32_soc__DOT____VlemCond_0 = vlSelfRef.rv32_soc__DOT__regs
                [(0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                 >> 0x0000000fU))];
        }
        vlSelfRef.rv32_soc__DOT__rs1_val = vlSelfRef.rv32_soc__DOT____VlemCond_0;
        if ((0U == (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x0000000fU)))) {
            vlCovHit29 |= 1073741824U;
        }
        if ((0U != (0x0000001fU & (vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x0000000fU)))) {
            vlCovHit29 |= 2147483648U;
        }
        __VdfgRegularize_h6e95ff9d_0_0 = ((0U == (7U 
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
        if (vlcu_cov_any((0U != (rv32_soc__DOT__ld_shift 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__ld_shift)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1427, rv32_soc__DOT__ld_shift, vlSelfRef.rv32_soc__DOT____Vtogcov__ld_shift);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ld_shift 
                = rv32_soc__DOT__ld_shift;
        }
        if ((0U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit48 |= 2097152U;
            vlSelfRef.rv32_soc__DOT____VlemCond_7 = 
                (((- (IData)((1U & (rv32_soc__DOT__ld_shift 
                                    >> 7U)))) << 8U) 
                 | (0x000000ffU & rv32_soc__DOT__ld_shift));
        } else {
            vlCovHit48 |= 268435456U;
            if ((1U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU)))) {
                vlCovHit48 |= 4194304U;
                vlSelfRef.rv32_soc__DOT____VlemCond_7 
                    = (((- (IData)((1U & (rv32_soc__DOT__ld_shift 
                                          >> 0x0fU)))) 
                        << 0x00000010U) | (0x0000ffffU 
                                           & rv32_soc__DOT__ld_shift));
            } else {
                vlCovHit48 |= 134217728U;
                if ((4U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                  >> 0x0000000cU)))) {
                    vlCovHit48 |= 8388608U;
                    vlSelfRef.rv32_soc__DOT____VlemCond_7 
                        = (0x000000ffU & rv32_soc__DOT__ld_shift);
                } else {
                    vlCovHit48 |= 67108864U;
                    if ((5U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                                      >> 0x0000000cU)))) {
                        vlCovHit48 |= 16777216U;
                        vlSelfRef.rv32_soc__DOT____VlemCond_7 
                            = (0x0000ffffU & rv32_soc__DOT__ld_shift);
                    } else {
                        vlCovHit48 |= 33554432U;
                        vlSelfRef.rv32_soc__DOT____VlemCond_7 
                            = vlSelfRef.rv32_soc__DOT__mem
                            [(0x000003ffU & (vlSelfRef.rv32_soc__DOT__ld_addr 
                                             >> 2U))];
                    }
                }
            }
        }
        vlSelfRef.rv32_soc__DOT__ld_val = vlSelfRef.rv32_soc__DOT____VlemCond_7;
        if ((0U == (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit48 |= 524288U;
        }
        if ((0U != (7U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit48 |= 1048576U;
        }
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            vlCovHit49 |= 16384U;
        }
        if (vlSelfRef.rv32_soc__DOT__pc_bad) {
            vlCovHit49 |= 32768U;
        }
        if ((0U != (IData)(vlSelfRef.rv32_soc__DOT__state))) {
            vlCovHit49 |= 65536U;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__pc_bad) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__pc_bad)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1565, vlSelfRef.rv32_soc__DOT__pc_bad, vlSelfRef.rv32_soc__DOT____Vtogcov__pc_bad);
            vlSelfRef.rv32_soc__DOT____Vtogcov__pc_bad 
                = vlSelfRef.rv32_soc__DOT__pc_bad;
        }
        vlSelfRef.mem_instr = ((~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)) 
                               & (0U == (IData)(vlSelfRef.rv32_soc__DOT__state)));
        if (vlcu_cov_any(((IData)(vlSelfRef.trap) ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__trap)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 4, vlSelfRef.trap, vlSelfRef.rv32_soc__DOT____Vtogcov__trap);
            vlSelfRef.rv32_soc__DOT____Vtogcov__trap 
                = vlSelfRef.trap;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__imm_b 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__imm_b)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 650, vlSelfRef.rv32_soc__DOT__imm_b, vlSelfRef.rv32_soc__DOT____Vtogcov__imm_b);
            vlSelfRef.rv32_soc__DOT____Vtogcov__imm_b 
                = vlSelfRef.rv32_soc__DOT__imm_b;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__imm_j 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__imm_j)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 778, vlSelfRef.rv32_soc__DOT__imm_j, vlSelfRef.rv32_soc__DOT____Vtogcov__imm_j);
            vlSelfRef.rv32_soc__DOT____Vtogcov__imm_j 
                = vlSelfRef.rv32_soc__DOT__imm_j;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_jalr) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_jalr)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 848, vlSelfRef.rv32_soc__DOT__is_jalr, vlSelfRef.rv32_soc__DOT____Vtogcov__is_jalr);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_jalr 
                = vlSelfRef.rv32_soc__DOT__is_jalr;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_branch) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_branch)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 853, vlSelfRef.rv32_soc__DOT__is_branch, vlSelfRef.rv32_soc__DOT____Vtogcov__is_branch);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_branch 
                = vlSelfRef.rv32_soc__DOT__is_branch;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_alui) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_alui)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 875, vlSelfRef.rv32_soc__DOT__is_alui, vlSelfRef.rv32_soc__DOT____Vtogcov__is_alui);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_alui 
                = vlSelfRef.rv32_soc__DOT__is_alui;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__imm_i 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__imm_i)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 522, vlSelfRef.rv32_soc__DOT__imm_i, vlSelfRef.rv32_soc__DOT____Vtogcov__imm_i);
            vlSelfRef.rv32_soc__DOT____Vtogcov__imm_i 
                = vlSelfRef.rv32_soc__DOT__imm_i;
        }
        if (((IData)(vlSelfRef.rv32_soc__DOT__is_alur) 
             & (vlSelfRef.rv32_soc__DOT__insn >> 0x0000001eU))) {
            vlCovHit34 |= 4194304U;
        }
        if ((1U & (~ (vlSelfRef.rv32_soc__DOT__insn 
                      >> 0x0000001eU)))) {
            vlCovHit34 |= 8388608U;
        }
        if ((1U & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_alur)))) {
            vlCovHit34 |= 16777216U;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_alur) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_alur)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 886, vlSelfRef.rv32_soc__DOT__is_alur, vlSelfRef.rv32_soc__DOT____Vtogcov__is_alur);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_alur 
                = vlSelfRef.rv32_soc__DOT__is_alur;
        }
        rv32_soc__DOT__alu_sub = ((vlSelfRef.rv32_soc__DOT__insn 
                                   >> 0x0000001eU) 
                                  & (IData)(vlSelfRef.rv32_soc__DOT__is_alur));
        if (vlcu_cov_any((0U != (rv32_soc__DOT__rs2_val 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__rs2_val)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 962, rv32_soc__DOT__rs2_val, vlSelfRef.rv32_soc__DOT____Vtogcov__rs2_val);
            vlSelfRef.rv32_soc__DOT____Vtogcov__rs2_val 
                = rv32_soc__DOT__rs2_val;
        }
        if ((0U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit42 |= 2U;
            vlSelfRef.rv32_soc__DOT____VlemCond_5 = 
                ((rv32_soc__DOT__rs2_val << 0x00000018U) 
                 | ((0x00ff0000U & (rv32_soc__DOT__rs2_val 
                                    << 0x00000010U)) 
                    | ((0x0000ff00U & (rv32_soc__DOT__rs2_val 
                                       << 8U)) | (0x000000ffU 
                                                  & rv32_soc__DOT__rs2_val))));
        } else {
            vlCovHit42 |= 16U;
            if ((1U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU)))) {
                vlCovHit42 |= 4U;
                vlSelfRef.rv32_soc__DOT____VlemCond_5 
                    = ((rv32_soc__DOT__rs2_val << 0x00000010U) 
                       | (0x0000ffffU & rv32_soc__DOT__rs2_val));
            } else {
                vlCovHit42 |= 8U;
                vlSelfRef.rv32_soc__DOT____VlemCond_5 
                    = rv32_soc__DOT__rs2_val;
            }
        }
        vlSelfRef.rv32_soc__DOT__st_data = vlSelfRef.rv32_soc__DOT____VlemCond_5;
        if ((0U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit41 |= 2147483648U;
        }
        if ((0U != (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit42 |= 1U;
        }
        if (vlSelfRef.rv32_soc__DOT__is_alur) {
            vlCovHit34 |= 256U;
            vlSelfRef.rv32_soc__DOT____VlemCond_2 = rv32_soc__DOT__rs2_val;
        } else {
            vlCovHit34 |= 512U;
            vlSelfRef.rv32_soc__DOT____VlemCond_2 = vlSelfRef.rv32_soc__DOT__imm_i;
        }
        rv32_soc__DOT__alu_b = vlSelfRef.rv32_soc__DOT____VlemCond_2;
        if (vlSelfRef.rv32_soc__DOT__is_alur) {
            vlCovHit34 |= 64U;
        }
        if ((1U & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_alur)))) {
            vlCovHit34 |= 128U;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__rs1_val 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__rs1_val)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 894, vlSelfRef.rv32_soc__DOT__rs1_val, vlSelfRef.rv32_soc__DOT____Vtogcov__rs1_val);
            vlSelfRef.rv32_soc__DOT____Vtogcov__rs1_val 
                = vlSelfRef.rv32_soc__DOT__rs1_val;
        }
        if ((0x00004000U & vlSelfRef.rv32_soc__DOT__insn)) {
            if ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)) {
                if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
                    vlSelfRef.rv32_soc__DOT__take_branch 
                        = (vlSelfRef.rv32_soc__DOT__rs1_val 
                           >= rv32_soc__DOT__rs2_val);
                    ++(vlSelf->__Vcoverage[1201]);
                } else {
                    vlSelfRef.rv32_soc__DOT__take_branch 
                        = (vlSelfRef.rv32_soc__DOT__rs1_val 
                           < rv32_soc__DOT__rs2_val);
                    vlCovHit37 |= 65536U;
                }
            } else if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
                vlSelfRef.rv32_soc__DOT__take_branch 
                    = VL_GTES_III(32, vlSelfRef.rv32_soc__DOT__rs1_val, rv32_soc__DOT__rs2_val);
                vlCovHit37 |= 32768U;
            } else {
                vlSelfRef.rv32_soc__DOT__take_branch 
                    = VL_LTS_III(32, vlSelfRef.rv32_soc__DOT__rs1_val, rv32_soc__DOT__rs2_val);
                vlCovHit37 |= 16384U;
            }
        } else if ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)) {
            vlSelfRef.rv32_soc__DOT__take_branch = 
                (vlSelfRef.rv32_soc__DOT__rs1_val >= rv32_soc__DOT__rs2_val);
            ++(vlSelf->__Vcoverage[1201]);
        } else if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
            vlSelfRef.rv32_soc__DOT__take_branch = 
                (vlSelfRef.rv32_soc__DOT__rs1_val != rv32_soc__DOT__rs2_val);
            vlCovHit37 |= 8192U;
        } else {
            vlSelfRef.rv32_soc__DOT__take_branch = 
                (vlSelfRef.rv32_soc__DOT__rs1_val == rv32_soc__DOT__rs2_val);
            vlCovHit37 |= 4096U;
        }
        vlCovHit37 |= 262144U;
        vlSelfRef.rv32_soc__DOT__is_load = ((3U == 
                                             (0x0000007fU 
                                              & vlSelfRef.rv32_soc__DOT__insn)) 
                                            & ((IData)(__VdfgRegularize_h6e95ff9d_0_0) 
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
                                             & (IData)(__VdfgRegularize_h6e95ff9d_0_0));
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__ld_val 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__ld_val)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1491, vlSelfRef.rv32_soc__DOT__ld_val, vlSelfRef.rv32_soc__DOT____Vtogcov__ld_val);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ld_val 
                = vlSelfRef.rv32_soc__DOT__ld_val;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.mem_instr) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__mem_instr)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 8, vlSelfRef.mem_instr, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_instr);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_instr 
                = vlSelfRef.mem_instr;
        }
        if (vlcu_cov_any(((IData)(rv32_soc__DOT__alu_sub) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__alu_sub)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1108, rv32_soc__DOT__alu_sub, vlSelfRef.rv32_soc__DOT____Vtogcov__alu_sub);
            vlSelfRef.rv32_soc__DOT____Vtogcov__alu_sub 
                = rv32_soc__DOT__alu_sub;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__st_data 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__st_data)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1279, vlSelfRef.rv32_soc__DOT__st_data, vlSelfRef.rv32_soc__DOT____Vtogcov__st_data);
            vlSelfRef.rv32_soc__DOT____Vtogcov__st_data 
                = vlSelfRef.rv32_soc__DOT__st_data;
        }
        if (vlcu_cov_any((0U != (rv32_soc__DOT__alu_b 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__alu_b)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1030, rv32_soc__DOT__alu_b, vlSelfRef.rv32_soc__DOT____Vtogcov__alu_b);
            vlSelfRef.rv32_soc__DOT____Vtogcov__alu_b 
                = rv32_soc__DOT__alu_b;
        }
        if (vlcu_cov_any((0U != (0x0000001fU & (rv32_soc__DOT__alu_b 
                                                ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__shamt)))))) {
            VL_COV_TOGGLE_CHG_ST_I(5, vlSelf->__Vcoverage + 1098, rv32_soc__DOT__alu_b, vlSelfRef.rv32_soc__DOT____Vtogcov__shamt);
            vlSelfRef.rv32_soc__DOT____Vtogcov__shamt 
                = (0x0000001fU & rv32_soc__DOT__alu_b);
        }
        if ((0x00004000U & vlSelfRef.rv32_soc__DOT__insn)) {
            if ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)) {
                if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
                    vlSelfRef.rv32_soc__DOT__alu_out 
                        = (vlSelfRef.rv32_soc__DOT__rs1_val 
                           & rv32_soc__DOT__alu_b);
                    vlCovHit37 |= 256U;
                } else {
                    vlSelfRef.rv32_soc__DOT__alu_out 
                        = (vlSelfRef.rv32_soc__DOT__rs1_val 
                           | rv32_soc__DOT__alu_b);
                    vlCovHit37 |= 128U;
                }
            } else if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
                if ((0x40000000U & vlSelfRef.rv32_soc__DOT__insn)) {
                    vlCovHit37 |= 16U;
                    vlSelfRef.rv32_soc__DOT____VlemCond_4 
                        = VL_SHIFTRS_III(32,32,5, vlSelfRef.rv32_soc__DOT__rs1_val, 
                                         (0x0000001fU 
                                          & rv32_soc__DOT__alu_b));
                } else {
                    vlCovHit37 |= 32U;
                    vlSelfRef.rv32_soc__DOT____VlemCond_4 
                        = (vlSelfRef.rv32_soc__DOT__rs1_val 
                           >> (0x0000001fU & rv32_soc__DOT__alu_b));
                }
                vlSelfRef.rv32_soc__DOT__alu_out = vlSelfRef.rv32_soc__DOT____VlemCond_4;
                vlCovHit37 |= 64U;
            } else {
                vlSelfRef.rv32_soc__DOT__alu_out = 
                    (vlSelfRef.rv32_soc__DOT__rs1_val 
                     ^ rv32_soc__DOT__alu_b);
                vlCovHit37 |= 2U;
            }
        } else if ((0x00002000U & vlSelfRef.rv32_soc__DOT__insn)) {
            if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
                vlSelfRef.rv32_soc__DOT__alu_out = 
                    (vlSelfRef.rv32_soc__DOT__rs1_val 
                     < rv32_soc__DOT__alu_b);
                vlCovHit37 |= 1U;
            } else {
                vlSelfRef.rv32_soc__DOT__alu_out = 
                    VL_LTS_III(32, vlSelfRef.rv32_soc__DOT__rs1_val, rv32_soc__DOT__alu_b);
                vlCovHit36 |= 2147483648U;
            }
        } else if ((0x00001000U & vlSelfRef.rv32_soc__DOT__insn)) {
            vlSelfRef.rv32_soc__DOT__alu_out = (vlSelfRef.rv32_soc__DOT__rs1_val 
                                                << 
                                                (0x0000001fU 
                                                 & rv32_soc__DOT__alu_b));
            vlCovHit36 |= 1073741824U;
        } else {
            if (rv32_soc__DOT__alu_sub) {
                vlCovHit36 |= 134217728U;
                vlSelfRef.rv32_soc__DOT____VlemCond_3 
                    = (vlSelfRef.rv32_soc__DOT__rs1_val 
                       - rv32_soc__DOT__alu_b);
            } else {
                vlCovHit36 |= 268435456U;
                vlSelfRef.rv32_soc__DOT____VlemCond_3 
                    = (vlSelfRef.rv32_soc__DOT__rs1_val 
                       + rv32_soc__DOT__alu_b);
            }
            vlSelfRef.rv32_soc__DOT__alu_out = vlSelfRef.rv32_soc__DOT____VlemCond_3;
            vlCovHit36 |= 536870912U;
        }
        if (rv32_soc__DOT__alu_sub) {
            vlCovHit36 |= 33554432U;
        }
        if ((1U & (~ (IData)(rv32_soc__DOT__alu_sub)))) {
            vlCovHit36 |= 67108864U;
        }
        if ((0x40000000U & vlSelfRef.rv32_soc__DOT__insn)) {
            vlCovHit37 |= 4U;
        }
        if ((1U & (~ (vlSelfRef.rv32_soc__DOT__insn 
                      >> 0x0000001eU)))) {
            vlCovHit37 |= 8U;
        }
        vlCovHit37 |= 512U;
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__take_branch) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__take_branch)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1194, vlSelfRef.rv32_soc__DOT__take_branch, vlSelfRef.rv32_soc__DOT____Vtogcov__take_branch);
            vlSelfRef.rv32_soc__DOT____Vtogcov__take_branch 
                = vlSelfRef.rv32_soc__DOT__take_branch;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_load) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_load)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 859, vlSelfRef.rv32_soc__DOT__is_load, vlSelfRef.rv32_soc__DOT____Vtogcov__is_load);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_load 
                = vlSelfRef.rv32_soc__DOT__is_load;
        }
        if (vlSelfRef.rv32_soc__DOT__is_store) {
            vlCovHit39 |= 524288U;
        }
        if ((1U & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_store)))) {
            vlCovHit39 |= 1048576U;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__is_store) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__is_store)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 868, vlSelfRef.rv32_soc__DOT__is_store, vlSelfRef.rv32_soc__DOT____Vtogcov__is_store);
            vlSelfRef.rv32_soc__DOT____Vtogcov__is_store 
                = vlSelfRef.rv32_soc__DOT__is_store;
        }
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
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__alu_out 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__alu_out)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1113, vlSelfRef.rv32_soc__DOT__alu_out, vlSelfRef.rv32_soc__DOT____Vtogcov__alu_out);
            vlSelfRef.rv32_soc__DOT____Vtogcov__alu_out 
                = vlSelfRef.rv32_soc__DOT__alu_out;
        }
        if ((IData)(((0x00002000U == (0x00003000U & vlSelfRef.rv32_soc__DOT__insn)) 
                     & (0U != (3U & vlSelfRef.rv32_soc__DOT__ls_addr))))) {
            vlCovHit39 |= 8388608U;
        }
        if ((IData)(((0x00001000U == (0x00003000U & vlSelfRef.rv32_soc__DOT__insn)) 
                     & vlSelfRef.rv32_soc__DOT__ls_addr))) {
            vlCovHit39 |= 16777216U;
        }
        if ((IData)((0U == (3U & vlSelfRef.rv32_soc__DOT__ls_addr)))) {
            vlCovHit39 |= 33554432U;
        }
        if (((~ vlSelfRef.rv32_soc__DOT__ls_addr) & 
             (2U != (3U & (vlSelfRef.rv32_soc__DOT__insn 
                           >> 0x0000000cU))))) {
            vlCovHit39 |= 67108864U;
        }
        if (((1U != (3U & (vlSelfRef.rv32_soc__DOT__insn 
                           >> 0x0000000cU))) & (0U 
                                                == 
                                                (3U 
                                                 & vlSelfRef.rv32_soc__DOT__ls_addr)))) {
            vlCovHit39 |= 134217728U;
        }
        if (((1U != (3U & (vlSelfRef.rv32_soc__DOT__insn 
                           >> 0x0000000cU))) & (2U 
                                                != 
                                                (3U 
                                                 & (vlSelfRef.rv32_soc__DOT__insn 
                                                    >> 0x0000000cU))))) {
            vlCovHit39 |= 268435456U;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.rv32_soc__DOT__ls_addr 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__ls_addr)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 1203, vlSelfRef.rv32_soc__DOT__ls_addr, vlSelfRef.rv32_soc__DOT____Vtogcov__ls_addr);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ls_addr 
                = vlSelfRef.rv32_soc__DOT__ls_addr;
        }
        if (vlcu_cov_any(((0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                           >> 2U)) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__ls_oob)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1277, 
                                   (0x00000400U <= 
                                    (vlSelfRef.rv32_soc__DOT__ls_addr 
                                     >> 2U)), vlSelfRef.rv32_soc__DOT____Vtogcov__ls_oob);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ls_oob 
                = (0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                   >> 2U));
        }
        if ((0U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit42 |= 32768U;
            vlSelfRef.rv32_soc__DOT____VlemCond_6 = 
                (0x0000000fU & ((IData)(1U) << (3U 
                                                & vlSelfRef.rv32_soc__DOT__ls_addr)));
        } else {
            vlCovHit42 |= 262144U;
            if ((1U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                              >> 0x0000000cU)))) {
                vlCovHit42 |= 65536U;
                vlSelfRef.rv32_soc__DOT____VlemCond_6 
                    = (0x0000000fU & ((IData)(3U) << 
                                      (3U & vlSelfRef.rv32_soc__DOT__ls_addr)));
            } else {
                vlCovHit42 |= 131072U;
                vlSelfRef.rv32_soc__DOT____VlemCond_6 = 0x0fU;
            }
        }
        vlSelfRef.rv32_soc__DOT__st_strb = vlSelfRef.rv32_soc__DOT____VlemCond_6;
        if ((0U == (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit42 |= 8192U;
        }
        if ((0U != (3U & (vlSelfRef.rv32_soc__DOT__insn 
                          >> 0x0000000cU)))) {
            vlCovHit42 |= 16384U;
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
        if (vlcu_cov_any((0U != ((IData)(vlSelfRef.rv32_soc__DOT__st_strb) 
                                 ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__st_strb))))) {
            VL_COV_TOGGLE_CHG_ST_I(4, vlSelf->__Vcoverage + 1349, vlSelfRef.rv32_soc__DOT__st_strb, vlSelfRef.rv32_soc__DOT____Vtogcov__st_strb);
            vlSelfRef.rv32_soc__DOT____Vtogcov__st_strb 
                = vlSelfRef.rv32_soc__DOT__st_strb;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__ls_misaligned)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1269, vlSelfRef.rv32_soc__DOT__ls_misaligned, vlSelfRef.rv32_soc__DOT____Vtogcov__ls_misaligned);
            vlSelfRef.rv32_soc__DOT____Vtogcov__ls_misaligned 
                = vlSelfRef.rv32_soc__DOT__ls_misaligned;
        }
        if (((((1U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
               & (IData)(vlSelfRef.rv32_soc__DOT__is_store)) 
              & (~ (IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned))) 
             & (0x00000400U > (vlSelfRef.rv32_soc__DOT__ls_addr 
                               >> 2U)))) {
            vlCovHit49 |= 16U;
        }
        if (((((1U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
               & (IData)(vlSelfRef.rv32_soc__DOT__is_load)) 
              & (~ (IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned))) 
             & (0x00000400U > (vlSelfRef.rv32_soc__DOT__ls_addr 
                               >> 2U)))) {
            vlCovHit49 |= 32U;
        }
        if ((0x00000400U <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                             >> 2U))) {
            vlCovHit49 |= 64U;
        }
        if (vlSelfRef.rv32_soc__DOT__ls_misaligned) {
            vlCovHit49 |= 128U;
        }
        if ((1U & ((~ (IData)(vlSelfRef.rv32_soc__DOT__is_load)) 
                   & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_store))))) {
            vlCovHit49 |= 256U;
        }
        if ((1U != (IData)(vlSelfRef.rv32_soc__DOT__state))) {
            vlCovHit49 |= 512U;
        }
        rv32_soc__DOT__exec_ls = ((1U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
                                  & ((~ ((0x00000400U 
                                          <= (vlSelfRef.rv32_soc__DOT__ls_addr 
                                              >> 2U)) 
                                         | (IData)(vlSelfRef.rv32_soc__DOT__ls_misaligned))) 
                                     & ((IData)(vlSelfRef.rv32_soc__DOT__is_load) 
                                        | (IData)(vlSelfRef.rv32_soc__DOT__is_store))));
        if (rv32_soc__DOT__exec_ls) {
            vlCovHit49 |= 1024U;
        }
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            vlCovHit49 |= 2048U;
        }
        if (((IData)(vlSelfRef.rv32_soc__DOT__pc_bad) 
             & (~ (IData)(rv32_soc__DOT__exec_ls)))) {
            vlCovHit49 |= 4096U;
        }
        if (((0U != (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(rv32_soc__DOT__exec_ls)))) {
            vlCovHit49 |= 8192U;
        }
        if (vlcu_cov_any(((IData)(rv32_soc__DOT__exec_ls) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__exec_ls)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 1570, rv32_soc__DOT__exec_ls, vlSelfRef.rv32_soc__DOT____Vtogcov__exec_ls);
            vlSelfRef.rv32_soc__DOT____Vtogcov__exec_ls 
                = rv32_soc__DOT__exec_ls;
        }
        vlSelfRef.mem_valid = ((IData)(vlSelfRef.mem_instr) 
                               | (IData)(rv32_soc__DOT__exec_ls));
        if (((IData)(rv32_soc__DOT__exec_ls) & (IData)(vlSelfRef.rv32_soc__DOT__is_store))) {
            vlCovHit49 |= 134217728U;
            vlSelfRef.rv32_soc__DOT____VlemCond_9 = vlSelfRef.rv32_soc__DOT__st_data;
        } else {
            vlCovHit49 |= 268435456U;
            vlSelfRef.rv32_soc__DOT____VlemCond_9 = 0U;
        }
        vlSelfRef.mem_wdata = vlSelfRef.rv32_soc__DOT____VlemCond_9;
        if (((IData)(rv32_soc__DOT__exec_ls) & (IData)(vlSelfRef.rv32_soc__DOT__is_store))) {
            vlCovHit49 |= 16777216U;
        }
        if ((1U & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_store)))) {
            vlCovHit49 |= 33554432U;
        }
        if ((1U & (~ (IData)(rv32_soc__DOT__exec_ls)))) {
            vlCovHit49 |= 67108864U;
        }
        if (((IData)(rv32_soc__DOT__exec_ls) & (IData)(vlSelfRef.rv32_soc__DOT__is_store))) {
            vlCovHit50 |= 1U;
            vlSelfRef.rv32_soc__DOT____VlemCond_10 
                = vlSelfRef.rv32_soc__DOT__st_strb;
        } else {
            vlCovHit50 |= 2U;
            vlSelfRef.rv32_soc__DOT____VlemCond_10 = 0U;
        }
        vlSelfRef.mem_wstrb = vlSelfRef.rv32_soc__DOT____VlemCond_10;
        if (((IData)(rv32_soc__DOT__exec_ls) & (IData)(vlSelfRef.rv32_soc__DOT__is_store))) {
            vlCovHit49 |= 536870912U;
        }
        if ((1U & (~ (IData)(vlSelfRef.rv32_soc__DOT__is_store)))) {
            vlCovHit49 |= 1073741824U;
        }
        if ((1U & (~ (IData)(rv32_soc__DOT__exec_ls)))) {
            vlCovHit49 |= 2147483648U;
        }
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            vlCovHit49 |= 1048576U;
            vlSelfRef.rv32_soc__DOT____VlemCond_8 = vlSelfRef.rv32_soc__DOT__pc;
        } else {
            vlCovHit49 |= 8388608U;
            if (rv32_soc__DOT__exec_ls) {
                vlCovHit49 |= 2097152U;
                vlSelfRef.rv32_soc__DOT____VlemCond_8 
                    = (0xfffffffcU & vlSelfRef.rv32_soc__DOT__ls_addr);
            } else {
                vlCovHit49 |= 4194304U;
                vlSelfRef.rv32_soc__DOT____VlemCond_8 = 0U;
            }
        }
        vlSelfRef.mem_addr = vlSelfRef.rv32_soc__DOT____VlemCond_8;
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            vlCovHit49 |= 131072U;
        }
        if (vlSelfRef.rv32_soc__DOT__pc_bad) {
            vlCovHit49 |= 262144U;
        }
        if ((0U != (IData)(vlSelfRef.rv32_soc__DOT__state))) {
            vlCovHit49 |= 524288U;
        }
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            vlCovHit50 |= 32U;
            vlSelfRef.rv32_soc__DOT____VlemCond_11 
                = vlSelfRef.rv32_soc__DOT__mem[(0x000003ffU 
                                                & (vlSelfRef.rv32_soc__DOT__pc 
                                                   >> 2U))];
        } else {
            vlCovHit50 |= 256U;
            if (((IData)(rv32_soc__DOT__exec_ls) & (IData)(vlSelfRef.rv32_soc__DOT__is_load))) {
                vlCovHit50 |= 64U;
                vlSelfRef.rv32_soc__DOT____VlemCond_11 
                    = vlSelfRef.rv32_soc__DOT__mem[
                    (0x000003ffU & (vlSelfRef.rv32_soc__DOT__ls_addr 
                                    >> 2U))];
            } else {
                vlCovHit50 |= 128U;
                vlSelfRef.rv32_soc__DOT____VlemCond_11 = 0U;
            }
        }
        vlSelfRef.mem_rdata = vlSelfRef.rv32_soc__DOT____VlemCond_11;
        if (((0U == (IData)(vlSelfRef.rv32_soc__DOT__state)) 
             & (~ (IData)(vlSelfRef.rv32_soc__DOT__pc_bad)))) {
            vlCovHit50 |= 4U;
        }
        if (vlSelfRef.rv32_soc__DOT__pc_bad) {
            vlCovHit50 |= 8U;
        }
        if ((0U != (IData)(vlSelfRef.rv32_soc__DOT__state))) {
            vlCovHit50 |= 16U;
        }
        if (vlcu_cov_any(((IData)(vlSelfRef.mem_valid) 
                          ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__mem_valid)))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 6, vlSelfRef.mem_valid, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_valid);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_valid 
                = vlSelfRef.mem_valid;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.mem_wdata 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wdata)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 74, vlSelfRef.mem_wdata, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wdata);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wdata 
                = vlSelfRef.mem_wdata;
        }
        if (vlcu_cov_any((0U != ((IData)(vlSelfRef.mem_wstrb) 
                                 ^ (IData)(vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wstrb))))) {
            VL_COV_TOGGLE_CHG_ST_I(4, vlSelf->__Vcoverage + 138, vlSelfRef.mem_wstrb, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wstrb);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_wstrb 
                = vlSelfRef.mem_wstrb;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.mem_addr 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__mem_addr)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 10, vlSelfRef.mem_addr, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_addr);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_addr 
                = vlSelfRef.mem_addr;
        }
        if (vlcu_cov_any((0U != (vlSelfRef.mem_rdata 
                                 ^ vlSelfRef.rv32_soc__DOT____Vtogcov__mem_rdata)))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 146, vlSelfRef.mem_rdata, vlSelfRef.rv32_soc__DOT____Vtogcov__mem_rdata);
            vlSelfRef.rv32_soc__DOT____Vtogcov__mem_rdata 
                = vlSelfRef.mem_rdata;
        }
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 832, vlCovHit26);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 864, vlCovHit27);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 928, vlCovHit29);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 960, vlCovHit30);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1024, vlCovHit32);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1088, vlCovHit34);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1152, vlCovHit36);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1184, vlCovHit37);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1248, vlCovHit39);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1312, vlCovHit41);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1344, vlCovHit42);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1536, vlCovHit48);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1568, vlCovHit49);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1600, vlCovHit50);
        vlcu_cov_flush_hits(vlSelf->__Vcoverage + 1632, vlCovHit51);
    }

    VLCU_D static void Vrv32_cov___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vrv32_cov___024root___eval_phase__act(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__act
            vlSelfRef.__VactTriggered[0U] = (QData)((IData)(
                                                            ((IData)(vlSelfRef.clk) 
                                                             & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__1)))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__1 
                = vlSelfRef.clk;
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrv32_cov___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vrv32_cov___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vrv32_cov___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vrv32_cov___024root___eval_phase__nba(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vrv32_cov___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vrv32_cov___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vrv32_cov___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vrv32_cov___024root___eval(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __VicoIterCount;
        IData/*31:0*/ __VnbaIterCount;
        // Body
        __VicoIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VicoIterCount)))) {
#ifdef VL_DEBUG
                Vrv32_cov___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 16, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vrv32_cov___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vrv32_cov___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 16, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vrv32_cov___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 16, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vrv32_cov___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vrv32_cov___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vrv32_cov___024root___ctor_var_reset(Vrv32_cov___024root* vlSelf, VlcuContext& vlCtx) {
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
        vlSelf->rv32_soc__DOT__imm_b = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 8292269646142298134ull);
        vlSelf->rv32_soc__DOT__imm_u = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 9881345184067898959ull);
        vlSelf->rv32_soc__DOT__imm_j = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 1258952058935086896ull);
        vlSelf->rv32_soc__DOT__is_jalr = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 14499488222146277783ull);
        vlSelf->rv32_soc__DOT__is_branch = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 578062292577486727ull);
        vlSelf->rv32_soc__DOT__is_load = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 2994363835360905129ull);
        vlSelf->rv32_soc__DOT__is_store = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 7894171225882899832ull);
        vlSelf->rv32_soc__DOT__is_alui = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 17494699088985813507ull);
        vlSelf->rv32_soc__DOT__is_alur = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 11965279755098390934ull);
        vlSelf->rv32_soc__DOT__rs1_val = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 4619132075120175768ull);
        vlSelf->rv32_soc__DOT__alu_out = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 6575204758842702413ull);
        vlSelf->rv32_soc__DOT__take_branch = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 13704323896175277489ull);
        vlSelf->rv32_soc__DOT__ls_addr = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 327928301318749977ull);
        vlSelf->rv32_soc__DOT__ls_misaligned = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 13860169767913532206ull);
        vlSelf->rv32_soc__DOT__st_data = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 620096875113001671ull);
        vlSelf->rv32_soc__DOT__st_strb = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 9732364407211009767ull);
        vlSelf->rv32_soc__DOT__ld_val = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 7331941586127922612ull);
        vlSelf->rv32_soc__DOT__pc_bad = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16411224859674924017ull);
        vlSelf->rv32_soc__DOT____Vtogcov__clk = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__resetn = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__trap = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__mem_valid = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__mem_instr = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__mem_addr = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__mem_wdata = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__mem_wstrb = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__mem_rdata = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__retired = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__pc = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__insn = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__ld_addr = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__state = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__opcode = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__rd = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__funct3 = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__rs1 = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__rs2 = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__f7b5 = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__imm_i = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__imm_s = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__imm_b = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__imm_u = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__imm_j = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__is_lui = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__is_auipc = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__is_jal = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__is_jalr = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__is_branch = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__is_load = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__is_store = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__is_alui = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__is_alur = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__rs1_val = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__rs2_val = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__alu_b = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__shamt = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__alu_sub = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__alu_out = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__take_branch = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__ls_addr = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__ls_misaligned = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__ls_oob = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__st_data = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__st_strb = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__ld_word = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__ld_shift = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__ld_val = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__pc_bad = 0;
        vlSelf->rv32_soc__DOT____Vtogcov__exec_ls = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__resetn__0 = 0;
        vlSelf->__VicoDidInit = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__1 = 0;
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
        vlSelf->__Vcoverage.zero(1664U);
        Vrv32_cov___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vrv32_cov___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vrv32_cov___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vrv32_cov___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vrv32_cov___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vrv32_cov___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vrv32_cov__cuda::info() {
    typedef Vrv32_cov__cuda T;
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
        {"rv32_soc__DOT____VlemCond_10", "rv32_soc.__VlemCond_10", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_rv32_soc__DOT____VlemCond_10},
        {"rv32_soc__DOT____VlemCond_6", "rv32_soc.__VlemCond_6", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_rv32_soc__DOT____VlemCond_6},
        {"rv32_soc__DOT__state", "rv32_soc.state", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_rv32_soc__DOT__state},
        {"rv32_soc__DOT__is_jalr", "rv32_soc.is_jalr", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__is_jalr},
        {"rv32_soc__DOT__is_branch", "rv32_soc.is_branch", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__is_branch},
        {"rv32_soc__DOT__is_load", "rv32_soc.is_load", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__is_load},
        {"rv32_soc__DOT__is_store", "rv32_soc.is_store", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__is_store},
        {"rv32_soc__DOT__is_alui", "rv32_soc.is_alui", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__is_alui},
        {"rv32_soc__DOT__is_alur", "rv32_soc.is_alur", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__is_alur},
        {"rv32_soc__DOT__take_branch", "rv32_soc.take_branch", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__take_branch},
        {"rv32_soc__DOT__ls_misaligned", "rv32_soc.ls_misaligned", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__ls_misaligned},
        {"rv32_soc__DOT__st_strb", "rv32_soc.st_strb", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_rv32_soc__DOT__st_strb},
        {"rv32_soc__DOT__pc_bad", "rv32_soc.pc_bad", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT__pc_bad},
        {"rv32_soc__DOT____Vtogcov__clk", "rv32_soc.__Vtogcov__clk", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__clk},
        {"rv32_soc__DOT____Vtogcov__resetn", "rv32_soc.__Vtogcov__resetn", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__resetn},
        {"rv32_soc__DOT____Vtogcov__trap", "rv32_soc.__Vtogcov__trap", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__trap},
        {"rv32_soc__DOT____Vtogcov__mem_valid", "rv32_soc.__Vtogcov__mem_valid", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__mem_valid},
        {"rv32_soc__DOT____Vtogcov__mem_instr", "rv32_soc.__Vtogcov__mem_instr", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__mem_instr},
        {"rv32_soc__DOT____Vtogcov__mem_wstrb", "rv32_soc.__Vtogcov__mem_wstrb", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_rv32_soc__DOT____Vtogcov__mem_wstrb},
        {"rv32_soc__DOT____Vtogcov__state", "rv32_soc.__Vtogcov__state", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_rv32_soc__DOT____Vtogcov__state},
        {"rv32_soc__DOT____Vtogcov__opcode", "rv32_soc.__Vtogcov__opcode", VLCU_MEMBER_STATE, 1, 0, 1, 7, 6, 0, T::OFF_rv32_soc__DOT____Vtogcov__opcode},
        {"rv32_soc__DOT____Vtogcov__rd", "rv32_soc.__Vtogcov__rd", VLCU_MEMBER_STATE, 1, 0, 1, 5, 4, 0, T::OFF_rv32_soc__DOT____Vtogcov__rd},
        {"rv32_soc__DOT____Vtogcov__funct3", "rv32_soc.__Vtogcov__funct3", VLCU_MEMBER_STATE, 1, 0, 1, 3, 2, 0, T::OFF_rv32_soc__DOT____Vtogcov__funct3},
        {"rv32_soc__DOT____Vtogcov__rs1", "rv32_soc.__Vtogcov__rs1", VLCU_MEMBER_STATE, 1, 0, 1, 5, 4, 0, T::OFF_rv32_soc__DOT____Vtogcov__rs1},
        {"rv32_soc__DOT____Vtogcov__rs2", "rv32_soc.__Vtogcov__rs2", VLCU_MEMBER_STATE, 1, 0, 1, 5, 4, 0, T::OFF_rv32_soc__DOT____Vtogcov__rs2},
        {"rv32_soc__DOT____Vtogcov__f7b5", "rv32_soc.__Vtogcov__f7b5", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__f7b5},
        {"rv32_soc__DOT____Vtogcov__is_lui", "rv32_soc.__Vtogcov__is_lui", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__is_lui},
        {"rv32_soc__DOT____Vtogcov__is_auipc", "rv32_soc.__Vtogcov__is_auipc", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__is_auipc},
        {"rv32_soc__DOT____Vtogcov__is_jal", "rv32_soc.__Vtogcov__is_jal", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__is_jal},
        {"rv32_soc__DOT____Vtogcov__is_jalr", "rv32_soc.__Vtogcov__is_jalr", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__is_jalr},
        {"rv32_soc__DOT____Vtogcov__is_branch", "rv32_soc.__Vtogcov__is_branch", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__is_branch},
        {"rv32_soc__DOT____Vtogcov__is_load", "rv32_soc.__Vtogcov__is_load", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__is_load},
        {"rv32_soc__DOT____Vtogcov__is_store", "rv32_soc.__Vtogcov__is_store", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__is_store},
        {"rv32_soc__DOT____Vtogcov__is_alui", "rv32_soc.__Vtogcov__is_alui", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__is_alui},
        {"rv32_soc__DOT____Vtogcov__is_alur", "rv32_soc.__Vtogcov__is_alur", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__is_alur},
        {"rv32_soc__DOT____Vtogcov__shamt", "rv32_soc.__Vtogcov__shamt", VLCU_MEMBER_STATE, 1, 0, 1, 5, 4, 0, T::OFF_rv32_soc__DOT____Vtogcov__shamt},
        {"rv32_soc__DOT____Vtogcov__alu_sub", "rv32_soc.__Vtogcov__alu_sub", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__alu_sub},
        {"rv32_soc__DOT____Vtogcov__take_branch", "rv32_soc.__Vtogcov__take_branch", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__take_branch},
        {"rv32_soc__DOT____Vtogcov__ls_misaligned", "rv32_soc.__Vtogcov__ls_misaligned", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__ls_misaligned},
        {"rv32_soc__DOT____Vtogcov__ls_oob", "rv32_soc.__Vtogcov__ls_oob", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__ls_oob},
        {"rv32_soc__DOT____Vtogcov__st_strb", "rv32_soc.__Vtogcov__st_strb", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_rv32_soc__DOT____Vtogcov__st_strb},
        {"rv32_soc__DOT____Vtogcov__pc_bad", "rv32_soc.__Vtogcov__pc_bad", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__pc_bad},
        {"rv32_soc__DOT____Vtogcov__exec_ls", "rv32_soc.__Vtogcov__exec_ls", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rv32_soc__DOT____Vtogcov__exec_ls},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__resetn__0", "__Vtrigprevexpr___TOP__resetn__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__resetn__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"rv32_soc__DOT____VlemCond_11", "rv32_soc.__VlemCond_11", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____VlemCond_11},
        {"rv32_soc__DOT____VlemCond_9", "rv32_soc.__VlemCond_9", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____VlemCond_9},
        {"rv32_soc__DOT____VlemCond_8", "rv32_soc.__VlemCond_8", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____VlemCond_8},
        {"rv32_soc__DOT____VlemCond_7", "rv32_soc.__VlemCond_7", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____VlemCond_7},
        {"rv32_soc__DOT____VlemCond_5", "rv32_soc.__VlemCond_5", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____VlemCond_5},
        {"rv32_soc__DOT____VlemCond_4", "rv32_soc.__VlemCond_4", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____VlemCond_4},
        {"rv32_soc__DOT____VlemCond_3", "rv32_soc.__VlemCond_3", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____VlemCond_3},
        {"rv32_soc__DOT____VlemCond_2", "rv32_soc.__VlemCond_2", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____VlemCond_2},
        {"rv32_soc__DOT____VlemCond_1", "rv32_soc.__VlemCond_1", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____VlemCond_1},
        {"rv32_soc__DOT____VlemCond_0", "rv32_soc.__VlemCond_0", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____VlemCond_0},
        {"rv32_soc__DOT__pc", "rv32_soc.pc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__pc},
        {"rv32_soc__DOT__insn", "rv32_soc.insn", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__insn},
        {"rv32_soc__DOT__ld_addr", "rv32_soc.ld_addr", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__ld_addr},
        {"rv32_soc__DOT__instret", "rv32_soc.instret", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__instret},
        {"rv32_soc__DOT__imm_i", "rv32_soc.imm_i", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__imm_i},
        {"rv32_soc__DOT__imm_b", "rv32_soc.imm_b", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__imm_b},
        {"rv32_soc__DOT__imm_u", "rv32_soc.imm_u", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__imm_u},
        {"rv32_soc__DOT__imm_j", "rv32_soc.imm_j", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__imm_j},
        {"rv32_soc__DOT__rs1_val", "rv32_soc.rs1_val", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__rs1_val},
        {"rv32_soc__DOT__alu_out", "rv32_soc.alu_out", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__alu_out},
        {"rv32_soc__DOT__ls_addr", "rv32_soc.ls_addr", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__ls_addr},
        {"rv32_soc__DOT__st_data", "rv32_soc.st_data", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__st_data},
        {"rv32_soc__DOT__ld_val", "rv32_soc.ld_val", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT__ld_val},
        {"rv32_soc__DOT____Vtogcov__mem_addr", "rv32_soc.__Vtogcov__mem_addr", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__mem_addr},
        {"rv32_soc__DOT____Vtogcov__mem_wdata", "rv32_soc.__Vtogcov__mem_wdata", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__mem_wdata},
        {"rv32_soc__DOT____Vtogcov__mem_rdata", "rv32_soc.__Vtogcov__mem_rdata", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__mem_rdata},
        {"rv32_soc__DOT____Vtogcov__retired", "rv32_soc.__Vtogcov__retired", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__retired},
        {"rv32_soc__DOT____Vtogcov__pc", "rv32_soc.__Vtogcov__pc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__pc},
        {"rv32_soc__DOT____Vtogcov__insn", "rv32_soc.__Vtogcov__insn", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__insn},
        {"rv32_soc__DOT____Vtogcov__ld_addr", "rv32_soc.__Vtogcov__ld_addr", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__ld_addr},
        {"rv32_soc__DOT____Vtogcov__imm_i", "rv32_soc.__Vtogcov__imm_i", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__imm_i},
        {"rv32_soc__DOT____Vtogcov__imm_s", "rv32_soc.__Vtogcov__imm_s", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__imm_s},
        {"rv32_soc__DOT____Vtogcov__imm_b", "rv32_soc.__Vtogcov__imm_b", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__imm_b},
        {"rv32_soc__DOT____Vtogcov__imm_u", "rv32_soc.__Vtogcov__imm_u", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__imm_u},
        {"rv32_soc__DOT____Vtogcov__imm_j", "rv32_soc.__Vtogcov__imm_j", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__imm_j},
        {"rv32_soc__DOT____Vtogcov__rs1_val", "rv32_soc.__Vtogcov__rs1_val", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__rs1_val},
        {"rv32_soc__DOT____Vtogcov__rs2_val", "rv32_soc.__Vtogcov__rs2_val", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__rs2_val},
        {"rv32_soc__DOT____Vtogcov__alu_b", "rv32_soc.__Vtogcov__alu_b", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__alu_b},
        {"rv32_soc__DOT____Vtogcov__alu_out", "rv32_soc.__Vtogcov__alu_out", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__alu_out},
        {"rv32_soc__DOT____Vtogcov__ls_addr", "rv32_soc.__Vtogcov__ls_addr", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__ls_addr},
        {"rv32_soc__DOT____Vtogcov__st_data", "rv32_soc.__Vtogcov__st_data", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__st_data},
        {"rv32_soc__DOT____Vtogcov__ld_word", "rv32_soc.__Vtogcov__ld_word", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__ld_word},
        {"rv32_soc__DOT____Vtogcov__ld_shift", "rv32_soc.__Vtogcov__ld_shift", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__ld_shift},
        {"rv32_soc__DOT____Vtogcov__ld_val", "rv32_soc.__Vtogcov__ld_val", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rv32_soc__DOT____Vtogcov__ld_val},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"rv32_soc__DOT__mem", "rv32_soc.mem", VLCU_MEMBER_MEMORY, 4, 0, 1024, 32, 31, 0, T::OFF_rv32_soc__DOT__mem},
        {"rv32_soc__DOT__regs", "rv32_soc.regs", VLCU_MEMBER_STATE, 4, 0, 32, 32, 31, 0, T::OFF_rv32_soc__DOT__regs},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VicoTriggered", "__VicoTriggered", VLCU_MEMBER_STATE, 8, 0, 2, 64, 63, 0, T::OFF___VicoTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
        {"__Vcoverage", nullptr, VLCU_MEMBER_COVERAGE, 4, 0, 104, 32, 31, 0, T::OFF___Vcoverage},
    };
    static VlcuCoverTable t;
    if (t.points.empty()) {
        t.toggle(0, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 19, 24, ".rv32_soc", "v_toggle/rv32_soc", "clk");
        t.toggle(2, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 20, 24, ".rv32_soc", "v_toggle/rv32_soc", "resetn");
        t.toggle(4, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 21, 24, ".rv32_soc", "v_toggle/rv32_soc", "trap");
        t.toggle(6, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 22, 24, ".rv32_soc", "v_toggle/rv32_soc", "mem_valid");
        t.toggle(8, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 23, 24, ".rv32_soc", "v_toggle/rv32_soc", "mem_instr");
        t.toggle(10, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 24, 24, ".rv32_soc", "v_toggle/rv32_soc", "mem_addr");
        t.toggle(74, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 25, 24, ".rv32_soc", "v_toggle/rv32_soc", "mem_wdata");
        t.toggle(138, 0, 3, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 26, 24, ".rv32_soc", "v_toggle/rv32_soc", "mem_wstrb");
        t.toggle(146, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 27, 24, ".rv32_soc", "v_toggle/rv32_soc", "mem_rdata");
        t.toggle(210, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 28, 24, ".rv32_soc", "v_toggle/rv32_soc", "retired");
        t.toggle(274, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 34, 14, ".rv32_soc", "v_toggle/rv32_soc", "pc");
        t.toggle(338, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 35, 14, ".rv32_soc", "v_toggle/rv32_soc", "insn");
        t.toggle(402, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 36, 14, ".rv32_soc", "v_toggle/rv32_soc", "ld_addr");
        t.toggle(210, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 37, 14, ".rv32_soc", "v_toggle/rv32_soc", "instret");
        t.toggle(466, 0, 1, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 38, 14, ".rv32_soc", "v_toggle/rv32_soc", "state");
        t.toggle(470, 0, 6, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 41, 14, ".rv32_soc", "v_toggle/rv32_soc", "opcode");
        t.toggle(484, 0, 4, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 42, 14, ".rv32_soc", "v_toggle/rv32_soc", "rd");
        t.toggle(494, 0, 2, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 43, 14, ".rv32_soc", "v_toggle/rv32_soc", "funct3");
        t.toggle(500, 0, 4, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 44, 14, ".rv32_soc", "v_toggle/rv32_soc", "rs1");
        t.toggle(510, 0, 4, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 45, 14, ".rv32_soc", "v_toggle/rv32_soc", "rs2");
        t.toggle(520, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 46, 14, ".rv32_soc", "v_toggle/rv32_soc", "f7b5");
        t.toggle(522, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 47, 15, ".rv32_soc", "v_toggle/rv32_soc", "imm_i");
        t.toggle(586, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 48, 15, ".rv32_soc", "v_toggle/rv32_soc", "imm_s");
        t.toggle(650, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 49, 15, ".rv32_soc", "v_toggle/rv32_soc", "imm_b");
        t.toggle(714, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 50, 15, ".rv32_soc", "v_toggle/rv32_soc", "imm_u");
        t.toggle(778, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 51, 15, ".rv32_soc", "v_toggle/rv32_soc", "imm_j");
        t.toggle(842, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 53, 8, ".rv32_soc", "v_toggle/rv32_soc", "is_lui");
        t.toggle(844, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 54, 8, ".rv32_soc", "v_toggle/rv32_soc", "is_auipc");
        t.toggle(846, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 55, 8, ".rv32_soc", "v_toggle/rv32_soc", "is_jal");
        t.toggle(848, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 56, 8, ".rv32_soc", "v_toggle/rv32_soc", "is_jalr");
        t.other(850, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 56, 39, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h67)==1 && (funct3 == 3'h0)==1) => 1", "");
        t.other(851, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 56, 39, ".rv32_soc", "v_expr/rv32_soc", "((funct3 == 3'h0)==0) => 0", "");
        t.other(852, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 56, 39, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h67)==0) => 0", "");
        t.toggle(853, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 57, 8, ".rv32_soc", "v_toggle/rv32_soc", "is_branch");
        t.other(855, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 57, 61, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h63)==1 && (funct3 != 3'h2)==1 && (funct3 != 3'h3)==1) => 1", "");
        t.other(856, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 57, 61, ".rv32_soc", "v_expr/rv32_soc", "((funct3 != 3'h3)==0) => 0", "");
        t.other(857, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 57, 61, ".rv32_soc", "v_expr/rv32_soc", "((funct3 != 3'h2)==0) => 0", "");
        t.other(858, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 57, 61, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h63)==0) => 0", "");
        t.toggle(859, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 58, 8, ".rv32_soc", "v_toggle/rv32_soc", "is_load");
        t.other(861, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 58, 39, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h3)==1 && (funct3 == 3'h5)==1) => 1", "");
        t.other(862, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 58, 39, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h3)==1 && (funct3 == 3'h4)==1) => 1", "");
        t.other(863, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 58, 39, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h3)==1 && (funct3 == 3'h2)==1) => 1", "");
        t.other(864, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 58, 39, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h3)==1 && (funct3 == 3'h1)==1) => 1", "");
        t.other(865, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 58, 39, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h3)==1 && (funct3 == 3'h0)==1) => 1", "");
        t.other(866, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 58, 39, ".rv32_soc", "v_expr/rv32_soc", "((funct3 == 3'h0)==0 && (funct3 == 3'h1)==0 && (funct3 == 3'h2)==0 && (funct3 == 3'h4)==0 && (funct3 == 3'h5)==0) => 0", "");
        t.other(867, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 58, 39, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h3)==0) => 0", "");
        t.toggle(868, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 60, 8, ".rv32_soc", "v_toggle/rv32_soc", "is_store");
        t.other(870, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 60, 40, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h23)==1 && (funct3 == 3'h2)==1) => 1", "");
        t.other(871, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 60, 40, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h23)==1 && (funct3 == 3'h1)==1) => 1", "");
        t.other(872, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 60, 40, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h23)==1 && (funct3 == 3'h0)==1) => 1", "");
        t.other(873, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 60, 40, ".rv32_soc", "v_expr/rv32_soc", "((funct3 == 3'h0)==0 && (funct3 == 3'h1)==0 && (funct3 == 3'h2)==0) => 0", "");
        t.other(874, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 60, 40, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h23)==0) => 0", "");
        t.toggle(875, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 61, 8, ".rv32_soc", "v_toggle/rv32_soc", "is_alui");
        t.other(877, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 63, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h13)==1 && (insn[31:25] == 7'h0)==1 && (insn[31:25] == 7'h20)==1) => 1", "");
        t.other(878, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 63, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h13)==1 && (insn[31:25] == 7'h0)==1) => 1", "");
        t.other(879, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 63, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h13)==1 && (insn[31:25] == 7'h0)==1 && (funct3 != 3'h5)==1) => 1", "");
        t.other(880, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 63, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h13)==1 && (funct3 != 3'h1)==1 && (insn[31:25] == 7'h20)==1) => 1", "");
        t.other(881, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 63, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h13)==1 && (funct3 != 3'h1)==1 && (insn[31:25] == 7'h0)==1) => 1", "");
        t.other(882, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 63, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h13)==1 && (funct3 != 3'h1)==1 && (funct3 != 3'h5)==1) => 1", "");
        t.other(883, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 63, 18, ".rv32_soc", "v_expr/rv32_soc", "((funct3 != 3'h5)==0 && (insn[31:25] == 7'h0)==0 && (insn[31:25] == 7'h20)==0) => 0", "");
        t.other(884, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 63, 18, ".rv32_soc", "v_expr/rv32_soc", "((funct3 != 3'h1)==0 && (insn[31:25] == 7'h0)==0) => 0", "");
        t.other(885, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 63, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h13)==0) => 0", "");
        t.toggle(886, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 64, 8, ".rv32_soc", "v_toggle/rv32_soc", "is_alur");
        t.other(888, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 65, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h33)==1 && (insn[31:25] == 7'h20)==1 && (funct3 == 3'h5)==1) => 1", "");
        t.other(889, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 65, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h33)==1 && (insn[31:25] == 7'h20)==1 && (funct3 == 3'h0)==1) => 1", "");
        t.other(890, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 65, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h33)==1 && (insn[31:25] == 7'h0)==1) => 1", "");
        t.other(891, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 65, 18, ".rv32_soc", "v_expr/rv32_soc", "((insn[31:25] == 7'h0)==0 && (funct3 == 3'h0)==0 && (funct3 == 3'h5)==0) => 0", "");
        t.other(892, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 65, 18, ".rv32_soc", "v_expr/rv32_soc", "((insn[31:25] == 7'h0)==0 && (insn[31:25] == 7'h20)==0) => 0", "");
        t.other(893, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 65, 18, ".rv32_soc", "v_expr/rv32_soc", "((opcode == 7'h33)==0) => 0", "");
        t.toggle(894, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 68, 15, ".rv32_soc", "v_toggle/rv32_soc", "rs1_val");
        t.other(958, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 68, 30, ".rv32_soc", "v_expr/rv32_soc", "((rs1 == 5'h0)==1) => 1", "");
        t.other(959, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 68, 30, ".rv32_soc", "v_expr/rv32_soc", "((rs1 == 5'h0)==0) => 0", "");
        t.other(960, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 68, 41, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "68");
        t.other(961, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 68, 42, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "68");
        t.toggle(962, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 69, 15, ".rv32_soc", "v_toggle/rv32_soc", "rs2_val");
        t.other(1026, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 69, 30, ".rv32_soc", "v_expr/rv32_soc", "((rs2 == 5'h0)==1) => 1", "");
        t.other(1027, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 69, 30, ".rv32_soc", "v_expr/rv32_soc", "((rs2 == 5'h0)==0) => 0", "");
        t.other(1028, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 69, 41, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "69");
        t.other(1029, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 69, 42, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "69");
        t.toggle(1030, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 72, 15, ".rv32_soc", "v_toggle/rv32_soc", "alu_b");
        t.other(1094, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 72, 23, ".rv32_soc", "v_expr/rv32_soc", "(is_alur==1) => 1", "");
        t.other(1095, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 72, 23, ".rv32_soc", "v_expr/rv32_soc", "(is_alur==0) => 0", "");
        t.other(1096, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 72, 33, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "72");
        t.other(1097, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 72, 34, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "72");
        t.toggle(1098, 0, 4, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 73, 15, ".rv32_soc", "v_toggle/rv32_soc", "shamt");
        t.toggle(1108, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 74, 15, ".rv32_soc", "v_toggle/rv32_soc", "alu_sub");
        t.other(1110, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 74, 33, ".rv32_soc", "v_expr/rv32_soc", "(is_alur==1 && f7b5==1) => 1", "");
        t.other(1111, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 74, 33, ".rv32_soc", "v_expr/rv32_soc", "(f7b5==0) => 0", "");
        t.other(1112, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 74, 33, ".rv32_soc", "v_expr/rv32_soc", "(is_alur==0) => 0", "");
        t.toggle(1113, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 75, 15, ".rv32_soc", "v_toggle/rv32_soc", "alu_out");
        t.other(1177, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 78, 26, ".rv32_soc", "v_expr/rv32_soc", "(alu_sub==1) => 1", "");
        t.other(1178, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 78, 26, ".rv32_soc", "v_expr/rv32_soc", "(alu_sub==0) => 0", "");
        t.other(1179, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 78, 44, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "78");
        t.other(1180, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 78, 45, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "78");
        t.other(1181, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 78, 13, ".rv32_soc", "v_line/rv32_soc", "case", "78");
        t.other(1182, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 79, 13, ".rv32_soc", "v_line/rv32_soc", "case", "79");
        t.other(1183, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 80, 13, ".rv32_soc", "v_line/rv32_soc", "case", "80");
        t.other(1184, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 81, 13, ".rv32_soc", "v_line/rv32_soc", "case", "81");
        t.other(1185, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 82, 13, ".rv32_soc", "v_line/rv32_soc", "case", "82");
        t.other(1186, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 83, 26, ".rv32_soc", "v_expr/rv32_soc", "(f7b5==1) => 1", "");
        t.other(1187, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 83, 26, ".rv32_soc", "v_expr/rv32_soc", "(f7b5==0) => 0", "");
        t.other(1188, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 83, 60, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "83");
        t.other(1189, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 83, 61, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "83");
        t.other(1190, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 83, 13, ".rv32_soc", "v_line/rv32_soc", "case", "83");
        t.other(1191, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 84, 13, ".rv32_soc", "v_line/rv32_soc", "case", "84");
        t.other(1192, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 85, 7, ".rv32_soc", "v_line/rv32_soc", "case", "85");
        t.other(1193, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 76, 3, ".rv32_soc", "v_line/rv32_soc", "block", "76-77");
        t.toggle(1194, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 89, 7, ".rv32_soc", "v_toggle/rv32_soc", "take_branch");
        t.other(1196, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 92, 13, ".rv32_soc", "v_line/rv32_soc", "case", "92");
        t.other(1197, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 93, 13, ".rv32_soc", "v_line/rv32_soc", "case", "93");
        t.other(1198, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 94, 13, ".rv32_soc", "v_line/rv32_soc", "case", "94");
        t.other(1199, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 95, 13, ".rv32_soc", "v_line/rv32_soc", "case", "95");
        t.other(1200, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 96, 13, ".rv32_soc", "v_line/rv32_soc", "case", "96");
        t.other(1201, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 97, 7, ".rv32_soc", "v_line/rv32_soc", "case", "97");
        t.other(1202, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 90, 3, ".rv32_soc", "v_line/rv32_soc", "block", "90-91");
        t.toggle(1203, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 102, 15, ".rv32_soc", "v_toggle/rv32_soc", "ls_addr");
        t.other(1267, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 102, 36, ".rv32_soc", "v_expr/rv32_soc", "(is_store==1) => 1", "");
        t.other(1268, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 102, 36, ".rv32_soc", "v_expr/rv32_soc", "(is_store==0) => 0", "");
        t.toggle(1269, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 103, 15, ".rv32_soc", "v_toggle/rv32_soc", "ls_misaligned");
        t.other(1271, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 103, 68, ".rv32_soc", "v_expr/rv32_soc", "((funct3[1:0] == 2'h2)==1 && (ls_addr[1:0] != 2'h0)==1) => 1", "");
        t.other(1272, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 103, 68, ".rv32_soc", "v_expr/rv32_soc", "((funct3[1:0] == 2'h1)==1 && ls_addr[0]==1) => 1", "");
        t.other(1273, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 103, 68, ".rv32_soc", "v_expr/rv32_soc", "(ls_addr[0]==0 && (ls_addr[1:0] != 2'h0)==0) => 0", "");
        t.other(1274, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 103, 68, ".rv32_soc", "v_expr/rv32_soc", "(ls_addr[0]==0 && (funct3[1:0] == 2'h2)==0) => 0", "");
        t.other(1275, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 103, 68, ".rv32_soc", "v_expr/rv32_soc", "((funct3[1:0] == 2'h1)==0 && (ls_addr[1:0] != 2'h0)==0) => 0", "");
        t.other(1276, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 103, 68, ".rv32_soc", "v_expr/rv32_soc", "((funct3[1:0] == 2'h1)==0 && (funct3[1:0] == 2'h2)==0) => 0", "");
        t.toggle(1277, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 104, 15, ".rv32_soc", "v_toggle/rv32_soc", "ls_oob");
        t.toggle(1279, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 105, 15, ".rv32_soc", "v_toggle/rv32_soc", "st_data");
        t.other(1343, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 105, 38, ".rv32_soc", "v_expr/rv32_soc", "((funct3[1:0] == 2'h0)==1) => 1", "");
        t.other(1344, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 105, 38, ".rv32_soc", "v_expr/rv32_soc", "((funct3[1:0] == 2'h0)==0) => 0", "");
        t.other(1345, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 105, 52, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "105");
        t.other(1346, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 106, 54, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "106");
        t.other(1347, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 106, 55, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "106");
        t.other(1348, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 105, 53, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "");
        t.toggle(1349, 0, 3, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 107, 15, ".rv32_soc", "v_toggle/rv32_soc", "st_strb");
        t.other(1357, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 107, 38, ".rv32_soc", "v_expr/rv32_soc", "((funct3[1:0] == 2'h0)==1) => 1", "");
        t.other(1358, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 107, 38, ".rv32_soc", "v_expr/rv32_soc", "((funct3[1:0] == 2'h0)==0) => 0", "");
        t.other(1359, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 107, 59, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "107");
        t.other(1360, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 108, 61, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "108");
        t.other(1361, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 108, 62, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "108");
        t.other(1362, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 107, 60, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "");
        t.toggle(1363, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 110, 15, ".rv32_soc", "v_toggle/rv32_soc", "ld_word");
        t.toggle(1427, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 111, 15, ".rv32_soc", "v_toggle/rv32_soc", "ld_shift");
        t.toggle(1491, 0, 31, 1, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 112, 15, ".rv32_soc", "v_toggle/rv32_soc", "ld_val");
        t.other(1555, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 112, 32, ".rv32_soc", "v_expr/rv32_soc", "((funct3 == 3'h0)==1) => 1", "");
        t.other(1556, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 112, 32, ".rv32_soc", "v_expr/rv32_soc", "((funct3 == 3'h0)==0) => 0", "");
        t.other(1557, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 112, 45, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "112");
        t.other(1558, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 113, 47, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "113");
        t.other(1559, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 114, 47, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "114");
        t.other(1560, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 115, 47, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "115");
        t.other(1561, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 115, 48, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "115");
        t.other(1562, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 114, 48, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "");
        t.other(1563, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 113, 48, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "");
        t.other(1564, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 112, 46, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "");
        t.toggle(1565, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 117, 8, ".rv32_soc", "v_toggle/rv32_soc", "pc_bad");
        t.other(1567, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 117, 34, ".rv32_soc", "v_expr/rv32_soc", "((pc[31:2] >= MEM_WORDS)==1) => 1", "");
        t.other(1568, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 117, 34, ".rv32_soc", "v_expr/rv32_soc", "((pc[1:0] != 2'h0)==1) => 1", "");
        t.other(1569, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 117, 34, ".rv32_soc", "v_expr/rv32_soc", "((pc[1:0] != 2'h0)==0 && (pc[31:2] >= MEM_WORDS)==0) => 0", "");
        t.toggle(1570, 0, 0, 0, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 120, 8, ".rv32_soc", "v_toggle/rv32_soc", "exec_ls");
        t.other(1572, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 120, 77, ".rv32_soc", "v_expr/rv32_soc", "((state == S_EXEC)==1 && is_store==1 && ls_misaligned==0 && ls_oob==0) => 1", "");
        t.other(1573, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 120, 77, ".rv32_soc", "v_expr/rv32_soc", "((state == S_EXEC)==1 && is_load==1 && ls_misaligned==0 && ls_oob==0) => 1", "");
        t.other(1574, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 120, 77, ".rv32_soc", "v_expr/rv32_soc", "(ls_oob==1) => 0", "");
        t.other(1575, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 120, 77, ".rv32_soc", "v_expr/rv32_soc", "(ls_misaligned==1) => 0", "");
        t.other(1576, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 120, 77, ".rv32_soc", "v_expr/rv32_soc", "(is_load==0 && is_store==0) => 0", "");
        t.other(1577, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 120, 77, ".rv32_soc", "v_expr/rv32_soc", "((state == S_EXEC)==0) => 0", "");
        t.other(1578, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 122, 52, ".rv32_soc", "v_expr/rv32_soc", "(exec_ls==1) => 1", "");
        t.other(1579, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 122, 52, ".rv32_soc", "v_expr/rv32_soc", "((state == S_FETCH)==1 && pc_bad==0) => 1", "");
        t.other(1580, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 122, 52, ".rv32_soc", "v_expr/rv32_soc", "(pc_bad==1 && exec_ls==0) => 0", "");
        t.other(1581, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 122, 52, ".rv32_soc", "v_expr/rv32_soc", "((state == S_FETCH)==0 && exec_ls==0) => 0", "");
        t.other(1582, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 123, 39, ".rv32_soc", "v_expr/rv32_soc", "((state == S_FETCH)==1 && pc_bad==0) => 1", "");
        t.other(1583, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 123, 39, ".rv32_soc", "v_expr/rv32_soc", "(pc_bad==1) => 0", "");
        t.other(1584, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 123, 39, ".rv32_soc", "v_expr/rv32_soc", "((state == S_FETCH)==0) => 0", "");
        t.other(1585, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 124, 39, ".rv32_soc", "v_expr/rv32_soc", "((state == S_FETCH)==1 && pc_bad==0) => 1", "");
        t.other(1586, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 124, 39, ".rv32_soc", "v_expr/rv32_soc", "(pc_bad==1) => 0", "");
        t.other(1587, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 124, 39, ".rv32_soc", "v_expr/rv32_soc", "((state == S_FETCH)==0) => 0", "");
        t.other(1588, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 124, 53, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "124");
        t.other(1589, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 124, 68, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "124");
        t.other(1590, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 124, 69, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "124");
        t.other(1591, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 124, 54, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "");
        t.other(1592, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 125, 31, ".rv32_soc", "v_expr/rv32_soc", "(exec_ls==1 && is_store==1) => 1", "");
        t.other(1593, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 125, 31, ".rv32_soc", "v_expr/rv32_soc", "(is_store==0) => 0", "");
        t.other(1594, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 125, 31, ".rv32_soc", "v_expr/rv32_soc", "(exec_ls==0) => 0", "");
        t.other(1595, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 125, 46, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "125");
        t.other(1596, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 125, 47, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "125");
        t.other(1597, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 126, 31, ".rv32_soc", "v_expr/rv32_soc", "(exec_ls==1 && is_store==1) => 1", "");
        t.other(1598, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 126, 31, ".rv32_soc", "v_expr/rv32_soc", "(is_store==0) => 0", "");
        t.other(1599, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 126, 31, ".rv32_soc", "v_expr/rv32_soc", "(exec_ls==0) => 0", "");
        t.other(1600, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 126, 46, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "126");
        t.other(1601, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 126, 47, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "126");
        t.other(1602, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 127, 40, ".rv32_soc", "v_expr/rv32_soc", "((state == S_FETCH)==1 && pc_bad==0) => 1", "");
        t.other(1603, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 127, 40, ".rv32_soc", "v_expr/rv32_soc", "(pc_bad==1) => 0", "");
        t.other(1604, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 127, 40, ".rv32_soc", "v_expr/rv32_soc", "((state == S_FETCH)==0) => 0", "");
        t.other(1605, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 127, 57, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "127");
        t.other(1606, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 128, 50, ".rv32_soc", "v_branch/rv32_soc", "cond_then", "128");
        t.other(1607, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 128, 51, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "128");
        t.other(1608, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 127, 58, ".rv32_soc", "v_branch/rv32_soc", "cond_else", "");
        t.other(1609, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 140, 11, ".rv32_soc", "v_branch/rv32_soc", "if", "140");
        t.other(1610, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 140, 12, ".rv32_soc", "v_branch/rv32_soc", "else", "141-143");
        t.other(1611, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 139, 16, ".rv32_soc", "v_line/rv32_soc", "case", "139");
        t.other(1612, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 151, 13, ".rv32_soc", "v_branch/rv32_soc", "if", "151");
        t.other(1613, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 151, 14, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1614, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 153, 13, ".rv32_soc", "v_branch/rv32_soc", "if", "153");
        t.other(1615, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 153, 14, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1616, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 155, 13, ".rv32_soc", "v_branch/rv32_soc", "if", "155");
        t.other(1617, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 155, 14, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1618, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 158, 13, ".rv32_soc", "v_branch/rv32_soc", "if", "158");
        t.other(1619, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 158, 14, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1620, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 161, 13, ".rv32_soc", "v_branch/rv32_soc", "if", "161");
        t.other(1621, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 161, 14, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1622, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 163, 13, ".rv32_soc", "v_branch/rv32_soc", "if", "163");
        t.other(1623, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 163, 14, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1624, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 165, 13, ".rv32_soc", "v_branch/rv32_soc", "if", "165-168");
        t.other(1625, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 165, 14, ".rv32_soc", "v_branch/rv32_soc", "else", "169-173");
        t.other(1626, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 165, 31, ".rv32_soc", "v_expr/rv32_soc", "(ls_oob==1) => 1", "");
        t.other(1627, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 165, 31, ".rv32_soc", "v_expr/rv32_soc", "(ls_misaligned==1) => 1", "");
        t.other(1628, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 165, 31, ".rv32_soc", "v_expr/rv32_soc", "(ls_misaligned==0 && ls_oob==0) => 0", "");
        t.other(1629, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 181, 15, ".rv32_soc", "v_branch/rv32_soc", "if", "181");
        t.other(1630, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 181, 16, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1631, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 182, 15, ".rv32_soc", "v_branch/rv32_soc", "if", "182");
        t.other(1632, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 182, 16, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1633, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 183, 15, ".rv32_soc", "v_branch/rv32_soc", "if", "183");
        t.other(1634, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 183, 16, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1635, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 184, 15, ".rv32_soc", "v_branch/rv32_soc", "if", "184");
        t.other(1636, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 184, 16, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1637, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 176, 13, ".rv32_soc", "v_branch/rv32_soc", "if", "176-179");
        t.other(1638, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 176, 14, ".rv32_soc", "v_branch/rv32_soc", "else", "180");
        t.other(1639, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 176, 31, ".rv32_soc", "v_expr/rv32_soc", "(ls_oob==1) => 1", "");
        t.other(1640, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 176, 31, ".rv32_soc", "v_expr/rv32_soc", "(ls_misaligned==1) => 1", "");
        t.other(1641, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 176, 31, ".rv32_soc", "v_expr/rv32_soc", "(ls_misaligned==0 && ls_oob==0) => 0", "");
        t.other(1642, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 175, 20, ".rv32_soc", "v_line/rv32_soc", "if", "175");
        t.other(1643, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 175, 21, ".rv32_soc", "v_line/rv32_soc", "else", "186-189");
        t.other(1644, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 164, 20, ".rv32_soc", "v_line/rv32_soc", "elsif", "164");
        t.other(1645, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 162, 20, ".rv32_soc", "v_line/rv32_soc", "elsif", "162");
        t.other(1646, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 162, 32, ".rv32_soc", "v_expr/rv32_soc", "(is_alur==1) => 1", "");
        t.other(1647, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 162, 32, ".rv32_soc", "v_expr/rv32_soc", "(is_alui==1) => 1", "");
        t.other(1648, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 162, 32, ".rv32_soc", "v_expr/rv32_soc", "(is_alui==0 && is_alur==0) => 0", "");
        t.other(1649, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 160, 20, ".rv32_soc", "v_line/rv32_soc", "elsif", "160");
        t.other(1650, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 157, 20, ".rv32_soc", "v_line/rv32_soc", "elsif", "157,159");
        t.other(1651, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 154, 20, ".rv32_soc", "v_line/rv32_soc", "elsif", "154,156");
        t.other(1652, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 152, 20, ".rv32_soc", "v_line/rv32_soc", "elsif", "152");
        t.other(1653, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 150, 11, ".rv32_soc", "v_line/rv32_soc", "elsif", "150");
        t.other(1654, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 146, 15, ".rv32_soc", "v_line/rv32_soc", "case", "146-149");
        t.other(1655, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 193, 11, ".rv32_soc", "v_branch/rv32_soc", "if", "193");
        t.other(1656, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 193, 12, ".rv32_soc", "v_branch/rv32_soc", "else", "");
        t.other(1657, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 192, 15, ".rv32_soc", "v_line/rv32_soc", "case", "192,194-196");
        t.other(1658, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 198, 9, ".rv32_soc", "v_line/rv32_soc", "case", "198");
        t.other(1659, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 133, 5, ".rv32_soc", "v_branch/rv32_soc", "if", "133-136");
        t.other(1660, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 133, 6, ".rv32_soc", "v_branch/rv32_soc", "else", "137-138");
        t.other(1661, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 133, 9, ".rv32_soc", "v_expr/rv32_soc", "(resetn==0) => 1", "");
        t.other(1662, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 133, 9, ".rv32_soc", "v_expr/rv32_soc", "(resetn==1) => 0", "");
        t.other(1663, "../verilator_b200/csrc/designs/rv32i/rv32_soc.v", 132, 3, ".rv32_soc", "v_line/rv32_soc", "block", "132");
        t.finish();
    }
    // one uint64 counter per bin for every warp of 32 instances, [warp][bin] in the block of __Vcoverage
    static const VlcuCoverScope coverScopes[] = {{0, 1664}};
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 7, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 5, 6, 8, 9, 62, 63, 64, 65, 66, 67, 68, 69, 70, 71, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 91, 92, 93, 94, 95, 96, 97, 98, 99, 100, 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "rv32_soc", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        t.points.data(), static_cast<uint32_t>(t.points.size()), 1664, T::OFF___Vcoverage,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 113, 0, 0xef0e7ce4001d23d2ULL, nullptr, 12, 12, 0, coverScopes, sizeof(coverScopes) / sizeof(coverScopes[0])};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vrv32_cov__cuda VLCU_TESTBENCHES);
