// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_display_wide"
#endif

// Constant pool
static __device__ const VlWide<128>/*4095:0*/ Vt_display_wide__ConstPool__CONST_hfdfd7487_0 = VlWide<128>{{
        0x3df238a1, 0xaff1fa14, 0x29374ee9, 0x08d979af,
        0xd23af6ff, 0x6e5ea313, 0x80569ffe, 0xd6cf5693,
        0x9df66488, 0xfe3438ee, 0x31995d7d, 0x3e6adfaa,
        0x90a81dc6, 0x4ca62961, 0x9452c81a, 0x56c4b42b,
        0xb4fe6756, 0xb849910e, 0xeab63ab7, 0x8fef89db,
        0xd829b7e7, 0x722cca2d, 0x475541fd, 0x9a6f82a8,
        0xe8efc18c, 0x677d5128, 0x7ee21205, 0xc0f63011,
        0xadcb6e42, 0xa5a2ddc2, 0x09714942, 0x1c48cf1c,
        0xe11165b7, 0x1beaa348, 0x273defe2, 0x98d02896,
        0x8c7329b8, 0xebc2e0aa, 0xb1a135a8, 0x9ae9e1db,
        0xcd6a21ca, 0x7907f2dc, 0x5a7d02c0, 0x27557c53,
        0xc628b447, 0x736e45e5, 0xcf859696, 0xffb317cb,
        0x8d0afd69, 0x537ae232, 0xae1578f9, 0xddce990c,
        0x24c1e97b, 0x704a63d8, 0xd2f7ff9e, 0xf2d0256a,
        0x7bbe37fa, 0xb6da1d89, 0xc3138e84, 0xc1ab7855,
        0x41c4e8ba, 0xe19f2857, 0xfd6ca77e, 0x42823eb3,
        0x82210e8e, 0x29b40c3f, 0xa734ea17, 0x5486df8a,
        0xb2679b40, 0x83ec2441, 0xc68ada6f, 0xd76d2cbb,
        0x772d1f3c, 0x01279ac2, 0x16d68caa, 0x46e03791,
        0xa1b66ef3, 0xa3a60366, 0x0a2ebfd1, 0x36f8cc3a,
        0xb02cf233, 0x8f6afa50, 0x2d22210d, 0x1774fc43,
        0x96dcdd3d, 0x72de6512, 0x8909a542, 0x05ebe3c0,
        0x6b62b1ad, 0xc27697b0, 0x341f0594, 0x3246b165,
        0xb2188663, 0xee5e2c6e, 0x4bda96fd, 0xd751bb27,
        0xa167f316, 0x1f93db5a, 0x2af1defd, 0xa922cfce,
        0x50bdb137, 0x857b9632, 0x38f37694, 0x40464674,
        0xe9bbb7e4, 0x1e7e5f41, 0x5a3323a3, 0xb2b64931,
        0x852d29a6, 0xb18da998, 0x02b21f49, 0x597121c5,
        0xfba1deab, 0xd3635b20, 0xbb9d5bb8, 0xb6fa7488,
        0xf7c0c147, 0xb21d9f97, 0x6244017b, 0x7aad0a65,
        0xdcad7cf1, 0x336fcc9e, 0x8062fa5c, 0xb778f971,
        0x4a58649b, 0xab990714, 0x5275c7b6, 0x9f51804b
}};

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_display_wide___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    VlWide<128>/*4095:0*/ t__DOT__crc;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_display_wide__cuda final {
    typedef Vt_display_wide___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 536,
        OFF___Vtrigprevexpr___TOP__clk__0 = 537,
        OFF___VactPhaseResult = 538,
        OFF___VnbaPhaseResult = 539,
        OFF_t__DOT__cyc = 16,
        OFF_t__DOT__crc = 20,
        OFF___VactIterCount = 532,
        OFF___VactTriggered = 0,
        OFF___VnbaTriggered = 8,
        kStateBytes = 544,  // S_persist
        kTouchedReadBytes = 540,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 539,  // S_state_wr
    };
    struct Locals final {
    };
    struct Shared final {
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        st.ldW(OFF_t__DOT__crc, i, s.t__DOT__crc);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.stW(OFF_t__DOT__crc, i, s.t__DOT__crc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 1, kOutWords = 0, kInBytes = 1, kOutBytes = 0, kTraceWords = 0 };

    VLCU_D static void Vt_display_wide___024root___eval_static(Vt_display_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_static__TOP
            vlSelfRef.t__DOT__cyc = 0U;
        }
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_display_wide___024root___eval_initial(Vt_display_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_display_wide___024root___eval_final(Vt_display_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_display_wide___024root___eval_settle(Vt_display_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_display_wide___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_display_wide___024root___nba_sequent__TOP__0(Vt_display_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        VlWide<128>/*4095:0*/ __Vdly__t__DOT__crc;
        VL_ZERO_W(4096, __Vdly__t__DOT__crc);
        VlWide<32>/*1023:0*/ __Vtemp_3;
        VlWide<32>/*1023:0*/ __Vtemp_4;
        VlWide<32>/*1023:0*/ __Vtemp_5;
        VlWide<32>/*1023:0*/ __Vtemp_6;
        VlWide<32>/*1023:0*/ __Vtemp_7;
        VlWide<32>/*1023:0*/ __Vtemp_8;
        VlWide<512>/*16383:0*/ __Vtemp_9;
        IData/*31:0*/ __Vilp1;
        IData/*31:0*/ __Vilp2;
        IData/*31:0*/ __Vdeeptemp_hcc1f4236__0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vilp1 = 0U;
        while ((__Vilp1 <= 0x0000007fU)) {
            __Vdly__t__DOT__crc[__Vilp1] = vlSelfRef.t__DOT__crc
                [__Vilp1];
            __Vilp1 = ((IData)(1U) + __Vilp1);
        }
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        __Vdly__t__DOT__crc[0U] = ((0xfffffffeU & __Vdly__t__DOT__crc[0U]) 
                                   | (1U & (((vlSelfRef.t__DOT__crc[1U] 
                                              >> 0x0000001fU) 
                                             ^ (vlSelfRef.t__DOT__crc[0U] 
                                                >> 2U)) 
                                            ^ vlSelfRef.t__DOT__crc[0U])));
        __Vdly__t__DOT__crc[0U] = ((1U & __Vdly__t__DOT__crc[0U]) 
                                   | (vlSelfRef.t__DOT__crc[0U] 
                                      << 1U));
        __Vdly__t__DOT__crc[1U] = ((vlSelfRef.t__DOT__crc[0U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.t__DOT__crc[1U] 
                                      << 1U));
        __Vdly__t__DOT__crc[2U] = ((vlSelfRef.t__DOT__crc[1U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.t__DOT__crc[2U] 
                                      << 1U));
        __Vdly__t__DOT__crc[3U] = ((vlSelfRef.t__DOT__crc[2U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.t__DOT__crc[3U] 
                                      << 1U));
        __Vdly__t__DOT__crc[4U] = ((vlSelfRef.t__DOT__crc[3U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.t__DOT__crc[4U] 
                                      << 1U));
        __Vdly__t__DOT__crc[5U] = ((vlSelfRef.t__DOT__crc[4U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.t__DOT__crc[5U] 
                                      << 1U));
        __Vdly__t__DOT__crc[6U] = ((vlSelfRef.t__DOT__crc[5U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.t__DOT__crc[6U] 
                                      << 1U));
        __Vdly__t__DOT__crc[7U] = ((vlSelfRef.t__DOT__crc[6U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.t__DOT__crc[7U] 
                                      << 1U));
        __Vdly__t__DOT__crc[8U] = ((vlSelfRef.t__DOT__crc[7U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.t__DOT__crc[8U] 
                                      << 1U));
        __Vdly__t__DOT__crc[9U] = ((vlSelfRef.t__DOT__crc[8U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.t__DOT__crc[9U] 
                                      << 1U));
        __Vdly__t__DOT__crc[10U] = ((vlSelfRef.t__DOT__crc[9U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[10U] 
                                       << 1U));
        __Vdly__t__DOT__crc[11U] = ((vlSelfRef.t__DOT__crc[10U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[11U] 
                                       << 1U));
        __Vdly__t__DOT__crc[12U] = ((vlSelfRef.t__DOT__crc[11U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[12U] 
                                       << 1U));
        __Vdly__t__DOT__crc[13U] = ((vlSelfRef.t__DOT__crc[12U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[13U] 
                                       << 1U));
        __Vdly__t__DOT__crc[14U] = ((vlSelfRef.t__DOT__crc[13U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[14U] 
                                       << 1U));
        __Vdly__t__DOT__crc[15U] = ((vlSelfRef.t__DOT__crc[14U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[15U] 
                                       << 1U));
        __Vdly__t__DOT__crc[16U] = ((vlSelfRef.t__DOT__crc[15U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[16U] 
                                       << 1U));
        __Vdly__t__DOT__crc[17U] = ((vlSelfRef.t__DOT__crc[16U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[17U] 
                                       << 1U));
        __Vdly__t__DOT__crc[18U] = ((vlSelfRef.t__DOT__crc[17U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[18U] 
                                       << 1U));
        __Vdly__t__DOT__crc[19U] = ((vlSelfRef.t__DOT__crc[18U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[19U] 
                                       << 1U));
        __Vdly__t__DOT__crc[20U] = ((vlSelfRef.t__DOT__crc[19U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[20U] 
                                       << 1U));
        __Vdly__t__DOT__crc[21U] = ((vlSelfRef.t__DOT__crc[20U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[21U] 
                                       << 1U));
        __Vdly__t__DOT__crc[22U] = ((vlSelfRef.t__DOT__crc[21U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[22U] 
                                       << 1U));
        __Vdly__t__DOT__crc[23U] = ((vlSelfRef.t__DOT__crc[22U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[23U] 
                                       << 1U));
        __Vdly__t__DOT__crc[24U] = ((vlSelfRef.t__DOT__crc[23U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[24U] 
                                       << 1U));
        __Vdly__t__DOT__crc[25U] = ((vlSelfRef.t__DOT__crc[24U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[25U] 
                                       << 1U));
        __Vdly__t__DOT__crc[26U] = ((vlSelfRef.t__DOT__crc[25U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[26U] 
                                       << 1U));
        __Vdly__t__DOT__crc[27U] = ((vlSelfRef.t__DOT__crc[26U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[27U] 
                                       << 1U));
        __Vdly__t__DOT__crc[28U] = ((vlSelfRef.t__DOT__crc[27U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[28U] 
                                       << 1U));
        __Vdly__t__DOT__crc[29U] = ((vlSelfRef.t__DOT__crc[28U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[29U] 
                                       << 1U));
        __Vdly__t__DOT__crc[30U] = ((vlSelfRef.t__DOT__crc[29U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[30U] 
                                       << 1U));
        __Vdly__t__DOT__crc[31U] = ((vlSelfRef.t__DOT__crc[30U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[31U] 
                                       << 1U));
        __Vdly__t__DOT__crc[32U] = ((vlSelfRef.t__DOT__crc[31U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[32U] 
                                       << 1U));
        __Vdly__t__DOT__crc[33U] = ((vlSelfRef.t__DOT__crc[32U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[33U] 
                                       << 1U));
        __Vdly__t__DOT__crc[34U] = ((vlSelfRef.t__DOT__crc[33U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[34U] 
                                       << 1U));
        __Vdly__t__DOT__crc[35U] = ((vlSelfRef.t__DOT__crc[34U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[35U] 
                                       << 1U));
        __Vdly__t__DOT__crc[36U] = ((vlSelfRef.t__DOT__crc[35U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[36U] 
                                       << 1U));
        __Vdly__t__DOT__crc[37U] = ((vlSelfRef.t__DOT__crc[36U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[37U] 
                                       << 1U));
        __Vdly__t__DOT__crc[38U] = ((vlSelfRef.t__DOT__crc[37U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[38U] 
                                       << 1U));
        __Vdly__t__DOT__crc[39U] = ((vlSelfRef.t__DOT__crc[38U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[39U] 
                                       << 1U));
        __Vdly__t__DOT__crc[40U] = ((vlSelfRef.t__DOT__crc[39U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[40U] 
                                       << 1U));
        __Vdly__t__DOT__crc[41U] = ((vlSelfRef.t__DOT__crc[40U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[41U] 
                                       << 1U));
        __Vdly__t__DOT__crc[42U] = ((vlSelfRef.t__DOT__crc[41U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[42U] 
                                       << 1U));
        __Vdly__t__DOT__crc[43U] = ((vlSelfRef.t__DOT__crc[42U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[43U] 
                                       << 1U));
        __Vdly__t__DOT__crc[44U] = ((vlSelfRef.t__DOT__crc[43U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[44U] 
                                       << 1U));
        __Vdly__t__DOT__crc[45U] = ((vlSelfRef.t__DOT__crc[44U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[45U] 
                                       << 1U));
        __Vdly__t__DOT__crc[46U] = ((vlSelfRef.t__DOT__crc[45U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[46U] 
                                       << 1U));
        __Vdly__t__DOT__crc[47U] = ((vlSelfRef.t__DOT__crc[46U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[47U] 
                                       << 1U));
        __Vdly__t__DOT__crc[48U] = ((vlSelfRef.t__DOT__crc[47U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[48U] 
                                       << 1U));
        __Vdly__t__DOT__crc[49U] = ((vlSelfRef.t__DOT__crc[48U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[49U] 
                                       << 1U));
        __Vdly__t__DOT__crc[50U] = ((vlSelfRef.t__DOT__crc[49U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[50U] 
                                       << 1U));
        __Vdly__t__DOT__crc[51U] = ((vlSelfRef.t__DOT__crc[50U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[51U] 
                                       << 1U));
        __Vdly__t__DOT__crc[52U] = ((vlSelfRef.t__DOT__crc[51U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[52U] 
                                       << 1U));
        __Vdly__t__DOT__crc[53U] = ((vlSelfRef.t__DOT__crc[52U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[53U] 
                                       << 1U));
        __Vdly__t__DOT__crc[54U] = ((vlSelfRef.t__DOT__crc[53U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[54U] 
                                       << 1U));
        __Vdly__t__DOT__crc[55U] = ((vlSelfRef.t__DOT__crc[54U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[55U] 
                                       << 1U));
        __Vdly__t__DOT__crc[56U] = ((vlSelfRef.t__DOT__crc[55U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[56U] 
                                       << 1U));
        __Vdly__t__DOT__crc[57U] = ((vlSelfRef.t__DOT__crc[56U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[57U] 
                                       << 1U));
        __Vdly__t__DOT__crc[58U] = ((vlSelfRef.t__DOT__crc[57U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[58U] 
                                       << 1U));
        __Vdly__t__DOT__crc[59U] = ((vlSelfRef.t__DOT__crc[58U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[59U] 
                                       << 1U));
        __Vdly__t__DOT__crc[60U] = ((vlSelfRef.t__DOT__crc[59U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[60U] 
                                       << 1U));
        __Vdly__t__DOT__crc[61U] = ((vlSelfRef.t__DOT__crc[60U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[61U] 
                                       << 1U));
        __Vdly__t__DOT__crc[62U] = ((vlSelfRef.t__DOT__crc[61U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[62U] 
                                       << 1U));
        __Vdly__t__DOT__crc[63U] = ((vlSelfRef.t__DOT__crc[62U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[63U] 
                                       << 1U));
        __Vdly__t__DOT__crc[64U] = ((vlSelfRef.t__DOT__crc[63U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[64U] 
                                       << 1U));
        __Vdly__t__DOT__crc[65U] = ((vlSelfRef.t__DOT__crc[64U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[65U] 
                                       << 1U));
        __Vdly__t__DOT__crc[66U] = ((vlSelfRef.t__DOT__crc[65U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[66U] 
                                       << 1U));
        __Vdly__t__DOT__crc[67U] = ((vlSelfRef.t__DOT__crc[66U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[67U] 
                                       << 1U));
        __Vdly__t__DOT__crc[68U] = ((vlSelfRef.t__DOT__crc[67U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[68U] 
                                       << 1U));
        __Vdly__t__DOT__crc[69U] = ((vlSelfRef.t__DOT__crc[68U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[69U] 
                                       << 1U));
        __Vdly__t__DOT__crc[70U] = ((vlSelfRef.t__DOT__crc[69U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[70U] 
                                       << 1U));
        __Vdly__t__DOT__crc[71U] = ((vlSelfRef.t__DOT__crc[70U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[71U] 
                                       << 1U));
        __Vdly__t__DOT__crc[72U] = ((vlSelfRef.t__DOT__crc[71U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[72U] 
                                       << 1U));
        __Vdly__t__DOT__crc[73U] = ((vlSelfRef.t__DOT__crc[72U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[73U] 
                                       << 1U));
        __Vdly__t__DOT__crc[74U] = ((vlSelfRef.t__DOT__crc[73U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[74U] 
                                       << 1U));
        __Vdly__t__DOT__crc[75U] = ((vlSelfRef.t__DOT__crc[74U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[75U] 
                                       << 1U));
        __Vdly__t__DOT__crc[76U] = ((vlSelfRef.t__DOT__crc[75U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[76U] 
                                       << 1U));
        __Vdly__t__DOT__crc[77U] = ((vlSelfRef.t__DOT__crc[76U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[77U] 
                                       << 1U));
        __Vdly__t__DOT__crc[78U] = ((vlSelfRef.t__DOT__crc[77U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[78U] 
                                       << 1U));
        __Vdly__t__DOT__crc[79U] = ((vlSelfRef.t__DOT__crc[78U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[79U] 
                                       << 1U));
        __Vdly__t__DOT__crc[80U] = ((vlSelfRef.t__DOT__crc[79U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[80U] 
                                       << 1U));
        __Vdly__t__DOT__crc[81U] = ((vlSelfRef.t__DOT__crc[80U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[81U] 
                                       << 1U));
        __Vdly__t__DOT__crc[82U] = ((vlSelfRef.t__DOT__crc[81U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[82U] 
                                       << 1U));
        __Vdly__t__DOT__crc[83U] = ((vlSelfRef.t__DOT__crc[82U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[83U] 
                                       << 1U));
        __Vdly__t__DOT__crc[84U] = ((vlSelfRef.t__DOT__crc[83U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[84U] 
                                       << 1U));
        __Vdly__t__DOT__crc[85U] = ((vlSelfRef.t__DOT__crc[84U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[85U] 
                                       << 1U));
        __Vdly__t__DOT__crc[86U] = ((vlSelfRef.t__DOT__crc[85U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[86U] 
                                       << 1U));
        __Vdly__t__DOT__crc[87U] = ((vlSelfRef.t__DOT__crc[86U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[87U] 
                                       << 1U));
        __Vdly__t__DOT__crc[88U] = ((vlSelfRef.t__DOT__crc[87U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[88U] 
                                       << 1U));
        __Vdly__t__DOT__crc[89U] = ((vlSelfRef.t__DOT__crc[88U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[89U] 
                                       << 1U));
        __Vdly__t__DOT__crc[90U] = ((vlSelfRef.t__DOT__crc[89U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[90U] 
                                       << 1U));
        __Vdly__t__DOT__crc[91U] = ((vlSelfRef.t__DOT__crc[90U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[91U] 
                                       << 1U));
        __Vdly__t__DOT__crc[92U] = ((vlSelfRef.t__DOT__crc[91U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[92U] 
                                       << 1U));
        __Vdly__t__DOT__crc[93U] = ((vlSelfRef.t__DOT__crc[92U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[93U] 
                                       << 1U));
        __Vdly__t__DOT__crc[94U] = ((vlSelfRef.t__DOT__crc[93U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[94U] 
                                       << 1U));
        __Vdly__t__DOT__crc[95U] = ((vlSelfRef.t__DOT__crc[94U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[95U] 
                                       << 1U));
        __Vdly__t__DOT__crc[96U] = ((vlSelfRef.t__DOT__crc[95U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[96U] 
                                       << 1U));
        __Vdly__t__DOT__crc[97U] = ((vlSelfRef.t__DOT__crc[96U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[97U] 
                                       << 1U));
        __Vdly__t__DOT__crc[98U] = ((vlSelfRef.t__DOT__crc[97U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[98U] 
                                       << 1U));
        __Vdly__t__DOT__crc[99U] = ((vlSelfRef.t__DOT__crc[98U] 
                                     >> 0x0000001fU) 
                                    | (vlSelfRef.t__DOT__crc[99U] 
                                       << 1U));
        __Vdly__t__DOT__crc[100U] = ((vlSelfRef.t__DOT__crc[99U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[100U] 
                                        << 1U));
        __Vdly__t__DOT__crc[101U] = ((vlSelfRef.t__DOT__crc[100U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[101U] 
                                        << 1U));
        __Vdly__t__DOT__crc[102U] = ((vlSelfRef.t__DOT__crc[101U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[102U] 
                                        << 1U));
        __Vdly__t__DOT__crc[103U] = ((vlSelfRef.t__DOT__crc[102U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[103U] 
                                        << 1U));
        __Vdly__t__DOT__crc[104U] = ((vlSelfRef.t__DOT__crc[103U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[104U] 
                                        << 1U));
        __Vdly__t__DOT__crc[105U] = ((vlSelfRef.t__DOT__crc[104U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[105U] 
                                        << 1U));
        __Vdly__t__DOT__crc[106U] = ((vlSelfRef.t__DOT__crc[105U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[106U] 
                                        << 1U));
        __Vdly__t__DOT__crc[107U] = ((vlSelfRef.t__DOT__crc[106U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[107U] 
                                        << 1U));
        __Vdly__t__DOT__crc[108U] = ((vlSelfRef.t__DOT__crc[107U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[108U] 
                                        << 1U));
        __Vdly__t__DOT__crc[109U] = ((vlSelfRef.t__DOT__crc[108U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[109U] 
                                        << 1U));
        __Vdly__t__DOT__crc[110U] = ((vlSelfRef.t__DOT__crc[109U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[110U] 
                                        << 1U));
        __Vdly__t__DOT__crc[111U] = ((vlSelfRef.t__DOT__crc[110U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[111U] 
                                        << 1U));
        __Vdly__t__DOT__crc[112U] = ((vlSelfRef.t__DOT__crc[111U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[112U] 
                                        << 1U));
        __Vdly__t__DOT__crc[113U] = ((vlSelfRef.t__DOT__crc[112U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[113U] 
                                        << 1U));
        __Vdly__t__DOT__crc[114U] = ((vlSelfRef.t__DOT__crc[113U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[114U] 
                                        << 1U));
        __Vdly__t__DOT__crc[115U] = ((vlSelfRef.t__DOT__crc[114U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[115U] 
                                        << 1U));
        __Vdly__t__DOT__crc[116U] = ((vlSelfRef.t__DOT__crc[115U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[116U] 
                                        << 1U));
        __Vdly__t__DOT__crc[117U] = ((vlSelfRef.t__DOT__crc[116U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[117U] 
                                        << 1U));
        __Vdly__t__DOT__crc[118U] = ((vlSelfRef.t__DOT__crc[117U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[118U] 
                                        << 1U));
        __Vdly__t__DOT__crc[119U] = ((vlSelfRef.t__DOT__crc[118U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[119U] 
                                        << 1U));
        __Vdly__t__DOT__crc[120U] = ((vlSelfRef.t__DOT__crc[119U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[120U] 
                                        << 1U));
        __Vdly__t__DOT__crc[121U] = ((vlSelfRef.t__DOT__crc[120U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[121U] 
                                        << 1U));
        __Vdly__t__DOT__crc[122U] = ((vlSelfRef.t__DOT__crc[121U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[122U] 
                                        << 1U));
        __Vdly__t__DOT__crc[123U] = ((vlSelfRef.t__DOT__crc[122U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[123U] 
                                        << 1U));
        __Vdly__t__DOT__crc[124U] = ((vlSelfRef.t__DOT__crc[123U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[124U] 
                                        << 1U));
        __Vdly__t__DOT__crc[125U] = ((vlSelfRef.t__DOT__crc[124U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[125U] 
                                        << 1U));
        __Vdly__t__DOT__crc[126U] = ((vlSelfRef.t__DOT__crc[125U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[126U] 
                                        << 1U));
        __Vdly__t__DOT__crc[127U] = ((vlSelfRef.t__DOT__crc[126U] 
                                      >> 0x0000001fU) 
                                     | (vlSelfRef.t__DOT__crc[127U] 
                                        << 1U));
        if ((0U == vlSelfRef.t__DOT__cyc)) {
            VL_ASSIGN_W(4096, __Vdly__t__DOT__crc, Vt_display_wide__ConstPool__CONST_hfdfd7487_0);
        } else if (VL_UNLIKELY(((0x00000063U == vlSelfRef.t__DOT__cyc)))) {
            __Vtemp_3[0U] = vlSelfRef.t__DOT__crc[96U];
            __Vtemp_3[1U] = vlSelfRef.t__DOT__crc[97U];
            __Vtemp_3[2U] = vlSelfRef.t__DOT__crc[98U];
            __Vtemp_3[3U] = vlSelfRef.t__DOT__crc[99U];
            __Vtemp_3[4U] = vlSelfRef.t__DOT__crc[100U];
            __Vtemp_3[5U] = vlSelfRef.t__DOT__crc[101U];
            __Vtemp_3[6U] = vlSelfRef.t__DOT__crc[102U];
            __Vtemp_3[7U] = vlSelfRef.t__DOT__crc[103U];
            __Vtemp_3[8U] = vlSelfRef.t__DOT__crc[104U];
            __Vtemp_3[9U] = vlSelfRef.t__DOT__crc[105U];
            __Vtemp_3[10U] = vlSelfRef.t__DOT__crc[106U];
            __Vtemp_3[11U] = vlSelfRef.t__DOT__crc[107U];
            __Vtemp_3[12U] = vlSelfRef.t__DOT__crc[108U];
            __Vtemp_3[13U] = vlSelfRef.t__DOT__crc[109U];
            __Vtemp_3[14U] = vlSelfRef.t__DOT__crc[110U];
            __Vtemp_3[15U] = vlSelfRef.t__DOT__crc[111U];
            __Vtemp_3[16U] = vlSelfRef.t__DOT__crc[112U];
            __Vtemp_3[17U] = vlSelfRef.t__DOT__crc[113U];
            __Vtemp_3[18U] = vlSelfRef.t__DOT__crc[114U];
            __Vtemp_3[19U] = vlSelfRef.t__DOT__crc[115U];
            __Vtemp_3[20U] = vlSelfRef.t__DOT__crc[116U];
            __Vtemp_3[21U] = vlSelfRef.t__DOT__crc[117U];
            __Vtemp_3[22U] = vlSelfRef.t__DOT__crc[118U];
            __Vtemp_3[23U] = vlSelfRef.t__DOT__crc[119U];
            __Vtemp_3[24U] = vlSelfRef.t__DOT__crc[120U];
            __Vtemp_3[25U] = vlSelfRef.t__DOT__crc[121U];
            __Vtemp_3[26U] = vlSelfRef.t__DOT__crc[122U];
            __Vtemp_3[27U] = vlSelfRef.t__DOT__crc[123U];
            __Vtemp_3[28U] = vlSelfRef.t__DOT__crc[124U];
            __Vtemp_3[29U] = vlSelfRef.t__DOT__crc[125U];
            __Vtemp_3[30U] = vlSelfRef.t__DOT__crc[126U];
            __Vtemp_3[31U] = vlSelfRef.t__DOT__crc[127U];
            __Vtemp_4[0U] = vlSelfRef.t__DOT__crc[32U];
            __Vtemp_4[1U] = vlSelfRef.t__DOT__crc[33U];
            __Vtemp_4[2U] = vlSelfRef.t__DOT__crc[34U];
            __Vtemp_4[3U] = vlSelfRef.t__DOT__crc[35U];
            __Vtemp_4[4U] = vlSelfRef.t__DOT__crc[36U];
            __Vtemp_4[5U] = vlSelfRef.t__DOT__crc[37U];
            __Vtemp_4[6U] = vlSelfRef.t__DOT__crc[38U];
            __Vtemp_4[7U] = vlSelfRef.t__DOT__crc[39U];
            __Vtemp_4[8U] = vlSelfRef.t__DOT__crc[40U];
            __Vtemp_4[9U] = vlSelfRef.t__DOT__crc[41U];
            __Vtemp_4[10U] = vlSelfRef.t__DOT__crc[42U];
            __Vtemp_4[11U] = vlSelfRef.t__DOT__crc[43U];
            __Vtemp_4[12U] = vlSelfRef.t__DOT__crc[44U];
            __Vtemp_4[13U] = vlSelfRef.t__DOT__crc[45U];
            __Vtemp_4[14U] = vlSelfRef.t__DOT__crc[46U];
            __Vtemp_4[15U] = vlSelfRef.t__DOT__crc[47U];
            __Vtemp_4[16U] = vlSelfRef.t__DOT__crc[48U];
            __Vtemp_4[17U] = vlSelfRef.t__DOT__crc[49U];
            __Vtemp_4[18U] = vlSelfRef.t__DOT__crc[50U];
            __Vtemp_4[19U] = vlSelfRef.t__DOT__crc[51U];
            __Vtemp_4[20U] = vlSelfRef.t__DOT__crc[52U];
            __Vtemp_4[21U] = vlSelfRef.t__DOT__crc[53U];
            __Vtemp_4[22U] = vlSelfRef.t__DOT__crc[54U];
            __Vtemp_4[23U] = vlSelfRef.t__DOT__crc[55U];
            __Vtemp_4[24U] = vlSelfRef.t__DOT__crc[56U];
            __Vtemp_4[25U] = vlSelfRef.t__DOT__crc[57U];
            __Vtemp_4[26U] = vlSelfRef.t__DOT__crc[58U];
            __Vtemp_4[27U] = vlSelfRef.t__DOT__crc[59U];
            __Vtemp_4[28U] = vlSelfRef.t__DOT__crc[60U];
            __Vtemp_4[29U] = vlSelfRef.t__DOT__crc[61U];
            __Vtemp_4[30U] = vlSelfRef.t__DOT__crc[62U];
            __Vtemp_4[31U] = vlSelfRef.t__DOT__crc[63U];
            __Vtemp_5[0U] = vlSelfRef.t__DOT__crc[0U];
            __Vtemp_5[1U] = vlSelfRef.t__DOT__crc[1U];
            __Vtemp_5[2U] = vlSelfRef.t__DOT__crc[2U];
            __Vtemp_5[3U] = vlSelfRef.t__DOT__crc[3U];
            __Vtemp_5[4U] = vlSelfRef.t__DOT__crc[4U];
            __Vtemp_5[5U] = vlSelfRef.t__DOT__crc[5U];
            __Vtemp_5[6U] = vlSelfRef.t__DOT__crc[6U];
            __Vtemp_5[7U] = vlSelfRef.t__DOT__crc[7U];
            __Vtemp_5[8U] = vlSelfRef.t__DOT__crc[8U];
            __Vtemp_5[9U] = vlSelfRef.t__DOT__crc[9U];
            __Vtemp_5[10U] = vlSelfRef.t__DOT__crc[10U];
            __Vtemp_5[11U] = vlSelfRef.t__DOT__crc[11U];
            __Vtemp_5[12U] = vlSelfRef.t__DOT__crc[12U];
            __Vtemp_5[13U] = vlSelfRef.t__DOT__crc[13U];
            __Vtemp_5[14U] = vlSelfRef.t__DOT__crc[14U];
            __Vtemp_5[15U] = vlSelfRef.t__DOT__crc[15U];
            __Vtemp_5[16U] = vlSelfRef.t__DOT__crc[16U];
            __Vtemp_5[17U] = vlSelfRef.t__DOT__crc[17U];
            __Vtemp_5[18U] = vlSelfRef.t__DOT__crc[18U];
            __Vtemp_5[19U] = vlSelfRef.t__DOT__crc[19U];
            __Vtemp_5[20U] = vlSelfRef.t__DOT__crc[20U];
            __Vtemp_5[21U] = vlSelfRef.t__DOT__crc[21U];
            __Vtemp_5[22U] = vlSelfRef.t__DOT__crc[22U];
            __Vtemp_5[23U] = vlSelfRef.t__DOT__crc[23U];
            __Vtemp_5[24U] = vlSelfRef.t__DOT__crc[24U];
            __Vtemp_5[25U] = vlSelfRef.t__DOT__crc[25U];
            __Vtemp_5[26U] = vlSelfRef.t__DOT__crc[26U];
            __Vtemp_5[27U] = vlSelfRef.t__DOT__crc[27U];
            __Vtemp_5[28U] = vlSelfRef.t__DOT__crc[28U];
            __Vtemp_5[29U] = vlSelfRef.t__DOT__crc[29U];
            __Vtemp_5[30U] = vlSelfRef.t__DOT__crc[30U];
            __Vtemp_5[31U] = vlSelfRef.t__DOT__crc[31U];
            vlCtx.writef(0U, 6U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__cyc), vlcu_msg_arg(__Vtemp_3), vlcu_msg_arg(
                                                                                (0x00ffffffU 
                                                                                & vlSelfRef.t__DOT__crc[64U])), vlcu_msg_arg(__Vtemp_4), vlcu_msg_arg(__Vtemp_5));
            __Vtemp_6[0U] = vlSelfRef.t__DOT__crc[96U];
            __Vtemp_6[1U] = vlSelfRef.t__DOT__crc[97U];
            __Vtemp_6[2U] = vlSelfRef.t__DOT__crc[98U];
            __Vtemp_6[3U] = vlSelfRef.t__DOT__crc[99U];
            __Vtemp_6[4U] = vlSelfRef.t__DOT__crc[100U];
            __Vtemp_6[5U] = vlSelfRef.t__DOT__crc[101U];
            __Vtemp_6[6U] = vlSelfRef.t__DOT__crc[102U];
            __Vtemp_6[7U] = vlSelfRef.t__DOT__crc[103U];
            __Vtemp_6[8U] = vlSelfRef.t__DOT__crc[104U];
            __Vtemp_6[9U] = vlSelfRef.t__DOT__crc[105U];
            __Vtemp_6[10U] = vlSelfRef.t__DOT__crc[106U];
            __Vtemp_6[11U] = vlSelfRef.t__DOT__crc[107U];
            __Vtemp_6[12U] = vlSelfRef.t__DOT__crc[108U];
            __Vtemp_6[13U] = vlSelfRef.t__DOT__crc[109U];
            __Vtemp_6[14U] = vlSelfRef.t__DOT__crc[110U];
            __Vtemp_6[15U] = vlSelfRef.t__DOT__crc[111U];
            __Vtemp_6[16U] = vlSelfRef.t__DOT__crc[112U];
            __Vtemp_6[17U] = vlSelfRef.t__DOT__crc[113U];
            __Vtemp_6[18U] = vlSelfRef.t__DOT__crc[114U];
            __Vtemp_6[19U] = vlSelfRef.t__DOT__crc[115U];
            __Vtemp_6[20U] = vlSelfRef.t__DOT__crc[116U];
            __Vtemp_6[21U] = vlSelfRef.t__DOT__crc[117U];
            __Vtemp_6[22U] = vlSelfRef.t__DOT__crc[118U];
            __Vtemp_6[23U] = vlSelfRef.t__DOT__crc[119U];
            __Vtemp_6[24U] = vlSelfRef.t__DOT__crc[120U];
            __Vtemp_6[25U] = vlSelfRef.t__DOT__crc[121U];
            __Vtemp_6[26U] = vlSelfRef.t__DOT__crc[122U];
            __Vtemp_6[27U] = vlSelfRef.t__DOT__crc[123U];
            __Vtemp_6[28U] = vlSelfRef.t__DOT__crc[124U];
            __Vtemp_6[29U] = vlSelfRef.t__DOT__crc[125U];
            __Vtemp_6[30U] = vlSelfRef.t__DOT__crc[126U];
            __Vtemp_6[31U] = vlSelfRef.t__DOT__crc[127U];
            __Vtemp_7[0U] = vlSelfRef.t__DOT__crc[32U];
            __Vtemp_7[1U] = vlSelfRef.t__DOT__crc[33U];
            __Vtemp_7[2U] = vlSelfRef.t__DOT__crc[34U];
            __Vtemp_7[3U] = vlSelfRef.t__DOT__crc[35U];
            __Vtemp_7[4U] = vlSelfRef.t__DOT__crc[36U];
            __Vtemp_7[5U] = vlSelfRef.t__DOT__crc[37U];
            __Vtemp_7[6U] = vlSelfRef.t__DOT__crc[38U];
            __Vtemp_7[7U] = vlSelfRef.t__DOT__crc[39U];
            __Vtemp_7[8U] = vlSelfRef.t__DOT__crc[40U];
            __Vtemp_7[9U] = vlSelfRef.t__DOT__crc[41U];
            __Vtemp_7[10U] = vlSelfRef.t__DOT__crc[42U];
            __Vtemp_7[11U] = vlSelfRef.t__DOT__crc[43U];
            __Vtemp_7[12U] = vlSelfRef.t__DOT__crc[44U];
            __Vtemp_7[13U] = vlSelfRef.t__DOT__crc[45U];
            __Vtemp_7[14U] = vlSelfRef.t__DOT__crc[46U];
            __Vtemp_7[15U] = vlSelfRef.t__DOT__crc[47U];
            __Vtemp_7[16U] = vlSelfRef.t__DOT__crc[48U];
            __Vtemp_7[17U] = vlSelfRef.t__DOT__crc[49U];
            __Vtemp_7[18U] = vlSelfRef.t__DOT__crc[50U];
            __Vtemp_7[19U] = vlSelfRef.t__DOT__crc[51U];
            __Vtemp_7[20U] = vlSelfRef.t__DOT__crc[52U];
            __Vtemp_7[21U] = vlSelfRef.t__DOT__crc[53U];
            __Vtemp_7[22U] = vlSelfRef.t__DOT__crc[54U];
            __Vtemp_7[23U] = vlSelfRef.t__DOT__crc[55U];
            __Vtemp_7[24U] = vlSelfRef.t__DOT__crc[56U];
            __Vtemp_7[25U] = vlSelfRef.t__DOT__crc[57U];
            __Vtemp_7[26U] = vlSelfRef.t__DOT__crc[58U];
            __Vtemp_7[27U] = vlSelfRef.t__DOT__crc[59U];
            __Vtemp_7[28U] = vlSelfRef.t__DOT__crc[60U];
            __Vtemp_7[29U] = vlSelfRef.t__DOT__crc[61U];
            __Vtemp_7[30U] = vlSelfRef.t__DOT__crc[62U];
            __Vtemp_7[31U] = vlSelfRef.t__DOT__crc[63U];
            __Vtemp_8[0U] = vlSelfRef.t__DOT__crc[0U];
            __Vtemp_8[1U] = vlSelfRef.t__DOT__crc[1U];
            __Vtemp_8[2U] = vlSelfRef.t__DOT__crc[2U];
            __Vtemp_8[3U] = vlSelfRef.t__DOT__crc[3U];
            __Vtemp_8[4U] = vlSelfRef.t__DOT__crc[4U];
            __Vtemp_8[5U] = vlSelfRef.t__DOT__crc[5U];
            __Vtemp_8[6U] = vlSelfRef.t__DOT__crc[6U];
            __Vtemp_8[7U] = vlSelfRef.t__DOT__crc[7U];
            __Vtemp_8[8U] = vlSelfRef.t__DOT__crc[8U];
            __Vtemp_8[9U] = vlSelfRef.t__DOT__crc[9U];
            __Vtemp_8[10U] = vlSelfRef.t__DOT__crc[10U];
            __Vtemp_8[11U] = vlSelfRef.t__DOT__crc[11U];
            __Vtemp_8[12U] = vlSelfRef.t__DOT__crc[12U];
            __Vtemp_8[13U] = vlSelfRef.t__DOT__crc[13U];
            __Vtemp_8[14U] = vlSelfRef.t__DOT__crc[14U];
            __Vtemp_8[15U] = vlSelfRef.t__DOT__crc[15U];
            __Vtemp_8[16U] = vlSelfRef.t__DOT__crc[16U];
            __Vtemp_8[17U] = vlSelfRef.t__DOT__crc[17U];
            __Vtemp_8[18U] = vlSelfRef.t__DOT__crc[18U];
            __Vtemp_8[19U] = vlSelfRef.t__DOT__crc[19U];
            __Vtemp_8[20U] = vlSelfRef.t__DOT__crc[20U];
            __Vtemp_8[21U] = vlSelfRef.t__DOT__crc[21U];
            __Vtemp_8[22U] = vlSelfRef.t__DOT__crc[22U];
            __Vtemp_8[23U] = vlSelfRef.t__DOT__crc[23U];
            __Vtemp_8[24U] = vlSelfRef.t__DOT__crc[24U];
            __Vtemp_8[25U] = vlSelfRef.t__DOT__crc[25U];
            __Vtemp_8[26U] = vlSelfRef.t__DOT__crc[26U];
            __Vtemp_8[27U] = vlSelfRef.t__DOT__crc[27U];
            __Vtemp_8[28U] = vlSelfRef.t__DOT__crc[28U];
            __Vtemp_8[29U] = vlSelfRef.t__DOT__crc[29U];
            __Vtemp_8[30U] = vlSelfRef.t__DOT__crc[30U];
            __Vtemp_8[31U] = vlSelfRef.t__DOT__crc[31U];
            vlCtx.writef(1U, 6U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__cyc), vlcu_msg_arg(__Vtemp_6), vlcu_msg_arg(
                                                                                (0x00ffffffU 
                                                                                & vlSelfRef.t__DOT__crc[64U])), vlcu_msg_arg(__Vtemp_7), vlcu_msg_arg(__Vtemp_8));
            VL_REPLICATE_WWI(4096, __Vtemp_9, vlSelfRef.t__DOT__crc, 4U);
            vlCtx.writef(2U, 3U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__cyc), vlcu_msg_arg(__Vtemp_9));
            __Vdeeptemp_hcc1f4236__0 = ((((((((((((
                                                   (((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((((0x33f564fcU 
                                                                                ^ vlSelfRef.t__DOT__crc[0U]) 
                                                                                | (0x85a977a0U 
                                                                                ^ vlSelfRef.t__DOT__crc[1U])) 
                                                                                | (0x2c717a1dU 
                                                                                ^ vlSelfRef.t__DOT__crc[2U])) 
                                                                                | (0xf7c8e284U 
                                                                                ^ vlSelfRef.t__DOT__crc[3U])) 
                                                                                | (0xbfc7e850U 
                                                                                ^ vlSelfRef.t__DOT__crc[4U])) 
                                                                                | (0xa4dd3ba6U 
                                                                                ^ vlSelfRef.t__DOT__crc[5U])) 
                                                                                | (0x2365e6bcU 
                                                                                ^ vlSelfRef.t__DOT__crc[6U])) 
                                                                                | (0x48ebdbfcU 
                                                                                ^ vlSelfRef.t__DOT__crc[7U])) 
                                                                                | (0xb97a8c4fU 
                                                                                ^ vlSelfRef.t__DOT__crc[8U])) 
                                                                                | (0x015a7ff9U 
                                                                                ^ vlSelfRef.t__DOT__crc[9U])) 
                                                                                | (0x5b3d5a4eU 
                                                                                ^ vlSelfRef.t__DOT__crc[10U])) 
                                                                                | (0x77d99223U 
                                                                                ^ vlSelfRef.t__DOT__crc[11U])) 
                                                                                | (0xf8d0e3baU 
                                                                                ^ vlSelfRef.t__DOT__crc[12U])) 
                                                                                | (0xc66575f7U 
                                                                                ^ vlSelfRef.t__DOT__crc[13U])) 
                                                                                | (0xf9ab7ea8U 
                                                                                ^ vlSelfRef.t__DOT__crc[14U])) 
                                                                                | (0x42a07718U 
                                                                                ^ vlSelfRef.t__DOT__crc[15U])) 
                                                                                | (0x3298a586U 
                                                                                ^ vlSelfRef.t__DOT__crc[16U])) 
                                                                                | (0x514b2069U 
                                                                                ^ vlSelfRef.t__DOT__crc[17U])) 
                                                                                | (0x5b12d0aeU 
                                                                                ^ vlSelfRef.t__DOT__crc[18U])) 
                                                                                | (0xd3f99d59U 
                                                                                ^ vlSelfRef.t__DOT__crc[19U])) 
                                                                                | (0xe126443aU 
                                                                                ^ vlSelfRef.t__DOT__crc[20U])) 
                                                                                | (0xaad8eadeU 
                                                                                ^ vlSelfRef.t__DOT__crc[21U])) 
                                                                                | (0x3fbe276fU 
                                                                                ^ vlSelfRef.t__DOT__crc[22U])) 
                                                                                | (0x60a6df9eU 
                                                                                ^ vlSelfRef.t__DOT__crc[23U])) 
                                                                                | (0xc8b328b7U 
                                                                                ^ vlSelfRef.t__DOT__crc[24U])) 
                                                                                | (0x1d5507f5U 
                                                                                ^ vlSelfRef.t__DOT__crc[25U])) 
                                                                                | (0x69be0aa1U 
                                                                                ^ vlSelfRef.t__DOT__crc[26U])) 
                                                                                | (0xa3bf0632U 
                                                                                ^ vlSelfRef.t__DOT__crc[27U])) 
                                                                                | (0x9df544a3U 
                                                                                ^ vlSelfRef.t__DOT__crc[28U])) 
                                                                                | (0xfb884815U 
                                                                                ^ vlSelfRef.t__DOT__crc[29U])) 
                                                                                | (0x03d8c045U 
                                                                                ^ vlSelfRef.t__DOT__crc[30U])) 
                                                                                | (0xb72db90bU 
                                                                                ^ vlSelfRef.t__DOT__crc[31U])) 
                                                                                | (0x968b770aU 
                                                                                ^ vlSelfRef.t__DOT__crc[32U])) 
                                                                                | (0x25c5250aU 
                                                                                ^ vlSelfRef.t__DOT__crc[33U])) 
                                                                                | (0x71233c70U 
                                                                                ^ vlSelfRef.t__DOT__crc[34U])) 
                                                                                | (0x844596dcU 
                                                                                ^ vlSelfRef.t__DOT__crc[35U])) 
                                                                                | (0x6faa8d23U 
                                                                                ^ vlSelfRef.t__DOT__crc[36U])) 
                                                                                | (0x9cf7bf88U 
                                                                                ^ vlSelfRef.t__DOT__crc[37U])) 
                                                                                | (0x6340a258U 
                                                                                ^ vlSelfRef.t__DOT__crc[38U])) 
                                                                                | (0x31cca6e2U 
                                                                                ^ vlSelfRef.t__DOT__crc[39U])) 
                                                                                | (0xaf0b82aaU 
                                                                                ^ vlSelfRef.t__DOT__crc[40U])) 
                                                                                | (0xc684d6a3U 
                                                                                ^ vlSelfRef.t__DOT__crc[41U])) 
                                                                                | (0x6ba7876eU 
                                                                                ^ vlSelfRef.t__DOT__crc[42U])) 
                                                                                | (0x35a8872aU 
                                                                                ^ vlSelfRef.t__DOT__crc[43U])) 
                                                                                | (0xe41fcb73U 
                                                                                ^ vlSelfRef.t__DOT__crc[44U])) 
                                                                                | (0x69f40b01U 
                                                                                ^ vlSelfRef.t__DOT__crc[45U])) 
                                                                                | (0x9d55f14dU 
                                                                                ^ vlSelfRef.t__DOT__crc[46U])) 
                                                                                | (0x18a2d11cU 
                                                                                ^ vlSelfRef.t__DOT__crc[47U])) 
                                                                                | (0xcdb91797U 
                                                                                ^ vlSelfRef.t__DOT__crc[48U])) 
                                                                                | (0x3e165a59U 
                                                                                ^ vlSelfRef.t__DOT__crc[49U])) 
                                                                                | (0xfecc5f2fU 
                                                                                ^ vlSelfRef.t__DOT__crc[50U])) 
                                                                                | (0x342bf5a7U 
                                                                                ^ vlSelfRef.t__DOT__crc[51U])) 
                                                                                | (0x4deb88caU 
                                                                                ^ vlSelfRef.t__DOT__crc[52U])) 
                                                                                | (0xb855e3e5U 
                                                                                ^ vlSelfRef.t__DOT__crc[53U])) 
                                                                                | (0x773a6432U 
                                                                                ^ vlSelfRef.t__DOT__crc[54U])) 
                                                                                | (0x9307a5efU 
                                                                                ^ vlSelfRef.t__DOT__crc[55U])) 
                                                                                | (0xc1298f60U 
                                                                                ^ vlSelfRef.t__DOT__crc[56U])) 
                                                                                | (0x4bdffe79U 
                                                                                ^ vlSelfRef.t__DOT__crc[57U])) 
                                                                                | (0xcb4095abU 
                                                                                ^ vlSelfRef.t__DOT__crc[58U])) 
                                                                                | (0xeef8dfebU 
                                                                                ^ vlSelfRef.t__DOT__crc[59U])) 
                                                                                | (0xdb687625U 
                                                                                ^ vlSelfRef.t__DOT__crc[60U])) 
                                                                                | (0x0c4e3a12U 
                                                                                ^ vlSelfRef.t__DOT__crc[61U])) 
                                                                                | (0x06ade157U 
                                                                                ^ vlSelfRef.t__DOT__crc[62U])) 
                                                                                | (0x0713a2ebU 
                                                                                ^ vlSelfRef.t__DOT__crc[63U])) 
                                                                                | (0x867ca15dU 
                                                                                ^ vlSelfRef.t__DOT__crc[64U])) 
                                                                                | (0xf5b29dfbU 
                                                                                ^ vlSelfRef.t__DOT__crc[65U])) 
                                                                                | (0x0a08facfU 
                                                                                ^ vlSelfRef.t__DOT__crc[66U])) 
                                                                                | (0x08843a39U 
                                                                                ^ vlSelfRef.t__DOT__crc[67U])) 
                                                                                | (0xa6d030feU 
                                                                                ^ vlSelfRef.t__DOT__crc[68U])) 
                                                                                | (0x9cd3a85cU 
                                                                                ^ vlSelfRef.t__DOT__crc[69U])) 
                                                                                | (0x521b7e2aU 
                                                                                ^ vlSelfRef.t__DOT__crc[70U])) 
                                                                                | (0xc99e6d01U 
                                                                                ^ vlSelfRef.t__DOT__crc[71U])) 
                                                                                | (0x0fb09106U 
                                                                                ^ vlSelfRef.t__DOT__crc[72U])) 
                                                                                | (0x1a2b69beU 
                                                                                ^ vlSelfRef.t__DOT__crc[73U])) 
                                                                                | (0x5db4b2efU 
                                                                                ^ vlSelfRef.t__DOT__crc[74U])) 
                                                                                | (0xdcb47cf3U 
                                                                                ^ vlSelfRef.t__DOT__crc[75U])) 
                                                                                | (0x049e6b09U 
                                                                                ^ vlSelfRef.t__DOT__crc[76U])) 
                                                                                | (0x5b5a32a8U 
                                                                                ^ vlSelfRef.t__DOT__crc[77U])) 
                                                                                | (0x1b80de44U 
                                                                                ^ vlSelfRef.t__DOT__crc[78U])) 
                                                                               | (0x86d9bbcdU 
                                                                                ^ vlSelfRef.t__DOT__crc[79U])) 
                                                                              | (0x8e980d9aU 
                                                                                ^ vlSelfRef.t__DOT__crc[80U])) 
                                                                             | (0x28baff46U 
                                                                                ^ vlSelfRef.t__DOT__crc[81U])) 
                                                                            | (0xdbe330e8U 
                                                                               ^ vlSelfRef.t__DOT__crc[82U])) 
                                                                           | (0xc0b3c8ccU 
                                                                              ^ vlSelfRef.t__DOT__crc[83U])) 
                                                                          | (0x3dabe942U 
                                                                             ^ vlSelfRef.t__DOT__crc[84U])) 
                                                                         | (0xb4888436U 
                                                                            ^ vlSelfRef.t__DOT__crc[85U])) 
                                                                        | (0x5dd3f10cU 
                                                                           ^ vlSelfRef.t__DOT__crc[86U])) 
                                                                       | (0x5b7374f4U 
                                                                          ^ vlSelfRef.t__DOT__crc[87U])) 
                                                                      | (0xcb79944aU 
                                                                         ^ vlSelfRef.t__DOT__crc[88U])) 
                                                                     | (0x24269509U 
                                                                        ^ vlSelfRef.t__DOT__crc[89U])) 
                                                                    | (0x17af8f02U 
                                                                       ^ vlSelfRef.t__DOT__crc[90U])) 
                                                                   | (0xad8ac6b4U 
                                                                      ^ vlSelfRef.t__DOT__crc[91U])) 
                                                                  | (0x09da5ec1U 
                                                                     ^ vlSelfRef.t__DOT__crc[92U])) 
                                                                 | (0xd07c1653U 
                                                                    ^ vlSelfRef.t__DOT__crc[93U])) 
                                                                | (0xc91ac594U 
                                                                   ^ vlSelfRef.t__DOT__crc[94U])) 
                                                               | (0xc862198cU 
                                                                  ^ vlSelfRef.t__DOT__crc[95U])) 
                                                              | (0xb978b1baU 
                                                                 ^ vlSelfRef.t__DOT__crc[96U])) 
                                                             | (0x2f6a5bf7U 
                                                                ^ vlSelfRef.t__DOT__crc[97U])) 
                                                            | (0x5d46ec9dU 
                                                               ^ vlSelfRef.t__DOT__crc[98U])) 
                                                           | (0x859fcc5bU 
                                                              ^ vlSelfRef.t__DOT__crc[99U])) 
                                                          | (0x7e4f6d6aU 
                                                             ^ vlSelfRef.t__DOT__crc[100U])) 
                                                         | (0xabc77bf4U 
                                                            ^ vlSelfRef.t__DOT__crc[101U])) 
                                                        | (0xa48b3f38U 
                                                           ^ vlSelfRef.t__DOT__crc[102U])) 
                                                       | (0x42f6c4deU 
                                                          ^ vlSelfRef.t__DOT__crc[103U])) 
                                                      | (0x15ee58c9U 
                                                         ^ vlSelfRef.t__DOT__crc[104U])) 
                                                     | (0xe3cdda52U 
                                                        ^ vlSelfRef.t__DOT__crc[105U])) 
                                                    | (0x011919d0U 
                                                       ^ vlSelfRef.t__DOT__crc[106U])) 
                                                   | (0xa6eedf91U 
                                                      ^ vlSelfRef.t__DOT__crc[107U])) 
                                                  | (0x79f97d07U 
                                                     ^ vlSelfRef.t__DOT__crc[108U])) 
                                                 | (0x68cc8e8cU 
                                                    ^ vlSelfRef.t__DOT__crc[109U])) 
                                                | (0xcad924c5U 
                                                   ^ vlSelfRef.t__DOT__crc[110U])) 
                                               | (0x14b4a69aU 
                                                  ^ vlSelfRef.t__DOT__crc[111U])) 
                                              | (0xc636a662U 
                                                 ^ vlSelfRef.t__DOT__crc[112U])) 
                                             | (0x0ac87d26U 
                                                ^ vlSelfRef.t__DOT__crc[113U])) 
                                            | (0x65c48714U 
                                               ^ vlSelfRef.t__DOT__crc[114U])) 
                                           | (0xee877aadU 
                                              ^ vlSelfRef.t__DOT__crc[115U])) 
                                          | (0x4d8d6c83U 
                                             ^ vlSelfRef.t__DOT__crc[116U])) 
                                         | (0xee756ee3U 
                                            ^ vlSelfRef.t__DOT__crc[117U])) 
                                        | (0xdbe9d222U 
                                           ^ vlSelfRef.t__DOT__crc[118U]));
            if (VL_UNLIKELY(((0U != (((((((((__Vdeeptemp_hcc1f4236__0 
                                             | (0xdf03051eU 
                                                ^ vlSelfRef.t__DOT__crc[119U])) 
                                            | (0xc8767e5fU 
                                               ^ vlSelfRef.t__DOT__crc[120U])) 
                                           | (0x891005eeU 
                                              ^ vlSelfRef.t__DOT__crc[121U])) 
                                          | (0xeab42995U 
                                             ^ vlSelfRef.t__DOT__crc[122U])) 
                                         | (0x72b5f3c5U 
                                            ^ vlSelfRef.t__DOT__crc[123U])) 
                                        | (0xcdbf327bU 
                                           ^ vlSelfRef.t__DOT__crc[124U])) 
                                       | (0x018be970U 
                                          ^ vlSelfRef.t__DOT__crc[125U])) 
                                      | (0xdde3e5c6U 
                                         ^ vlSelfRef.t__DOT__crc[126U])) 
                                     | (0x2961926eU 
                                        ^ vlSelfRef.t__DOT__crc[127U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_display_wide.v", 31, "");
            }
            vlCtx.writef(3U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_display_wide.v", 33, "");
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        __Vilp2 = 0U;
        while ((__Vilp2 <= 0x0000007fU)) {
            vlSelfRef.t__DOT__crc[__Vilp2] = __Vdly__t__DOT__crc
                [__Vilp2];
            __Vilp2 = ((IData)(1U) + __Vilp2);
        }
    }

    VLCU_D static void Vt_display_wide___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_display_wide___024root___eval_phase__act(Vt_display_wide___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_display_wide___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_display_wide___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_display_wide___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_display_wide___024root___eval_phase__nba(Vt_display_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_display_wide___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_display_wide___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_display_wide___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_display_wide___024root___eval(Vt_display_wide___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_display_wide___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_display_wide.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_display_wide___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_display_wide.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_display_wide___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_display_wide___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_display_wide___024root___ctor_var_reset(Vt_display_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        VL_SCOPED_RAND_RESET_W(4096, vlSelf->t__DOT__crc, __VscopeHash, 17680824324711171631ull);
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
        Vt_display_wide___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_display_wide___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_display_wide___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_display_wide___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_display_wide___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_display_wide___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_display_wide__cuda::info() {
    typedef Vt_display_wide__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 4, 0, 128, 4096, 4095, 0, T::OFF_t__DOT__crc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"[%0t] cyc==%0d crc=%x%x%x%x\n", "[%0t] cyc==%0d crc=%b%b%b%b\n", "[%0t] cyc==%0d crc=%d\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 4, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 9, 0, 0x2b2182f623f93839ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_display_wide__cuda VLCU_TESTBENCHES);
