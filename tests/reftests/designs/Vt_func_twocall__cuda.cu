// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_func_twocall"
#endif

// Constant pool
static __device__ const VlUnpacked<QData/*39:0*/, 256> Vt_func_twocall__ConstPool__TABLE_h3655e69b_0 = {{
    0x000000ffffffffffULL, 0x000000fdfbf7efdfULL,
    0x000000fbf7efdfbfULL, 0x000000f9f3e7cf9fULL,
    0x000000f7efdfbf7eULL, 0x000000f5ebd7af5eULL,
    0x000000f3e7cf9f3eULL, 0x000000f1e3c78f1eULL,
    0x000000efdfbf7efdULL, 0x000000eddbb76eddULL,
    0x000000ebd7af5ebdULL, 0x000000e9d3a74e9dULL,
    0x000000e7cf9f3e7cULL, 0x000000e5cb972e5cULL,
    0x000000e3c78f1e3cULL, 0x000000e1c3870e1cULL,
    0x000000dfbf7efdfbULL, 0x000000ddbb76eddbULL,
    0x000000dbb76eddbbULL, 0x000000d9b366cd9bULL,
    0x000000d7af5ebd7aULL, 0x000000d5ab56ad5aULL,
    0x000000d3a74e9d3aULL, 0x000000d1a3468d1aULL,
    0x000000cf9f3e7cf9ULL, 0x000000cd9b366cd9ULL,
    0x000000cb972e5cb9ULL, 0x000000c993264c99ULL,
    0x000000c78f1e3c78ULL, 0x000000c58b162c58ULL,
    0x000000c3870e1c38ULL, 0x000000c183060c18ULL,
    0x000000bf7efdfbf7ULL, 0x000000bd7af5ebd7ULL,
    0x000000bb76eddbb7ULL, 0x000000b972e5cb97ULL,
    0x000000b76eddbb76ULL, 0x000000b56ad5ab56ULL,
    0x000000b366cd9b36ULL, 0x000000b162c58b16ULL,
    0x000000af5ebd7af5ULL, 0x000000ad5ab56ad5ULL,
    0x000000ab56ad5ab5ULL, 0x000000a952a54a95ULL,
    0x000000a74e9d3a74ULL, 0x000000a54a952a54ULL,
    0x000000a3468d1a34ULL, 0x000000a142850a14ULL,
    0x0000009f3e7cf9f3ULL, 0x0000009d3a74e9d3ULL,
    0x0000009b366cd9b3ULL, 0x000000993264c993ULL,
    0x000000972e5cb972ULL, 0x000000952a54a952ULL,
    0x00000093264c9932ULL, 0x0000009122448912ULL,
    0x0000008f1e3c78f1ULL, 0x0000008d1a3468d1ULL,
    0x0000008b162c58b1ULL, 0x0000008912244891ULL,
    0x000000870e1c3870ULL, 0x000000850a142850ULL,
    0x00000083060c1830ULL, 0x0000008102040810ULL,
    0x0000007efdfbf7efULL, 0x0000007cf9f3e7cfULL,
    0x0000007af5ebd7afULL, 0x00000078f1e3c78fULL,
    0x00000076eddbb76eULL, 0x00000074e9d3a74eULL,
    0x00000072e5cb972eULL, 0x00000070e1c3870eULL,
    0x0000006eddbb76edULL, 0x0000006cd9b366cdULL,
    0x0000006ad5ab56adULL, 0x00000068d1a3468dULL,
    0x00000066cd9b366cULL, 0x00000064c993264cULL,
    0x00000062c58b162cULL, 0x00000060c183060cULL,
    0x0000005ebd7af5ebULL, 0x0000005cb972e5cbULL,
    0x0000005ab56ad5abULL, 0x00000058b162c58bULL,
    0x00000056ad5ab56aULL, 0x00000054a952a54aULL,
    0x00000052a54a952aULL, 0x00000050a142850aULL,
    0x0000004e9d3a74e9ULL, 0x0000004c993264c9ULL,
    0x0000004a952a54a9ULL, 0x0000004891224489ULL,
    0x000000468d1a3468ULL, 0x0000004489122448ULL,
    0x00000042850a1428ULL, 0x0000004081020408ULL,
    0x0000003e7cf9f3e7ULL, 0x0000003c78f1e3c7ULL,
    0x0000003a74e9d3a7ULL, 0x0000003870e1c387ULL,
    0x000000366cd9b366ULL, 0x0000003468d1a346ULL,
    0x0000003264c99326ULL, 0x0000003060c18306ULL,
    0x0000002e5cb972e5ULL, 0x0000002c58b162c5ULL,
    0x0000002a54a952a5ULL, 0x0000002850a14285ULL,
    0x000000264c993264ULL, 0x0000002448912244ULL,
    0x0000002244891224ULL, 0x0000002040810204ULL,
    0x0000001e3c78f1e3ULL, 0x0000001c3870e1c3ULL,
    0x0000001a3468d1a3ULL, 0x000000183060c183ULL,
    0x000000162c58b162ULL, 0x000000142850a142ULL,
    0x0000001224489122ULL, 0x0000001020408102ULL,
    0x0000000e1c3870e1ULL, 0x0000000c183060c1ULL,
    0x0000000a142850a1ULL, 0x0000000810204081ULL,
    0x000000060c183060ULL, 0x0000000408102040ULL,
    0x0000000204081020ULL, 0ULL,
    0x000000ffffffffffULL, 0x000000fdfbf7efdfULL,
    0x000000fbf7efdfbfULL, 0x000000f9f3e7cf9fULL,
    0x000000f7efdfbf7eULL, 0x000000f5ebd7af5eULL,
    0x000000f3e7cf9f3eULL, 0x000000f1e3c78f1eULL,
    0x000000efdfbf7efdULL, 0x000000eddbb76eddULL,
    0x000000ebd7af5ebdULL, 0x000000e9d3a74e9dULL,
    0x000000e7cf9f3e7cULL, 0x000000e5cb972e5cULL,
    0x000000e3c78f1e3cULL, 0x000000e1c3870e1cULL,
    0x000000dfbf7efdfbULL, 0x000000ddbb76eddbULL,
    0x000000dbb76eddbbULL, 0x000000d9b366cd9bULL,
    0x000000d7af5ebd7aULL, 0x000000d5ab56ad5aULL,
    0x000000d3a74e9d3aULL, 0x000000d1a3468d1aULL,
    0x000000cf9f3e7cf9ULL, 0x000000cd9b366cd9ULL,
    0x000000cb972e5cb9ULL, 0x000000c993264c99ULL,
    0x000000c78f1e3c78ULL, 0x000000c58b162c58ULL,
    0x000000c3870e1c38ULL, 0x000000c183060c18ULL,
    0x000000bf7efdfbf7ULL, 0x000000bd7af5ebd7ULL,
    0x000000bb76eddbb7ULL, 0x000000b972e5cb97ULL,
    0x000000b76eddbb76ULL, 0x000000b56ad5ab56ULL,
    0x000000b366cd9b36ULL, 0x000000b162c58b16ULL,
    0x000000af5ebd7af5ULL, 0x000000ad5ab56ad5ULL,
    0x000000ab56ad5ab5ULL, 0x000000a952a54a95ULL,
    0x000000a74e9d3a74ULL, 0x000000a54a952a54ULL,
    0x000000a3468d1a34ULL, 0x000000a142850a14ULL,
    0x0000009f3e7cf9f3ULL, 0x0000009d3a74e9d3ULL,
    0x0000009b366cd9b3ULL, 0x000000993264c993ULL,
    0x000000972e5cb972ULL, 0x000000952a54a952ULL,
    0x00000093264c9932ULL, 0x0000009122448912ULL,
    0x0000008f1e3c78f1ULL, 0x0000008d1a3468d1ULL,
    0x0000008b162c58b1ULL, 0x0000008912244891ULL,
    0x000000870e1c3870ULL, 0x000000850a142850ULL,
    0x00000083060c1830ULL, 0x0000008102040810ULL,
    0x0000007efdfbf7efULL, 0x0000007cf9f3e7cfULL,
    0x0000007af5ebd7afULL, 0x00000078f1e3c78fULL,
    0x00000076eddbb76eULL, 0x00000074e9d3a74eULL,
    0x00000072e5cb972eULL, 0x00000070e1c3870eULL,
    0x0000006eddbb76edULL, 0x0000006cd9b366cdULL,
    0x0000006ad5ab56adULL, 0x00000068d1a3468dULL,
    0x00000066cd9b366cULL, 0x00000064c993264cULL,
    0x00000062c58b162cULL, 0x00000060c183060cULL,
    0x0000005ebd7af5ebULL, 0x0000005cb972e5cbULL,
    0x0000005ab56ad5abULL, 0x00000058b162c58bULL,
    0x00000056ad5ab56aULL, 0x00000054a952a54aULL,
    0x00000052a54a952aULL, 0x00000050a142850aULL,
    0x0000004e9d3a74e9ULL, 0x0000004c993264c9ULL,
    0x0000004a952a54a9ULL, 0x0000004891224489ULL,
    0x000000468d1a3468ULL, 0x0000004489122448ULL,
    0x00000042850a1428ULL, 0x0000004081020408ULL,
    0x0000003e7cf9f3e7ULL, 0x0000003c78f1e3c7ULL,
    0x0000003a74e9d3a7ULL, 0x0000003870e1c387ULL,
    0x000000366cd9b366ULL, 0x0000003468d1a346ULL,
    0x0000003264c99326ULL, 0x0000003060c18306ULL,
    0x0000002e5cb972e5ULL, 0x0000002c58b162c5ULL,
    0x0000002a54a952a5ULL, 0x0000002850a14285ULL,
    0x000000264c993264ULL, 0x0000002448912244ULL,
    0x0000002244891224ULL, 0x0000002040810204ULL,
    0x0000001e3c78f1e3ULL, 0x0000001c3870e1c3ULL,
    0x0000001a3468d1a3ULL, 0x000000183060c183ULL,
    0x000000162c58b162ULL, 0x000000142850a142ULL,
    0x0000001224489122ULL, 0x0000001020408102ULL,
    0x0000000e1c3870e1ULL, 0x0000000c183060c1ULL,
    0x0000000a142850a1ULL, 0x0000000810204081ULL,
    0x000000060c183060ULL, 0x0000000408102040ULL,
    0x0000000204081020ULL, 0ULL
}};

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_func_twocall___024root final {
    CData/*0:0*/ clk;
    CData/*7:0*/ t__DOT__crc;
    CData/*0:0*/ t__DOT__sel;
    CData/*0:0*/ t__DOT__func__Vstatic__carry;
    CData/*2:0*/ __Vfunc_t__DOT__func__0__inh;
    CData/*2:0*/ __Vfunc_t__DOT__func__1__inh;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    SData/*14:0*/ t__DOT__func__Vstatic__func_mid;
    IData/*31:0*/ t__DOT__cyc;
    IData/*31:0*/ __VactIterCount;
    QData/*33:0*/ t__DOT____VlemCall_0__func;
    QData/*33:0*/ t__DOT__q;
    QData/*39:0*/ __Vfunc_t__DOT__func__0__inl;
    QData/*39:0*/ __Vfunc_t__DOT__func__1__inl;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_func_twocall__cuda final {
    typedef Vt_func_twocall___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 66,
        OFF_t__DOT__crc = 67,
        OFF_t__DOT__sel = 68,
        OFF_t__DOT__func__Vstatic__carry = 69,
        OFF___Vfunc_t__DOT__func__0__inh = 70,
        OFF___Vfunc_t__DOT__func__1__inh = 71,
        OFF___VstlFirstIteration = 72,
        OFF___VstlPhaseResult = 73,
        OFF___Vtrigprevexpr___TOP__clk__0 = 74,
        OFF___VactPhaseResult = 75,
        OFF___VnbaPhaseResult = 76,
        OFF_t__DOT__func__Vstatic__func_mid = 64,
        OFF_t__DOT__cyc = 56,
        OFF___VactIterCount = 60,
        OFF_t__DOT____VlemCall_0__func = 0,
        OFF_t__DOT__q = 8,
        OFF___Vfunc_t__DOT__func__0__inl = 16,
        OFF___Vfunc_t__DOT__func__1__inl = 24,
        OFF___VstlTriggered = 32,
        OFF___VactTriggered = 40,
        OFF___VnbaTriggered = 48,
        kStateBytes = 80,  // S_persist
        kTouchedReadBytes = 67,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 66,  // S_state_wr
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
        s.t__DOT__crc = st.ld<VlcuMem8>(OFF_t__DOT__crc, i);
        s.t__DOT__sel = st.ld<VlcuMem8>(OFF_t__DOT__sel, i);
        s.t__DOT__func__Vstatic__carry = st.ld<VlcuMem8>(OFF_t__DOT__func__Vstatic__carry, i);
        s.__Vfunc_t__DOT__func__0__inh = st.ld<VlcuMem8>(OFF___Vfunc_t__DOT__func__0__inh, i);
        s.__Vfunc_t__DOT__func__1__inh = st.ld<VlcuMem8>(OFF___Vfunc_t__DOT__func__1__inh, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__func__Vstatic__func_mid = st.ld<VlcuMem16>(OFF_t__DOT__func__Vstatic__func_mid, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT____VlemCall_0__func = st.ld<QData>(OFF_t__DOT____VlemCall_0__func, i);
        s.t__DOT__q = st.ld<QData>(OFF_t__DOT__q, i);
        s.__Vfunc_t__DOT__func__0__inl = st.ld<QData>(OFF___Vfunc_t__DOT__func__0__inl, i);
        s.__Vfunc_t__DOT__func__1__inl = st.ld<QData>(OFF___Vfunc_t__DOT__func__1__inl, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__crc, i, s.t__DOT__crc);
        st.st<VlcuMem8>(OFF_t__DOT__sel, i, s.t__DOT__sel);
        st.st<VlcuMem8>(OFF_t__DOT__func__Vstatic__carry, i, s.t__DOT__func__Vstatic__carry);
        st.st<VlcuMem8>(OFF___Vfunc_t__DOT__func__0__inh, i, s.__Vfunc_t__DOT__func__0__inh);
        st.st<VlcuMem8>(OFF___Vfunc_t__DOT__func__1__inh, i, s.__Vfunc_t__DOT__func__1__inh);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<VlcuMem16>(OFF_t__DOT__func__Vstatic__func_mid, i, s.t__DOT__func__Vstatic__func_mid);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT____VlemCall_0__func, i, s.t__DOT____VlemCall_0__func);
        st.st<QData>(OFF_t__DOT__q, i, s.t__DOT__q);
        st.st<QData>(OFF___Vfunc_t__DOT__func__0__inl, i, s.__Vfunc_t__DOT__func__0__inl);
        st.st<QData>(OFF___Vfunc_t__DOT__func__1__inl, i, s.__Vfunc_t__DOT__func__1__inl);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
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

    VLCU_D static void Vt_func_twocall___024root___eval_static(Vt_func_twocall___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_static__TOP
            const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
            vlSelfRef.t__DOT__func__Vstatic__func_mid = VL_SCOPED_RAND_RESET_I(15, __VscopeHash, 2156958800579108869ull);
            vlSelfRef.t__DOT__func__Vstatic__carry = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 12175148556233867061ull);
        }
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_func_twocall___024root___eval_initial(Vt_func_twocall___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlSelfRef.t__DOT__cyc = 1U;
        }
    }

    VLCU_D static void Vt_func_twocall___024root___eval_final(Vt_func_twocall___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_func_twocall___024root___eval_settle(Vt_func_twocall___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_func_twocall___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_func_twocall.v", 7, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_func_twocall___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_func_twocall___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_func_twocall___024root___eval_phase__stl(Vt_func_twocall___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_func_twocall___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_func_twocall___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _stl_sequent__TOP__0
                        if (vlSelfRef.t__DOT__sel) {
                            vlSelfRef.__Vfunc_t__DOT__func__0__inl 
                                = (0x000000ffffffffffULL 
                                   & (((QData)((IData)(vlSelfRef.t__DOT__crc)) 
                                       << 0x00000020U) 
                                      | (((QData)((IData)(vlSelfRef.t__DOT__crc)) 
                                          << 0x00000018U) 
                                         | (((QData)((IData)(vlSelfRef.t__DOT__crc)) 
                                             << 0x00000010U) 
                                            | (((QData)((IData)(vlSelfRef.t__DOT__crc)) 
                                                << 8U) 
                                               | (QData)((IData)(vlSelfRef.t__DOT__crc)))))));
                            vlSelfRef.__Vfunc_t__DOT__func__0__inh 
                                = (7U & ((IData)(vlSelfRef.t__DOT__crc) 
                                         >> 3U));
                            vlSelfRef.t__DOT__func__Vstatic__carry 
                                = (0x03ffffffU == (0x03ffffffU 
                                                   & (IData)(vlSelfRef.__Vfunc_t__DOT__func__0__inl)));
                            vlSelfRef.t__DOT__func__Vstatic__func_mid 
                                = (0x00007fffU & ((0x00003fffU 
                                                   & (IData)(
                                                             (vlSelfRef.__Vfunc_t__DOT__func__0__inl 
                                                              >> 0x1aU))) 
                                                  + (IData)(vlSelfRef.t__DOT__func__Vstatic__carry)));
                            vlSelfRef.t__DOT____VlemCall_0__func 
                                = ((0x000000007fffffffULL 
                                    & vlSelfRef.t__DOT____VlemCall_0__func) 
                                   | ((QData)((IData)(
                                                      (7U 
                                                       & ((IData)(vlSelfRef.__Vfunc_t__DOT__func__0__inh) 
                                                          + 
                                                          (1U 
                                                           & ((IData)(vlSelfRef.t__DOT__func__Vstatic__func_mid) 
                                                              >> 0x0eU)))))) 
                                      << 0x0000001fU));
                        } else {
                            vlSelfRef.__Vfunc_t__DOT__func__1__inl 
                                = Vt_func_twocall__ConstPool__TABLE_h3655e69b_0
                                [vlSelfRef.t__DOT__crc];
                            vlSelfRef.__Vfunc_t__DOT__func__1__inh 
                                = (7U & (~ ((IData)(vlSelfRef.t__DOT__crc) 
                                            >> 2U)));
                            vlSelfRef.t__DOT__func__Vstatic__carry 
                                = (0x03ffffffU == (0x03ffffffU 
                                                   & (IData)(vlSelfRef.__Vfunc_t__DOT__func__1__inl)));
                            vlSelfRef.t__DOT__func__Vstatic__func_mid 
                                = (0x00007fffU & ((0x00003fffU 
                                                   & (IData)(
                                                             (vlSelfRef.__Vfunc_t__DOT__func__1__inl 
                                                              >> 0x1aU))) 
                                                  + (IData)(vlSelfRef.t__DOT__func__Vstatic__carry)));
                            vlSelfRef.t__DOT____VlemCall_0__func 
                                = ((0x000000007fffffffULL 
                                    & vlSelfRef.t__DOT____VlemCall_0__func) 
                                   | ((QData)((IData)(
                                                      (7U 
                                                       & ((IData)(vlSelfRef.__Vfunc_t__DOT__func__1__inh) 
                                                          + 
                                                          (1U 
                                                           & ((IData)(vlSelfRef.t__DOT__func__Vstatic__func_mid) 
                                                              >> 0x0eU)))))) 
                                      << 0x0000001fU));
                        }
                        vlSelfRef.t__DOT____VlemCall_0__func 
                            = ((0x0000000380000000ULL 
                                & vlSelfRef.t__DOT____VlemCall_0__func) 
                               | (IData)((IData)(((0x7fffc000U 
                                                   & ((- (IData)(
                                                                 (1U 
                                                                  & ((IData)(vlSelfRef.t__DOT__func__Vstatic__func_mid) 
                                                                     >> 0x0dU)))) 
                                                      << 0x0000000eU)) 
                                                  | (0x00003fffU 
                                                     & (IData)(vlSelfRef.t__DOT__func__Vstatic__func_mid))))));
                        vlSelfRef.t__DOT__q = vlSelfRef.t__DOT____VlemCall_0__func;
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_func_twocall___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_func_twocall___024root___nba_sequent__TOP__0(Vt_func_twocall___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        if ((0U != vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
            vlSelfRef.t__DOT__sel = (1U & (~ (IData)(vlSelfRef.t__DOT__sel)));
            vlSelfRef.t__DOT__crc = ((0x000000feU & 
                                      ((IData)(vlSelfRef.t__DOT__crc) 
                                       << 1U)) | (1U 
                                                  & (~ 
                                                     (1U 
                                                      & VL_REDXOR_8(
                                                                    (0xb8U 
                                                                     & (IData)(vlSelfRef.t__DOT__crc)))))));
            if ((1U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__sel = 1U;
                vlSelfRef.t__DOT__crc = 0x12U;
            }
            if ((2U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0x0000000100000484ULL 
                                  != vlSelfRef.t__DOT__q)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_func_twocall.v", 47, "");
                }
            }
            if ((3U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0x000000037fffeddbULL 
                                  != vlSelfRef.t__DOT__q)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_func_twocall.v", 48, "");
                }
            }
            if ((4U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0x0000000080001212ULL 
                                  != vlSelfRef.t__DOT__q)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_func_twocall.v", 49, "");
                }
            }
            if ((5U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0x00000001fffff7efULL 
                                  != vlSelfRef.t__DOT__q)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_func_twocall.v", 50, "");
                }
            }
            if ((6U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0x0000000200000848ULL 
                                  != vlSelfRef.t__DOT__q)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_func_twocall.v", 51, "");
                }
            }
            if ((7U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0x0000000380001ebdULL 
                                  != vlSelfRef.t__DOT__q)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_func_twocall.v", 52, "");
                }
            }
            if ((8U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0x000000007fffe161ULL 
                                  != vlSelfRef.t__DOT__q)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_func_twocall.v", 53, "");
                }
            }
            if (VL_UNLIKELY(((9U == vlSelfRef.t__DOT__cyc)))) {
                vlCtx.writef(0U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_func_twocall.v", 56, "");
            }
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        if (vlSelfRef.t__DOT__sel) {
            vlSelfRef.__Vfunc_t__DOT__func__0__inl 
                = (0x000000ffffffffffULL & (((QData)((IData)(vlSelfRef.t__DOT__crc)) 
                                             << 0x00000020U) 
                                            | (((QData)((IData)(vlSelfRef.t__DOT__crc)) 
                                                << 0x00000018U) 
                                               | (((QData)((IData)(vlSelfRef.t__DOT__crc)) 
                                                   << 0x00000010U) 
                                                  | (((QData)((IData)(vlSelfRef.t__DOT__crc)) 
                                                      << 8U) 
                                                     | (QData)((IData)(vlSelfRef.t__DOT__crc)))))));
            vlSelfRef.__Vfunc_t__DOT__func__0__inh 
                = (7U & ((IData)(vlSelfRef.t__DOT__crc) 
                         >> 3U));
            vlSelfRef.t__DOT__func__Vstatic__carry 
                = (0x03ffffffU == (0x03ffffffU & (IData)(vlSelfRef.__Vfunc_t__DOT__func__0__inl)));
            vlSelfRef.t__DOT__func__Vstatic__func_mid 
                = (0x00007fffU & ((0x00003fffU & (IData)(
                                                         (vlSelfRef.__Vfunc_t__DOT__func__0__inl 
                                                          >> 0x1aU))) 
                                  + (IData)(vlSelfRef.t__DOT__func__Vstatic__carry)));
            vlSelfRef.t__DOT____VlemCall_0__func = 
                ((0x000000007fffffffULL & vlSelfRef.t__DOT____VlemCall_0__func) 
                 | ((QData)((IData)((7U & ((IData)(vlSelfRef.__Vfunc_t__DOT__func__0__inh) 
                                           + (1U & 
                                              ((IData)(vlSelfRef.t__DOT__func__Vstatic__func_mid) 
                                               >> 0x0eU)))))) 
                    << 0x0000001fU));
        } else {
            vlSelfRef.__Vfunc_t__DOT__func__1__inl 
                = Vt_func_twocall__ConstPool__TABLE_h3655e69b_0
                [vlSelfRef.t__DOT__crc];
            vlSelfRef.__Vfunc_t__DOT__func__1__inh 
                = (7U & (~ ((IData)(vlSelfRef.t__DOT__crc) 
                            >> 2U)));
            vlSelfRef.t__DOT__func__Vstatic__carry 
                = (0x03ffffffU == (0x03ffffffU & (IData)(vlSelfRef.__Vfunc_t__DOT__func__1__inl)));
            vlSelfRef.t__DOT__func__Vstatic__func_mid 
                = (0x00007fffU & ((0x00003fffU & (IData)(
                                                         (vlSelfRef.__Vfunc_t__DOT__func__1__inl 
                                                          >> 0x1aU))) 
                                  + (IData)(vlSelfRef.t__DOT__func__Vstatic__carry)));
            vlSelfRef.t__DOT____VlemCall_0__func = 
                ((0x000000007fffffffULL & vlSelfRef.t__DOT____VlemCall_0__func) 
                 | ((QData)((IData)((7U & ((IData)(vlSelfRef.__Vfunc_t__DOT__func__1__inh) 
                                           + (1U & 
                                              ((IData)(vlSelfRef.t__DOT__func__Vstatic__func_mid) 
                                               >> 0x0eU)))))) 
                    << 0x0000001fU));
        }
        vlSelfRef.t__DOT____VlemCall_0__func = ((0x0000000380000000ULL 
                                                 & vlSelfRef.t__DOT____VlemCall_0__func) 
                                                | (IData)((IData)(
                                                                  ((0x7fffc000U 
                                                                    & ((- (IData)(
                                                                                (1U 
                                                                                & ((IData)(vlSelfRef.t__DOT__func__Vstatic__func_mid) 
                                                                                >> 0x0dU)))) 
                                                                       << 0x0000000eU)) 
                                                                   | (0x00003fffU 
                                                                      & (IData)(vlSelfRef.t__DOT__func__Vstatic__func_mid))))));
        vlSelfRef.t__DOT__q = vlSelfRef.t__DOT____VlemCall_0__func;
    }

    VLCU_D static void Vt_func_twocall___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_func_twocall___024root___eval_phase__act(Vt_func_twocall___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_func_twocall___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_func_twocall___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_func_twocall___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_func_twocall___024root___eval_phase__nba(Vt_func_twocall___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_func_twocall___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_func_twocall___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_func_twocall___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_func_twocall___024root___eval(Vt_func_twocall___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_func_twocall___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_func_twocall.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_func_twocall___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_func_twocall.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_func_twocall___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_func_twocall___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_func_twocall___024root___ctor_var_reset(Vt_func_twocall___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 17680824324711171631ull);
        vlSelf->t__DOT__sel = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 6625750053862610701ull);
        vlSelf->t__DOT__q = VL_SCOPED_RAND_RESET_Q(34, __VscopeHash, 4790251044348089774ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15508623621340971111ull);
        vlSelf->__Vfunc_t__DOT__func__0__inh = 0;
        vlSelf->__Vfunc_t__DOT__func__0__inl = 0;
        vlSelf->__Vfunc_t__DOT__func__1__inh = 0;
        vlSelf->__Vfunc_t__DOT__func__1__inl = 0;
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
        Vt_func_twocall___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_twocall___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_twocall___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_twocall___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_twocall___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_twocall___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_func_twocall__cuda::info() {
    typedef Vt_func_twocall__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__crc},
        {"t__DOT__sel", "t.sel", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__sel},
        {"t__DOT__func__Vstatic__carry", "t.func__Vstatic__carry", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__func__Vstatic__carry},
        {"__Vfunc_t__DOT__func__0__inh", "__Vfunc_t.func__0__inh", VLCU_MEMBER_STATE, 1, 0, 1, 3, 61, 59, T::OFF___Vfunc_t__DOT__func__0__inh},
        {"__Vfunc_t__DOT__func__1__inh", "__Vfunc_t.func__1__inh", VLCU_MEMBER_STATE, 1, 0, 1, 3, 61, 59, T::OFF___Vfunc_t__DOT__func__1__inh},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__func__Vstatic__func_mid", "t.func__Vstatic__func_mid", VLCU_MEMBER_STATE, 2, 0, 1, 15, 42, 28, T::OFF_t__DOT__func__Vstatic__func_mid},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT____VlemCall_0__func", "t.__VlemCall_0__func", VLCU_MEMBER_STATE, 8, 0, 1, 34, 61, 28, T::OFF_t__DOT____VlemCall_0__func},
        {"t__DOT__q", "t.q", VLCU_MEMBER_STATE, 8, 0, 1, 34, 61, 28, T::OFF_t__DOT__q},
        {"__Vfunc_t__DOT__func__0__inl", "__Vfunc_t.func__0__inl", VLCU_MEMBER_STATE, 8, 0, 1, 40, 41, 2, T::OFF___Vfunc_t__DOT__func__0__inl},
        {"__Vfunc_t__DOT__func__1__inl", "__Vfunc_t.func__1__inl", VLCU_MEMBER_STATE, 8, 0, 1, 40, 41, 2, T::OFF___Vfunc_t__DOT__func__1__inl},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 21, 0, 0x1a019dfe181fbb9aULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_func_twocall__cuda VLCU_TESTBENCHES);
