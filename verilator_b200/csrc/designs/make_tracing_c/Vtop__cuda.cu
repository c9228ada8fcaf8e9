// -*- mode: C++; c-file-style: "cc-mode" -*-
// DESCRIPTION: examples/make_tracing_c as a batched CUDA model -- the translation unit
// V3EmitCuda would write as Vtop__cuda.cu: device code (Vtop__cuda_body.h), host tables
// (what V3EmitCSyms / __Vconfigure provide to the host runtime) and registration.
// Two builds of the same design, as Verilator would make them:
//   "make_tracing_c"      -cc -x-assign fast --assert            (no --coverage)
//   "make_tracing_c_cov"  ... --coverage  (the flags of examples/make_tracing_c/Makefile)

#include "verilated_cuda_launch.h"

#include <string>
#include <vector>

#include "sim_main_tb.h"

#define VTOP Vtop__cuda
#define VTOP_ROOT Vtop___024root
#define VTOP_COVERAGE 0
#include "Vtop__cuda_body.h"
#undef VTOP
#undef VTOP_ROOT
#undef VTOP_COVERAGE

#define VTOP Vtop_cov__cuda
#define VTOP_ROOT Vtop_cov___024root
#define VTOP_COVERAGE 1
#include "Vtop__cuda_body.h"
#undef VTOP
#undef VTOP_ROOT
#undef VTOP_COVERAGE

namespace {

const char* const kMessages[] = {
    "[%0t] Tracing to logs/vlt_dump.vcd...\n\n",  // top.v:38
    "[%0t] Model running...\n\n",  // top.v:42
    "*-* All Finished *-*\n",  // sub.v:27
    "[%0t] %%Error: sub.v:35: Assertion failed in TOP.top.sub.AssertionExample\n",  // sub.v:35
};
const char* const kPlusargs[] = {"trace"};

#define M(T, D) T::OFF_##D
template <class T>
std::vector<VlcuMemberDesc> makeMembers(bool cov) {
    std::vector<VlcuMemberDesc> m = {
        // name, hierName, kind, elemBytes, traceCode, elems, width, msb, lsb, soaOffset
        {"clk", "top.clk", VLCU_MEMBER_IN, 1, 1, 1, 1, 0, 0, M(T, clk)},
        {"reset_l", "top.reset_l", VLCU_MEMBER_IN, 1, 2, 1, 1, 0, 0, M(T, reset_l)},
        {"out_small", "top.out_small", VLCU_MEMBER_OUT, 1, 3, 1, 2, 1, 0, M(T, out_small)},
        {"out_quad", "top.out_quad", VLCU_MEMBER_OUT, 8, 4, 1, 40, 39, 0, M(T, out_quad)},
        {"out_wide", "top.out_wide", VLCU_MEMBER_OUT, 4, 6, 3, 70, 69, 0, M(T, out_wide)},
        {"in_small", "top.in_small", VLCU_MEMBER_IN, 1, 9, 1, 2, 1, 0, M(T, in_small)},
        {"in_quad", "top.in_quad", VLCU_MEMBER_IN, 8, 10, 1, 40, 39, 0, M(T, in_quad)},
        {"in_wide", "top.in_wide", VLCU_MEMBER_IN, 4, 12, 3, 70, 69, 0, M(T, in_wide)},
        {"top__DOT__sub__DOT__count_c", "top.sub.count_c", VLCU_MEMBER_STATE, 4, 17, 1, 32, 31, 0,
         M(T, top__DOT__sub__DOT__count_c)},
        {"__Vtrigprevexpr___TOP__clk__0", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0,
         M(T, __Vtrigprevexpr___TOP__clk__0)},
        {"__Vtrigprevexpr___TOP__reset_l__0", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0,
         M(T, __Vtrigprevexpr___TOP__reset_l__0)},
        {"__Vtrigprevexpr___TOP__in_small__0", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0,
         M(T, __Vtrigprevexpr___TOP__in_small__0)},
        {"__Vtrigprevexpr___TOP__in_quad__0", nullptr, VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0,
         M(T, __Vtrigprevexpr___TOP__in_quad__0)},
        {"__Vtrigprevexpr___TOP__in_wide__0", nullptr, VLCU_MEMBER_STATE, 4, 0, 3, 70, 69, 0,
         M(T, __Vtrigprevexpr___TOP__in_wide__0)},
        {"__Vtrigprevexpr___TOP__clk__1", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0,
         M(T, __Vtrigprevexpr___TOP__clk__1)},
        {"__VicoDidInit", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, M(T, __VicoDidInit)},
    };
    (void)cov;
    return m;
}
std::vector<VlcuMemberDesc> makeCovMembers() {
    typedef Vtop_cov__cuda T;
    std::vector<VlcuMemberDesc> m = makeMembers<T>(true);
    const VlcuMemberDesc extra[] = {
        {"top__DOT____Vtogcov__clk", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0,
         M(T, top__DOT____Vtogcov__clk)},
        {"top__DOT____Vtogcov__reset_l", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0,
         M(T, top__DOT____Vtogcov__reset_l)},
        {"top__DOT____Vtogcov__out_small", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0,
         M(T, top__DOT____Vtogcov__out_small)},
        {"top__DOT____Vtogcov__out_quad", nullptr, VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0,
         M(T, top__DOT____Vtogcov__out_quad)},
        {"top__DOT____Vtogcov__out_wide", nullptr, VLCU_MEMBER_STATE, 4, 0, 3, 70, 69, 0,
         M(T, top__DOT____Vtogcov__out_wide)},
        {"top__DOT____Vtogcov__in_small", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0,
         M(T, top__DOT____Vtogcov__in_small)},
        {"top__DOT____Vtogcov__in_quad", nullptr, VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0,
         M(T, top__DOT____Vtogcov__in_quad)},
        {"top__DOT____Vtogcov__in_wide", nullptr, VLCU_MEMBER_STATE, 4, 0, 3, 70, 69, 0,
         M(T, top__DOT____Vtogcov__in_wide)},
        {"top__DOT__sub__DOT____Vtogcov__clk", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0,
         M(T, top__DOT__sub__DOT____Vtogcov__clk)},
        {"top__DOT__sub__DOT____Vtogcov__reset_l", nullptr, VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0,
         M(T, top__DOT__sub__DOT____Vtogcov__reset_l)},
        {"top__DOT__sub__DOT____Vtogcov__count_c", nullptr, VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0,
         M(T, top__DOT__sub__DOT____Vtogcov__count_c)},
        {"__Vcoverage", nullptr, VLCU_MEMBER_COVERAGE, 4, 0, 527, 32, 31, 0, M(T, __Vcoverage)},
    };
    for (const VlcuMemberDesc& e : extra) m.push_back(e);
    return m;
}
#undef M

// trace_init_top (V3TraceDecl.cpp): ports at the model's root scope, again inside `top` (same
// codes: V3Trace de-duplicates identical values), then the inlined `sub`
#define SIG(name, code, member, bussed, msb, lsb) {VLCU_TRACE_SIG, bussed, name, code, member, -1, msb, lsb}
#define PORTS_ \
    SIG("clk", 1, 0, 0, 0, 0), SIG("reset_l", 2, 1, 0, 0, 0), SIG("out_small", 3, 2, 1, 1, 0), \
    SIG("out_quad", 4, 3, 1, 39, 0), SIG("out_wide", 6, 4, 1, 69, 0), SIG("in_small", 9, 5, 1, 1, 0), \
    SIG("in_quad", 10, 6, 1, 39, 0), SIG("in_wide", 12, 7, 1, 69, 0)
const VlcuTraceDecl kTraceDecls[] = {
    PORTS_,
    {VLCU_TRACE_PUSH, 0, "top", 0, -1, -1, 0, 0},
    PORTS_,
    {VLCU_TRACE_PUSH, 0, "sub", 0, -1, -1, 0, 0},
    SIG("clk", 1, 0, 0, 0, 0), SIG("reset_l", 2, 1, 0, 0, 0), SIG("count_c", 15, 8, 1, 31, 0),
    {VLCU_TRACE_POP, 0, nullptr, 0, -1, -1, 0, 0},
    {VLCU_TRACE_POP, 0, nullptr, 0, -1, -1, 0, 0},
};
#undef PORTS_
#undef SIG
const uint32_t kNTraceDecls = sizeof(kTraceDecls) / sizeof(kTraceDecls[0]);

// Coverage registrations, in bin order: what Vtop___024root::__Vconfigure would pass to
// __vlCoverToggleInsert / __vlCoverInsert (V3EmitCImp.cpp:179-262, V3EmitCFunc.h:179-215,895-927)
struct CoverTables final {
    std::vector<std::string> strings;  // owns the generated comment strings
    std::vector<VlcuCoverPoint> points;
    void toggle(const char* file, int line, int col, const char* hier, const char* page,
                const char* sig, int lsb, int msb, bool ranged) {
        for (int i = lsb; i <= msb; ++i) {
            for (int j = 0; j < 2; ++j) {
                std::string c = sig;
                if (ranged) c += '[' + std::to_string(i) + ']';
                c += j ? ":0->1" : ":1->0";
                strings.push_back(c);
                points.push_back({file, line, col, hier, page, nullptr, ""});
            }
        }
    }
    void other(const char* file, int line, int col, const char* hier, const char* page,
               const char* comment, const char* linescov) {
        strings.push_back(comment);
        points.push_back({file, line, col, hier, page, nullptr, linescov});
    }
    void finish() {
        for (size_t i = 0; i < points.size(); ++i) points[i].comment = strings[i].c_str();
    }
};
const CoverTables& coverTables() {
    static CoverTables t;
    if (t.points.empty()) {
        t.strings.reserve(600);
        const char* const topv = "top.v";
        const char* const subv = "sub.v";
        t.toggle(topv, 13, 11, ".top", "v_toggle/top", "clk", 0, 0, false);
        t.toggle(topv, 14, 11, ".top", "v_toggle/top", "reset_l", 0, 0, false);
        t.toggle(topv, 16, 23, ".top", "v_toggle/top", "out_small", 0, 1, true);
        t.toggle(topv, 17, 24, ".top", "v_toggle/top", "out_quad", 0, 39, true);
        t.toggle(topv, 18, 24, ".top", "v_toggle/top", "out_wide", 0, 69, true);
        t.toggle(topv, 19, 17, ".top", "v_toggle/top", "in_small", 0, 1, true);
        t.toggle(topv, 20, 18, ".top", "v_toggle/top", "in_quad", 0, 39, true);
        t.toggle(topv, 21, 18, ".top", "v_toggle/top", "in_wide", 0, 69, true);
        t.toggle(subv, 9, 11, ".top.sub", "v_toggle/sub", "clk", 0, 0, false);
        t.toggle(subv, 10, 11, ".top.sub", "v_toggle/sub", "reset_l", 0, 0, false);
        t.toggle(subv, 14, 14, ".top.sub", "v_toggle/sub", "count_c", 0, 31, true);
        t.other(subv, 15, 3, ".top.sub", "v_line/sub", "block", "15");
        t.other(subv, 16, 5, ".top.sub", "v_line/sub", "if", "16,19");
        t.other(subv, 16, 6, ".top.sub", "v_line/sub", "else", "23");
        t.other(subv, 24, 7, ".top.sub", "v_line/sub", "if", "24,27-28");
        t.other(subv, 24, 8, ".top.sub", "v_line/sub", "else", "");
        t.other(subv, 34, 3, ".top.sub", "v_line/sub", "block", "34-35");
        t.other(subv, 39, 3, ".top.sub", "v_user/sub", "cover", "");
        t.finish();
    }
    return t;
}

}  // namespace

const VlcuDesignInfo& Vtop__cuda::info() {
    static const std::vector<VlcuMemberDesc> members = makeMembers<Vtop__cuda>(false);
    static const VlcuDesignInfo s = {
        "make_tracing_c", "top", members.data(), static_cast<uint32_t>(members.size()),
        kStateBytes, kInWords, kOutWords, kInBytes, kOutBytes, 0, 0,
        nullptr, 0, 0,
        kMessages, 4, kPlusargs, 1, kTraceDecls, kNTraceDecls};
    return s;
}
const VlcuDesignInfo& Vtop_cov__cuda::info() {
    static const std::vector<VlcuMemberDesc> members = makeCovMembers();
    static const VlcuDesignInfo s = {
        "make_tracing_c_cov", "top", members.data(), static_cast<uint32_t>(members.size()),
        kStateBytes, kInWords, kOutWords, kInBytes, kOutBytes, 0, 0,
        coverTables().points.data(), static_cast<uint32_t>(coverTables().points.size()),
        OFF___Vcoverage,
        kMessages, 4, kPlusargs, 1, kTraceDecls, kNTraceDecls};
    return s;
}

VLCU_REGISTER_DESIGN(Vtop__cuda, TbSimMain);
VLCU_REGISTER_DESIGN(Vtop_cov__cuda, TbSimMain);
