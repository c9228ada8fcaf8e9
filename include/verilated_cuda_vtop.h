// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
//
// verilated_cuda_vtop.h -- the generated model class with its PUBLIC PORT MEMBERS, over the
// batched CUDA evaluator with one instance: a testbench written for the C++ model
// (examples/make_tracing_c/sim_main.cpp) compiles and runs UNMODIFIED.
//
// The reference's model class (src/V3EmitCModel.cpp:55-269) has one public reference member per
// primary port (`VL_IN8(&clk,0,0)` ..., :108-117) that the testbench pokes and reads between
// calls of eval().  Here a three-line "Vtop.h" gives the class its name, its design and its
// ports (the <prefix>__ports.inc V3EmitCuda prints beside <prefix>__cuda.cu):
//
//     #define VLCU_VTOP Vtop
//     #define VLCU_VTOP_DESIGN "make_tracing_c"
//     #define VLCU_VTOP_PORTS "Vtop__ports.inc"
//     #include "verilated_cuda_vtop.h"
//
// and eval() becomes: inputs -> device (vlcu_set_input), the context's time -> device, one
// vlcu_eval(), outputs <- device (vlcu_get_output), $display text printed with VL_PRINTF,
// $finish/$stop handed to the VerilatedContext the testbench polls (gotFinish()).  It is the
// K = 1, N = 1 corner of the evaluator -- a drop-in for bring-up and debugging, three PCIe round
// trips per eval -- not the way to get throughput (that is vlcu_run over a batch).
//
// Needs the reference's own verilated.h (VerilatedContext, VerilatedModel, VL_IN/VL_OUT types)
// and links against the reference runtime and libverilated_cuda.so.
//
//*************************************************************************
#ifndef VERILATOR_VERILATED_CUDA_VTOP_H_
#define VERILATOR_VERILATED_CUDA_VTOP_H_

#include "verilated.h"
#include "verilated_cuda_abi.h"

#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#if !defined(VLCU_VTOP) || !defined(VLCU_VTOP_DESIGN) || !defined(VLCU_VTOP_PORTS)
#error "define VLCU_VTOP (class name), VLCU_VTOP_DESIGN (registered design) and VLCU_VTOP_PORTS (<prefix>__ports.inc)"
#endif

// C type of a port by its width, as AstNodeDType::cTypeRecurse picks it (src/V3AstNodes.cpp:1175-1300)
template <int T_Width, int T_Words>
struct VlcuVtopPort final {
    typedef typename std::conditional<
        (T_Width <= 8), CData,
        typename std::conditional<
            (T_Width <= 16), SData,
            typename std::conditional<(T_Width <= 32), IData,
                                      typename std::conditional<(T_Width <= 64), QData, VlWide<T_Words>>::type>::type>::
            type>::type type;
};

class VLCU_VTOP final : public VerilatedModel {
    vlcu_model_t* m_h = nullptr;
    std::string m_name;
    uint64_t m_deviceTime = 0;
    uint32_t m_msgSeen = 0;

    static void check(int rc, const char* what) {
        if (rc != VLCU_OK) throw std::runtime_error(std::string{what} + ": " + vlcu_last_error());
    }
    // $display/$write text, formatted like VL_WRITEF for the conversions the designs here use
    void printMessage(const vlcu_msg_t& msg) {
        const char* fmt = vlcu_message_format(m_h, msg.id);
        if (!fmt) return;
        std::string out;
        uint32_t arg = 0;
        for (const char* p = fmt; *p; ++p) {
            if (*p != '%') {
                out += *p;
                continue;
            }
            ++p;
            while (*p && (std::isdigit(static_cast<unsigned char>(*p)) || *p == '-' || *p == '.')) ++p;
            if (!*p) break;
            if (*p == '%') {
                out += '%';
                continue;
            }
            const uint64_t v = arg < msg.nargs ? msg.args[arg] : 0;
            ++arg;
            char buf[72];
            switch (std::tolower(static_cast<unsigned char>(*p))) {
            case 'x':
            case 'h': snprintf(buf, sizeof(buf), "%llx", static_cast<unsigned long long>(v)); break;
            case 'b': {
                int k = 0;
                for (int b = 63; b >= 0; --b) {
                    if (k || ((v >> b) & 1U) || b == 0) buf[k++] = ((v >> b) & 1U) ? '1' : '0';
                }
                buf[k] = 0;
                break;
            }
            default: snprintf(buf, sizeof(buf), "%llu", static_cast<unsigned long long>(v)); break;  // d, t
            }
            out += buf;
        }
        VL_PRINTF("%s", out.c_str());
    }

public:
    // PORTS: the members the testbench pokes and reads (V3EmitCModel.cpp:108-117)
#define VLCU_PORT_IN(name, width, words) typename VlcuVtopPort<width, words>::type name{};
#define VLCU_PORT_OUT(name, width, words) typename VlcuVtopPort<width, words>::type name{};
#define VLCU_PORT_INOUT(name, width, words) typename VlcuVtopPort<width, words>::type name{};
#include VLCU_VTOP_PORTS
#undef VLCU_PORT_IN
#undef VLCU_PORT_OUT
#undef VLCU_PORT_INOUT
    VLCU_VTOP* const rootp = this;  // (testbenches reach internals through rootp in the reference)

    // CONSTRUCTORS (V3EmitCModel.cpp:159-161): randReset / randSeed come from the context, as the
    // C++ model's _ctor_var_reset takes them
    explicit VLCU_VTOP(VerilatedContext* contextp, const char* name = "TOP")
        : VerilatedModel{*contextp}
        , m_name{name} {
        m_h = vlcu_create(VLCU_VTOP_DESIGN, 1, 0, 0, static_cast<uint64_t>(contextp->randSeed()),
                          contextp->randReset(), 64, nullptr);
        if (!m_h) throw std::runtime_error(std::string{"vlcu_create: "} + vlcu_last_error());
    }
    explicit VLCU_VTOP(const char* name = "TOP")
        : VLCU_VTOP{Verilated::threadContextp(), name} {}
    ~VLCU_VTOP() override { vlcu_destroy(m_h); }

    // API METHODS (V3EmitCModel.cpp:171-208)
    void eval() { eval_step(); }
    void eval_step() {
        int port = 0;
#define VLCU_PORT_IN(name, width, words) check(vlcu_set_input(m_h, port++, &name, sizeof(name)), #name);
#define VLCU_PORT_OUT(name, width, words) ++port;
#define VLCU_PORT_INOUT(name, width, words) check(vlcu_set_input(m_h, port++, &name, sizeof(name)), #name);
#include VLCU_VTOP_PORTS
#undef VLCU_PORT_IN
#undef VLCU_PORT_OUT
#undef VLCU_PORT_INOUT
        const uint64_t now = contextp()->time();
        if (now != m_deviceTime) check(vlcu_time_inc(m_h, now - m_deviceTime), "vlcu_time_inc");
        m_deviceTime = now;
        check(vlcu_eval(m_h), "vlcu_eval");
        port = 0;
#define VLCU_PORT_IN(name, width, words) ++port;
#define VLCU_PORT_OUT(name, width, words) check(vlcu_get_output(m_h, port++, &name, sizeof(name)), #name);
#define VLCU_PORT_INOUT(name, width, words) check(vlcu_get_output(m_h, port++, &name, sizeof(name)), #name);
#include VLCU_VTOP_PORTS
#undef VLCU_PORT_IN
#undef VLCU_PORT_OUT
#undef VLCU_PORT_INOUT
        // what the eval printed and whether it finished or stopped
        vlcu_msg_t msgs[64];
        uint32_t count = 0;
        check(vlcu_messages(m_h, 0, msgs, 64, &count), "vlcu_messages");
        for (uint32_t k = m_msgSeen; k < count && k < 64; ++k) printMessage(msgs[k]);
        m_msgSeen = count;
        vlcu_status_t st;
        check(vlcu_status(m_h, &st), "vlcu_status");
        if (st.status & VLCU_STATUS_FINISH) contextp()->gotFinish(true);
        if (st.status & VLCU_STATUS_ERROR) contextp()->gotError(true);
    }
    void eval_end_step() {}
    void final() { check(vlcu_final(m_h), "vlcu_final"); }
    bool eventsPending() { return false; }
    uint64_t nextTimeSlot() { return 0; }
    // (waveforms of the batched model are sampled by vlcu_run; nothing to attach here)
    void trace(VerilatedTraceBaseC*, int, int = 0) {}
    const char* name() const { return m_name.c_str(); }
    const char* hierName() const override { return m_name.c_str(); }
    const char* modelName() const override { return VLCU_VTOP_DESIGN; }
    unsigned threads() const override { return 1; }
    vlcu_model_t* handle() const { return m_h; }
};

#endif  // guard
