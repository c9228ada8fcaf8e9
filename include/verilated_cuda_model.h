// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
//
// verilated_cuda_model.h -- header-only C++ binding of the C ABI
// (verilated_cuda_abi.h) with the verbs of the generated model class.
//
// This is the reference-side stub a testbench author uses instead of
// "Vtop.h": where the reference has
//     VerilatedContext ctx; ctx.randReset(2); Vtop top{&ctx, "TOP"};
//     top.in_quad = v; ctx.timeInc(1); top.eval(); v = top.out_quad; top.final();
// (src/V3EmitCModel.cpp:55-269, examples/make_tracing_c/sim_main.cpp:59-113)
// a batch of N models is
//     VlcuModel top{"design", N, device, seed, randReset};
//     top.set("in_quad", ptr, stride); top.timeInc(1); top.eval();
//     top.get("out_quad", ptr, stride); top.final();
// Errors of the C ABI become exceptions (std::runtime_error) carrying
// vlcu_last_error(); per-instance $finish/$stop are read with status().
//
//*************************************************************************
#ifndef VERILATOR_VERILATED_CUDA_MODEL_H_
#define VERILATOR_VERILATED_CUDA_MODEL_H_

#include "verilated_cuda_abi.h"

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

class VlcuModel final {
    vlcu_model_t* m_h = nullptr;

    static void check(int rc, const char* what) {
        if (rc != VLCU_OK) {
            throw std::runtime_error(std::string{what} + ": " + vlcu_last_error() + " (code "
                                     + std::to_string(rc) + ")");
        }
    }

public:
    // CONSTRUCTORS
    // instance i of this shard has global index instBase+i and randSeed = seed ^ (instBase+i)
    VlcuModel(const char* design, uint64_t nInstances, int device = 0, uint64_t seed = 1,
              int randReset = 0, uint64_t instBase = 0, uint32_t msgCap = 4,
              const char* plusargs = nullptr) {
        m_h = vlcu_create(design, nInstances, device, instBase, seed, randReset, msgCap, plusargs);
        if (!m_h) throw std::runtime_error(std::string{"vlcu_create: "} + vlcu_last_error());
    }
    ~VlcuModel() { vlcu_destroy(m_h); }
    VlcuModel(const VlcuModel&) = delete;
    VlcuModel& operator=(const VlcuModel&) = delete;

    // ACCESSORS
    vlcu_model_t* handle() const { return m_h; }
    uint64_t instances() const { return vlcu_instances(m_h); }
    uint64_t paddedInstances() const { return vlcu_padded_instances(m_h); }
    int port(const char* name) const {
        const int p = vlcu_port_index(m_h, name);
        check(p < 0 ? p : VLCU_OK, name);
        return p;
    }
    vlcu_port_info_t portInfo(int port) const {
        vlcu_port_info_t info;
        check(vlcu_port_info(m_h, port, &info), "vlcu_port_info");
        return info;
    }

    // PORTS  (top->port = v  /  v = top->port), one value per instance at src + i*stride
    void set(int port, const void* src, size_t strideBytes) {
        check(vlcu_set_input(m_h, port, src, strideBytes), "vlcu_set_input");
    }
    void set(const char* name, const void* src, size_t strideBytes) { set(port(name), src, strideBytes); }
    template <typename T>
    void setAll(const char* name, T value) {  // the same value for every instance
        set(port(name), &value, 0);
    }
    void get(int port, void* dst, size_t strideBytes) {
        check(vlcu_get_output(m_h, port, dst, strideBytes), "vlcu_get_output");
    }
    void get(const char* name, void* dst, size_t strideBytes) { get(port(name), dst, strideBytes); }

    // EVALUATION
    void timeInc(uint64_t add) { check(vlcu_time_inc(m_h, add), "vlcu_time_inc"); }
    void eval() { check(vlcu_eval(m_h), "vlcu_eval"); }  // Vtop::eval(), every instance
    void eval_step() { eval(); }
    void run(const vlcu_run_desc_t& desc) { check(vlcu_run(m_h, &desc), "vlcu_run"); }
    void final() { check(vlcu_final(m_h), "vlcu_final"); }
    void sync() { check(vlcu_sync(m_h), "vlcu_sync"); }
    void reset() { check(vlcu_reset(m_h), "vlcu_reset"); }

    // CONTEXT  (contextp->time() / gotFinish() / gotError() per instance)
    std::vector<vlcu_status_t> status() {
        std::vector<vlcu_status_t> out(instances());
        check(vlcu_status(m_h, out.data()), "vlcu_status");
        return out;
    }
    bool allFinished() {
        for (const vlcu_status_t& s : status()) {
            if (!(s.status & VLCU_STATUS_FINISH)) return false;
        }
        return true;
    }
    std::vector<uint64_t> signatures() {
        std::vector<uint64_t> out(instances());
        check(vlcu_signatures(m_h, out.data()), "vlcu_signatures");
        return out;
    }

    // PIPELINED CAMPAIGNS: the next batch's inputs travel while this batch computes, and this
    // batch's results while the next one computes (page-locked buffers: vlcu_host_alloc)
    int member(const char* name) const {
        const int m = vlcu_member_index(m_h, name);
        check(m < 0 ? m : VLCU_OK, name);
        return m;
    }
    void stage(int member, const void* src, size_t strideBytes, bool maskWidth = true) {
        check(vlcu_stage_member(m_h, member, src, strideBytes, maskWidth ? 1 : 0), "vlcu_stage_member");
    }
    void commitStaged() { check(vlcu_commit_staged(m_h), "vlcu_commit_staged"); }
    void resultsBegin(uint64_t* sigs, vlcu_status_t* status, const std::vector<int>& members = {},
                      const std::vector<void*>& dsts = {}) {
        check(vlcu_results_begin(m_h, sigs, status, static_cast<int>(members.size()), members.data(),
                                 dsts.data()),
              "vlcu_results_begin");
    }
    void resultsWait() { check(vlcu_results_wait(m_h), "vlcu_results_wait"); }

    // SAVE / RESTORE  (VerilatedSave os; os << *top;  /  VerilatedRestore os; os >> *top;  for one
    // instance, in the reference's file format: include/verilated_save.h:193-223)
    void save(uint64_t instance, const char* filename) {
        check(vlcu_save_instance(m_h, instance, filename), "vlcu_save_instance");
    }
    void restore(uint64_t instance, const char* filename) {
        check(vlcu_restore_instance(m_h, instance, filename), "vlcu_restore_instance");
    }

    // WAVEFORMS  (tfp->dump(time) of a sampled instance; `trace` / `streams` as vlcu_run filled them)
    void vcdWrite(const void* trace, uint64_t traceEvals, uint32_t nSamples, uint32_t sample,
                  const char* filename) {
        check(vlcu_vcd_write(m_h, trace, traceEvals, nSamples, sample, filename), "vlcu_vcd_write");
    }
    void vcdWriteChanges(const uint32_t* streams, uint32_t streamWords, uint32_t nSamples, uint32_t sample,
                         const char* filename) {
        check(vlcu_vcd_write_changes(m_h, streams, streamWords, nSamples, sample, filename),
              "vlcu_vcd_write_changes");
    }

    // COVERAGE  (contextp->coveragep()->write(filename) for the sum over all instances)
    std::vector<uint64_t> coverageReduce() {
        std::vector<uint64_t> out(vlcu_coverage_bins(m_h));
        check(vlcu_coverage_reduce(m_h, out.data()), "vlcu_coverage_reduce");
        return out;
    }
    void coverageWrite(const char* filename) {
        const std::vector<uint64_t> sums = coverageReduce();
        check(vlcu_coverage_write(m_h, sums.data(), filename), "vlcu_coverage_write");
    }
};

#endif  // guard
