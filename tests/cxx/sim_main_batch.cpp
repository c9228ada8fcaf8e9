// TEST PROGRAM: examples/make_tracing_c/sim_main.cpp:52-113 rewritten against the batched model
// binding (include/verilated_cuda_model.h): the same host-driven loop -- poke ports, timeInc,
// eval(), read outputs -- for N instances at once, every port access going through the C ABI.
// Usage: sim_main_batch <n_instances> <n_evals> <out.bin>
//   out.bin: per eval, per instance: u8 out_small, u64 out_quad, u32 out_wide[3]  (packed)
#include "verilated_cuda_model.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const uint64_t n = strtoull(argv[1], nullptr, 0);
    const uint64_t nEvals = strtoull(argv[2], nullptr, 0);
    try {
        // contextp->randReset(2); ... new Vtop{contextp.get(), "TOP"}
        VlcuModel top{"make_tracing_c", n, /*device=*/0, /*seed=*/1, /*randReset=*/2, 0, 8, "+trace"};
        // Set Vtop's input signals (sim_main.cpp:62-68)
        top.setAll<uint8_t>("reset_l", 1);
        top.setAll<uint8_t>("clk", 0);
        top.setAll<uint8_t>("in_small", 1);
        const uint32_t inWide[3] = {0x11111111, 0x22222222, 0x3};
        top.set("in_wide", inWide, 0);
        std::vector<uint64_t> inQuad(n, 0x1234);
        top.set("in_quad", inQuad.data(), sizeof(uint64_t));
        std::vector<uint8_t> outSmall(n);
        std::vector<uint64_t> outQuad(n);
        std::vector<uint32_t> outWide(3 * n);
        FILE* const fp = fopen(argv[3], "wb");
        if (!fp) return 3;
        uint8_t clk = 0;
        uint64_t time = 0;
        for (uint64_t e = 0; e < nEvals; ++e) {
            top.timeInc(1);
            ++time;
            clk = !clk;
            top.setAll<uint8_t>("clk", clk);
            if (!clk) {
                top.setAll<uint8_t>("reset_l", (time > 1 && time < 10) ? 0 : 1);
                for (uint64_t i = 0; i < n; ++i) inQuad[i] += 0x12;  // top->in_quad += 0x12
                top.set("in_quad", inQuad.data(), sizeof(uint64_t));
            }
            top.eval();
            top.get("out_small", outSmall.data(), 1);
            top.get("out_quad", outQuad.data(), sizeof(uint64_t));
            top.get("out_wide", outWide.data(), 3 * sizeof(uint32_t));
            for (uint64_t i = 0; i < n; ++i) {
                fwrite(&outSmall[i], 1, 1, fp);
                fwrite(&outQuad[i], 8, 1, fp);
                fwrite(&outWide[3 * i], 4, 3, fp);
            }
        }
        top.final();
        fclose(fp);
        const std::vector<vlcu_status_t> st = top.status();
        printf("finished=%d time=%llu first_finish=%llu\n", top.allFinished() ? 1 : 0,
               static_cast<unsigned long long>(st[0].time),
               static_cast<unsigned long long>(st[0].first_finish_time));
    } catch (const std::exception& ex) {
        fprintf(stderr, "sim_main_batch: %s\n", ex.what());
        return 1;
    }
    return 0;
}
