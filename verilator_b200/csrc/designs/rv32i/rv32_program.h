// -*- mode: C++; c-file-style: "cc-mode" -*-
// Synthetic program generator of the rv32i workload (BASELINE.json config 2: "each instance
// running a random-seeded synthetic program").  ONE source, compiled twice: as device code
// into the CUDA testbench (rv32_tb.h) and as host code into the oracle harness
// (oracle/models/rv32i/harness.cpp), so both sides load bit-identical programs.  It only
// needs a VlRNG with rand64() (include/verilated_cuda.h resp. the reference's verilated.h).
#ifndef RV32_PROGRAM_H_
#define RV32_PROGRAM_H_
#ifndef RV32_GEN_FN
#define RV32_GEN_FN inline
#endif

//=============================================================================
// Synthetic stimulus shared by the device testbench and the oracle harness: a
// random *valid* RV32I program (ALU / forward branch / load / store mix) in words
// 0..511 whose last instruction jumps back to word 2, random data in words 512..1023, drawn from
// VlRNG(seed) = the reference's splitmix64 -> xoroshiro128** (SURVEY 8d config 2).
// x1 holds the data-region base (2048) and is never a destination.

template <class T_Mem>
RV32_GEN_FN void rv32_gen_program(T_Mem& mem, uint64_t seed) {
    VlRNG rng{seed};
    const IData PROG = 512;
    mem[0] = 0x7ff00093U;  // addi x1, x0, 2047
    mem[1] = 0x00108093U;  // addi x1, x1, 1      -> x1 = 2048
    for (IData i = 2; i < PROG - 1; ++i) {
        const uint64_t r = rng.rand64();
        const IData kind = static_cast<IData>(r & 15U);
        IData rd = static_cast<IData>((r >> 4) & 31U);
        if (rd == 1U) rd = 0U;  // never clobber the data base register
        const IData rs1 = static_cast<IData>((r >> 9) & 31U);
        const IData rs2 = static_cast<IData>((r >> 14) & 31U);
        const IData f3 = static_cast<IData>((r >> 19) & 7U);
        const IData imm12 = static_cast<IData>((r >> 22) & 0xfffU);
        const IData bit = static_cast<IData>((r >> 34) & 1U);
        IData insn;
        if (kind < 6U) {  // ALU register-register
            const IData f7 = (bit && (f3 == 0U || f3 == 5U)) ? 0x20U : 0U;
            insn = (f7 << 25) | (rs2 << 20) | (rs1 << 15) | (f3 << 12) | (rd << 7) | 0x33U;
        } else if (kind < 10U) {  // ALU immediate
            IData imm = imm12;
            if (f3 == 1U) imm &= 0x1fU;
            if (f3 == 5U) imm = (imm & 0x1fU) | (bit ? 0x400U : 0U);
            insn = (imm << 20) | (rs1 << 15) | (f3 << 12) | (rd << 7) | 0x13U;
        } else if (kind == 10U) {  // LUI
            insn = (static_cast<IData>(r >> 36) << 12) | (rd << 7) | 0x37U;
        } else if (kind == 11U) {  // AUIPC
            insn = (static_cast<IData>(r >> 36) << 12) | (rd << 7) | 0x17U;
        } else if (kind < 14U) {  // conditional branch, forward only (no tiny endless loops)
            const IData bsel = static_cast<IData>((r >> 36) % 6U);
            const IData bf3 = (bsel < 2U) ? bsel : (bsel + 2U);  // {0,1,4,5,6,7}
            IData words = static_cast<IData>((r >> 40) & 7U) + 1U;  // +1..+8 instructions
            if (i + words > PROG - 1U) words = PROG - 1U - i;
            const IData off = words * 4U;
            insn = (((off >> 12) & 1U) << 31) | (((off >> 5) & 0x3fU) << 25) | (rs2 << 20) | (rs1 << 15)
                   | (bf3 << 12) | (((off >> 1) & 0xfU) << 8) | (((off >> 11) & 1U) << 7) | 0x63U;
        } else if (kind == 14U) {  // load from the data region: base x1
            const IData lsel = static_cast<IData>((r >> 36) % 5U);
            const IData lf3 = (lsel < 3U) ? lsel : (lsel + 1U);  // {0,1,2,4,5}
            IData off = imm12 & 0x7ffU;
            if ((lf3 & 3U) == 1U) off &= ~1U;
            if ((lf3 & 3U) == 2U) off &= ~3U;
            insn = (off << 20) | (1U << 15) | (lf3 << 12) | (rd << 7) | 0x03U;
        } else {  // store to the data region: base x1
            const IData sf3 = static_cast<IData>((r >> 36) % 3U);
            IData off = imm12 & 0x7ffU;
            if (sf3 == 1U) off &= ~1U;
            if (sf3 == 2U) off &= ~3U;
            insn = ((off >> 5) << 25) | (rs2 << 20) | (1U << 15) | (sf3 << 12) | ((off & 0x1fU) << 7) | 0x23U;
        }
        mem[i] = insn;
    }
    {  // jal x0, -(PROG-3)*4: back to word 2; registers and data have changed, so every pass
        // through the program takes different branches and touches different data
        const IData off = 0U - (PROG - 3U) * 4U;
        mem[PROG - 1] = (((off >> 20) & 1U) << 31) | (((off >> 1) & 0x3ffU) << 21) | (((off >> 11) & 1U) << 20)
                        | (((off >> 12) & 0xffU) << 12) | 0x6fU;
    }
    for (IData i = PROG; i < 1024U; ++i) mem[i] = static_cast<IData>(rng.rand64());
}

#endif  // guard
