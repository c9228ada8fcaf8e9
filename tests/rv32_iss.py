"""Independent RV32I instruction-set simulator with the cycle timing of rv32_soc.v
(TEST INFRASTRUCTURE).  Written from the RISC-V unprivileged spec, not from the scheduled
model: it pins the *design logic* of the hand-scheduled rv32 model, which the oracle and the
GPU path share (see Vrv32_soc__body.inc)."""
M32 = 0xffffffff
S_FETCH, S_EXEC, S_LOAD, S_TRAP = 0, 1, 2, 3
MEM_WORDS = 1024


def sx(v, bits):
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


class Rv32:
    def __init__(self, mem, regs, pc=0, insn=0, ld_addr=0, instret=0, state=0):
        self.mem = [int(x) for x in mem]
        self.regs = [int(x) for x in regs]
        self.pc, self.insn, self.ld_addr, self.instret, self.state = pc, insn, ld_addr, instret, state

    def rd_reg(self, r):
        return 0 if r == 0 else self.regs[r]

    def wr_reg(self, r, v):
        if r:
            self.regs[r] = v & M32

    def posedge(self, resetn):
        if not resetn:
            self.pc, self.state, self.instret = 0, S_FETCH, 0
            return
        if self.state == S_FETCH:
            if self.pc & 3 or (self.pc >> 2) >= MEM_WORDS:
                self.state = S_TRAP
            else:
                self.insn = self.mem[self.pc >> 2]
                self.state = S_EXEC
        elif self.state == S_EXEC:
            self.exec()
        elif self.state == S_LOAD:
            i = self.insn
            f3, rd = (i >> 12) & 7, (i >> 7) & 31
            word = self.mem[(self.ld_addr >> 2) & (MEM_WORDS - 1)]
            sh = word >> (8 * (self.ld_addr & 3))
            val = {0: sx(sh, 8), 1: sx(sh, 16), 2: word, 4: sh & 0xff, 5: sh & 0xffff}[f3]
            self.wr_reg(rd, val)
            self.pc = (self.pc + 4) & M32
            self.instret = (self.instret + 1) & M32
            self.state = S_FETCH
        else:
            self.state = S_TRAP

    def exec(self):
        i, pc = self.insn, self.pc
        op, rd, f3, rs1, rs2, f7 = i & 0x7f, (i >> 7) & 31, (i >> 12) & 7, (i >> 15) & 31, (i >> 20) & 31, i >> 25
        a, b = self.rd_reg(rs1), self.rd_reg(rs2)
        imm_i = sx(i >> 20, 12)
        next_pc, retire, state = (pc + 4) & M32, True, S_FETCH

        def alu(f3, x, y, sub, sra):
            sh = y & 31
            return {0: (x - y) if sub else (x + y), 1: x << sh, 2: int(sx(x, 32) < sx(y, 32)),
                    3: int(x < (y & M32)), 4: x ^ y, 5: (sx(x, 32) >> sh) if sra else (x >> sh),
                    6: x | y, 7: x & y}[f3] & M32

        if op == 0x37:
            self.wr_reg(rd, i & 0xfffff000)
        elif op == 0x17:
            self.wr_reg(rd, pc + (i & 0xfffff000))
        elif op == 0x6f:
            imm = sx(((i >> 31) << 20) | (((i >> 12) & 0xff) << 12) | (((i >> 20) & 1) << 11)
                     | (((i >> 21) & 0x3ff) << 1), 21)
            self.wr_reg(rd, pc + 4)
            next_pc = (pc + imm) & M32
        elif op == 0x67 and f3 == 0:
            self.wr_reg(rd, pc + 4)
            next_pc = (a + imm_i) & M32 & ~1
        elif op == 0x63 and f3 not in (2, 3):
            imm = sx(((i >> 31) << 12) | (((i >> 7) & 1) << 11) | (((i >> 25) & 0x3f) << 5)
                     | (((i >> 8) & 0xf) << 1), 13)
            take = {0: a == b, 1: a != b, 4: sx(a, 32) < sx(b, 32), 5: sx(a, 32) >= sx(b, 32),
                    6: a < b, 7: a >= b}[f3]
            if take:
                next_pc = (pc + imm) & M32
        elif op == 0x13 and (f3 != 1 or f7 == 0) and (f3 != 5 or f7 in (0, 0x20)):
            self.wr_reg(rd, alu(f3, a, imm_i & M32, False, f3 == 5 and f7 == 0x20))
        elif op == 0x33 and (f7 == 0 or (f7 == 0x20 and f3 in (0, 5))):
            self.wr_reg(rd, alu(f3, a, b, f7 == 0x20 and f3 == 0, f7 == 0x20 and f3 == 5))
        elif op == 0x03 and f3 in (0, 1, 2, 4, 5):
            addr = (a + imm_i) & M32
            size = f3 & 3
            retire, next_pc = False, pc
            if (size == 1 and addr & 1) or (size == 2 and addr & 3) or (addr >> 2) >= MEM_WORDS:
                state = S_TRAP
            else:
                self.ld_addr, state = addr, S_LOAD
        elif op == 0x23 and f3 in (0, 1, 2):
            imm = sx(((i >> 25) << 5) | rd, 12)
            addr = (a + imm) & M32
            if (f3 == 1 and addr & 1) or (f3 == 2 and addr & 3) or (addr >> 2) >= MEM_WORDS:
                state, retire, next_pc = S_TRAP, False, pc
            else:
                w = addr >> 2
                if f3 == 0:
                    sh = 8 * (addr & 3)
                    self.mem[w] = (self.mem[w] & ~(0xff << sh) | ((b & 0xff) << sh)) & M32
                elif f3 == 1:
                    sh = 8 * (addr & 3)
                    self.mem[w] = (self.mem[w] & ~(0xffff << sh) | ((b & 0xffff) << sh)) & M32
                else:
                    self.mem[w] = b
        else:
            state, retire, next_pc = S_TRAP, False, pc
        self.pc, self.state = next_pc, state
        if retire:
            self.instret = (self.instret + 1) & M32

    def run(self, n_evals, reset_time=16):
        """The testbench of TbRv32Rand: time += 1 and clk toggles per eval; the first eval only
        initialises edge detection; resetn = time > reset_time."""
        for t in range(1, n_evals + 1):
            clk = t & 1
            if clk and t > 1:
                self.posedge(t > reset_time)
