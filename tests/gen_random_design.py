#!/usr/bin/env python3
"""Random design generator for differential testing of V3EmitCuda + the device operator library
against the genuine Verilated C++ model (both come from the same scheduled AST, so any legal
design is a test of the emitter and of include/verilated_cuda_funcs.h, not of the front end).

    python tests/gen_random_design.py SEED NAME > NAME.v

verilator_b200/csrc/designs/torture/rnd0N.v are this script's output for the seed in their first
line; they go through the same flow as the hand-written torture designs (oracle/Makefile TORTURE
list, tests/test_torture_gpu.py)."""
import random
import sys

WIDTHS = [1, 2, 3, 5, 7, 8, 9, 13, 16, 17, 24, 31, 32, 33, 40, 48, 63, 64, 65, 72, 96, 100, 128, 130, 150]


class Gen:
    def __init__(self, seed):
        self.r = random.Random(seed)
        self.sigs = []  # (name, width) usable as operands

    def const(self, w):
        v = self.r.choice([0, 1, (1 << w) - 1, 1 << (w - 1), self.r.getrandbits(w)])
        return f"{w}'h{v:x}"

    def operand(self, depth):
        r = self.r
        if depth <= 0 or r.random() < 0.25:
            if r.random() < 0.2:
                return self.const(r.choice(WIDTHS))
            name, w = r.choice(self.sigs)
            k = r.random()
            if k < 0.15 and w > 1:  # constant part-select
                hi = r.randrange(w)
                lo = r.randrange(hi + 1)
                return f"{name}[{hi}:{lo}]"
            if k < 0.25 and w > 1:  # variable bit-select, index always in range
                bits = (w - 1).bit_length() - (0 if (w & (w - 1)) == 0 else 1)
                if bits >= 1:
                    iname, iw = r.choice(self.sigs)
                    if iw >= bits:
                        return f"{name}[{iname}[{bits - 1}:0]]"
            if k < 0.33 and w > 8:  # indexed part-select, base + width always in range
                pw = r.choice([1, 4, 8])
                room = w - pw
                bits = room.bit_length() - 1
                if bits >= 1:
                    iname, iw = r.choice(self.sigs)
                    if iw >= bits:
                        return f"{name}[{iname}[{bits - 1}:0]+:{pw}]"
            return name
        return "(" + self.expr(depth - 1) + ")"

    def expr(self, depth):
        r = self.r
        a, b = self.operand(depth), self.operand(depth)
        k = r.randrange(30)
        if k < 4:
            return f"{a} {r.choice(['+', '-', '*'])} {b}"
        if k < 5:  # (wide division is t05_wide's subject; here operands of at most 64 bits, it
            #  is expensive to compile when inlined into every kernel variant)
            return f"{self.narrow_operand(64)} {r.choice(['/', '%'])} {self.narrow_operand(64)}"
        if k < 10:
            return f"{a} {r.choice(['&', '|', '^', '^~'])} {b}"
        if k < 13:
            return f"{a} {r.choice(['<<', '>>', '>>>'])} {self.narrow(depth)}"
        if k < 15:
            return f"$signed({a}) >>> {self.narrow(depth)}"
        if k < 18:
            return f"{a} {r.choice(['==', '!=', '<', '<=', '>', '>='])} {b}"
        if k < 20:
            return f"$signed({a}) {r.choice(['<', '<=', '>', '>='])} $signed({b})"
        if k < 21:
            return f"$signed({self.narrow_operand(64)}) {r.choice(['*', '/', '%'])} $signed({self.narrow_operand(64)})"
        if k < 23:
            return f"{self.operand(depth)} ? {a} : {b}"
        if k < 25:
            return "{" + a + ", " + b + "}"
        if k < 26:
            return "{" + str(r.randrange(1, 4)) + "{" + a + "}}"
        if k < 27:
            return f"{r.choice(['&', '|', '^', '~&', '~|', '~^'])}{a}"
        if k < 28:
            return f"{r.choice(['~', '-', '!'])}{a}"
        if k < 29:
            return f"{a} {r.choice(['&&', '||'])} {b}"
        return f"$countones({a}) + $onehot({b})"

    def narrow_operand(self, maxw):
        for _ in range(20):
            name, w = self.r.choice(self.sigs)
            if w <= maxw:
                return name
        name, w = self.r.choice(self.sigs)
        return f"{name}[{min(w, maxw) - 1}:0]"

    def narrow(self, depth):
        """A shift amount: mostly small, sometimes anything (shifts >= width give 0 / sign)."""
        if self.r.random() < 0.8:
            name, w = self.r.choice(self.sigs)
            return f"{name}[{min(w, 8) - 1}:0]" if w > 1 else name
        return self.operand(depth)


def generate(seed, name, wires=12, depth=2):
    g = Gen(seed)
    r = g.r
    out = [f"// {name}: random design, python tests/gen_random_design.py {seed} {name}",
           f"module {name} ("]
    ports = ["    input  wire clk"]
    for i in range(6):
        w = r.choice(WIDTHS)
        ports.append(f"    input  wire [{w - 1}:0] i{i}")
        g.sigs.append((f"i{i}", w))
    outs = []
    for i, w in enumerate([r.choice([1, 3, 8]), 16, 32, r.choice([40, 64]), r.choice([65, 96, 130])]):
        ports.append(f"    output wire [{w - 1}:0] o{i}")
        outs.append((f"o{i}", w))
    out.append(",\n".join(ports))
    out.append(");")
    body = []
    # registers first (their old values are operands of everything)
    regs = []
    for i in range(4):
        w = r.choice(WIDTHS)
        body.append(f"  reg [{w - 1}:0] r{i};")
        regs.append((f"r{i}", w))
    body.append("  reg [23:0] mem [0:15];")
    g.sigs.extend(regs)
    for i in range(wires):
        w = r.choice(WIDTHS)
        body.append(f"  wire [{w - 1}:0] w{i} = {g.expr(depth)};")
        g.sigs.append((f"w{i}", w))
    body.append("  always @(posedge clk) begin")
    for name_, w in regs:
        if r.random() < 0.3:
            body.append(f"    if ({g.operand(1)}) {name_} <= {g.expr(2)};")
            body.append(f"    else {name_} <= {g.expr(1)};")
        else:
            body.append(f"    {name_} <= {g.expr(2)};")
    a0, a1 = r.choice(g.sigs), r.choice(g.sigs)
    body.append(f"    mem[{a0[0]}[{min(a0[1], 4) - 1}:0]] <= {g.expr(1)};")
    body.append("  end")
    g.sigs.append((f"mem[{a1[0]}[{min(a1[1], 4) - 1}:0]]", 24))
    for name_, w in outs:
        body.append(f"  assign {name_} = {g.expr(2)};")
    out.extend(body)
    out.append("endmodule")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    # optional: number of wires and expression depth (the committed designs use the defaults)
    sys.stdout.write(generate(int(sys.argv[1]), sys.argv[2], *(int(a) for a in sys.argv[3:5])))
