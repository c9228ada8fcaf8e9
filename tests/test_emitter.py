"""V3EmitCuda (src/V3EmitCuda.cpp) run for real inside the reference's own compiler
(oracle/_ref/verilator_lite): the models committed under verilator_b200/csrc/designs/ must be
exactly what the emitter prints today, and the emitted text must keep the shape the runtime
relies on.  Needs the oracle build (present wherever /root/reference is)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VLITE = os.path.join(ROOT, "oracle", "_ref", "verilator_lite")
DESIGNS = os.path.join(ROOT, "verilator_b200", "csrc", "designs")
REF = "/root/reference"

CASES = [  # (committed model, top-level source, extra flags)
    ("make_tracing_c/Vtop__cuda.cu", f"{REF}/examples/make_tracing_c/top.v",
     ["--trace-vcd", "--assert", "--coverage", "-cuda-cover-per-instance", "--prefix", "Vtop", "-y",
      f"{REF}/examples/make_tracing_c"]),
    ("rv32i/Vrv32_cov__cuda.cu", "../verilator_b200/csrc/designs/rv32i/rv32_soc.v",  # counters per warp
     ["--coverage", "--prefix", "Vrv32_cov"]),
    ("make_tracing_c/Vtop_nocov__cuda.cu", f"{REF}/examples/make_tracing_c/top.v",
     ["--trace-vcd", "--assert", "--prefix", "Vtop_nocov", "-y", f"{REF}/examples/make_tracing_c"]),
    ("rv32i/Vrv32_soc__cuda.cu", "../verilator_b200/csrc/designs/rv32i/rv32_soc.v",
     ["--trace-vcd", "--savable", "--prefix", "Vrv32_soc"]),
    ("wide_bus/Vwide_bus__cuda.cu", "../verilator_b200/csrc/designs/wide_bus/wide_bus.v",
     ["--trace-vcd", "--prefix", "Vwide_bus"]),
    ("veer_class/Vveer_class__cuda.cu", "../verilator_b200/csrc/designs/veer_class/veer_class.v",
     ["-Wno-WIDTH", "-Wno-UNOPTFLAT", "-Wno-lint", "-Wno-fatal", "-fno-dfg-scoped", "--prefix", "Vveer_class"]),
]
TILED = {"veer_class/Vveer_class__cuda.cu"}

needs_vlite = pytest.mark.skipif(not (os.path.exists(VLITE) and os.path.isdir(REF)),
                                 reason="oracle/_ref/verilator_lite or /root/reference not present")


@needs_vlite
@pytest.mark.parametrize("model,source,flags", CASES)
def test_committed_model_is_what_the_emitter_prints(tmp_path, model, source, flags):
    mdir = str(tmp_path / "obj")
    cmd = [VLITE, "--cc", "--cuda", "--no-std", "--no-std-waiver", "-x-assign", "fast", "--Mdir", mdir, *flags, source]
    res = subprocess.run(cmd, cwd=os.path.join(ROOT, "oracle"), capture_output=True, text=True,
                         env=dict(os.environ, VERILATOR_ROOT=REF))
    assert res.returncode == 0, res.stdout + res.stderr
    fresh = open(os.path.join(mdir, os.path.basename(model))).read()
    committed = open(os.path.join(DESIGNS, model)).read()
    assert fresh == committed, f"{model} is stale: run `make -C oracle regen`"
    # the genuine C++ model was written by the same run
    prefix = flags[flags.index("--prefix") + 1]
    assert os.path.exists(os.path.join(mdir, f"{prefix}___024root__0.cpp"))


@pytest.mark.parametrize("model", [c[0] for c in CASES if c[0] not in TILED])
def test_emitted_model_shape(model):
    text = open(os.path.join(DESIGNS, model)).read()
    for needle in ("#include \"verilated_cuda_launch.h\"", "kStateBytes", "kTouchedReadBytes", "kTraceWords",
                   "VLCU_D static void load(Root& s", "VLCU_D static void store(const Root& s",
                   "VLCU_D static void trace_sample(", "VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx)",
                   "VLCU_REGISTER_DESIGN_X("):
        assert needle in text, needle
    # every scheduled function takes the per-instance context right after the self pointer
    heads = re.findall(r"VLCU_D static \w+ (\w+___eval\w*)\(([^)]*)\)", text)
    assert heads and all("vlSelf, VlcuContext& vlCtx" in args for _, args in heads)
    # host-only constructs never reach device code
    assert "std::string" not in text and "VL_WRITEF" not in text and "->dumpfile(" not in text
    # offsets are unique and 8-byte members come first
    offs = dict(re.findall(r"OFF_(\w+) = (\d+),", text))
    assert len(set(offs.values())) == len(offs)


@needs_vlite
def test_two_module_design_keeps_its_hierarchy(tmp_path):
    """Hierarchy: a child module marked no_inline_module stays a class of its own
    (V3Inline), its scope is addressed through vlSymsp / the cell (V3Descope.cpp:95-125), and the
    emitter prints the tiled form: one struct per module class, Syms as the tile image."""
    src = tmp_path / "two.v"
    src.write_text("""
module leaf (input clk, input [7:0] d, output reg [7:0] q);
  /* verilator no_inline_module */
  always @(posedge clk) q <= d + 8'd1;
endmodule
module two (input clk, input [7:0] d, output [7:0] q);
  leaf u (.clk(clk), .d(d), .q(q));
endmodule
""")
    res = subprocess.run([VLITE, "--cc", "--cuda", "--no-std", "--no-std-waiver", "--Mdir", str(tmp_path / "o"),
                          "--top-module", "two", str(src)],
                         capture_output=True, text=True, env=dict(os.environ, VERILATOR_ROOT=REF))
    assert res.returncode == 0, res.stdout + res.stderr
    text = (tmp_path / "o" / "Vtwo__cuda.cu").read_text()
    assert os.path.exists(tmp_path / "o" / "Vtwo_leaf.h")  # the C++ model keeps the class too
    assert "struct alignas(256) Vtwo_leaf final {" in text and "struct alignas(256) Vtwo__Syms final {" in text
    assert "Vtwo_leaf TOP__two__DOT__u;" in text and "static constexpr bool kTiled = true;" in text
    assert re.search(r"Vtwo_leaf\* vlSelf, Vtwo__Syms\* vlSymsp, VlcuContext& vlCtx", text)
    assert ".v[vlLane]" in text
    # --no-cuda-stream asks for the register form, which cannot hold a second class
    res = subprocess.run([VLITE, "--cc", "--cuda", "--no-cuda-stream", "--no-std", "--no-std-waiver", "--Mdir",
                          str(tmp_path / "o2"), "--top-module", "two", str(src)],
                         capture_output=True, text=True, env=dict(os.environ, VERILATOR_ROOT=REF))
    assert res.returncode != 0 and "was not inlined" in res.stdout + res.stderr


@needs_vlite
def test_cuda_option_gates_the_emitter(tmp_path):
    """--cuda is a real option of the compiler (integration/verilator-cuda.patch): without it
    only the C++ model is written."""
    src = tmp_path / "g.v"
    src.write_text("module g (input clk, input d, output reg q);\n  always @(posedge clk) q <= d;\nendmodule\n")
    for flag, expect in ((None, False), ("--cuda", True), ("--no-cuda", False)):
        mdir = tmp_path / ("o" + (flag or "none").strip("-"))
        cmd = [VLITE, "--cc", "--no-std", "--no-std-waiver", "--Mdir", str(mdir)] + ([flag] if flag else []) + [str(src)]
        res = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, VERILATOR_ROOT=REF))
        assert res.returncode == 0, res.stdout + res.stderr
        assert os.path.exists(mdir / "Vg.cpp")
        assert os.path.exists(mdir / "Vg__cuda.cu") == expect
        assert os.path.exists(mdir / "Vg__cuda.mk") == expect


@needs_vlite
def test_verilate_and_build_in_one_step(tmp_path):
    """bin/verilator_cuda = `verilator --cc --cuda` + `make -f <prefix>__cuda.mk`
    (rules in include/verilated_cuda.mk) gives a loadable lib<prefix>__cuda.so that registers
    the design and exports the C ABI."""
    import shutil
    import sys
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not on PATH")
    src = tmp_path / "cnt.v"
    src.write_text("""
module cnt (input clk, input rst, input [15:0] step, output reg [31:0] acc);
  always @(posedge clk) if (rst) acc <= 0; else acc <= acc + {16'd0, step};
endmodule
""")
    mdir = tmp_path / "obj_dir"
    # (the lite front end cannot parse include/verilated_std.sv: --no-std)
    res = subprocess.run([os.path.join(ROOT, "bin", "verilator_cuda"), "--design", "counter", "--Mdir", str(mdir),
                          "--prefix", "Vcnt", "--no-std", "--no-std-waiver", str(src)], capture_output=True,
                         text=True, env=dict(os.environ, VERILATOR_BIN=VLITE, VERILATOR_ROOT=REF))
    assert res.returncode == 0, res.stdout + res.stderr
    # a fresh process: this one may already hold libverilated_cuda.so with the same symbols
    probe = (
        "import ctypes, sys\n"
        "lib = ctypes.CDLL(sys.argv[1])\n"
        "lib.vlcu_design_name.restype = ctypes.c_char_p\n"
        "assert lib.vlcu_design_count() == 1 and lib.vlcu_design_name(0) == b'counter'\n"
        "for sym in ('vlcu_create', 'vlcu_run', 'vlcu_set_input', 'vlcu_get_output', 'vlcu_coverage_reduce'):\n"
        "    getattr(lib, sym)\n"
    )
    res = subprocess.run([sys.executable, "-c", probe, str(mdir / "libVcnt__cuda.so")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
