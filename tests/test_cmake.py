"""SURVEY 8(f) N1, the CMake half: integration/verilator-config-cuda.patch adds `CUDA` (and
CUDA_DESIGN / CUDA_TESTBENCH) to the reference's verilate() function
(verilator-config.cmake.in:259).  The patch must apply to the reference's file as it is, and a
project that calls verilate(<target> CUDA SOURCES ...) must configure -- running the compiler with
--cuda -- and, with VLCU_SLOW_TESTS=1, build a library that registers the design."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
VLITE = os.path.join(ROOT, "oracle", "_ref", "verilator_lite")
PATCH = os.path.join(ROOT, "integration", "verilator-config-cuda.patch")

needs = pytest.mark.skipif(not (os.path.exists(VLITE) and os.path.isdir(REF)),
                           reason="oracle/_ref/verilator_lite or /root/reference not present")


def _cmake():
    try:
        import cmake
        return os.path.join(cmake.CMAKE_BIN_DIR, "cmake")
    except ImportError:
        return shutil.which("cmake")


def _kit(tmp_path):
    """A VERILATOR_ROOT as `make install` would lay it out: the reference's include/, this repo's
    runtime beside it, bin/verilator_bin = the compiler built with V3EmitCuda, and
    verilator-config.cmake = the reference's template with the patch applied."""
    kit = tmp_path / "kit"
    (kit / "bin").mkdir(parents=True)
    shutil.copytree(os.path.join(REF, "include"), kit / "include")
    for name in os.listdir(os.path.join(ROOT, "include")):
        shutil.copy(os.path.join(ROOT, "include", name), kit / "include" / name)
    shutil.copy(os.path.join(ROOT, "verilator_b200", "csrc", "verilated_cuda.cu"), kit / "include" / "verilated_cuda.cu")
    shutil.copy(os.path.join(ROOT, "oracle", "_ref", "verilated_config.h"), kit / "include" / "verilated_config.h")
    os.symlink(VLITE, kit / "bin" / "verilator_bin")
    shutil.copy(os.path.join(REF, "verilator-config.cmake.in"), kit / "verilator-config.cmake.in")
    res = subprocess.run(["patch", "-p1", "-d", str(kit), "-i", PATCH], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    text = (kit / "verilator-config.cmake.in").read_text()
    for key, value in (("@CFG_CXX_FLAGS_CMAKE@", ""), ("@CFG_LDFLAGS_THREADS_CMAKE@", "-pthread"),
                       ("@CFG_CXXFLAGS_COROUTINES@", "-std=gnu++20")):
        text = text.replace(key, value)
    (kit / "verilator-config.cmake").write_text(text)
    return kit


@needs
def test_verilate_cuda_option(tmp_path):
    cmake = _cmake()
    if not cmake or not shutil.which("nvcc"):
        pytest.skip("cmake or nvcc not available")
    kit = _kit(tmp_path)
    proj = tmp_path / "proj"
    proj.mkdir()
    (proj / "cnt.v").write_text("""
module cnt (input clk, input rst, input [15:0] step, output reg [31:0] acc);
  always @(posedge clk) if (rst) acc <= 0; else acc <= acc + {16'd0, step};
endmodule
""")
    (proj / "CMakeLists.txt").write_text("""
cmake_minimum_required(VERSION 3.19)
project(cnt_batch LANGUAGES CXX CUDA)
find_package(verilator REQUIRED HINTS $ENV{VERILATOR_ROOT})
add_library(cnt_batch SHARED)
verilate(cnt_batch CUDA CUDA_DESIGN counter PREFIX Vcnt SOURCES cnt.v VERILATOR_ARGS --no-std --no-std-waiver)
""")
    env = dict(os.environ, VERILATOR_ROOT=str(kit), CUDACXX=shutil.which("nvcc"))
    build = tmp_path / "build"
    res = subprocess.run([cmake, "-S", str(proj), "-B", str(build), "-DCMAKE_CUDA_ARCHITECTURES=100a",
                          "-DCMAKE_BUILD_TYPE=Release", "-DOBJCACHE_ENABLED=OFF"], capture_output=True, text=True, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    vdir = build / "CMakeFiles" / "cnt_batch.dir" / "Vcnt.dir"
    assert (vdir / "Vcnt__cuda.cu").exists() and (vdir / "Vcnt.cpp").exists()  # --cuda reached the compiler
    if os.environ.get("VLCU_SLOW_TESTS") != "1":
        return
    res = subprocess.run([cmake, "--build", str(build), "-j", "8"], capture_output=True, text=True, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    probe = ("import ctypes, sys\nlib = ctypes.CDLL(sys.argv[1])\nlib.vlcu_design_name.restype = ctypes.c_char_p\n"
             "assert lib.vlcu_design_count() == 1 and lib.vlcu_design_name(0) == b'counter'\n")
    res = subprocess.run([sys.executable, "-c", probe, str(build / "libcnt_batch.so")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr


@needs
def test_patch_applies_to_the_reference_file(tmp_path):
    shutil.copy(os.path.join(REF, "verilator-config.cmake.in"), tmp_path / "verilator-config.cmake.in")
    res = subprocess.run(["patch", "-p1", "--dry-run", "-d", str(tmp_path), "-i", PATCH], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
