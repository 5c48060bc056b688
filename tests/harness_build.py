"""Builds tests/harness/libexact_host.so: the product's __host__ __device__ arithmetic compiled for the CPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "harness", "exact_math_host.cpp")
LIB = os.path.join(HERE, "harness", "libexact_host.so")
DEPS = [SRC] + [os.path.join(HERE, "..", "vi_depth_completion_b200", "csrc", f) for f in ("exact_math.cuh", "frame_params.cuh")] + [
    os.path.join(HERE, "..", "include", "vidc_b200.h")]


def build():
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        # -x c++: the .cuh headers are plain C++ on the host; -ffp-contract=off mirrors nvcc -fmad=false
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-mfma", "-o", LIB, SRC, "-lm"],
                       check=True)
    return LIB
