"""TEST-ONLY: the chain engine's source compiled with W2B_HOST_EMU (one lane per chain, g++).

Lets the CPU-only test run (`-m "not gpu"`) exercise the *logic* of the CUDA kernels against the
oracle.  It is not a product path: w2dynamics_b200.CTQMC loads the CUDA library only.
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "_build", "libw2b_hostemu.so")
SRC = os.path.join(ROOT, "w2dynamics_b200", "csrc")


def build():
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)]
    newest = max(os.path.getmtime(f) for f in srcs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-DW2B_HOST_EMU", "-shared", "-fPIC",
                               "-Wno-unknown-pragmas", "-o", LIB, os.path.join(SRC, "capi.cu")])
    return LIB


def make():
    from w2dynamics_b200.CTQMC import MCTQMC
    return MCTQMC(libpath=build(), device_kind="host-emu")
