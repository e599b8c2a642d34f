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


_HOOKS = []          # keep the ctypes callbacks alive


def make():
    """An MCTQMC object on the host-emulation library.  The product class knows nothing about it: the collective the CUDA library
    performs with NCCL is replaced here, through the test-only symbol w2b_test_set_allreduce_hook, by a torch.distributed (gloo)
    all-reduce of the host accumulator slab."""
    import ctypes as C
    import numpy as np
    from w2dynamics_b200.CTQMC import MCTQMC

    class HostEmuMCTQMC(MCTQMC):
        def enable_distributed(self, group=None):
            import torch
            import torch.distributed as dist
            rank, size = dist.get_rank(group), dist.get_world_size(group)

            def hook(buf, n, flags):
                t = torch.from_numpy(np.ctypeslib.as_array(buf, shape=(n,)))
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                f = torch.tensor([flags[0], flags[1]], dtype=torch.int32)
                dist.all_reduce(f, op=dist.ReduceOp.MAX, group=group)
                flags[0], flags[1] = int(f[0]), int(f[1])
            cb = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_int))(hook)
            _HOOKS.append(cb)
            self._load().w2b_test_set_allreduce_hook(cb)
            self.comm_init(b"\0" * 128, rank, size)

    return HostEmuMCTQMC(libpath=build())
