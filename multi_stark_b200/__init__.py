"""multi-stark commitment hot path on B200 (sm_100a).

Host-side mirror of the reference's `TwoAdicSubgroupDft` / `Pcs` / `Mmcs` surface (reference
src/types.rs:82-85,199-223; src/config.rs:64-123) over the C ABI of libmsgpu.so (include/msgpu.h).
There is no CPU fallback: importing works without a GPU (so the ABI can be checked), but every
compute call needs a CUDA device and raises `MsgpuError` otherwise."""
from ._ffi import MsgpuError, lib, lib_path  # noqa: F401
from .pcs import Challenger, GpuContext, GpuDft, GpuMmcs, GpuPcs, ProverData, pcs_open  # noqa: F401
from .system import (Program, Prover, System, claims_accumulator, fib_trace, multi_workload, shifted_quotient_slices,  # noqa: F401
                     u32_add_workload, wide_trace)

P = 2**64 - 2**32 + 1
GENERATOR = 7
