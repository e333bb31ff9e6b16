// Build shim for oracle/build_ref_dcn.sh (TEST INFRASTRUCTURE).  The reference's deform_conv_cuda_kernel.cu includes
// <THC/THCAtomics.cuh>, a header PyTorch removed; the reference only needs the atomicAdd overloads, which ATen provides.
// The reference sources themselves are compiled unmodified from /root/reference.
#pragma once
#include <ATen/cuda/Atomic.cuh>
