// Host-side declarations of the per-format kernel launchers (one translation unit per family).
#pragma once
#include "common.cuh"

namespace cfx {

// Number of CTAs that fill the device once for `kernel` (SM count x resident CTAs per SM).
uint32_t persistent_ctas(const void* kernel, int threads, size_t dyn_smem = 0);

// Each launcher enqueues the kernels for one surface on `stream` and returns how many kernels
// it launched (>0) or a negative CFX_ERR_* code.
int launch_bc45(const EncodeParams& p, cudaStream_t stream);
int launch_bc7(const EncodeParams& p, cudaStream_t stream);
int launch_astc(const EncodeParams& p, cudaStream_t stream);
int launch_bc6h(const EncodeParams& p, cudaStream_t stream);
int launch_bc123(const EncodeParams& p, cudaStream_t stream);
int launch_etc(const EncodeParams& p, cudaStream_t stream);
bool bc1_color_is_exact(uint32_t quality);
bool etc1_is_exact(uint32_t quality);

} // namespace cfx
