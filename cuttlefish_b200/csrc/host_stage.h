// Host side of the upload pipeline: a small worker pool and the row transforms that move a caller's PAGEABLE source
// surface into the library's pinned staging slots. A pageable buffer cannot be DMA'd directly, so every byte of it has to
// pass through the CPU once anyway; that one pass also narrows the texels to what the encoder's load stage would make of
// them, so that 4 (or 8) instead of 16 bytes per texel cross PCIe:
//   RGBA32F -> RGBA8   u8 = round(clamp(v,0,1)*255), half away from zero -- toColorBlock(), lib/src/S3tcConverter.cpp:97-111;
//                      identical to common.cuh f32_to_unorm8(), used for the formats whose kernels only ever see that view
//   RGBA32F -> RGBA16F round to nearest even -- packHalfFloatBlockHardware(), lib/src/HalfFloat.h:96-134 (F16C imm 0)
//   anything else      a plain row copy
// Pinned (cudaHostAlloc / cudaHostRegister) sources skip all of this and are DMA'd in place.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace cfx {

enum StageOp { STAGE_COPY = 0, STAGE_F32_TO_U8 = 1, STAGE_F32_TO_F16 = 2 };

// dst/src rows of `texels` RGBA texels; src_texel_bytes only matters for STAGE_COPY.
void stage_row(StageOp op, void* dst, const void* src, size_t texels, size_t src_texel_bytes);

// Runs fn(i) for i in [0, n) on the pool's threads and the calling thread; returns when all are done.
// The pool is created on first use (min(hardware threads, 16) - 1 workers) and lives until process exit.
void parallel_for(size_t n, const std::function<void(size_t)>& fn);
unsigned stage_threads();

} // namespace cfx
