// Host side of the ASTC encoder: the per-device cache of the footprint tables (block modes, infill/decimation,
// quantisation, partitions -- built from the ASTC specification by astc_tables.hpp / astc3_tables.hpp once per
// footprint, resident in global memory) and the launcher that hands a surface to the two-phase kernel of astc3.cu.
//
// Replaces the per-thread astcenc_context cache of AstcContextManager (lib/src/AstcConverter.cpp:39-101).
// (Two earlier kernels -- lane = candidate and an exhaustive warp-cooperative search -- live on as developer
// cross-checks under tools/legacy/; they are not part of the library.)
#include "astc3_tables.hpp"
#include "astc_core.cuh"
#include "common.cuh"
#include "kernels.h"

#include <map>
#include <mutex>

namespace cfx {

using namespace astc;

namespace {

struct DeviceTables {
    Ctx ctx;
    Astc3Tab t3;
};

std::mutex g_mutex;
std::map<std::pair<int, int>, DeviceTables> g_tables;   // (device, footprint) -> tables in that device's memory

} // namespace

int launch_astc3(const EncodeParams& p, const Ctx& ctx, const Astc3Tab& t3, cudaStream_t stream);   // astc3.cu

int launch_astc(const EncodeParams& p, cudaStream_t stream)
{
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return -4;
    Ctx ctx;
    Astc3Tab t3;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto key = std::make_pair(device, static_cast<int>(p.block_w*16 + p.block_h));
        auto it = g_tables.find(key);
        if (it == g_tables.end()) {
            Built b = build_tables(static_cast<int>(p.block_w), static_cast<int>(p.block_h));
            const Astc3Tab b3 = build_tables3(b);
            if (b.tab.n_grids > static_cast<uint32_t>(kMaxGrids3)) return -2;
            uint8_t* d = nullptr;
            if (cudaMalloc(&d, b.blob.size()) != cudaSuccess) return -4;
            if (cudaMemcpy(d, b.blob.data(), b.blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return -4; }
            DeviceTables dt;
            dt.ctx.blob = d; dt.ctx.tab = b.tab; dt.t3 = b3;
            it = g_tables.insert(std::make_pair(key, dt)).first;
        }
        ctx = it->second.ctx;
        t3 = it->second.t3;
    }
    return launch_astc3(p, ctx, t3, stream);
}

// cfx_shutdown(): the blobs must not outlive the contexts (a cudaDeviceReset would leave them dangling).
void astc_release_tables()
{
    std::lock_guard<std::mutex> lock(g_mutex);
    for (auto& kv : g_tables) {
        if (cudaSetDevice(kv.first.first) == cudaSuccess) cudaFree(const_cast<uint8_t*>(kv.second.ctx.blob));
    }
    g_tables.clear();
}

} // namespace cfx
