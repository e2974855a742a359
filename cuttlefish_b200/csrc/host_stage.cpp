// See host_stage.h. Plain C++ (no CUDA): compiled by the host compiler only.
#include "host_stage.h"

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define CFX_X86 1
#endif

namespace cfx {

namespace {

// ---- RGBA32F -> RGBA8 ------------------------------------------------------------------------

inline uint32_t unorm8(float v)
{
    // fminf(fmaxf(v,0),1) of the device code: NaN -> 0
    v = v > 0.0f ? v : 0.0f;
    v = v < 1.0f ? v : 1.0f;
    const float t = v*255.0f;
    uint32_t r = static_cast<uint32_t>(t);                 // t >= 0: truncation
    if (t - static_cast<float>(r) >= 0.5f) ++r;            // the difference is exact
    return r;
}

void f32_to_u8_scalar(uint32_t* dst, const float* src, size_t texels)
{
    for (size_t i = 0; i < texels; ++i, src += 4)
        dst[i] = unorm8(src[0]) | (unorm8(src[1]) << 8) | (unorm8(src[2]) << 16) | (unorm8(src[3]) << 24);
}

#ifdef CFX_X86
__attribute__((target("avx2"))) inline __m256i unorm8x8(__m256 v)
{
    const __m256 zero = _mm256_setzero_ps(), one = _mm256_set1_ps(1.0f);
    v = _mm256_max_ps(v, zero);                            // NaN in the first operand -> second operand (0)
    v = _mm256_min_ps(v, one);
    const __m256 t = _mm256_mul_ps(v, _mm256_set1_ps(255.0f));
    __m256i r = _mm256_cvttps_epi32(t);
    const __m256 frac = _mm256_sub_ps(t, _mm256_cvtepi32_ps(r));
    const __m256 up = _mm256_cmp_ps(frac, _mm256_set1_ps(0.5f), _CMP_GE_OQ);
    return _mm256_sub_epi32(r, _mm256_castps_si256(up));   // mask is -1 where the fraction rounds up
}

__attribute__((target("avx2"))) void f32_to_u8_avx2(uint32_t* dst, const float* src, size_t texels)
{
    size_t i = 0;
    const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    for (; i + 8 <= texels; i += 8, src += 32) {
        const __m256i a = unorm8x8(_mm256_loadu_ps(src)), b = unorm8x8(_mm256_loadu_ps(src + 8));
        const __m256i c = unorm8x8(_mm256_loadu_ps(src + 16)), d = unorm8x8(_mm256_loadu_ps(src + 24));
        const __m256i ab = _mm256_packus_epi32(a, b), cd = _mm256_packus_epi32(c, d);
        const __m256i q = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(ab, cd), order);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), q);
    }
    f32_to_u8_scalar(dst + i, src, texels - i);
}
#endif

// ---- RGBA32F -> RGBA16F ----------------------------------------------------------------------

inline uint16_t half_rne(float f)
{
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) return static_cast<uint16_t>(sign | (x > 0x7F800000u ? 0x7E00u | ((x >> 13) & 0x3FFu) : 0x7C00u));
    if (x >= 0x477FF000u) return static_cast<uint16_t>(sign | 0x7C00u);           // rounds to >= 65520 -> inf
    if (x < 0x33000001u) return static_cast<uint16_t>(sign);                       // < 2^-25 (or exactly 2^-25: ties to even 0)
    if (x < 0x38800000u) {                                                          // subnormal half
        const uint32_t shift = 126u - (x >> 23);                                    // 14 .. 24
        const uint32_t m = (x & 0x7FFFFFu) | 0x800000u;
        uint32_t h = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1u);
        if (rem > halfway || (rem == halfway && (h & 1u))) ++h;
        return static_cast<uint16_t>(sign | h);
    }
    uint32_t h = (x - 0x38000000u) >> 13;
    const uint32_t rem = x & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    return static_cast<uint16_t>(sign | h);
}

void f32_to_f16_scalar(uint16_t* dst, const float* src, size_t n)
{
    for (size_t i = 0; i < n; ++i) dst[i] = half_rne(src[i]);
}

#ifdef CFX_X86
__attribute__((target("avx2,f16c"))) void f32_to_f16_f16c(uint16_t* dst, const float* src, size_t n)
{
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i),
            _mm256_cvtps_ph(_mm256_loadu_ps(src + i), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
    f32_to_f16_scalar(dst + i, src + i, n - i);
}
#endif

struct Cpu {
    bool avx2 = false, f16c = false;
    Cpu()
    {
#ifdef CFX_X86
        __builtin_cpu_init();
        avx2 = __builtin_cpu_supports("avx2");
        f16c = avx2 && __builtin_cpu_supports("f16c");
#endif
    }
};
const Cpu g_cpu;

// ---- worker pool -----------------------------------------------------------------------------

class Pool {
public:
    Pool()
    {
        unsigned hw = std::thread::hardware_concurrency();
        if (hw == 0) hw = 4;
        threads_ = hw < 16 ? hw : 16;
        for (unsigned i = 1; i < threads_; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> lock(m_);
            quit_ = true;
        }
        wake_.notify_all();
        for (auto& t : workers_) t.join();
    }
    unsigned threads() const { return threads_; }

    void run(size_t n, const std::function<void(size_t)>& fn)
    {
        if (n == 0) return;
        if (n == 1 || threads_ <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
        std::lock_guard<std::mutex> serial(run_);              // one parallel_for at a time
        {
            std::lock_guard<std::mutex> lock(m_);
            fn_ = &fn; n_ = n; next_.store(0); pending_ = n; ++epoch_;
        }
        wake_.notify_all();
        drain();
        std::unique_lock<std::mutex> lock(m_);
        // also wait for every worker that joined this run to have left drain(): it must not look at n_/fn_ of the next one
        done_.wait(lock, [this] { return pending_ == 0 && inflight_ == 0; });
        fn_ = nullptr;
    }

private:
    void drain(bool worker = false)
    {
        size_t did = 0;
        for (;;) {
            const size_t i = next_.fetch_add(1);
            if (i >= n_) break;
            (*fn_)(i);
            ++did;
        }
        std::lock_guard<std::mutex> lock(m_);
        pending_ -= did;
        if (worker) --inflight_;
        if (pending_ == 0 && inflight_ == 0) done_.notify_all();
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(m_);
                wake_.wait(lock, [&] { return quit_ || (epoch_ != seen && fn_ != nullptr); });
                if (quit_) return;
                seen = epoch_;
                ++inflight_;
            }
            drain(true);
        }
    }

    std::mutex m_, run_;
    std::condition_variable wake_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t n_ = 0, pending_ = 0, inflight_ = 0;
    std::atomic<size_t> next_{0};
    uint64_t epoch_ = 0;
    unsigned threads_ = 1;
    bool quit_ = false;
};

Pool& pool()
{
    static Pool* p = new Pool();      // leaked on purpose: worker threads must not be joined from a static destructor
    return *p;                        // of a dlclose()d library
}

} // namespace

void stage_row(StageOp op, void* dst, const void* src, size_t texels, size_t src_texel_bytes)
{
    switch (op) {
        case STAGE_F32_TO_U8:
#ifdef CFX_X86
            if (g_cpu.avx2) { f32_to_u8_avx2(static_cast<uint32_t*>(dst), static_cast<const float*>(src), texels); return; }
#endif
            f32_to_u8_scalar(static_cast<uint32_t*>(dst), static_cast<const float*>(src), texels);
            return;
        case STAGE_F32_TO_F16:
#ifdef CFX_X86
            if (g_cpu.f16c) { f32_to_f16_f16c(static_cast<uint16_t*>(dst), static_cast<const float*>(src), texels*4); return; }
#endif
            f32_to_f16_scalar(static_cast<uint16_t*>(dst), static_cast<const float*>(src), texels*4);
            return;
        default:
            std::memcpy(dst, src, texels*src_texel_bytes);
    }
}

void parallel_for(size_t n, const std::function<void(size_t)>& fn) { pool().run(n, fn); }
unsigned stage_threads() { return pool().threads(); }

} // namespace cfx
