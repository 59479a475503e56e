// afb_host.cpp - host-side stage of the upload path (plain C++: nvcc hands this file to the host compiler as it is).
//
// The reference keeps its label maps as int64 on the host (datasets hand `batch['label']` as torch.long, run_dl.py:261) and
// moves them to the GPU as they are: 8 bytes per voxel over PCIe for values < num_classes <= 256.  At 64 volumes of 128^3 that
// is 1.07 of the 1.61 GB of a step's upload, and the end-to-end path is PCIe-bound (55 GB/s).  afb_host_narrow_labels packs
// an integer label map to uint8 on the host cores - a streaming pass that several threads run at memory speed - so that one
// byte per voxel crosses the link; the device-side one-hot expansion takes uint8 maps as they are (afb_onehot_expand).
// Values outside [0, 255] are reported, never wrapped silently.
#include <immintrin.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../include/afb200.h"

namespace {

// int64 -> uint8, 16 elements per iteration (SSE2 only: the library is built on one machine and runs on another).
// `bad` collects the bits above the low byte of every element (negative values have them set).
void narrow_i64(const int64_t* src, uint8_t* dst, int64_t n, uint64_t* bad) {
    int64_t i = 0;
    __m128i acc = _mm_setzero_si128();
    if ((((uintptr_t)src) & 15u) == 0) {
        for (; i + 16 <= n; i += 16) {
            const __m128i* p = reinterpret_cast<const __m128i*>(src + i);
            const __m128i a0 = _mm_load_si128(p + 0), a1 = _mm_load_si128(p + 1), a2 = _mm_load_si128(p + 2), a3 = _mm_load_si128(p + 3);
            const __m128i a4 = _mm_load_si128(p + 4), a5 = _mm_load_si128(p + 5), a6 = _mm_load_si128(p + 6), a7 = _mm_load_si128(p + 7);
            acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(_mm_or_si128(a0, a1), _mm_or_si128(a2, a3)),
                                                 _mm_or_si128(_mm_or_si128(a4, a5), _mm_or_si128(a6, a7))));
            // low 32 bits of each int64: lanes 0 and 2 of every vector
            const __m128i l0 = _mm_castps_si128(_mm_shuffle_ps(_mm_castsi128_ps(a0), _mm_castsi128_ps(a1), _MM_SHUFFLE(2, 0, 2, 0)));
            const __m128i l1 = _mm_castps_si128(_mm_shuffle_ps(_mm_castsi128_ps(a2), _mm_castsi128_ps(a3), _MM_SHUFFLE(2, 0, 2, 0)));
            const __m128i l2 = _mm_castps_si128(_mm_shuffle_ps(_mm_castsi128_ps(a4), _mm_castsi128_ps(a5), _MM_SHUFFLE(2, 0, 2, 0)));
            const __m128i l3 = _mm_castps_si128(_mm_shuffle_ps(_mm_castsi128_ps(a6), _mm_castsi128_ps(a7), _MM_SHUFFLE(2, 0, 2, 0)));
            // in-range values survive the saturating packs unchanged; out-of-range ones are caught through `acc`
            const __m128i bytes = _mm_packus_epi16(_mm_packs_epi32(l0, l1), _mm_packs_epi32(l2, l3));
            _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), bytes);
        }
    }
    alignas(16) uint64_t lanes[2];
    _mm_store_si128(reinterpret_cast<__m128i*>(lanes), acc);
    uint64_t b = (lanes[0] | lanes[1]) & ~0xFFull;
    for (; i < n; ++i) {
        b |= ((uint64_t)src[i]) & ~0xFFull;
        dst[i] = (uint8_t)src[i];
    }
    *bad = b;
}

// The same with 32-byte vectors where the CPU has AVX2 (checked at run time; the build machine's ISA is not assumed).
__attribute__((target("avx2"))) void narrow_i64_avx2(const int64_t* src, uint8_t* dst, int64_t n, uint64_t* bad) {
    int64_t i = 0;
    __m256i acc = _mm256_setzero_si256();
    // software prefetch, bytes ahead (0 = none): one core streams ~12 GB/s on the hardware prefetcher alone, ~17 GB/s with it
    // (64 volumes, one thread: 85 ms -> 61 ms; 16 threads 10.8 -> 7.3-9.7 ms on the B200 box's host)
    static const int pf = [] { const char* e = getenv("AFB_HOST_PREFETCH"); return e ? atoi(e) : 2048; }();
    for (; i + 16 <= n; i += 16) {
        const __m256i* p = reinterpret_cast<const __m256i*>(src + i);
        if (pf) {
            _mm_prefetch(reinterpret_cast<const char*>(p) + pf, _MM_HINT_T0);
            _mm_prefetch(reinterpret_cast<const char*>(p) + pf + 64, _MM_HINT_T0);
        }
        const __m256i a0 = _mm256_loadu_si256(p + 0), a1 = _mm256_loadu_si256(p + 1), a2 = _mm256_loadu_si256(p + 2), a3 = _mm256_loadu_si256(p + 3);
        acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(a0, a1), _mm256_or_si256(a2, a3)));
        // low dwords of the four int64 of a vector -> its low 128 bits
        const __m256i idx = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
        const __m128i l0 = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a0, idx));
        const __m128i l1 = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a1, idx));
        const __m128i l2 = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a2, idx));
        const __m128i l3 = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a3, idx));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_packus_epi16(_mm_packs_epi32(l0, l1), _mm_packs_epi32(l2, l3)));
    }
    alignas(32) uint64_t lanes[4];
    _mm256_store_si256(reinterpret_cast<__m256i*>(lanes), acc);
    uint64_t b = (lanes[0] | lanes[1] | lanes[2] | lanes[3]) & ~0xFFull;
    for (; i < n; ++i) {
        b |= ((uint64_t)src[i]) & ~0xFFull;
        dst[i] = (uint8_t)src[i];
    }
    *bad = b;
}

template <typename T>
void narrow_small(const T* src, uint8_t* dst, int64_t n, uint64_t* bad) {
    uint64_t b = 0;
    for (int64_t i = 0; i < n; ++i) {
        b |= ((uint64_t)(int64_t)src[i]) & ~0xFFull;
        dst[i] = (uint8_t)src[i];
    }
    *bad = b;
}

bool have_avx2() {
    static const bool yes = __builtin_cpu_supports("avx2") && getenv("AFB_HOST_NO_AVX2") == nullptr;
    return yes;
}

void narrow_range(const void* src, int dtype, uint8_t* dst, int64_t lo, int64_t hi, uint64_t* bad) {
    switch (dtype) {
        case AFB_I64:
            if (have_avx2()) narrow_i64_avx2((const int64_t*)src + lo, dst + lo, hi - lo, bad);
            else narrow_i64((const int64_t*)src + lo, dst + lo, hi - lo, bad);
            break;
        case AFB_I32: narrow_small((const int32_t*)src + lo, dst + lo, hi - lo, bad); break;
        case AFB_I16: narrow_small((const int16_t*)src + lo, dst + lo, hi - lo, bad); break;
        default: *bad = 0; break;
    }
}

}  // namespace

extern "C" int afb_host_narrow_labels(const void* src, int src_dtype, int64_t n, uint8_t* dst, int n_threads, int* out_of_range) {
    if (!src || !dst || n < 0) return AFB_EINVAL;
    if (src_dtype != AFB_I64 && src_dtype != AFB_I32 && src_dtype != AFB_I16) return AFB_EDTYPE;
    if (out_of_range) *out_of_range = 0;
    if (n == 0) return AFB_OK;
    int nt = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    const int64_t min_per_thread = 1 << 16;
    if ((int64_t)nt * min_per_thread > n) nt = (int)((n + min_per_thread - 1) / min_per_thread);
    std::vector<uint64_t> bad((size_t)nt, 0);
    // equal 64-element-aligned shares (the int64 kernel wants 16-byte aligned loads: src itself is checked inside)
    const int64_t share = (((n + nt - 1) / nt) + 63) & ~(int64_t)63;
    std::vector<std::thread> pool;
    pool.reserve((size_t)nt);
    for (int t = 1; t < nt; ++t) {
        const int64_t lo = t * share, hi = lo + share < n ? lo + share : n;
        if (lo >= n) break;
        pool.emplace_back(narrow_range, src, src_dtype, dst, lo, hi, &bad[(size_t)t]);
    }
    narrow_range(src, src_dtype, dst, 0, share < n ? share : n, &bad[0]);
    for (auto& th : pool) th.join();
    uint64_t b = 0;
    for (uint64_t v : bad) b |= v;
    if (out_of_range) *out_of_range = b != 0 ? 1 : 0;
    return AFB_OK;
}
