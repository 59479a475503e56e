// afb_minmask.cu - the reference's min-shift bookkeeping with HALF the HBM traffic in the backward.
//
// MinBackward of `volume.min()` (utils/nifti_utils.py:200) spreads d(min) evenly over ALL voxels equal to the minimum,
// so the backward needs the predicate (vol == min) for every voxel: afb_min_grad_fill re-reads the whole volume for it
// (4 B read + 4 B write per voxel).  The forward's min pass already streams the volume once; this variant makes it
// also leave behind a 1-bit-per-voxel record that the backward can use instead of the volume:
//
//   forward  : per CHUNK of 512 voxels (one warp, 16 per lane held in registers) compute the chunk minimum by
//              shuffles, then store chunk_min[chunk] and the bitmask (v == chunk_min): +4 % write traffic.
//   backward : a voxel equals the GLOBAL minimum iff its chunk's minimum equals the global minimum AND its bit is set
//              (chunks with a larger minimum contain no global minimum at all).  The fill therefore reads
//              4 B + 64 B per chunk instead of 2 KiB: 4.13 B/voxel of traffic instead of 8.
//
// Exact, no speculation.  fp32 volumes only (the only dtype for which the reference's autograd produces dVolume).
#include "afb_device.cuh"

namespace afb {

constexpr int MM_THREADS = 256;
constexpr int MM_LOADS = 4;                                   // 16-byte loads per lane per chunk
constexpr int MM_CHUNK = 32 * MM_LOADS * 4;                   // 512 voxels per WARP-chunk: the chunk minimum is a pure
                                                              // shuffle reduction, no CTA barrier in the streaming loop
constexpr int MM_BLOCKS = 148 * 8;

struct MinCountF { float m, n; };

__device__ __forceinline__ void mm_merge(float& m, float& n, float m2, float n2) {
    if (m2 < m) { m = m2; n = n2; }
    else if (m2 == m) { n += n2; }
}

// mask layout: [n_chunks] float chunk minima, then [n_chunks][32] uint16 (bit 4*j+e of lane l <-> element
// chunk*512 + j*128 + l*4 + e)
// CTA reduction of each lane's (min, multiplicity), then the last CTA to arrive reduces the per-CTA partials into out[0..1]
// and re-arms the counter.  Called by all threads of the CTA.
template <int NT = MM_THREADS>
__device__ __forceinline__ void mm_finish(float gm, float gc, MinCountF* __restrict__ partial, unsigned* __restrict__ counter,
                                          float* __restrict__ out) {
    __shared__ float sm[NT / 32], sn[NT / 32];
    __shared__ bool last;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, gm, o), n2 = __shfl_xor_sync(0xffffffffu, gc, o);
        mm_merge(gm, gc, m2, n2);
    }
    if (lane == 0) { sm[w] = gm; sn[w] = gc; }
    __syncthreads();
    if (t == 0) {
        for (int k = 1; k < NT / 32; ++k) mm_merge(gm, gc, sm[k], sn[k]);
        partial[blockIdx.x].m = gm;
        partial[blockIdx.x].n = gc;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    gm = INFINITY; gc = 0.0f;
    for (int i = t; i < (int)gridDim.x; i += NT) {
        const float2 pc = __ldcg(reinterpret_cast<const float2*>(partial) + i);
        mm_merge(gm, gc, pc.x, pc.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, gm, o), n2 = __shfl_xor_sync(0xffffffffu, gc, o);
        mm_merge(gm, gc, m2, n2);
    }
    __syncthreads();
    if (lane == 0) { sm[w] = gm; sn[w] = gc; }
    __syncthreads();
    if (t == 0) {
        for (int k = 1; k < NT / 32; ++k) mm_merge(gm, gc, sm[k], sn[k]);
        out[0] = gm;
        out[1] = gc;
        *counter = 0u;
    }
}

__device__ __forceinline__ void mm_load_chunk(const float* __restrict__ data, long long n, long long ch, int lane, float4* v) {
    const long long base = ch * MM_CHUNK;
#pragma unroll
    for (int j = 0; j < MM_LOADS; ++j) {
        const long long e = base + j * 128 + lane * 4;
        if (e + 3 < n) {
            v[j] = __ldcs(reinterpret_cast<const float4*>(data + e));
        } else {
            v[j].x = e + 0 < n ? data[e + 0] : INFINITY;
            v[j].y = e + 1 < n ? data[e + 1] : INFINITY;
            v[j].z = e + 2 < n ? data[e + 2] : INFINITY;
            v[j].w = INFINITY;
        }
    }
}

// chunk minimum + "== minimum" bits of the 16 values this lane holds; writes the record, merges into the lane's running pair
__device__ __forceinline__ void mm_chunk_record(const float4* v, long long ch, int lane, float* __restrict__ chunk_min,
                                                unsigned short* __restrict__ bits, float& gm, float& gc) {
    float c = INFINITY;
#pragma unroll
    for (int j = 0; j < MM_LOADS; ++j) c = fminf(fminf(fminf(c, v[j].x), fminf(v[j].y, v[j].z)), v[j].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = fminf(c, __shfl_xor_sync(0xffffffffu, c, o));
    unsigned b = 0u;
#pragma unroll
    for (int j = 0; j < MM_LOADS; ++j) {
        const unsigned q = (v[j].x == c ? 1u : 0u) | (v[j].y == c ? 2u : 0u) | (v[j].z == c ? 4u : 0u) | (v[j].w == c ? 8u : 0u);
        b |= q << (4 * j);
    }
    bits[ch * 32 + lane] = (unsigned short)b;
    if (lane == 0) chunk_min[ch] = c;
    // chunk minimum with this lane's share of its multiplicity (lanes that do not attain c contribute 0)
    mm_merge(gm, gc, c, (float)__popc(b));
}

// A register double buffer (next chunk requested before the current one is reduced) was measured SLOWER on the
// B200 (0.740 vs 0.719 ms per 4.3 GB: 63 registers cost two resident CTAs per SM), as was sizing the grid to exactly
// one resident wave (0.731): the plain loop below at 148 x 8 CTAs reads at 6.1-6.2 TB/s.
__global__ void __launch_bounds__(MM_THREADS)
volume_min_mask_kernel(const float* __restrict__ data, long long n, long long n_chunks, float* __restrict__ chunk_min,
                       unsigned short* __restrict__ bits, MinCountF* __restrict__ partial, unsigned* __restrict__ counter,
                       float* __restrict__ out) {
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const long long warps = (long long)gridDim.x * (MM_THREADS / 32);
    float gm = INFINITY, gc = 0.0f;                           // this lane's running (min, multiplicity)
    for (long long ch = (long long)blockIdx.x * (MM_THREADS / 32) + w; ch < n_chunks; ch += warps) {
        float4 v[MM_LOADS];
        mm_load_chunk(data, n, ch, lane, v);
        mm_chunk_record(v, ch, lane, chunk_min, bits, gm, gc);
    }
    mm_finish(gm, gc, partial, counter, out);
}

// d_vol = (vol == min) ? d_pad / count : 0 from the chunk record (never touches the volume).
// Loop-free: one warp per group of G consecutive chunks, the group's records (chunk minima + bit words, independent
// loads) fetched up front, then G x 2 KiB of 16-byte stores.  Measured on the B200 over 4.3 GB of dVolume: the earlier
// persistent grid-stride loop (record load -> stores chained in every iteration) 0.790 ms; loop-free G = 8 / 4 / 2 / 1:
// 0.688 / 0.690 / 0.670 / 0.665 ms (6.7 TB/s).  torch's plain zero_() of the same tensor takes 0.578 ms and so does this
// kernel with the record loads compiled out (0.585 ms): the remaining 0.08 ms is what the 3 % of record READS cost
// when they are interleaved into a pure HBM write stream.
constexpr int MM_GROUP = 1;

template <int G>
__global__ void __launch_bounds__(MM_THREADS)
min_grad_fill_mask_kernel(const float* __restrict__ chunk_min, const unsigned short* __restrict__ bits, long long n,
                          long long n_chunks, const float* __restrict__ min_count, const float* __restrict__ d_pad,
                          float* __restrict__ d_vol) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ch0 = ((long long)blockIdx.x * (MM_THREADS / 32) + w) * G;
    if (ch0 >= n_chunks) return;
    float cm[G];
    unsigned braw[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const bool ok = ch0 + g < n_chunks;
        cm[g] = ok ? __ldg(chunk_min + ch0 + g) : INFINITY;
        braw[g] = ok ? (unsigned)__ldg(bits + (ch0 + g) * 32 + lane) : 0u;
    }
    const float m = __ldg(min_count), share = __ldg(d_pad) / __ldg(min_count + 1);
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (ch0 + g >= n_chunks) break;
        const unsigned b = cm[g] == m ? braw[g] : 0u;
        const long long base = (ch0 + g) * MM_CHUNK;
#pragma unroll
        for (int j = 0; j < MM_LOADS; ++j) {
            const long long e = base + j * 128 + lane * 4;
            const unsigned q = (b >> (4 * j)) & 15u;
            const float4 o = make_float4((q & 1u) ? share : 0.0f, (q & 2u) ? share : 0.0f, (q & 4u) ? share : 0.0f, (q & 8u) ? share : 0.0f);
            if (e + 3 < n) {
                *reinterpret_cast<float4*>(d_vol + e) = o;
            } else {
                if (e + 0 < n) d_vol[e + 0] = o.x;
                if (e + 1 < n) d_vol[e + 1] = o.y;
                if (e + 2 < n) d_vol[e + 2] = o.z;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 16-bit storage (bf16 / fp16 volumes, BASELINE configs[4]): same record, lane-major for 8-element vectors.
// One chunk = 512 elements = two 16-byte loads per lane; bit 8j+q of lane l <-> element chunk*512 + j*256 + l*8 + q.
// The fill writes fp32 dVolume (the backward accumulates in fp32) with one 32-byte store per lane (st.global.v8.f32).
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float half_to_f(T v);
template <> __device__ __forceinline__ float half_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float half_to_f<__half>(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(MM_THREADS)
volume_min_mask_half_kernel(const T* __restrict__ data, long long n, long long n_chunks, float* __restrict__ chunk_min,
                            unsigned short* __restrict__ bits, MinCountF* __restrict__ partial, unsigned* __restrict__ counter,
                            float* __restrict__ out) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (MM_THREADS / 32);
    float gm = INFINITY, gc = 0.0f;
    for (long long ch = (long long)blockIdx.x * (MM_THREADS / 32) + w; ch < n_chunks; ch += warps) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const long long e = ch * MM_CHUNK + j * 256 + lane * 8;
            if (e + 7 < n) {
                const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(data + e));
                const T* t = reinterpret_cast<const T*>(&raw);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[8 * j + q] = half_to_f<T>(t[q]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[8 * j + q] = e + q < n ? half_to_f<T>(data[e + q]) : INFINITY;
            }
        }
        float c = INFINITY;
#pragma unroll
        for (int q = 0; q < 16; ++q) c = fminf(c, v[q]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = fminf(c, __shfl_xor_sync(0xffffffffu, c, o));
        unsigned b = 0u;
#pragma unroll
        for (int q = 0; q < 16; ++q) b |= (v[q] == c ? 1u : 0u) << q;
        bits[ch * 32 + lane] = (unsigned short)b;
        if (lane == 0) chunk_min[ch] = c;
        mm_merge(gm, gc, c, (float)__popc(b));
    }
    mm_finish(gm, gc, partial, counter, out);
}

__global__ void __launch_bounds__(MM_THREADS)
min_grad_fill_mask_half_kernel(const float* __restrict__ chunk_min, const unsigned short* __restrict__ bits, long long n,
                               long long n_chunks, const float* __restrict__ min_count, const float* __restrict__ d_pad,
                               float* __restrict__ d_vol) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ch = (long long)blockIdx.x * (MM_THREADS / 32) + w;
    if (ch >= n_chunks) return;
    const float cm = __ldg(chunk_min + ch);
    const unsigned braw = (unsigned)__ldg(bits + ch * 32 + lane);
    const float m = __ldg(min_count), share = __ldg(d_pad) / __ldg(min_count + 1);
    const unsigned b = cm == m ? braw : 0u;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const long long e = ch * MM_CHUNK + j * 256 + lane * 8;
        float o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = ((b >> (8 * j + q)) & 1u) ? share : 0.0f;
        if (e + 7 < n) {
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d_vol + e), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                         "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (e + q < n) d_vol[e + q] = o[q];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One-hot materialisation of running/run_dl.py:261-264 fused with the min record.
//   label map (integer, one value per voxel) -> int64 one-hot [voxel][C] and/or fp32 one-hot [voxel][C]
//   (channels-last in memory, i.e. exactly the strides of `one_hot(lab).permute(0,4,1,2,3)` and of its `.float()`).
// While the fp32 values are in registers the kernel also writes their chunk record (chunk minimum + "== minimum" bits,
// same layout as volume_min_mask_kernel), so the separate 4 B/voxel min pass over the soft-label volume disappears for
// volumes that are produced here.  Works on a RANGE of a larger tensor (labels / outputs point at the range, chunk0 is
// the range's first chunk in the whole tensor's record) so that a host batch can be expanded chunk by chunk while the next chunk is still in flight
// over PCIe; afb_min_count_from_mask then reduces the whole record to (min, multiplicity).
// Labels outside [0, C) give an all-zero voxel (torch's one_hot raises instead).
// ------------------------------------------------------------------------------------------------
template <typename TL>
__global__ void __launch_bounds__(MM_THREADS)
onehot_expand_kernel(const TL* __restrict__ labels, long long n_elem, int C, long long chunk0,
                     long long* __restrict__ onehot, float* __restrict__ soft, float* __restrict__ chunk_min,
                     unsigned short* __restrict__ bits) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long lc = (long long)blockIdx.x * (MM_THREADS / 32) + w;       // chunk within the range
    const long long base = lc * MM_CHUNK;                                      // first element of the chunk within the range
    if (base >= n_elem) return;
    const bool pow2 = (C & (C - 1)) == 0;
    const int shift = 31 - __clz(C);
    auto value_at = [&](long long e) -> int {                                  // one-hot value of range element e
        const long long vox = pow2 ? (e >> shift) : (e / C);
        const int ch = (int)(pow2 ? (e & (C - 1)) : (e - vox * C));
        return (long long)__ldg(labels + vox) == (long long)ch ? 1 : 0;
    };
    float4 v[MM_LOADS];
#pragma unroll
    for (int j = 0; j < MM_LOADS; ++j) {
        const long long e = base + j * 128 + lane * 4;
        v[j].x = e + 0 < n_elem ? (float)value_at(e + 0) : INFINITY;
        v[j].y = e + 1 < n_elem ? (float)value_at(e + 1) : INFINITY;
        v[j].z = e + 2 < n_elem ? (float)value_at(e + 2) : INFINITY;
        v[j].w = e + 3 < n_elem ? (float)value_at(e + 3) : INFINITY;
    }
    if (soft) {
#pragma unroll
        for (int j = 0; j < MM_LOADS; ++j) {
            const long long e = base + j * 128 + lane * 4;
            if (e + 3 < n_elem) {
                *reinterpret_cast<float4*>(soft + e) = v[j];
            } else {
                if (e + 0 < n_elem) soft[e + 0] = v[j].x;
                if (e + 1 < n_elem) soft[e + 1] = v[j].y;
                if (e + 2 < n_elem) soft[e + 2] = v[j].z;
            }
        }
    }
    if (chunk_min) {
        float c = INFINITY;
#pragma unroll
        for (int j = 0; j < MM_LOADS; ++j) c = fminf(fminf(fminf(c, v[j].x), fminf(v[j].y, v[j].z)), v[j].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = fminf(c, __shfl_xor_sync(0xffffffffu, c, o));
        unsigned b = 0u;
#pragma unroll
        for (int j = 0; j < MM_LOADS; ++j) {
            const unsigned q = (v[j].x == c ? 1u : 0u) | (v[j].y == c ? 2u : 0u) | (v[j].z == c ? 4u : 0u) | (v[j].w == c ? 8u : 0u);
            b |= q << (4 * j);
        }
        const long long gch = chunk0 + lc;
        bits[gch * 32 + lane] = (unsigned short)b;
        if (lane == 0) chunk_min[gch] = c;
    }
    if (onehot) {
        // int64 output: 16-byte pieces laid out so that every store instruction of the warp covers 512 contiguous bytes
#pragma unroll
        for (int k = 0; k < 2 * MM_LOADS; ++k) {
            const long long e = base + ((long long)k * 32 + lane) * 2;
            if (e + 1 < n_elem) {
                longlong2 o;
                o.x = value_at(e); o.y = value_at(e + 1);
                *reinterpret_cast<longlong2*>(onehot + e) = o;
            } else if (e < n_elem) {
                onehot[e] = value_at(e);
            }
        }
    }
}

// (min, multiplicity) of a whole tensor from its chunk record alone (4.1 % of the tensor's bytes)
__global__ void __launch_bounds__(MM_THREADS)
mask_reduce_kernel(const float* __restrict__ chunk_min, const unsigned short* __restrict__ bits, long long n_chunks,
                   MinCountF* __restrict__ partial, unsigned* __restrict__ counter, float* __restrict__ out) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (MM_THREADS / 32);
    float gm = INFINITY, gc = 0.0f;
    for (long long ch = (long long)blockIdx.x * (MM_THREADS / 32) + w; ch < n_chunks; ch += warps)
        mm_merge(gm, gc, __ldg(chunk_min + ch), (float)__popc((unsigned)__ldg(bits + ch * 32 + lane)));
    mm_finish(gm, gc, partial, counter, out);
}

}  // namespace afb

using namespace afb;

static long long mm_chunks(long long n) { return (n + MM_CHUNK - 1) / MM_CHUNK; }

extern "C" int64_t afb_min_mask_bytes(int64_t n_elements) {
    const long long c = mm_chunks(n_elements);
    return (int64_t)(((c * (long long)sizeof(float) + 15) / 16) * 16 + c * 32 * (long long)sizeof(unsigned short));
}

extern "C" int afb_volume_min_mask(const float* data, int64_t n, float* out_min_count, void* mask, void* workspace, void* stream) {
    if (!data || !out_min_count || !mask || !workspace || n <= 0) return AFB_EINVAL;
    if (((uintptr_t)data & 15u) || ((uintptr_t)mask & 15u)) return AFB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const long long c = mm_chunks(n);
    float* chunk_min = (float*)mask;
    unsigned short* bits = (unsigned short*)((char*)mask + ((c * (long long)sizeof(float) + 15) / 16) * 16);
    unsigned* counter = (unsigned*)workspace;
    MinCountF* partial = (MinCountF*)((char*)workspace + 16);
    cudaError_t e = cudaMemsetAsync(counter, 0, 16, st);
    if (e != cudaSuccess) return (int)e;
    const long long wantb = (c + MM_THREADS / 32 - 1) / (MM_THREADS / 32);
    const int blocks = (int)(wantb < MM_BLOCKS ? wantb : MM_BLOCKS);
    volume_min_mask_kernel<<<blocks, MM_THREADS, 0, st>>>(data, n, c, chunk_min, bits, partial, counter, out_min_count);
    return (int)cudaGetLastError();
}

extern "C" int afb_min_grad_fill_mask(const void* mask, int64_t n, const float* min_count, const float* d_pad, float* d_vol, void* stream) {
    if (!mask || !min_count || !d_pad || !d_vol || n <= 0) return AFB_EINVAL;
    if (((uintptr_t)d_vol & 15u) || ((uintptr_t)mask & 15u)) return AFB_EINVAL;
    const long long c = mm_chunks(n);
    const float* chunk_min = (const float*)mask;
    const unsigned short* bits = (const unsigned short*)((const char*)mask + ((c * (long long)sizeof(float) + 15) / 16) * 16);
    const long long per_cta = (long long)(MM_THREADS / 32) * MM_GROUP;
    const long long blocks = (c + per_cta - 1) / per_cta;
    if (blocks > 2147483647ll) return AFB_EUNSUPPORTED;
    min_grad_fill_mask_kernel<MM_GROUP><<<(unsigned)blocks, MM_THREADS, 0, (cudaStream_t)stream>>>(chunk_min, bits, n, c, min_count,
                                                                                                  d_pad, d_vol);
    return (int)cudaGetLastError();
}

extern "C" int afb_onehot_expand(const void* labels, int label_dtype, int64_t n_voxels, int num_classes, int64_t* onehot_i64,
                                 float* soft_f32, void* mask, int64_t total_elements, int64_t elem_offset, void* stream) {
    if (!labels || n_voxels <= 0 || num_classes <= 0 || (!onehot_i64 && !soft_f32)) return AFB_EINVAL;
    if (((uintptr_t)onehot_i64 & 15u) || ((uintptr_t)soft_f32 & 15u) || ((uintptr_t)mask & 15u)) return AFB_EINVAL;
    const long long n_elem = (long long)n_voxels * num_classes;
    float* chunk_min = nullptr;
    unsigned short* bits = nullptr;
    long long chunk0 = 0;
    if (mask) {
        if (!soft_f32 || elem_offset < 0 || elem_offset % MM_CHUNK != 0 || elem_offset + n_elem > total_elements) return AFB_EINVAL;
        // a range that does not end the tensor must end on a chunk boundary, or its last chunk's record would be partial
        if (elem_offset + n_elem != total_elements && n_elem % MM_CHUNK != 0) return AFB_EINVAL;
        const long long c = mm_chunks(total_elements);
        chunk_min = (float*)mask;
        bits = (unsigned short*)((char*)mask + ((c * (long long)sizeof(float) + 15) / 16) * 16);
        chunk0 = elem_offset / MM_CHUNK;
    }
    const long long chunks = mm_chunks(n_elem);
    const long long blocks = (chunks + MM_THREADS / 32 - 1) / (MM_THREADS / 32);
    if (blocks > 2147483647ll) return AFB_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
#define AFB_EXP(TL) onehot_expand_kernel<TL><<<(unsigned)blocks, MM_THREADS, 0, st>>>((const TL*)labels, n_elem, num_classes, \
        chunk0, (long long*)onehot_i64, soft_f32, chunk_min, bits)
    switch (label_dtype) {
        case AFB_I64: AFB_EXP(int64_t); break;
        case AFB_I32: AFB_EXP(int32_t); break;
        case AFB_I16: AFB_EXP(int16_t); break;
        case AFB_U8: AFB_EXP(uint8_t); break;
        default: return AFB_EDTYPE;
    }
#undef AFB_EXP
    return (int)cudaGetLastError();
}

extern "C" int afb_min_count_from_mask(const void* mask, int64_t n, float* out_min_count, void* workspace, void* stream) {
    if (!mask || !out_min_count || !workspace || n <= 0 || ((uintptr_t)mask & 15u)) return AFB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const long long c = mm_chunks(n);
    const float* chunk_min = (const float*)mask;
    const unsigned short* bits = (const unsigned short*)((const char*)mask + ((c * (long long)sizeof(float) + 15) / 16) * 16);
    unsigned* counter = (unsigned*)workspace;
    MinCountF* partial = (MinCountF*)((char*)workspace + 16);
    cudaError_t e = cudaMemsetAsync(counter, 0, 16, st);
    if (e != cudaSuccess) return (int)e;
    const long long wantb = (c + MM_THREADS / 32 - 1) / (MM_THREADS / 32);
    const int blocks = (int)(wantb < MM_BLOCKS ? wantb : MM_BLOCKS);
    mask_reduce_kernel<<<blocks, MM_THREADS, 0, st>>>(chunk_min, bits, c, partial, counter, out_min_count);
    return (int)cudaGetLastError();
}

// bf16 / fp16 volumes: same record size and meaning, 8-element lane vectors (see volume_min_mask_half_kernel)
extern "C" int afb_volume_min_mask_half(const void* data, int dtype, int64_t n, float* out_min_count, void* mask, void* workspace,
                                        void* stream) {
    if (!data || !out_min_count || !mask || !workspace || n <= 0) return AFB_EINVAL;
    if (((uintptr_t)data & 15u) || ((uintptr_t)mask & 15u)) return AFB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const long long c = mm_chunks(n);
    float* chunk_min = (float*)mask;
    unsigned short* bits = (unsigned short*)((char*)mask + ((c * (long long)sizeof(float) + 15) / 16) * 16);
    unsigned* counter = (unsigned*)workspace;
    MinCountF* partial = (MinCountF*)((char*)workspace + 16);
    cudaError_t e = cudaMemsetAsync(counter, 0, 16, st);
    if (e != cudaSuccess) return (int)e;
    const long long wantb = (c + MM_THREADS / 32 - 1) / (MM_THREADS / 32);
    const int blocks = (int)(wantb < MM_BLOCKS ? wantb : MM_BLOCKS);
    if (dtype == AFB_BF16)
        volume_min_mask_half_kernel<__nv_bfloat16><<<blocks, MM_THREADS, 0, st>>>((const __nv_bfloat16*)data, n, c, chunk_min, bits, partial,
                                                                                  counter, out_min_count);
    else if (dtype == AFB_F16)
        volume_min_mask_half_kernel<__half><<<blocks, MM_THREADS, 0, st>>>((const __half*)data, n, c, chunk_min, bits, partial, counter,
                                                                           out_min_count);
    else
        return AFB_EDTYPE;
    return (int)cudaGetLastError();
}

extern "C" int afb_min_grad_fill_mask_half(const void* mask, int64_t n, const float* min_count, const float* d_pad, float* d_vol,
                                           void* stream) {
    if (!mask || !min_count || !d_pad || !d_vol || n <= 0) return AFB_EINVAL;
    if (((uintptr_t)d_vol & 31u) || ((uintptr_t)mask & 15u)) return AFB_EINVAL;
    const long long c = mm_chunks(n);
    const float* chunk_min = (const float*)mask;
    const unsigned short* bits = (const unsigned short*)((const char*)mask + ((c * (long long)sizeof(float) + 15) / 16) * 16);
    const long long blocks = (c + MM_THREADS / 32 - 1) / (MM_THREADS / 32);
    if (blocks > 2147483647ll) return AFB_EUNSUPPORTED;
    min_grad_fill_mask_half_kernel<<<(unsigned)blocks, MM_THREADS, 0, (cudaStream_t)stream>>>(chunk_min, bits, n, c, min_count, d_pad, d_vol);
    return (int)cudaGetLastError();
}
