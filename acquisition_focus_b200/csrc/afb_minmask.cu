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
__global__ void __launch_bounds__(MM_THREADS)
volume_min_mask_kernel(const float* __restrict__ data, long long n, long long n_chunks, float* __restrict__ chunk_min,
                       unsigned short* __restrict__ bits, MinCountF* __restrict__ partial, unsigned* __restrict__ counter,
                       float* __restrict__ out) {
    __shared__ float sm[MM_THREADS / 32], sn[MM_THREADS / 32];
    __shared__ bool last;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const long long warps = (long long)gridDim.x * (MM_THREADS / 32);
    float gm = INFINITY, gc = 0.0f;                           // this lane's running (min, multiplicity)
    for (long long ch = (long long)blockIdx.x * (MM_THREADS / 32) + w; ch < n_chunks; ch += warps) {
        const long long base = ch * MM_CHUNK;
        float4 v[MM_LOADS];
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
    // CTA reduction of (min, multiplicity), then the last CTA reduces the per-CTA partials
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, gm, o), n2 = __shfl_xor_sync(0xffffffffu, gc, o);
        mm_merge(gm, gc, m2, n2);
    }
    if (lane == 0) { sm[w] = gm; sn[w] = gc; }
    __syncthreads();
    if (t == 0) {
        for (int k = 1; k < MM_THREADS / 32; ++k) mm_merge(gm, gc, sm[k], sn[k]);
        partial[blockIdx.x].m = gm;
        partial[blockIdx.x].n = gc;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    gm = INFINITY; gc = 0.0f;
    for (int i = t; i < (int)gridDim.x; i += MM_THREADS) {
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
        for (int k = 1; k < MM_THREADS / 32; ++k) mm_merge(gm, gc, sm[k], sn[k]);
        out[0] = gm;
        out[1] = gc;
        *counter = 0u;
    }
}

// d_vol = (vol == min) ? d_pad / count : 0 from the chunk record (never touches the volume)
__global__ void __launch_bounds__(MM_THREADS)
min_grad_fill_mask_kernel(const float* __restrict__ chunk_min, const unsigned short* __restrict__ bits, long long n,
                          long long n_chunks, const float* __restrict__ min_count, const float* __restrict__ d_pad,
                          float* __restrict__ d_vol) {
    const float m = __ldg(min_count), share = __ldg(d_pad) / __ldg(min_count + 1);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (MM_THREADS / 32);
    for (long long ch = (long long)blockIdx.x * (MM_THREADS / 32) + w; ch < n_chunks; ch += warps) {
        const bool hot = __ldg(chunk_min + ch) == m;
        const unsigned b = hot ? (unsigned)__ldg(bits + ch * 32 + lane) : 0u;
        const long long base = ch * MM_CHUNK;
#pragma unroll
        for (int j = 0; j < MM_LOADS; ++j) {
            const long long e = base + j * 128 + lane * 4;
            const unsigned q = (b >> (4 * j)) & 15u;
            const float4 o = make_float4((q & 1u) ? share : 0.0f, (q & 2u) ? share : 0.0f, (q & 4u) ? share : 0.0f, (q & 8u) ? share : 0.0f);
            if (e + 3 < n) {
                __stcs(reinterpret_cast<float4*>(d_vol + e), o);
            } else {
                if (e + 0 < n) d_vol[e + 0] = o.x;
                if (e + 1 < n) d_vol[e + 1] = o.y;
                if (e + 2 < n) d_vol[e + 2] = o.z;
            }
        }
    }
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
    const long long wantb = (c + MM_THREADS / 32 - 1) / (MM_THREADS / 32);
    const int blocks = (int)(wantb < 148 * 16 ? wantb : 148 * 16);
    min_grad_fill_mask_kernel<<<blocks, MM_THREADS, 0, (cudaStream_t)stream>>>(chunk_min, bits, n, c, min_count, d_pad, d_vol);
    return (int)cudaGetLastError();
}
