// afb_misc.cu - whole-volume passes of the bilinear path and the stand-alone R6 op.
//   afb_volume_min : `volume.min()` of nifti_utils.py:200 (+ multiplicity of the minimum, which
//                    torch's MinBackward needs: the gradient is spread evenly over all minima)
//   afb_min_grad   : that MinBackward
//   afb_r6_fwd/bwd : utils/transform_utils.py:27-58
// The two volume passes are pure HBM streams: 16-byte loads, grid = 148 SMs x 8 CTAs.
#include "afb_device.cuh"

namespace afb {

constexpr int MIN_BLOCKS = 148 * 8;
constexpr int MIN_THREADS = 256;

struct MinCount {
    float m;
    float n;
};

__device__ __forceinline__ void mc_merge(float& m, float& n, float m2, float n2) {
    if (m2 < m) { m = m2; n = n2; }
    else if (m2 == m) { n += n2; }
}

template <typename T>
__device__ __forceinline__ float to_f(T v) { return (float)v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(MIN_THREADS)
volume_min_kernel(const T* __restrict__ data, long long n, MinCount* __restrict__ partial, unsigned* __restrict__ counter,
                  float* __restrict__ out) {
    constexpr int VEC = 16 / sizeof(T);
    float m = INFINITY, cnt = 0.0f;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const uint4* __restrict__ d4 = reinterpret_cast<const uint4*>(data);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 raw = __ldcs(d4 + i);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int k = 0; k < VEC; ++k) mc_merge(m, cnt, to_f<T>(e[k]), 1.0f);
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) mc_merge(m, cnt, to_f<T>(data[i]), 1.0f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float m2 = __shfl_xor_sync(0xffffffffu, m, o), n2 = __shfl_xor_sync(0xffffffffu, cnt, o);
        mc_merge(m, cnt, m2, n2);
    }
    __shared__ float sm[MIN_THREADS / 32], sn[MIN_THREADS / 32];
    __shared__ bool last;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sm[w] = m; sn[w] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < MIN_THREADS / 32; ++k) mc_merge(m, cnt, sm[k], sn[k]);
        partial[blockIdx.x].m = m;
        partial[blockIdx.x].n = cnt;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    m = INFINITY; cnt = 0.0f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        const float2 pc = __ldcg(reinterpret_cast<const float2*>(partial) + i);
        mc_merge(m, cnt, pc.x, pc.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float m2 = __shfl_xor_sync(0xffffffffu, m, o), n2 = __shfl_xor_sync(0xffffffffu, cnt, o);
        mc_merge(m, cnt, m2, n2);
    }
    __syncthreads();
    if (lane == 0) { sm[w] = m; sn[w] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < MIN_THREADS / 32; ++k) mc_merge(m, cnt, sm[k], sn[k]);
        out[0] = m;
        out[1] = cnt;
        *counter = 0u;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
min_grad_kernel(const T* __restrict__ vol, long long n, const float* __restrict__ min_count, const float* __restrict__ d_pad,
                float* __restrict__ d_vol) {
    const float m = __ldg(min_count), share = __ldg(d_pad) / __ldg(min_count + 1);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (to_f<T>(vol[i]) == m) d_vol[i] += share;
    }
}

// d_vol = (vol == min) ? share : 0 : zero-fill and MinBackward in ONE pass (8 B/voxel instead of 4 + 12)
template <typename T>
__global__ void __launch_bounds__(256)
min_grad_fill_kernel(const T* __restrict__ vol, long long n, const float* __restrict__ min_count, const float* __restrict__ d_pad,
                     float* __restrict__ d_vol) {
    constexpr int VEC = 16 / sizeof(T);          // elements per 16-byte load of the volume
    const float m = __ldg(min_count), share = __ldg(d_pad) / __ldg(min_count + 1);
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const uint4* __restrict__ v4 = reinterpret_cast<const uint4*>(vol);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 raw = __ldcs(v4 + i);
        const T* e = reinterpret_cast<const T*>(&raw);
        float o[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = (to_f<T>(e[k]) == m) ? share : 0.0f;
#pragma unroll
        for (int k = 0; k < VEC; k += 4)
            *reinterpret_cast<float4*>(d_vol + i * VEC + k) = make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]);
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) d_vol[i] = (to_f<T>(vol[i]) == m) ? share : 0.0f;
}

// fp32 -> storage dtype over a dense block (dVolume of a bf16 / fp16 volume is accumulated in fp32 and handed back in the
// volume's dtype and strides, like the reference's autograd would): 32 bytes in, 16 bytes out per thread per iteration.
template <typename T>
__global__ void __launch_bounds__(256)
cast_from_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n) {
    const long long nvec = n / 8;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i), b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
        T o[8];
        o[0] = Store<T>::from_float(a.x); o[1] = Store<T>::from_float(a.y); o[2] = Store<T>::from_float(a.z); o[3] = Store<T>::from_float(a.w);
        o[4] = Store<T>::from_float(b.x); o[5] = Store<T>::from_float(b.y); o[6] = Store<T>::from_float(b.z); o[7] = Store<T>::from_float(b.w);
        *reinterpret_cast<uint4*>(dst + 8 * i) = *reinterpret_cast<const uint4*>(o);
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * 8 + threadIdx.x; i < n; i += blockDim.x) dst[i] = Store<T>::from_float(src[i]);
}

// Measurement helper: read an (L2-resident) buffer `passes` times with 16-byte ld.global.cg loads; used by
// bench.py to measure the L2 read-bandwidth denominator of the gather kernels' roofline on the same box.
__global__ void __launch_bounds__(512)
probe_read_kernel(const uint4* __restrict__ buf, long long nvec, int passes, float* __restrict__ sink) {
    unsigned acc = 0u;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p) {
#pragma unroll 4
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345678u) sink[0] = 1.0f;     // keeps the loads alive
}

__global__ void r6_fwd_kernel(const float* __restrict__ ortho, int N, float* __restrict__ mat) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float a[3] = {ortho[n * 6 + 0], ortho[n * 6 + 1], ortho[n * 6 + 2]};
    float b[3] = {ortho[n * 6 + 3], ortho[n * 6 + 4], ortho[n * 6 + 5]};
    float rot[9];
    r6_to_rot(a, b, rot, nullptr, nullptr);
    float* m = mat + (size_t)n * 16;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        m[r * 4 + 0] = rot[r * 3 + 0]; m[r * 4 + 1] = rot[r * 3 + 1]; m[r * 4 + 2] = rot[r * 3 + 2]; m[r * 4 + 3] = 0.0f;
    }
    m[12] = 0.0f; m[13] = 0.0f; m[14] = 0.0f; m[15] = 1.0f;
}

__device__ __forceinline__ void crossd(const double* u, const double* v, double* o) {
    o[0] = u[1] * v[2] - u[2] * v[1];
    o[1] = u[2] * v[0] - u[0] * v[2];
    o[2] = u[0] * v[1] - u[1] * v[0];
}

__global__ void r6_bwd_kernel(const float* __restrict__ ortho, const float* __restrict__ gmat, int N, float* __restrict__ d_ortho) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double a[3], b[3];
    for (int k = 0; k < 3; ++k) { a[k] = ortho[n * 6 + k]; b[k] = ortho[n * 6 + 3 + k]; }
    const double na = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    double x[3] = {a[0] / na, a[1] / na, a[2] / na}, zr[3], z[3];
    crossd(x, b, zr);
    const double nz = sqrt(zr[0] * zr[0] + zr[1] * zr[1] + zr[2] * zr[2]);
    for (int k = 0; k < 3; ++k) z[k] = zr[k] / nz;
    const float* g = gmat + (size_t)n * 16;
    double dx[3], dy[3], dz[3], t1[3], t2[3];
    for (int r = 0; r < 3; ++r) { dx[r] = g[r * 4 + 0]; dy[r] = g[r * 4 + 1]; dz[r] = g[r * 4 + 2]; }
    crossd(x, dy, t1); crossd(dy, z, t2);
    for (int r = 0; r < 3; ++r) { dz[r] += t1[r]; dx[r] += t2[r]; }
    const double zdz = z[0] * dz[0] + z[1] * dz[1] + z[2] * dz[2];
    double dzr[3], db[3];
    for (int r = 0; r < 3; ++r) dzr[r] = (dz[r] - z[r] * zdz) / nz;
    crossd(b, dzr, t1); crossd(dzr, x, db);
    for (int r = 0; r < 3; ++r) dx[r] += t1[r];
    const double xdx = x[0] * dx[0] + x[1] * dx[1] + x[2] * dx[2];
    for (int r = 0; r < 3; ++r) {
        d_ortho[n * 6 + r] = (float)((dx[r] - x[r] * xdx) / na);
        d_ortho[n * 6 + 3 + r] = (float)db[r];
    }
}

}  // namespace afb

using namespace afb;

extern "C" int64_t afb_volume_min_workspace_bytes(void) {
    return (int64_t)MIN_BLOCKS * sizeof(MinCount) + 16;
}

extern "C" int afb_volume_min(const void* data, int dtype, int64_t n, float* out_min_count, void* workspace, void* stream) {
    if (!data || !out_min_count || !workspace || n <= 0) return AFB_EINVAL;
    if (((uintptr_t)data & 15u) != 0) return AFB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned* counter = (unsigned*)workspace;
    MinCount* partial = (MinCount*)((char*)workspace + 16);
    cudaError_t e = cudaMemsetAsync(counter, 0, 16, st);
    if (e != cudaSuccess) return (int)e;
    long long want = (n / 4 + MIN_THREADS - 1) / MIN_THREADS;
    int blocks = (int)(want < 1 ? 1 : (want > MIN_BLOCKS ? MIN_BLOCKS : want));
    switch (dtype) {
        case AFB_F32: volume_min_kernel<float><<<blocks, MIN_THREADS, 0, st>>>((const float*)data, n, partial, counter, out_min_count); break;
        case AFB_BF16: volume_min_kernel<__nv_bfloat16><<<blocks, MIN_THREADS, 0, st>>>((const __nv_bfloat16*)data, n, partial, counter, out_min_count); break;
        case AFB_F16: volume_min_kernel<__half><<<blocks, MIN_THREADS, 0, st>>>((const __half*)data, n, partial, counter, out_min_count); break;
        case AFB_I64: volume_min_kernel<int64_t><<<blocks, MIN_THREADS, 0, st>>>((const int64_t*)data, n, partial, counter, out_min_count); break;
        case AFB_I32: volume_min_kernel<int32_t><<<blocks, MIN_THREADS, 0, st>>>((const int32_t*)data, n, partial, counter, out_min_count); break;
        case AFB_I16: volume_min_kernel<int16_t><<<blocks, MIN_THREADS, 0, st>>>((const int16_t*)data, n, partial, counter, out_min_count); break;
        case AFB_U8: volume_min_kernel<uint8_t><<<blocks, MIN_THREADS, 0, st>>>((const uint8_t*)data, n, partial, counter, out_min_count); break;
        default: return AFB_EDTYPE;
    }
    return (int)cudaGetLastError();
}

extern "C" int afb_min_grad(const void* vol, int dtype, int64_t n, const float* min_count, const float* d_pad, float* d_vol, void* stream) {
    if (!vol || !min_count || !d_pad || !d_vol || n <= 0) return AFB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    long long want = (n + 255) / 256;
    int blocks = (int)(want > 148 * 16 ? 148 * 16 : want);
    switch (dtype) {
        case AFB_F32: min_grad_kernel<float><<<blocks, 256, 0, st>>>((const float*)vol, n, min_count, d_pad, d_vol); break;
        case AFB_BF16: min_grad_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)vol, n, min_count, d_pad, d_vol); break;
        case AFB_F16: min_grad_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)vol, n, min_count, d_pad, d_vol); break;
        default: return AFB_EDTYPE;
    }
    return (int)cudaGetLastError();
}

extern "C" int afb_min_grad_fill(const void* vol, int dtype, int64_t n, const float* min_count, const float* d_pad, float* d_vol, void* stream) {
    if (!vol || !min_count || !d_pad || !d_vol || n <= 0) return AFB_EINVAL;
    if (((uintptr_t)vol & 15u) || ((uintptr_t)d_vol & 15u)) return AFB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    long long want = (n / 4 + 255) / 256;
    int blocks = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    switch (dtype) {
        case AFB_F32: min_grad_fill_kernel<float><<<blocks, 256, 0, st>>>((const float*)vol, n, min_count, d_pad, d_vol); break;
        case AFB_BF16: min_grad_fill_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)vol, n, min_count, d_pad, d_vol); break;
        case AFB_F16: min_grad_fill_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)vol, n, min_count, d_pad, d_vol); break;
        default: return AFB_EDTYPE;
    }
    return (int)cudaGetLastError();
}

extern "C" int afb_cast_from_f32(const float* src, void* dst, int dst_dtype, int64_t n, void* stream) {
    if (!src || !dst || n <= 0) return AFB_EINVAL;
    if (((uintptr_t)src & 15u) || ((uintptr_t)dst & 15u)) return AFB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    long long want = (n / 8 + 255) / 256;
    int blocks = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    switch (dst_dtype) {
        case AFB_BF16: cast_from_f32_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(src, (__nv_bfloat16*)dst, n); break;
        case AFB_F16: cast_from_f32_kernel<__half><<<blocks, 256, 0, st>>>(src, (__half*)dst, n); break;
        default: return AFB_EDTYPE;
    }
    return (int)cudaGetLastError();
}

extern "C" int afb_r6_fwd(const float* ortho, int N, float* mat, void* stream) {
    if (!ortho || !mat || N <= 0) return AFB_EINVAL;
    r6_fwd_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ortho, N, mat);
    return (int)cudaGetLastError();
}

extern "C" int afb_r6_bwd(const float* ortho, const float* grad_mat, int N, float* d_ortho, void* stream) {
    if (!ortho || !grad_mat || !d_ortho || N <= 0) return AFB_EINVAL;
    r6_bwd_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ortho, grad_mat, N, d_ortho);
    return (int)cudaGetLastError();
}

extern "C" int afb_probe_read(const void* buf, int64_t n_bytes, int passes, float* sink, void* stream) {
    if (!buf || !sink || n_bytes < 16 || passes <= 0 || ((uintptr_t)buf & 15u)) return AFB_EINVAL;
    probe_read_kernel<<<148 * 4, 512, 0, (cudaStream_t)stream>>>((const uint4*)buf, n_bytes / 16, passes, sink);
    return (int)cudaGetLastError();
}

extern "C" int afb_version(void) { return AFB_VERSION; }

extern "C" const char* afb_error_string(int code) {
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    switch (code) {
        case AFB_OK: return "ok";
        case AFB_EINVAL: return "invalid argument (null pointer, non-positive size or bad enum)";
        case AFB_EDTYPE: return "dtype / mode combination not supported";
        case AFB_ESHAPE: return "inconsistent shapes";
        case AFB_EUNSUPPORTED: return "unsupported";
        default: return "unknown afb error";
    }
}
