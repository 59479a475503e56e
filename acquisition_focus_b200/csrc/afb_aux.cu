// afb_aux.cu - the small callers either side of the samplers, one kernel each instead of a
// dozen eager torch ops:
//   afb_compose_pre_affine : Gpre = base_affine^-1 @ view_affine [@ augmentation]      running/run_dl.py:208-234
//   afb_upsample2d_fwd/bwd : F.interpolate(slice, size=hires[:2]+[1], 'trilinear')      running/run_dl.py:193-197
//   afb_rot3_fwd/bwd       : angle-axis / normal-vector -> rotation 4x4                 utils/transform_utils.py:62-178
#include "afb_device.cuh"

namespace afb {

// ------------------------------------------------------------------------------------------------
// a6: clinical composition.  The reference computes it in fp64 (base_affine is cast to the NIfTI affine's dtype,
// run_dl.py:248) with torch's LU inverse and casts to fp32 only inside the ATM (learnable_transform.py:284).  One thread
// per batch element: 4x4 Gauss-Jordan with partial pivoting in fp64, two 4x4 products, one rounding to fp32.
// ------------------------------------------------------------------------------------------------
__device__ inline bool inverse4x4(const double* a, double* inv) {
    double m[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { m[r][c] = a[r * 4 + c]; m[r][4 + c] = (r == c) ? 1.0 : 0.0; }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        double best = fabs(m[col][col]);
        for (int r = col + 1; r < 4; ++r)
            if (fabs(m[r][col]) > best) { best = fabs(m[r][col]); piv = r; }
        if (best == 0.0) return false;
        if (piv != col)
            for (int c = 0; c < 8; ++c) { const double t = m[col][c]; m[col][c] = m[piv][c]; m[piv][c] = t; }
        const double d = 1.0 / m[col][col];
        for (int c = 0; c < 8; ++c) m[col][c] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const double f = m[r][col];
            if (f != 0.0)
                for (int c = 0; c < 8; ++c) m[r][c] -= f * m[col][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) inv[r * 4 + c] = m[r][4 + c];
    return true;
}

__device__ inline void matmul4d(const double* a, const double* b, double* c) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += a[i * 4 + k] * b[k * 4 + j];
            c[i * 4 + j] = acc;
        }
}

__global__ void compose_pre_affine_kernel(const double* __restrict__ base, const void* __restrict__ view, int view_is_f64,
                                          const float* __restrict__ aug, int B, float* __restrict__ out, int* __restrict__ singular) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double A[16], Ai[16], Vw[16], P[16], Q[16];
    for (int i = 0; i < 16; ++i) {
        A[i] = base[(size_t)b * 16 + i];
        Vw[i] = view_is_f64 ? ((const double*)view)[(size_t)b * 16 + i] : (double)((const float*)view)[(size_t)b * 16 + i];
    }
    if (!inverse4x4(A, Ai)) {
        if (singular) atomicExch(singular, 1);
        for (int i = 0; i < 16; ++i) out[(size_t)b * 16 + i] = nanf("");
        return;
    }
    matmul4d(Ai, Vw, P);
    if (aug) {
        double G[16];
        for (int i = 0; i < 16; ++i) G[i] = (double)aug[(size_t)b * 16 + i];
        matmul4d(P, G, Q);
        for (int i = 0; i < 16; ++i) out[(size_t)b * 16 + i] = (float)Q[i];
    } else {
        for (int i = 0; i < 16; ++i) out[(size_t)b * 16 + i] = (float)P[i];
    }
}

// ------------------------------------------------------------------------------------------------
// f3: up-sampling of low-resolution slices to the hires in-plane size.  F.interpolate(x[N,C,h,w,1], size=[H,W,1],
// mode='trilinear', align_corners=False) = ATen upsample_trilinear3d with a singleton last axis: source index
// max(0, scale*(dst+0.5)-0.5), scale = in/out (fp32 division), i1 = i0 + (i0 < in-1), weights (1-l, l), value
// t0*(h0*x00 + h1*x01) + t1*(h0*x10 + h1*x11) in that order (the singleton axis contributes 1*x + 0*x = x).
// ------------------------------------------------------------------------------------------------
struct Src1 { int i0, i1; float l0, l1; };

__device__ __forceinline__ Src1 up_source(int dst, float scale, int in_size) {
    float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    if (s < 0.0f) s = 0.0f;
    Src1 r;
    r.i0 = (int)s;
    r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
    r.l1 = __fsub_rn(s, (float)r.i0);
    r.l0 = __fsub_rn(1.0f, r.l1);
    return r;
}

__global__ void __launch_bounds__(256)
upsample2d_fwd_kernel(const float* __restrict__ x, int h, int w, int H, int W, float sh, float sw, float* __restrict__ out) {
    const int ow = blockIdx.x * blockDim.x + threadIdx.x, oh = blockIdx.y;
    if (ow >= W) return;
    const size_t n = blockIdx.z;
    const Src1 a = up_source(oh, sh, h), b = up_source(ow, sw, w);
    const float* __restrict__ p = x + n * (size_t)h * w;
    const float x00 = __ldg(p + a.i0 * w + b.i0), x01 = __ldg(p + a.i0 * w + b.i1);
    const float x10 = __ldg(p + a.i1 * w + b.i0), x11 = __ldg(p + a.i1 * w + b.i1);
    const float r0 = __fadd_rn(__fmul_rn(b.l0, x00), __fmul_rn(b.l1, x01));
    const float r1 = __fadd_rn(__fmul_rn(b.l0, x10), __fmul_rn(b.l1, x11));
    out[(n * H + oh) * (size_t)W + ow] = __fadd_rn(__fmul_rn(a.l0, r0), __fmul_rn(a.l1, r1));
}

// backward, gather form (deterministic, no atomics): an input pixel collects from the output rows / columns whose two taps
// include it; the candidate range follows from the source-index formula, membership is tested with the exact forward code.
__global__ void __launch_bounds__(256)
upsample2d_bwd_kernel(const float* __restrict__ go, int h, int w, int H, int W, float sh, float sw, float* __restrict__ dx) {
    const int iw = blockIdx.x * blockDim.x + threadIdx.x, ih = blockIdx.y;
    if (iw >= w) return;
    const size_t n = blockIdx.z;
    const float fh = (float)H / (float)h, fw = (float)W / (float)w;
    const int oh0 = max(0, (int)floorf(((float)ih - 1.0f) * fh) - 1), oh1 = min(H - 1, (int)ceilf(((float)ih + 1.5f) * fh) + 1);
    const int ow0 = max(0, (int)floorf(((float)iw - 1.0f) * fw) - 1), ow1 = min(W - 1, (int)ceilf(((float)iw + 1.5f) * fw) + 1);
    const float* __restrict__ g = go + n * (size_t)H * W;
    float acc = 0.0f;
    for (int oh = oh0; oh <= oh1; ++oh) {
        const Src1 a = up_source(oh, sh, h);
        float wa = 0.0f;
        if (a.i0 == ih) wa += a.l0;
        if (a.i1 == ih) wa += a.l1;
        if (wa == 0.0f && a.i0 != ih && a.i1 != ih) continue;
        for (int ow = ow0; ow <= ow1; ++ow) {
            const Src1 b = up_source(ow, sw, w);
            float wb = 0.0f;
            if (b.i0 == iw) wb += b.l0;
            if (b.i1 == iw) wb += b.l1;
            if (b.i0 != iw && b.i1 != iw) continue;
            acc = fmaf(wa * wb, __ldg(g + (size_t)oh * W + ow), acc);
        }
    }
    dx[(n * h + ih) * (size_t)w + iw] = acc;
}

// ------------------------------------------------------------------------------------------------
// f4: the two non-default rotation parameterisations (optim_method 'angle-axis' / 'normal-vector').
// ------------------------------------------------------------------------------------------------
constexpr float AA_EPS = 1e-6f;

__global__ void rot3_fwd_kernel(int kind, const float* __restrict__ in, int N, float* __restrict__ mat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float a = in[i * 3 + 0], b = in[i * 3 + 1], c = in[i * 3 + 2];
    float R[9];
    if (kind == AFB_ROT_ANGLE_AXIS) {
        // transform_utils.py:106-178: w = r / (sqrt(|r|^2 + eps) + eps) where |r|^2 > eps, first-order I + [r]x elsewhere
        const float t2 = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
        if (t2 > AA_EPS) {
            const float th = __fsqrt_rn(__fadd_rn(t2, AA_EPS));
            const float d = __fadd_rn(th, AA_EPS);
            const float wx = __fdiv_rn(a, d), wy = __fdiv_rn(b, d), wz = __fdiv_rn(c, d);
            const float co = cosf(th), si = sinf(th), k = __fsub_rn(1.0f, co);
            R[0] = __fadd_rn(co, __fmul_rn(__fmul_rn(wx, wx), k));
            R[1] = __fsub_rn(__fmul_rn(__fmul_rn(wx, wy), k), __fmul_rn(wz, si));
            R[2] = __fadd_rn(__fmul_rn(wy, si), __fmul_rn(__fmul_rn(wx, wz), k));
            R[3] = __fadd_rn(__fmul_rn(wz, si), __fmul_rn(__fmul_rn(wx, wy), k));
            R[4] = __fadd_rn(co, __fmul_rn(__fmul_rn(wy, wy), k));
            R[5] = __fadd_rn(__fmul_rn(-wx, si), __fmul_rn(__fmul_rn(wy, wz), k));
            R[6] = __fadd_rn(__fmul_rn(-wy, si), __fmul_rn(__fmul_rn(wx, wz), k));
            R[7] = __fadd_rn(__fmul_rn(wx, si), __fmul_rn(__fmul_rn(wy, wz), k));
            R[8] = __fadd_rn(co, __fmul_rn(__fmul_rn(wz, wz), k));
        } else {
            R[0] = 1.0f; R[1] = -c; R[2] = b; R[3] = c; R[4] = 1.0f; R[5] = -a; R[6] = -b; R[7] = a; R[8] = 1.0f;
        }
    } else {
        // transform_utils.py:62-103: input columns are (nz, ny, nx); d = sqrt(nx^2 + ny^2) (d = 0 divides by zero as the reference)
        const float nz = a, ny = b, nx = c;
        const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)));
        R[0] = __fdiv_rn(ny, d); R[1] = __fdiv_rn(-nx, d); R[2] = 0.0f;
        R[3] = __fdiv_rn(__fmul_rn(nx, nz), d); R[4] = __fdiv_rn(__fmul_rn(ny, nz), d); R[5] = -d;
        R[6] = nx; R[7] = ny; R[8] = nz;
    }
    float* m = mat + (size_t)i * 16;
    for (int r = 0; r < 3; ++r) {
        for (int q = 0; q < 3; ++q) m[r * 4 + q] = R[r * 3 + q];
        m[r * 4 + 3] = 0.0f;
    }
    m[12] = 0.0f; m[13] = 0.0f; m[14] = 0.0f; m[15] = 1.0f;
}

__global__ void rot3_bwd_kernel(int kind, const float* __restrict__ in, const float* __restrict__ gmat, int N, float* __restrict__ d_in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double a = in[i * 3 + 0], b = in[i * 3 + 1], c = in[i * 3 + 2];
    double G[9];
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) G[r * 3 + q] = (double)gmat[(size_t)i * 16 + r * 4 + q];
    double d0, d1, d2;
    if (kind == AFB_ROT_ANGLE_AXIS) {
        // R = c I + (1-c) w w^T + s [w]x ;  skew part of G: q = (G21-G12, G02-G20, G10-G01)
        const double q0 = G[7] - G[5], q1 = G[2] - G[6], q2 = G[3] - G[1];
        const double t2 = (double)(float)((float)(a * a) + (float)(b * b) + (float)(c * c));
        if (t2 > (double)AA_EPS) {
            const double th = sqrt(t2 + (double)AA_EPS), d = th + (double)AA_EPS;
            const double w[3] = {a / d, b / d, c / d};
            const double co = cos(th), si = sin(th), k = 1.0 - co;
            double wGw = 0.0, Sw[3] = {0.0, 0.0, 0.0};
            for (int r = 0; r < 3; ++r)
                for (int q = 0; q < 3; ++q) {
                    wGw += G[r * 3 + q] * w[r] * w[q];
                    Sw[r] += (G[r * 3 + q] + G[q * 3 + r]) * w[q];
                }
            const double dco = (G[0] + G[4] + G[8]) - wGw;
            const double dsi = w[0] * q0 + w[1] * q1 + w[2] * q2;
            const double dw[3] = {k * Sw[0] + si * q0, k * Sw[1] + si * q1, k * Sw[2] + si * q2};
            double dth = -si * dco + co * dsi;
            const double dd = -(dw[0] * a + dw[1] * b + dw[2] * c) / (d * d);
            dth += dd;
            const double dt2 = dth / (2.0 * th);
            d0 = dw[0] / d + 2.0 * a * dt2;
            d1 = dw[1] / d + 2.0 * b * dt2;
            d2 = dw[2] / d + 2.0 * c * dt2;
        } else {
            d0 = q0; d1 = q1; d2 = q2;
        }
    } else {
        const double nz = a, ny = b, nx = c;
        const double d = sqrt(nx * nx + ny * ny), id = 1.0 / d;
        const double dd = (-G[0] * ny + G[1] * nx - G[3] * nx * nz - G[4] * ny * nz) * id * id - G[5];
        const double dnx = -G[1] * id + G[3] * nz * id + G[6] + dd * nx * id;
        const double dny = G[0] * id + G[4] * nz * id + G[7] + dd * ny * id;
        const double dnz = (G[3] * nx + G[4] * ny) * id + G[8];
        d0 = dnz; d1 = dny; d2 = dnx;
    }
    d_in[i * 3 + 0] = (float)d0; d_in[i * 3 + 1] = (float)d1; d_in[i * 3 + 2] = (float)d2;
}

}  // namespace afb

using namespace afb;

extern "C" int afb_compose_pre_affine(const double* base, const void* view, int view_is_f64, const float* aug, int B, float* out,
                                      int* singular_flag, void* stream) {
    if (!base || !view || !out || B <= 0) return AFB_EINVAL;
    compose_pre_affine_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(base, view, view_is_f64, aug, B, out, singular_flag);
    return (int)cudaGetLastError();
}

extern "C" int afb_upsample2d_fwd(const float* x, int64_t n_planes, int h, int w, int H, int W, float* out, void* stream) {
    if (!x || !out) return AFB_EINVAL;
    if (n_planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || H > 65535 || n_planes > 2147483647ll) return AFB_ESHAPE;
    const float sh = (float)h / (float)H, sw = (float)w / (float)W;          // area_pixel_compute_scale (align_corners=False)
    for (int64_t n0 = 0; n0 < n_planes; n0 += 65535) {
        const unsigned nz = (unsigned)(n_planes - n0 < 65535 ? n_planes - n0 : 65535);
        upsample2d_fwd_kernel<<<dim3((W + 255) / 256, H, nz), 256, 0, (cudaStream_t)stream>>>(x + n0 * (int64_t)h * w, h, w, H, W, sh, sw,
                                                                                               out + n0 * (int64_t)H * W);
    }
    return (int)cudaGetLastError();
}

extern "C" int afb_upsample2d_bwd(const float* grad_out, int64_t n_planes, int h, int w, int H, int W, float* d_x, void* stream) {
    if (!grad_out || !d_x) return AFB_EINVAL;
    if (n_planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || h > 65535 || n_planes > 2147483647ll) return AFB_ESHAPE;
    const float sh = (float)h / (float)H, sw = (float)w / (float)W;
    for (int64_t n0 = 0; n0 < n_planes; n0 += 65535) {
        const unsigned nz = (unsigned)(n_planes - n0 < 65535 ? n_planes - n0 : 65535);
        upsample2d_bwd_kernel<<<dim3((w + 255) / 256, h, nz), 256, 0, (cudaStream_t)stream>>>(grad_out + n0 * (int64_t)H * W, h, w, H, W, sh, sw,
                                                                                               d_x + n0 * (int64_t)h * w);
    }
    return (int)cudaGetLastError();
}

extern "C" int afb_rot3_fwd(int kind, const float* params, int N, float* mat, void* stream) {
    if (!params || !mat || N <= 0 || (kind != AFB_ROT_ANGLE_AXIS && kind != AFB_ROT_NORMAL)) return AFB_EINVAL;
    rot3_fwd_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(kind, params, N, mat);
    return (int)cudaGetLastError();
}

extern "C" int afb_rot3_bwd(int kind, const float* params, const float* grad_mat, int N, float* d_params, void* stream) {
    if (!params || !grad_mat || !d_params || N <= 0 || (kind != AFB_ROT_ANGLE_AXIS && kind != AFB_ROT_NORMAL)) return AFB_EINVAL;
    rot3_bwd_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(kind, params, grad_mat, N, d_params);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// f4 (rest): the voxel passes of get_clinical_cardiac_view_affines (functional/clinical_cardiac_views.py:223-364), which the
// reference runs on sparse CPU tensors: (1) count / first / second moments of the voxel index cloud of several label GROUPS
// in ONE pass over the integer label map (exact 64-bit integer sums -> centre and inertia tensor of
// utils/torch_sparse_tensor_utils.py:34-56), (2) the bisection search for the extent of a group along an axis (:36-60).
// The 3x3 eigenproblems and the frame algebra in between are a few dozen flops and stay on the host, as in the reference.
// ------------------------------------------------------------------------------------------------
namespace afb {

constexpr int MOM_THREADS = 256;
constexpr int MOM_MAX_GROUPS = 8;

template <typename L>
__global__ void __launch_bounds__(MOM_THREADS)
label_group_moments_kernel(const L* __restrict__ lab, int D, int H, int W, const unsigned* __restrict__ masks, int G,
                           unsigned long long* __restrict__ out /*[G][10]*/) {
    __shared__ unsigned long long red[MOM_THREADS / 32][10];
    const long long n = (long long)D * H * W;
    unsigned gm[MOM_MAX_GROUPS];
#pragma unroll
    for (int g = 0; g < MOM_MAX_GROUPS; ++g) gm[g] = g < G ? masks[g] : 0u;
    for (int g = 0; g < G; ++g) {
        unsigned long long s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (long long i = (long long)blockIdx.x * MOM_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * MOM_THREADS) {
            const long long l = (long long)lab[i];
            if (l <= 0 || l > 31 || !((gm[g] >> l) & 1u)) continue;
            const unsigned long long w = (unsigned long long)(i % W), h = (unsigned long long)((i / W) % H), d = (unsigned long long)(i / ((long long)W * H));
            s[0] += 1; s[1] += d; s[2] += h; s[3] += w;
            s[4] += d * d; s[5] += d * h; s[6] += d * w; s[7] += h * h; s[8] += h * w; s[9] += w * w;
        }
        const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int q = 0; q < 10; ++q) {
            unsigned long long v = s[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[wi][q] = v;
        }
        __syncthreads();
        if (threadIdx.x < 10) {
            unsigned long long t = 0;
            for (int ww = 0; ww < MOM_THREADS / 32; ++ww) t += red[ww][threadIdx.x];
            if (t) atomicAdd(out + g * 10 + threadIdx.x, t);
        }
        __syncthreads();
    }
}

// bisection of get_extent_vect (:36-48) for +dir and -dir, one CTA: fact in fp64 like the reference's Python floats, the end
// point and the distances in fp32 like its tensors; out[0], out[1] = the returned factors ((start+end)/2) for +dir / -dir.
constexpr int EXT_THREADS = 1024;

template <typename L>
__global__ void __launch_bounds__(EXT_THREADS)
label_extent_search_kernel(const L* __restrict__ lab, int D, int H, int W, unsigned mask, const float* __restrict__ center,
                           const float* __restrict__ dir, double init_end, double* __restrict__ out) {
    __shared__ float red[EXT_THREADS / 32];
    __shared__ float dmin_s;
    const long long n = (long long)D * H * W;
    const float MIN_DIST = (float)(1.73 / 2.0);
    for (int sgn = 0; sgn < 2; ++sgn) {
        const float dd = sgn ? -dir[0] : dir[0], dh = sgn ? -dir[1] : dir[1], dw = sgn ? -dir[2] : dir[2];
        double start = 0.0, end = init_end;
        while ((end - start) > 1.73 / 2.0) {
            const double new_end = end - (end - start) / 2.0;
            const float f = (float)new_end;
            const float ed = __fadd_rn(center[0], __fmul_rn(f, dd)), eh = __fadd_rn(center[1], __fmul_rn(f, dh)),
                        ew = __fadd_rn(center[2], __fmul_rn(f, dw));
            float best = INFINITY;
            for (long long i = threadIdx.x; i < n; i += EXT_THREADS) {
                const long long l = (long long)lab[i];
                if (l <= 0 || l > 31 || !((mask >> l) & 1u)) continue;
                const float w = (float)(i % W), h = (float)((i / W) % H), d = (float)(i / ((long long)W * H));
                const float a = __fsub_rn(d, ed), b = __fsub_rn(h, eh), c = __fsub_rn(w, ew);
                const float r = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
                best = fminf(best, r);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
            __syncthreads();
            if (threadIdx.x == 0) {
                float m = red[0];
                for (int q = 1; q < EXT_THREADS / 32; ++q) m = fminf(m, red[q]);
                dmin_s = m;
            }
            __syncthreads();
            if (dmin_s > MIN_DIST) end = new_end;
            else start += (end - start) / 2.0;
            __syncthreads();
        }
        if (threadIdx.x == 0) out[sgn] = (start + end) / 2.0;
    }
}

}  // namespace afb

extern "C" int afb_label_group_moments(const void* labels, int dtype, int D, int H, int W, const unsigned* group_masks_dev, int n_groups,
                                       unsigned long long* out_dev /*[n_groups][10], zeroed by the caller*/, void* stream) {
    if (!labels || !group_masks_dev || !out_dev) return AFB_EINVAL;
    if (D <= 0 || H <= 0 || W <= 0 || n_groups <= 0 || n_groups > MOM_MAX_GROUPS) return AFB_ESHAPE;
    const long long n = (long long)D * H * W;
    long long want = (n + MOM_THREADS - 1) / MOM_THREADS;
    const unsigned grid = (unsigned)(want > 148 * 8 ? 148 * 8 : want);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case AFB_U8: label_group_moments_kernel<uint8_t><<<grid, MOM_THREADS, 0, st>>>((const uint8_t*)labels, D, H, W, group_masks_dev, n_groups, out_dev); break;
        case AFB_I16: label_group_moments_kernel<int16_t><<<grid, MOM_THREADS, 0, st>>>((const int16_t*)labels, D, H, W, group_masks_dev, n_groups, out_dev); break;
        case AFB_I32: label_group_moments_kernel<int32_t><<<grid, MOM_THREADS, 0, st>>>((const int32_t*)labels, D, H, W, group_masks_dev, n_groups, out_dev); break;
        case AFB_I64: label_group_moments_kernel<int64_t><<<grid, MOM_THREADS, 0, st>>>((const int64_t*)labels, D, H, W, group_masks_dev, n_groups, out_dev); break;
        default: return AFB_EDTYPE;
    }
    return (int)cudaGetLastError();
}

extern "C" int afb_label_extent_search(const void* labels, int dtype, int D, int H, int W, unsigned group_mask, const float* center_dev,
                                       const float* dir_dev, double init_end, double* out_dev /*[2]*/, void* stream) {
    if (!labels || !center_dev || !dir_dev || !out_dev) return AFB_EINVAL;
    if (D <= 0 || H <= 0 || W <= 0 || !(init_end > 0.0)) return AFB_ESHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case AFB_U8: label_extent_search_kernel<uint8_t><<<1, EXT_THREADS, 0, st>>>((const uint8_t*)labels, D, H, W, group_mask, center_dev, dir_dev, init_end, out_dev); break;
        case AFB_I16: label_extent_search_kernel<int16_t><<<1, EXT_THREADS, 0, st>>>((const int16_t*)labels, D, H, W, group_mask, center_dev, dir_dev, init_end, out_dev); break;
        case AFB_I32: label_extent_search_kernel<int32_t><<<1, EXT_THREADS, 0, st>>>((const int32_t*)labels, D, H, W, group_mask, center_dev, dir_dev, init_end, out_dev); break;
        case AFB_I64: label_extent_search_kernel<int64_t><<<1, EXT_THREADS, 0, st>>>((const int64_t*)labels, D, H, W, group_mask, center_dev, dir_dev, init_end, out_dev); break;
        default: return AFB_EDTYPE;
    }
    return (int)cudaGetLastError();
}
