// afb_device.cuh - device-side building blocks shared by the slice and embedding kernels.
//
// Arithmetic contract (see DESIGN.md "Numerics"): everything that decides WHICH voxel is read
// (base coordinates, grid coordinates, un-normalisation, floor / nearbyint) and the forward
// trilinear weights/accumulation are written with explicit round-to-nearest intrinsics
// (__fmul_rn/__fadd_rn/__fmaf_rn/__fdiv_rn) so that nvcc can neither fuse nor re-associate them.
// The sequence reproduces torch-CPU ATen bit for bit (oracle/aten_np.py documents and tests it):
//   linspace:    fma(step, i, -1) | fma(-step, K-1-i, 1)          (ATen RangeFactories, one FMA)
//   base:        (lin * (K-1)) / K                                (ATen make_base_grid, 2 ops)
//   affine_grid: fadd(fma(z,t2, fma(y,t1, fmul(x,t0))), t3)       (bmm, FMA chain in k order)
//   unnormalise: ((g + 1) * size - 1) / 2                         (GridSampler.h:27-36, no FMA)
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/afb200.h"

namespace afb {

// ------------------------------------------------------------------------------------------
// coordinates
// ------------------------------------------------------------------------------------------
struct AxisConst {  // per output axis constants, computed on the host in fp32 (IEEE, same as device)
    float step;     // 2/(K-1)   (0 for K==1)
    float km1;      // K-1
    float kf;       // K
    float inv_k;    // 1/K when K is a power of two (then x / K == x * inv_k bit for bit), else 0
    int K;
};

__host__ inline AxisConst make_axis(int K) {
    AxisConst a;
    a.K = K;
    a.km1 = (float)(K - 1);
    a.kf = (float)K;
    a.step = K > 1 ? 2.0f / (float)(K - 1) : 0.0f;
    a.inv_k = (K & (K - 1)) == 0 ? 1.0f / (float)K : 0.0f;
    return a;
}

// Normalised base coordinate of F.affine_grid(align_corners=False), bitwise as torch.
__device__ __forceinline__ float base_coord(int idx, const AxisConst& a) {
    float lin;
    if (a.K == 1) {
        lin = -1.0f;
    } else if (idx < (a.K >> 1)) {
        lin = __fmaf_rn(a.step, (float)idx, -1.0f);
    } else {
        lin = __fmaf_rn(-a.step, (float)(a.K - 1 - idx), 1.0f);
    }
    // division by a power of two is an exact exponent shift (|lin * (K-1)| is 0 or >= 2/K, far from the subnormals),
    // so the multiply gives the same bits as ATen's division without the ~20-instruction IEEE divide
    const float num = __fmul_rn(lin, a.km1);
    return a.inv_k != 0.0f ? __fmul_rn(num, a.inv_k) : __fdiv_rn(num, a.kf);
}

__device__ __forceinline__ float grid_coord(const float* t /*row of 4*/, float x, float y, float z) {
    return __fadd_rn(__fmaf_rn(z, t[2], __fmaf_rn(y, t[1], __fmul_rn(x, t[0]))), t[3]);
}

__device__ __forceinline__ float unnormalize(float g, float size) {
    return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), size), 1.0f), 0.5f);
}

// ------------------------------------------------------------------------------------------
// storage <-> float
// ------------------------------------------------------------------------------------------
template <typename T> struct Store;
template <> struct Store<float> {
    static __device__ __forceinline__ float load(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ float from_float(float v) { return v; }
};
template <> struct Store<__nv_bfloat16> {
    static __device__ __forceinline__ float load(const __nv_bfloat16* p) {
        return __bfloat162float(__ldg(p));
    }
    static __device__ __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Store<__half> {
    static __device__ __forceinline__ float load(const __half* p) { return __half2float(__ldg(p)); }
    static __device__ __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
};
// integer storage: bilinear results are truncated toward zero like torch's .to(int) (nifti_utils.py:205)
template <> struct Store<int64_t> {
    static __device__ __forceinline__ float load(const int64_t* p) { return (float)__ldg((const long long*)p); }
    static __device__ __forceinline__ int64_t from_float(float v) { return (int64_t)v; }
};
template <> struct Store<int32_t> {
    static __device__ __forceinline__ float load(const int32_t* p) { return (float)__ldg(p); }
    static __device__ __forceinline__ int32_t from_float(float v) { return (int32_t)v; }
};
template <> struct Store<int16_t> {
    static __device__ __forceinline__ float load(const int16_t* p) { return (float)__ldg(p); }
    static __device__ __forceinline__ int16_t from_float(float v) { return (int16_t)v; }
};
template <> struct Store<uint8_t> {
    static __device__ __forceinline__ float load(const uint8_t* p) { return (float)__ldg(p); }
    static __device__ __forceinline__ uint8_t from_float(float v) { return (uint8_t)v; }
};

// ------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------
// view prologue: (raw parameters | P | theta) -> grid affine G'      [per slice, per block]
// ------------------------------------------------------------------------------------------
// Kernel-argument copy of afb_views plus derived constants (POD, passed by value).
struct ViewArgs {
    int kind, V;
    const float* theta;
    const void* pre;
    int pre_is_f64;
    const float* params;
    const float* gpre;
    const float* init;
    int R, spat;
    float offset_clip, zoom_clip;
    const double* nii_affine;
    double fov_mm[3];      // D,H,W; <=0 => input FOV
    int D, H, W;           // input volume size
    int Do, Ho, Wo;        // output size
    const void* state;     // ViewState[S] written by afb_view_prologue (NULL: compute in the sampler's own prologue)
};

// Everything the forward prologue produces; lives in shared memory.  The backward chain reads it.
struct alignas(16) ViewState {
    float g[16];           // G' (fp32, what the sampler uses and nifti_grid_sample returns)
    double P[16];          // pre_grid_sample_affine as fp64
    double n[3];           // column norms of P[:3,:3]
    double s[3];           // column scale s_j = rho_j / n_{2-j}
    double rho[3];         // FOV ratio in x,y,z order
    double zin[3];         // zooms of the NIfTI affine (D,H,W order)
    // PARAMS kind only
    float theta[16];       // T@R@Z (fp32)
    float gpre[16];
    float R0[9], Rb[9], Rm[9];
    float a[3], b[3], na, nz;  // R6 inputs (+init), |a|, |x cross b|
    float zm, zb, tanh_z, init_zp;
    float pos[3], offs[3];
};

// 3x3 helpers (row-major)
__device__ __forceinline__ void cross3(const float* u, const float* v, float* o) {
    o[0] = __fsub_rn(__fmul_rn(u[1], v[2]), __fmul_rn(u[2], v[1]));
    o[1] = __fsub_rn(__fmul_rn(u[2], v[0]), __fmul_rn(u[0], v[2]));
    o[2] = __fsub_rn(__fmul_rn(u[0], v[1]), __fmul_rn(u[1], v[0]));
}
__device__ __forceinline__ float norm3(const float* u) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(u[0], u[0]), __fmul_rn(u[1], u[1])), __fmul_rn(u[2], u[2])));
}

// R6 -> rotation, columns (x,y,z), row-major 3x3 out.  utils/transform_utils.py:27-58
__device__ __forceinline__ void r6_to_rot(const float* a, const float* b, float* rot, float* na_out, float* nz_out) {
    float na = norm3(a);
    float x[3] = {__fdiv_rn(a[0], na), __fdiv_rn(a[1], na), __fdiv_rn(a[2], na)};
    float zr[3];
    cross3(x, b, zr);
    float nz = norm3(zr);
    float z[3] = {__fdiv_rn(zr[0], nz), __fdiv_rn(zr[1], nz), __fdiv_rn(zr[2], nz)};
    float y[3];
    cross3(z, x, y);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        rot[r * 3 + 0] = x[r];
        rot[r * 3 + 1] = y[r];
        rot[r * 3 + 2] = z[r];
    }
    if (na_out) *na_out = na;
    if (nz_out) *nz_out = nz;
}

// fp32 matmul as torch-CPU does it for tiny matrices: plain mul+add, k ascending, accumulator 0.
template <int N>
__device__ __forceinline__ void matmul_rn(const float* A, const float* B, float* C) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < N; ++k) acc = __fadd_rn(acc, __fmul_rn(A[i * N + k], B[k * N + j]));
            C[i * N + j] = acc;
        }
}

// nifti_utils.py:36-58 collapsed: G' = P @ diag(s,1), s_j = rho_j / |P[:3, 2-j]|,
// rho_j = fov_mm_out[2-j] / (zoom_in[2-j] * vox_in[2-j]).  One thread.
__device__ inline void normalise_pre_affine(const ViewArgs& va, int b, ViewState& st) {
    double zin[3] = {1.0, 1.0, 1.0};
    if (va.nii_affine) {
        const double* A = va.nii_affine + (size_t)b * 16;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            zin[k] = sqrt(A[0 * 4 + k] * A[0 * 4 + k] + A[1 * 4 + k] * A[1 * 4 + k] + A[2 * 4 + k] * A[2 * 4 + k]);
    }
    const double vox_in[3] = {(double)va.D, (double)va.H, (double)va.W};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        st.zin[k] = zin[k];
        double fov_in = zin[k] * vox_in[k];
        double fov_out = va.fov_mm[k] > 0.0 ? va.fov_mm[k] : fov_in;
        st.rho[2 - k] = fov_out / fov_in;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        st.n[k] = sqrt(st.P[0 * 4 + k] * st.P[0 * 4 + k] + st.P[1 * 4 + k] * st.P[1 * 4 + k] + st.P[2 * 4 + k] * st.P[2 * 4 + k]);
#pragma unroll
    for (int j = 0; j < 3; ++j) st.s[j] = (1.0 / st.n[2 - j]) * st.rho[j];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int j = 0; j < 3; ++j) st.g[r * 4 + j] = (float)(st.P[r * 4 + j] * st.s[j]);
        st.g[r * 4 + 3] = (float)st.P[r * 4 + 3];
    }
}

// NIfTI affine of the resampled array, nifti_utils.py:60-71.  One thread; fp64.
__device__ inline void nii_affine_of_result(const ViewArgs& va, int b, const ViewState& st, double* out /*16*/) {
    double A[16];
    if (va.nii_affine) {
        for (int i = 0; i < 16; ++i) A[i] = va.nii_affine[(size_t)b * 16 + i];
    } else {
        for (int i = 0; i < 16; ++i) A[i] = (i % 5 == 0) ? 1.0 : 0.0;
    }
    const double vin[3] = {(double)va.D, (double)va.H, (double)va.W};
    const double vout[3] = {(double)va.Do, (double)va.Ho, (double)va.Wo};
    // G' in fp64, then swap axes 0<->2 on rows and columns
    double G[16], N[16];
    for (int r = 0; r < 4; ++r) {
        for (int j = 0; j < 3; ++j) G[r * 4 + j] = st.P[r * 4 + j] * st.s[j];
        G[r * 4 + 3] = st.P[r * 4 + 3];
    }
    const int perm[4] = {2, 1, 0, 3};
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) N[r * 4 + c] = G[perm[r] * 4 + perm[c]];
    for (int k = 0; k < 3; ++k) {
        double sc = (st.zin[k] * vin[k]) / (vout[k] * st.zin[k]);
        for (int r = 0; r < 4; ++r) N[r * 4 + k] *= sc;
    }
    for (int k = 0; k < 3; ++k) N[k * 4 + 3] = (N[k * 4 + 3] + 1.0) / 2.0 * vin[k];
    double half_vec[3] = {-(vout[0] - 1.0) / 2.0, -(vout[1] - 1.0) / 2.0, -(vout[2] - 1.0) / 2.0};
    double AN3[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) acc += A[r * 4 + k] * N[k * 4 + c];
            AN3[r * 3 + c] = acc;
        }
    double shift[3];
    for (int r = 0; r < 3; ++r) shift[r] = AN3[r * 3 + 0] * half_vec[0] + AN3[r * 3 + 1] * half_vec[1] + AN3[r * 3 + 2] * half_vec[2];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += A[r * 4 + k] * N[k * 4 + c];
            out[r * 4 + c] = acc;
        }
    for (int r = 0; r < 3; ++r) out[r * 4 + 3] += shift[r];
}

// Forward prologue executed by warp 0 of every block (lane-parallel soft-argmax, lane 0 for the
// 4x4 algebra).  Result in shared `st`.  Caller must __syncthreads() afterwards.
__device__ inline void view_prologue_warp0(const ViewArgs& va, int s, ViewState& st) {
    const int lane = threadIdx.x & 31;
    const int b = s / va.V, v = s % va.V;
    if (va.kind == AFB_AFFINE_GRID) {
        if (lane < 12) st.g[lane] = va.theta[(size_t)s * 12 + lane];
        if (lane >= 12 && lane < 16) st.g[lane] = (lane == 15) ? 1.0f : 0.0f;
        return;
    }
    if (va.kind == AFB_AFFINE_PRE) {
        if (lane < 16) {
            st.P[lane] = va.pre_is_f64 ? ((const double*)va.pre)[(size_t)s * 16 + lane]
                                       : (double)((const float*)va.pre)[(size_t)s * 16 + lane];
        }
        __syncwarp();
        if (lane == 0) normalise_pre_affine(va, b, st);
        return;
    }
    // ---- AFB_AFFINE_PARAMS: learnable_transform.py:144-230, 262-289 ----
    const int NP = 6 + 3 * va.R + 1;
    const float* prm = va.params + (size_t)s * NP;
    const float* ini = va.init + (size_t)v * 10;
    // soft-argmax offsets (:163-176): softmax over R logits, expectation of arra, (2p+1)/spat-1
    const int arra0 = (va.spat - va.R) / 2;
    for (int c = 0; c < 3; ++c) {
        const float* lg = prm + 6 + c * va.R;
        float m = -INFINITY;
        for (int i = lane; i < va.R; i += 32) m = fmaxf(m, lg[i]);
        m = warp_max(m);
        float se = 0.0f, sw = 0.0f;
        for (int i = lane; i < va.R; i += 32) {
            float e = expf(lg[i] - m);
            se += e;
            sw += e * (float)(arra0 + i);
        }
        se = warp_sum(se);
        sw = warp_sum(sw);
        if (lane == 0) {
            float pos = sw / se;
            st.pos[c] = pos;
            float off = __fsub_rn(__fdiv_rn(__fadd_rn(__fmul_rn(2.0f, pos), 1.0f), (float)va.spat), 1.0f);
            st.offs[c] = (va.offset_clip == 0.0f) ? 0.0f : off;
        }
    }
    if (lane < 16) st.gpre[lane] = va.gpre[(size_t)s * 16 + lane];
    __syncwarp();
    if (lane == 0) {
        float a0[3] = {ini[0], ini[1], ini[2]}, b0[3] = {ini[3], ini[4], ini[5]};
        r6_to_rot(a0, b0, st.R0, nullptr, nullptr);                     // init rotation (:150)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            st.a[k] = __fadd_rn(prm[k], ini[k]);                        // :198
            st.b[k] = __fadd_rn(prm[3 + k], ini[3 + k]);
        }
        r6_to_rot(st.a, st.b, st.Rb, &st.na, &st.nz);                   // :207
        matmul_rn<3>(st.R0, st.Rb, st.Rm);                              // :268
        st.init_zp = ini[9];
        float zp = __fadd_rn(prm[NP - 1], ini[9]);                      // :199
        st.tanh_z = tanhf(zp);
        st.zb = __fadd_rn(__fmul_rn(va.zoom_clip, -st.tanh_z), 1.0f);   // :220
        st.zm = __fmul_rn(ini[9], st.zb);                               // :270
        float t[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = __fadd_rn(st.offs[k], ini[6 + k]);   // :269
        // theta = T @ R @ Z (:272): rotation block scaled by zm, translation t
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) st.theta[r * 4 + c] = __fmul_rn(st.Rm[r * 3 + c], st.zm);
            st.theta[r * 4 + 3] = t[r];
        }
        st.theta[12] = 0.0f; st.theta[13] = 0.0f; st.theta[14] = 0.0f; st.theta[15] = 1.0f;
        float Pf[16];
        matmul_rn<4>(st.gpre, st.theta, Pf);                            // :289
#pragma unroll
        for (int i = 0; i < 16; ++i) st.P[i] = (double)Pf[i];
        normalise_pre_affine(va, b, st);
    }
}

// ------------------------------------------------------------------------------------------
// host helpers shared by the translation units
// ------------------------------------------------------------------------------------------
inline int make_view_args(const afb_views* views, int B, int D, int H, int W, int Do, int Ho, int Wo, bool need_state,
                          ViewArgs& a) {
    if (!views) return AFB_EINVAL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || Do <= 0 || Ho <= 0 || Wo <= 0 || views->V <= 0) return AFB_ESHAPE;
    if ((long long)B * views->V > 65535) return AFB_ESHAPE;
    if (views->kind < AFB_AFFINE_GRID || views->kind > AFB_AFFINE_PARAMS) return AFB_EINVAL;
    if (need_state && !views->state) return AFB_EINVAL;
    if (views->kind == AFB_AFFINE_PARAMS && (views->R < 0 || views->spat <= 0)) return AFB_ESHAPE;
    a.kind = views->kind; a.V = views->V; a.theta = views->theta; a.pre = views->pre; a.pre_is_f64 = views->pre_is_f64;
    a.params = views->params; a.gpre = views->gpre; a.init = views->init; a.R = views->R; a.spat = views->spat;
    a.offset_clip = views->offset_clip; a.zoom_clip = views->zoom_clip; a.nii_affine = views->nii_affine;
    for (int k = 0; k < 3; ++k) a.fov_mm[k] = views->fov_mm[k];
    a.D = D; a.H = H; a.W = W; a.Do = Do; a.Ho = Ho; a.Wo = Wo;
    a.state = views->state;
    return AFB_OK;
}

// afb_views.cu: dG' (fp64 sums in ws_acc + upstream) -> gradient of the view input; re-zeroes ws_acc
int launch_view_chain(const ViewArgs& a, int S, double* ws_acc, const float* grad_grid_affine, float* d_affine,
                      float* d_gpre, cudaStream_t st);

}  // namespace afb
