// afb_slice.cu - slice / volume extraction (F.affine_grid + F.grid_sample of the reference,
// utils/nifti_utils.py:182-203), forward and backward samplers.
//
// One launch covers all S = B*V slices: grid = (tiles per slice, V, B), 256 threads per CTA, one
// 16x16 tile of output locations per CTA.  Lanes of a warp cover an 8x4 patch (not 32x1) so that an
// oblique plane touches few distinct 128-byte lines per load instruction.  The sampling grid is never
// materialised: each thread rebuilds its coordinate from the per-slice grid affine (12 floats read from
// the ViewState written by afb_view_prologue) with the bit-exact arithmetic of afb_device.cuh.
//
// Two kernel families:
//   generic  - arbitrary element strides and storage dtype, loop over channels (image volumes, C = 1)
//   channels-last (sC == 1) - what one_hot(...).permute(...) of running/run_dl.py:261-264 hands the
//              sampler: the C channels of a voxel are contiguous, so a corner is one 16-byte load per
//              4 fp32 channels in the forward, one 32-byte load (LDG.256, sm_100) per 8 channels in the
//              backward, and dVolume uses 16-byte vector reductions (red.global.add.v4.f32) - 4x fewer
//              L2 atomic sectors than scalar REDs.
// Per-pixel instruction count matters here (the C = 1 and nearest kernels were issue-bound on coordinate and
// address arithmetic): no integer or IEEE divisions on the common path, 32-bit offsets, see DESIGN.md section 5.
#include <cstdlib>
#include <type_traits>

#include "afb_sampler.cuh"

namespace afb {

// ------------------------------------------------------------------------------------------------
// forward, generic strides (any layout, any dtype): one thread per output location, loop over C
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(NTHREADS, 4)
slice_fwd_kernel(VolArgs vol, ViewArgs va, OutGeom g, int pad_mode, float pad_value, const float* pad_device,
                 T* __restrict__ out) {
    const int s = blockIdx.z * va.V + blockIdx.y;      // grid = (tiles, V, B): no integer division for the batch index
    const Pix p = pixel_of_thread(g);
    if (!p.valid) return;
    const Sample sm = sample_coords(g, p, va, s, vol);
    const int b = blockIdx.z;
    const T* __restrict__ src = (const T*)vol.data + (long long)b * vol.sB;
    const int plane = g.Do * g.Ho * g.Wo;      // C * plane < 2^31 (checked on the host): 32-bit channel offsets
    T* __restrict__ dst = out + (size_t)s * (size_t)(vol.C * plane) + ((p.i * g.Ho + p.j) * g.Wo + p.k);

    if (MODE == AFB_NEAREST) {
        // nearbyint = round half to even (cvt.rni), ATen grid_sampler_3d nearest
        const int xn = __float2int_rn(sm.ix), yn = __float2int_rn(sm.iy), zn = __float2int_rn(sm.iz);
        const bool in = xn >= 0 && xn < vol.W && yn >= 0 && yn < vol.H && zn >= 0 && zn < vol.D;
        const int off = in ? (zn * (int)vol.sD + yn * (int)vol.sH + xn * (int)vol.sW) : 0;
#pragma unroll 4
        for (int c = 0; c < vol.C; ++c) {
            T v = T(0);
            if (in) v = __ldg(src + (long long)c * vol.sC + off);
            dst[c * plane] = v;
        }
        return;
    }

    const Corners cn = corners_of(sm, vol);
    const float pad = pad_of(pad_mode, pad_value, pad_device);
#pragma unroll 2
    for (int c = 0; c < vol.C; ++c) {
        const T* __restrict__ sc = src + (long long)c * vol.sC;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = cn.in(k) ? Store<T>::load(sc + cn.off(k, vol)) : pad;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (cn.in(k)) acc = __fadd_rn(acc, __fmul_rn(__fsub_rn(v[k], pad), cn.w(k)));
        dst[c * plane] = Store<T>::from_float(__fadd_rn(acc, pad));
    }
}

// ------------------------------------------------------------------------------------------------
// forward, channels-last volumes: VB-byte vector gathers (VB = 32: one LDG.256 per corner of an 8-channel fp32
// voxel; VB = 16: LDG.128).  Arithmetic per channel is identical to the generic kernel (bitwise the same results).
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE, int VB, int BATCH = 8>
__global__ void __launch_bounds__(NTHREADS, 3)
slice_fwd_cl_kernel(VolArgs vol, ViewArgs va, OutGeom g, int pad_mode, float pad_value, const float* pad_device,
                    T* __restrict__ out) {
    const int s = blockIdx.z * va.V + blockIdx.y;      // grid = (tiles, V, B): no integer division for the batch index
    const Pix p = pixel_of_thread(g);
    if (!p.valid) return;
    const Sample sm = sample_coords(g, p, va, s, vol);
    const int b = blockIdx.z;
    const T* __restrict__ src = (const T*)vol.data + (long long)b * vol.sB;
    const int plane = g.Do * g.Ho * g.Wo;      // C * plane < 2^31 (checked on the host): 32-bit channel offsets
    T* __restrict__ dst = out + (size_t)s * (size_t)(vol.C * plane) + ((p.i * g.Ho + p.j) * g.Wo + p.k);
    constexpr int N = VB / (int)sizeof(T);

    if (MODE == AFB_NEAREST) {
        const int xn = __float2int_rn(sm.ix), yn = __float2int_rn(sm.iy), zn = __float2int_rn(sm.iz);
        const bool in = xn >= 0 && xn < vol.W && yn >= 0 && yn < vol.H && zn >= 0 && zn < vol.D;
        const int off = in ? (zn * (int)vol.sD + yn * (int)vol.sH + xn * (int)vol.sW) : 0;
        for (int c0 = 0; c0 < vol.C; c0 += N) {
            Raw<VB> raw = raw_zero<VB>();
            if (in) raw = gather_nc<VB>(src + off + c0);
            const T* e = reinterpret_cast<const T*>(raw.w);
#pragma unroll
            for (int q = 0; q < N; ++q) dst[(c0 + q) * plane] = e[q];
        }
        return;
    }
    if constexpr (Widen<T>::is_float && MODE == AFB_BILINEAR) {
        // fp32 coordinates, weights and accumulation whatever the storage type
        const Corners cn = corners_of(sm, vol);
        const float pad = pad_of(pad_mode, pad_value, pad_device);
        for (int c0 = 0; c0 < vol.C; c0 += N) {
            float acc[N];
#pragma unroll
            for (int q = 0; q < N; ++q) acc[q] = 0.0f;
            // BATCH corners gathered at a time (8: all in flight; 4: z0 plane then z1 plane, half the registers);
            // the accumulation order k = 0..7 (ATen's) is the same either way
#pragma unroll
            for (int h = 0; h < 8; h += BATCH) {
                Raw<VB> raw[BATCH];
#pragma unroll
                for (int k = 0; k < BATCH; ++k)
                    raw[k] = cn.in(h + k) ? gather_nc<VB>(src + cn.off(h + k, vol) + c0) : raw_zero<VB>();
#pragma unroll
                for (int k = 0; k < BATCH; ++k)
                    if (cn.in(h + k)) {
                        float vv[N];
                        Widen<T>::template decode<VB>(raw[k], vv);
                        const float wk = cn.w(h + k);
#pragma unroll
                        for (int q = 0; q < N; ++q) acc[q] = __fadd_rn(acc[q], __fmul_rn(__fsub_rn(vv[q], pad), wk));
                    }
            }
#pragma unroll
            for (int q = 0; q < N; ++q) dst[(c0 + q) * plane] = Store<T>::from_float(__fadd_rn(acc[q], pad));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward, the three slicings of one acquisition in ONE launch (models/learnable_transform.py:287-306: soft label bilinear,
// one-hot label nearest, image bilinear with the same pre-affine): the coordinates, corners and weights of an output
// location are computed once and used for three gathers.  soft: channels-last fp32 (LDG.128 per 4 channels and corner),
// label: channels-last integers (16-byte vectors, nearest), image: any strides (scalar gathers, loop over its channels).
// Per-volume arithmetic is the same sequence as slice_fwd_cl_kernel / slice_fwd_kernel: bitwise the same results.
// ------------------------------------------------------------------------------------------------
template <typename LT>
__global__ void __launch_bounds__(NTHREADS, 3)
slice_fwd3_kernel(VolArgs soft, VolArgs lab, VolArgs img, ViewArgs va, OutGeom g,
                  int pad_mode_s, float pad_value_s, const float* pad_device_s,
                  int pad_mode_i, float pad_value_i, const float* pad_device_i,
                  float* __restrict__ y_soft, LT* __restrict__ y_lab, float* __restrict__ y_img) {
    const int s = blockIdx.z * va.V + blockIdx.y;
    const Pix p = pixel_of_thread(g);
    if (!p.valid) return;
    const Sample sm = sample_coords(g, p, va, s, soft);              // the three volumes share D, H, W (checked on the host)
    const int b = blockIdx.z;
    const int plane = g.Do * g.Ho * g.Wo;
    const int pix = (p.i * g.Ho + p.j) * g.Wo + p.k;
    // ---- label: nearest ----
    if (y_lab) {
        constexpr int NL = 16 / (int)sizeof(LT);
        const int xn = __float2int_rn(sm.ix), yn = __float2int_rn(sm.iy), zn = __float2int_rn(sm.iz);
        const bool in = xn >= 0 && xn < lab.W && yn >= 0 && yn < lab.H && zn >= 0 && zn < lab.D;
        const LT* __restrict__ src = (const LT*)lab.data + (long long)b * lab.sB;
        const int off = in ? (zn * (int)lab.sD + yn * (int)lab.sH + xn * (int)lab.sW) : 0;
        LT* __restrict__ dst = y_lab + (size_t)s * (size_t)(lab.C * plane) + pix;
        for (int c0 = 0; c0 < lab.C; c0 += NL) {
            Raw<16> raw = raw_zero<16>();
            if (in) raw = gather_nc<16>(src + off + c0);
            const LT* e = reinterpret_cast<const LT*>(raw.w);
#pragma unroll
            for (int q = 0; q < NL; ++q) dst[(c0 + q) * plane] = e[q];
        }
    }
    const Corners cn = corners_of(sm, soft);
    // ---- image: bilinear, generic strides ----
    if (y_img) {
        const float pad = pad_of(pad_mode_i, pad_value_i, pad_device_i);
        const float* __restrict__ src = (const float*)img.data + (long long)b * img.sB;
        float* __restrict__ dst = y_img + (size_t)s * (size_t)(img.C * plane) + pix;
        // corner offsets in the image's own strides
        const int x0 = max(-2, min(__float2int_rd(sm.ix), img.W + 1));
        const int y0 = max(-2, min(__float2int_rd(sm.iy), img.H + 1));
        const int z0 = max(-2, min(__float2int_rd(sm.iz), img.D + 1));
        const int ibase = z0 * (int)img.sD + y0 * (int)img.sH + x0 * (int)img.sW;
        for (int c = 0; c < img.C; ++c) {
            const float* __restrict__ sc = src + (long long)c * img.sC;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int o = ibase + ((k & 1) ? (int)img.sW : 0) + (((k >> 1) & 1) ? (int)img.sH : 0) + ((k >> 2) ? (int)img.sD : 0);
                v[k] = cn.in(k) ? __ldg(sc + o) : pad;
            }
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (cn.in(k)) acc = __fadd_rn(acc, __fmul_rn(__fsub_rn(v[k], pad), cn.w(k)));
            dst[c * plane] = __fadd_rn(acc, pad);
        }
    }
    // ---- soft label: bilinear, channels-last fp32 ----
    {
        const float pad = pad_of(pad_mode_s, pad_value_s, pad_device_s);
        const float* __restrict__ src = (const float*)soft.data + (long long)b * soft.sB;
        float* __restrict__ dst = y_soft + (size_t)s * (size_t)(soft.C * plane) + pix;
        for (int c0 = 0; c0 < soft.C; c0 += 4) {
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            Raw<16> raw[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) raw[k] = cn.in(k) ? gather_nc<16>(src + cn.off(k, soft) + c0) : raw_zero<16>();
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (cn.in(k)) {
                    const float wk = cn.w(k);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        acc[q] = __fadd_rn(acc[q], __fmul_rn(__fsub_rn(__uint_as_float(raw[k].w[q]), pad), wk));
                }
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[(c0 + q) * plane] = __fadd_rn(acc[q], pad);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward: re-gather, dVolume scatter (RED), dGrid -> 12 sums of dgrid (x) base per slice.
// CTA reduction (shuffle -> smem) then one fp64 atomic per sum per CTA into the per-slice workspace;
// view_chain_kernel (afb_views.cu) turns the totals into the gradient of the view input.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NTHREADS, 2)
slice_bwd_kernel(VolArgs vol, ViewArgs va, OutGeom g, int pad_mode, float pad_value, const float* pad_device,
                 const float* __restrict__ grad_out, float* __restrict__ d_vol, float* __restrict__ d_pad,
                 double* __restrict__ ws_acc) {
    const int s = blockIdx.z * va.V + blockIdx.y;      // grid = (tiles, V, B): no integer division for the batch index
    float part[13];
#pragma unroll
    for (int q = 0; q < 13; ++q) part[q] = 0.0f;
    const Pix p = pixel_of_thread(g);
    if (p.valid) {
        const Sample sm = sample_coords(g, p, va, s, vol);
        const Corners cn = corners_of(sm, vol);
        const float pad = pad_of(pad_mode, pad_value, pad_device);
        const int b = blockIdx.z;
        const T* __restrict__ src = (const T*)vol.data + (long long)b * vol.sB;
        float* __restrict__ dv = d_vol ? d_vol + (long long)b * vol.sB : nullptr;
        const int plane = g.Do * g.Ho * g.Wo;      // C * plane < 2^31 (checked on the host): 32-bit channel offsets
        const float* __restrict__ go_p = grad_out + (size_t)s * (size_t)(vol.C * plane) + ((p.i * g.Ho + p.j) * g.Wo + p.k);
        float dot[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) dot[k] = 0.0f;
        float gsum = 0.0f;
#pragma unroll 2
        for (int c = 0; c < vol.C; ++c) {
            const float go = __ldg(go_p + c * plane);
            const long long coff = (long long)c * vol.sC;
            gsum += go;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (cn.in(k)) {
                    const float v = Store<T>::load(src + coff + cn.off(k, vol)) - pad;
                    dot[k] = fmaf(v, go, dot[k]);
                    if (dv) atomicAdd(dv + coff + cn.off(k, vol), cn.w(k) * go);
                }
            }
        }
        grid_grad_parts(dot, cn, sm, vol, gsum, part);
    }
    bwd_reduce(s, part, pad_mode, d_pad, ws_acc);
}

// channels-last volumes: VB-byte gathers (LDG.256 for 8 fp32 channels) and 16-byte vector reductions
// (red.global.add.v4.f32 - the widest vector RED sm_100 has)
template <typename T, int VB, int BATCH, int MINB = 2>
__global__ void __launch_bounds__(NTHREADS, MINB)
slice_bwd_cl_kernel(VolArgs vol, ViewArgs va, OutGeom g, int pad_mode, float pad_value, const float* pad_device,
                    const float* __restrict__ grad_out, float* __restrict__ d_vol, float* __restrict__ d_pad,
                    double* __restrict__ ws_acc) {
    constexpr int N = VB / (int)sizeof(T);
    const int s = blockIdx.z * va.V + blockIdx.y;      // grid = (tiles, V, B): no integer division for the batch index
    float part[13];
#pragma unroll
    for (int q = 0; q < 13; ++q) part[q] = 0.0f;
    const Pix p = pixel_of_thread(g);
    if (p.valid) {
        const Sample sm = sample_coords(g, p, va, s, vol);
        const Corners cn = corners_of(sm, vol);
        const float pad = pad_of(pad_mode, pad_value, pad_device);
        const int b = blockIdx.z;
        const T* __restrict__ src = (const T*)vol.data + (long long)b * vol.sB;
        float* __restrict__ dv = d_vol ? d_vol + (long long)b * vol.sB : nullptr;
        const int plane = g.Do * g.Ho * g.Wo;      // C * plane < 2^31 (checked on the host): 32-bit channel offsets
        const float* __restrict__ go_p = grad_out + (size_t)s * (size_t)(vol.C * plane) + ((p.i * g.Ho + p.j) * g.Wo + p.k);
        float dot[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) dot[k] = 0.0f;
        float gsum = 0.0f;
        for (int c0 = 0; c0 < vol.C; c0 += N) {
            float go[N];
#pragma unroll
            for (int q = 0; q < N; ++q) { go[q] = __ldg(go_p + (c0 + q) * plane); gsum += go[q]; }
            // two batches of 4 corners (z0 plane, z1 plane): 4 vector gathers in flight per thread, half the registers
#pragma unroll
            for (int h = 0; h < 8; h += BATCH) {
                Raw<VB> raw[BATCH];
#pragma unroll
                for (int k = 0; k < BATCH; ++k)
                    raw[k] = cn.in(h + k) ? gather_nc<VB>(src + cn.off(h + k, vol) + c0) : raw_zero<VB>();
#pragma unroll
                for (int k = 0; k < BATCH; ++k) {
                    if (cn.in(h + k)) {
                        float vv[N];
                        Widen<T>::template decode<VB>(raw[k], vv);
#pragma unroll
                        for (int q = 0; q < N; ++q) dot[h + k] = fmaf(vv[q] - pad, go[q], dot[h + k]);
                        if (dv) {
                            const float wk = cn.w(h + k);
#pragma unroll
                            for (int q = 0; q < N; q += 4)
                                atomicAdd(reinterpret_cast<float4*>(dv + cn.off(h + k, vol) + c0 + q),
                                          make_float4(wk * go[q], wk * go[q + 1], wk * go[q + 2], wk * go[q + 3]));
                        }
                    }
                }
            }
        }
        grid_grad_parts(dot, cn, sm, vol, gsum, part);
    }
    bwd_reduce(s, part, pad_mode, d_pad, ws_acc);
}

// ------------------------------------------------------------------------------------------------
// scatter-only half of the backward: dVolume += w_k * grad_out at the 8 corners.  Needs the geometry and grad_out only
// (no volume gather), so the backward can be split: the dTheta half (slice_bwd* with d_vol == NULL: re-gather + dot
// products + reduction) is independent of the MinBackward fill and runs on a side stream UNDER it, this half follows the
// fill.  Issue-bound on the REDs: 16 x red.global.add.v4.f32 per pixel for C = 8.
// ------------------------------------------------------------------------------------------------
template <bool CL>
__global__ void __launch_bounds__(NTHREADS, 4)
slice_scatter_kernel(VolArgs vol, ViewArgs va, OutGeom g, const float* __restrict__ grad_out, float* __restrict__ d_vol) {
    const int s = blockIdx.z * va.V + blockIdx.y;
    const Pix p = pixel_of_thread(g);
    if (!p.valid) return;
    const Sample sm = sample_coords(g, p, va, s, vol);
    const Corners cn = corners_of(sm, vol);
    if (cn.inb == 0u) return;
    float* __restrict__ dv = d_vol + (long long)blockIdx.z * vol.sB;
    const int plane = g.Do * g.Ho * g.Wo;
    const float* __restrict__ go_p = grad_out + (size_t)s * (size_t)(vol.C * plane) + ((p.i * g.Ho + p.j) * g.Wo + p.k);
    float w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = cn.w(k);
    if (CL) {
        for (int c0 = 0; c0 < vol.C; c0 += 4) {
            float go[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) go[q] = __ldg(go_p + (c0 + q) * plane);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (cn.in(k))
                    atomicAdd(reinterpret_cast<float4*>(dv + cn.off(k, vol) + c0),
                              make_float4(w[k] * go[0], w[k] * go[1], w[k] * go[2], w[k] * go[3]));
        }
    } else {
        for (int c = 0; c < vol.C; ++c) {
            const float go = __ldg(go_p + c * plane);
            const long long coff = (long long)c * vol.sC;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (cn.in(k)) atomicAdd(dv + coff + cn.off(k, vol), w[k] * go);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// EXPERIMENT (profiles/ab_kernels.py, not on the product path): shared-memory privatisation of the dVolume scatter.
// An UPPER BOUND on what privatisation could buy: every corner update is accumulated into a CTA-private table with
// shared-memory atomics (channel-major layout, so that the lanes of a warp hit different banks), collisions of the
// direct-mapped table are ignored (two voxels may share a slot: results are WRONG on purpose - the point is the cost of the
// 64 shared atomics per pixel that any exact scheme also pays), then the table is flushed with coalesced vector REDs.
// Measured on the B200 against slice_scatter_kernel<true> (16 x red.global.add.v4.f32 per pixel): see DESIGN.md section 5.
// ------------------------------------------------------------------------------------------------
constexpr int PRIV_SLOTS = 1024;

__global__ void __launch_bounds__(NTHREADS, 3)
slice_scatter_priv_probe_kernel(VolArgs vol, ViewArgs va, OutGeom g, const float* __restrict__ grad_out, float* __restrict__ d_vol) {
    __shared__ float table[8][PRIV_SLOTS];                     // [channel][slot]: 32 KB
    for (int i = threadIdx.x; i < 8 * PRIV_SLOTS; i += NTHREADS) (&table[0][0])[i] = 0.0f;
    __syncthreads();
    const int s = blockIdx.z * va.V + blockIdx.y;
    const Pix p = pixel_of_thread(g);
    if (p.valid) {
        const Sample sm = sample_coords(g, p, va, s, vol);
        const Corners cn = corners_of(sm, vol);
        const int plane = g.Do * g.Ho * g.Wo;
        const float* __restrict__ go_p = grad_out + (size_t)s * (size_t)(vol.C * plane) + ((p.i * g.Ho + p.j) * g.Wo + p.k);
        float go[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) go[q] = __ldg(go_p + q * plane);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (cn.in(k)) {
                const float wk = cn.w(k);
                const unsigned vox = (unsigned)cn.off(k, vol) >> 3;          // voxel index (C = 8 channels-last)
                const int slot = (int)((vox ^ (vox >> 10) ^ (vox >> 20)) & (PRIV_SLOTS - 1));
#pragma unroll
                for (int q = 0; q < 8; ++q) atomicAdd(&table[q][slot], wk * go[q]);
            }
    }
    __syncthreads();
    // flush: one slot per thread and pass, two 16-byte REDs (coalesced: far cheaper than the real, scattered flush would be)
    float* __restrict__ dv = d_vol + (long long)blockIdx.z * vol.sB + (size_t)(blockIdx.x % 1024) * PRIV_SLOTS * 8;
    for (int sl = threadIdx.x; sl < PRIV_SLOTS; sl += NTHREADS) {
        const float4 a = make_float4(table[0][sl], table[1][sl], table[2][sl], table[3][sl]);
        const float4 b = make_float4(table[4][sl], table[5][sl], table[6][sl], table[7][sl]);
        if (a.x != 0.0f || a.y != 0.0f || b.x != 0.0f) {
            atomicAdd(reinterpret_cast<float4*>(dv + sl * 8), a);
            atomicAdd(reinterpret_cast<float4*>(dv + sl * 8 + 4), b);
        }
    }
}

// d(out)/d(pad) = sum go * (1 - sum of in-bounds weights): depends on geometry and grad_out only, so it can run
// BEFORE the dVolume fill, which lets MinBackward be fused with the zero-fill (afb_min_grad_fill).
// A pure stream over grad_out: each CTA covers PAD_TILES tiles with all their loads independent (enough bytes in
// flight to reach HBM speed) and issues one atomic for all of them.
constexpr int PAD_TILES = 4;

__global__ void __launch_bounds__(NTHREADS, 4)
slice_pad_grad_kernel(VolArgs vol, ViewArgs va, OutGeom g, int ntiles, const float* __restrict__ grad_out,
                      float* __restrict__ d_pad) {
    __shared__ float red[NTHREADS / 32];
    const int s = blockIdx.z * va.V + blockIdx.y;      // grid = (tiles, V, B): no integer division for the batch index
    const int plane = g.Do * g.Ho * g.Wo;      // C * plane < 2^31 (checked on the host): 32-bit channel offsets
    const float* __restrict__ go_s = grad_out + (size_t)s * (size_t)(vol.C * plane);
    // phase 1: request every grad_out element of the PAD_TILES tiles (predicated, no control flow between the loads)
    Pix px[PAD_TILES];
    float gsum[PAD_TILES];
    if (vol.C == 8) {
        float v[PAD_TILES][8];
#pragma unroll
        for (int u = 0; u < PAD_TILES; ++u) {
            const int tile = blockIdx.x * PAD_TILES + u;
            px[u] = pixel_of_tile(g, tile < ntiles ? tile : 0);
            px[u].valid = px[u].valid && tile < ntiles;
            const float* __restrict__ go_p = go_s + ((size_t)px[u].i * g.Ho + px[u].j) * g.Wo + px[u].k;
#pragma unroll
            for (int c = 0; c < 8; ++c) v[u][c] = px[u].valid ? __ldg(go_p + c * plane) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < PAD_TILES; ++u) {
            gsum[u] = 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) gsum[u] += v[u][c];
        }
    } else {
#pragma unroll
        for (int u = 0; u < PAD_TILES; ++u) {
            const int tile = blockIdx.x * PAD_TILES + u;
            px[u] = pixel_of_tile(g, tile < ntiles ? tile : 0);
            px[u].valid = px[u].valid && tile < ntiles;
            const float* __restrict__ go_p = go_s + ((size_t)px[u].i * g.Ho + px[u].j) * g.Wo + px[u].k;
            gsum[u] = 0.0f;
            if (px[u].valid)
                for (int c = 0; c < vol.C; ++c) gsum[u] += __ldg(go_p + c * plane);
        }
    }
    // phase 2: geometry (no memory traffic besides the 12 floats of the slice's grid affine)
    float part = 0.0f;
#pragma unroll
    for (int u = 0; u < PAD_TILES; ++u) {
        if (!px[u].valid) continue;
        const Sample sm = sample_coords(g, px[u], va, s, vol);
        const Corners cn = corners_of(sm, vol);
        float wsum = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) wsum += cn.in(k) ? cn.w(k) : 0.0f;
        part += gsum[u] * (1.0f - wsum);
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < NTHREADS / 32; ++w) t += (double)red[w];
        if (t != 0.0) atomicAdd(d_pad, (float)t);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int make_args(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo, VolArgs& v, ViewArgs& a) {
    if (!vol || !vol->data) return AFB_EINVAL;
    if (vol->C <= 0) return AFB_ESHAPE;
    v.data = vol->data; v.B = vol->B; v.C = vol->C; v.D = vol->D; v.H = vol->H; v.W = vol->W;
    v.sB = vol->sB; v.sC = vol->sC; v.sD = vol->sD; v.sH = vol->sH; v.sW = vol->sW;
    // corner offsets inside one (batch, channel) sub-volume are 32-bit in the kernels
    if (vol->sD < 0 || vol->sH < 0 || vol->sW < 0) return AFB_EUNSUPPORTED;
    if ((long long)(vol->D + 2) * vol->sD + (long long)(vol->H + 2) * vol->sH + (long long)(vol->W + 2) * vol->sW >= 2147483647ll)
        return AFB_EUNSUPPORTED;
    // channel offsets inside one slice's output are 32-bit in the kernels
    if ((long long)vol->C * Do * Ho * Wo >= 2147483647ll) return AFB_EUNSUPPORTED;
    return make_view_args(views, vol->B, vol->D, vol->H, vol->W, Do, Ho, Wo, /*need_state=*/true, a);
}

// channels-last fast path: channels contiguous, every voxel's channel vector VB-byte aligned (n = VB / sizeof(T)
// elements, at most 8 per vector).  Returns the widest usable vector width in bytes (<= max_vb) or 0.
static int channels_last_vec(const afb_volume* vol, int elem_bytes, const void* extra_ptr, int max_vb) {
    if (vol->sC != 1) return 0;
    for (int vb = max_vb; vb >= 16; vb >>= 1) {
        const int n = vb / elem_bytes;
        if (n < 1 || n > 8 || vol->C % n != 0) continue;
        if (((uintptr_t)vol->data % vb) || ((uintptr_t)extra_ptr % 16)) continue;
        if (vol->sW % n || vol->sH % n || vol->sD % n || vol->sB % n) continue;
        return vb;
    }
    return 0;
}

// Forward gathers stay at 16 bytes (LDG.128): measured on the B200, one LDG.256 per corner is SLOWER in the forward
// (0.340 vs 0.303 ms soft labels, 0.245 vs 0.217 ms int64 labels at 384 slices) - the kernel is bound by DRAM-miss
// latency x loads in flight, and the wider vectors cost a resident CTA per SM.  The backward keeps LDG.256.
template <typename T>
static bool Do3d_vb32(const afb_volume* vol, const OutGeom& g) {
    if (sizeof(T) != 4 || !Widen<T>::is_float || g.Do <= 1) return false;
    const char* e = getenv("AFB_FWD3D_VB");          // A/B knob (profiles/ab_f1_resample.py): 16 = the slices' LDG.128 pairs
    if (e && atoi(e) == 16) return false;
    return channels_last_vec(vol, 4, nullptr, 32) == 32;
}

template <typename T>
static int launch_fwd(const afb_volume* vol, const VolArgs& v, const ViewArgs& a, const OutGeom& g, int mode, int pad_mode,
                      float pad_value, const float* pad_device, void* out, cudaStream_t st) {
    const int S = v.B * a.V;
    const dim3 grid = slice_grid(g, v.B, a.V);
    const int vb = (mode == AFB_NEAREST || Widen<T>::is_float) ? channels_last_vec(vol, (int)sizeof(T), nullptr, 16) : 0;
    // 3-D outputs (Do > 1: the 128^3 -> 128^3 prescan resample, learnable_transform.py:252-255) of fp32 volumes with C % 8 == 0:
    // one LDG.256 per corner and lane, 4 corners in flight.  Unlike the single slices (see above) the resample samples every
    // input voxel ~8 times from L1/L2, it is bound by load instructions and L1 wavefronts rather than by DRAM-miss latency:
    // measured 0.675 vs 0.79 ms at B = 8 (profiles/r2_ab_f1_resample.json), bitwise the same output
    if (vb == 16 && mode == AFB_BILINEAR && Do3d_vb32<T>(vol, g)) {
        slice_fwd_cl_kernel<T, AFB_BILINEAR, Widen<T>::is_float && sizeof(T) == 4 ? 32 : 16, 4><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out);
    } else
    if (vb == 16 && mode == AFB_NEAREST) {
        slice_fwd_cl_kernel<T, AFB_NEAREST, 16><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out);
    } else if (vb == 16) {
        slice_fwd_cl_kernel<T, AFB_BILINEAR, 16><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out);
    } else if (mode == AFB_NEAREST) {
        slice_fwd_kernel<T, AFB_NEAREST><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out);
    } else {
        slice_fwd_kernel<T, AFB_BILINEAR><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out);
    }
    return (int)cudaGetLastError();
}

template <typename T>
static int launch_bwd(const afb_volume* vol, const VolArgs& v, const ViewArgs& a, const OutGeom& g, int pad_mode, float pad_value,
                      const float* pad_device, const float* grad_out, float* d_vol, float* d_pad, double* acc, cudaStream_t st) {
    const dim3 grid = slice_grid(g, v.B, a.V);
    const int vb = channels_last_vec(vol, (int)sizeof(T), d_vol, 32);
    // measured on the B200 (384 slices, C = 8 fp32): LDG.256 x 4 corners in flight 0.689 ms, LDG.128 x 4: 0.718,
    // LDG.128 x 8 (the earlier kernel): 0.702
    constexpr int WIDE = sizeof(T) >= 4 ? 32 : 16;
#define AFB_CLB(VB) slice_bwd_cl_kernel<T, VB, 4><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, grad_out, d_vol, d_pad, acc)
    // A/B knob for profiles/ab_kernels.py (fp32 only): 1 = <=85 registers / 3 CTAs per SM, 2 = LDG.128 gathers at 3 CTAs per SM
    const char* ev = getenv("AFB_BWD_VARIANT");
    const int variant = ev ? atoi(ev) : 0;
    if (variant != 0 && vb == 32 && std::is_same<T, float>::value) {
        if (variant == 1) slice_bwd_cl_kernel<T, WIDE, 4, 3><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, grad_out, d_vol, d_pad, acc);
        else slice_bwd_cl_kernel<T, 16, 4, 3><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, grad_out, d_vol, d_pad, acc);
    } else
    if (vb == 32) AFB_CLB(WIDE);
    else if (vb == 16) AFB_CLB(16);
#undef AFB_CLB
    else
        slice_bwd_kernel<T><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, grad_out, d_vol, d_pad, acc);
    return (int)cudaGetLastError();
}

}  // namespace afb

using namespace afb;

extern "C" int afb_slice_fwd(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo, int mode,
                             int pad_mode, float pad_value, const float* pad_device, void* out, void* stream) {
    VolArgs v; ViewArgs a;
    int rc = make_args(vol, views, Do, Ho, Wo, v, a);
    if (rc != AFB_OK) return rc;
    if (!out) return AFB_EINVAL;
    if (mode != AFB_BILINEAR && mode != AFB_NEAREST) return AFB_EINVAL;
    if (pad_mode < AFB_PAD_ZERO || pad_mode > AFB_PAD_DEVICE) return AFB_EINVAL;
    if (pad_mode == AFB_PAD_DEVICE && !pad_device) return AFB_EINVAL;
    const OutGeom g = make_geom(Do, Ho, Wo);
    cudaStream_t st = (cudaStream_t)stream;
#define AFB_FWD(T) return launch_fwd<T>(vol, v, a, g, mode, pad_mode, pad_value, pad_device, out, st)
    switch (vol->dtype) {
        case AFB_F32: AFB_FWD(float);
        case AFB_BF16: AFB_FWD(__nv_bfloat16);
        case AFB_F16: AFB_FWD(__half);
        case AFB_I64: AFB_FWD(int64_t);
        case AFB_I32: AFB_FWD(int32_t);
        case AFB_I16: AFB_FWD(int16_t);
        case AFB_U8: AFB_FWD(uint8_t);
        default: return AFB_EDTYPE;
    }
#undef AFB_FWD
}

extern "C" int64_t afb_slice_bwd_workspace_bytes(int S) { return (int64_t)S * 16 * sizeof(double); }

extern "C" int afb_slice_bwd(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo,
                             int pad_mode, float pad_value, const float* pad_device,
                             const float* grad_out, const float* grad_grid_affine,
                             float* d_vol, float* d_affine, float* d_gpre, float* d_pad,
                             void* workspace, void* stream) {
    VolArgs v; ViewArgs a;
    int rc = make_args(vol, views, Do, Ho, Wo, v, a);
    if (rc != AFB_OK) return rc;
    if (!workspace || (!grad_out && !grad_grid_affine)) return AFB_EINVAL;
    if (pad_mode < AFB_PAD_ZERO || pad_mode > AFB_PAD_DEVICE) return AFB_EINVAL;
    if (pad_mode == AFB_PAD_DEVICE && !pad_device) return AFB_EINVAL;
    if (a.kind == AFB_AFFINE_PARAMS && d_affine && !views->params) return AFB_EINVAL;
    const int S = v.B * a.V;
    const OutGeom g = make_geom(Do, Ho, Wo);
    double* acc = (double*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_out) {     // grad_out == NULL: chain-only (nearest / integer volumes): only the upstream grad is propagated
        switch (vol->dtype) {
            case AFB_F32: rc = launch_bwd<float>(vol, v, a, g, pad_mode, pad_value, pad_device, grad_out, d_vol, d_pad, acc, st); break;
            case AFB_BF16: rc = launch_bwd<__nv_bfloat16>(vol, v, a, g, pad_mode, pad_value, pad_device, grad_out, d_vol, d_pad, acc, st); break;
            case AFB_F16: rc = launch_bwd<__half>(vol, v, a, g, pad_mode, pad_value, pad_device, grad_out, d_vol, d_pad, acc, st); break;
            default: return AFB_EDTYPE;
        }
        if (rc != 0) return rc;
    }
    if (!d_affine && !d_gpre) {
        // nobody wants the view gradient: just re-zero the workspace the sampler accumulated into
        return grad_out ? (int)cudaMemsetAsync(acc, 0, (size_t)S * 16 * sizeof(double), st) : AFB_OK;
    }
    return launch_view_chain(a, S, acc, grad_grid_affine, d_affine, d_gpre, st);
}

extern "C" int afb_slice_pad_grad(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo,
                                  const float* grad_out, float* d_pad, void* stream) {
    VolArgs v; ViewArgs a;
    int rc = make_args(vol, views, Do, Ho, Wo, v, a);
    if (rc != AFB_OK) return rc;
    if (!grad_out || !d_pad) return AFB_EINVAL;
    const OutGeom g = make_geom(Do, Ho, Wo);
    const dim3 tiles = slice_grid(g, v.B, a.V);
    const dim3 grid((tiles.x + PAD_TILES - 1) / PAD_TILES, tiles.y, tiles.z);
    slice_pad_grad_kernel<<<grid, NTHREADS, 0, (cudaStream_t)stream>>>(v, a, g, (int)tiles.x, grad_out, d_pad);
    return (int)cudaGetLastError();
}

extern "C" int afb_slice_scatter(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo, const float* grad_out,
                                 float* d_vol, void* stream) {
    VolArgs v; ViewArgs a;
    int rc = make_args(vol, views, Do, Ho, Wo, v, a);
    if (rc != AFB_OK) return rc;
    if (!grad_out || !d_vol) return AFB_EINVAL;
    const OutGeom g = make_geom(Do, Ho, Wo);
    const dim3 grid = slice_grid(g, v.B, a.V);
    cudaStream_t st = (cudaStream_t)stream;
    // dVolume is fp32 whatever the storage type of the volume; 16-byte vector REDs need 4-channel groups, 16-byte aligned
    const bool cl = vol->sC == 1 && vol->C % 4 == 0 && ((uintptr_t)d_vol % 16) == 0 && vol->sW % 4 == 0 && vol->sH % 4 == 0 &&
                    vol->sD % 4 == 0 && vol->sB % 4 == 0;
    if (cl) slice_scatter_kernel<true><<<grid, NTHREADS, 0, st>>>(v, a, g, grad_out, d_vol);
    else slice_scatter_kernel<false><<<grid, NTHREADS, 0, st>>>(v, a, g, grad_out, d_vol);
    return (int)cudaGetLastError();
}


extern "C" int afb_slice_fwd3(const afb_volume* soft, const afb_volume* label, const afb_volume* image, const afb_views* views,
                              int Do, int Ho, int Wo, int pad_mode_soft, float pad_value_soft, const float* pad_device_soft,
                              int pad_mode_image, float pad_value_image, const float* pad_device_image,
                              float* y_soft, void* y_label, float* y_image, void* stream) {
    VolArgs vs, vl, vi; ViewArgs a, a2;
    int rc = make_args(soft, views, Do, Ho, Wo, vs, a);
    if (rc != AFB_OK) return rc;
    if (!y_soft || soft->dtype != AFB_F32) return soft && soft->dtype != AFB_F32 ? AFB_EDTYPE : AFB_EINVAL;
    if (channels_last_vec(soft, 4, nullptr, 16) != 16) return AFB_EUNSUPPORTED;      // caller falls back to afb_slice_fwd x3
    for (int w = 0; w < 2; ++w) {
        const int pm = w ? pad_mode_image : pad_mode_soft;
        const float* pd = w ? pad_device_image : pad_device_soft;
        if (pm < AFB_PAD_ZERO || pm > AFB_PAD_DEVICE || (pm == AFB_PAD_DEVICE && !pd)) return AFB_EINVAL;
    }
    vl = vs; vi = vs;
    if (label) {
        if (!y_label) return AFB_EINVAL;
        rc = make_args(label, views, Do, Ho, Wo, vl, a2);
        if (rc != AFB_OK) return rc;
        if (label->B != soft->B || label->D != soft->D || label->H != soft->H || label->W != soft->W) return AFB_ESHAPE;
        const int eb = label->dtype == AFB_I64 ? 8 : label->dtype == AFB_I32 ? 4 : label->dtype == AFB_I16 ? 2 : label->dtype == AFB_U8 ? 1 : 0;
        if (!eb) return AFB_EDTYPE;
        const int nl = 16 / eb;            // elements per 16-byte channel vector (up to 16 for u8: own check, not channels_last_vec)
        if (label->sC != 1 || label->C % nl || ((uintptr_t)label->data % 16) || label->sW % nl || label->sH % nl || label->sD % nl ||
            label->sB % nl)
            return AFB_EUNSUPPORTED;
    }
    if (image) {
        if (!y_image) return AFB_EINVAL;
        rc = make_args(image, views, Do, Ho, Wo, vi, a2);
        if (rc != AFB_OK) return rc;
        if (image->dtype != AFB_F32) return AFB_EUNSUPPORTED;
        if (image->B != soft->B || image->D != soft->D || image->H != soft->H || image->W != soft->W) return AFB_ESHAPE;
    }
    const OutGeom g = make_geom(Do, Ho, Wo);
    const dim3 grid = slice_grid(g, vs.B, a.V);
    cudaStream_t st = (cudaStream_t)stream;
    float* yi = image ? y_image : nullptr;
#define AFB_F3(LT) slice_fwd3_kernel<LT><<<grid, NTHREADS, 0, st>>>(vs, vl, vi, a, g, pad_mode_soft, pad_value_soft, pad_device_soft, \
        pad_mode_image, pad_value_image, pad_device_image, y_soft, label ? (LT*)y_label : (LT*)nullptr, yi)
    switch (label ? label->dtype : AFB_I64) {
        case AFB_I64: AFB_F3(int64_t); break;
        case AFB_I32: AFB_F3(int32_t); break;
        case AFB_I16: AFB_F3(int16_t); break;
        case AFB_U8: AFB_F3(uint8_t); break;
        default: return AFB_EDTYPE;
    }
#undef AFB_F3
    return (int)cudaGetLastError();
}


/* EXPERIMENT entry (not declared in afb200.h): cost of shared-memory privatisation of the scatter, see the kernel's comment. */
extern "C" int afbx_slice_scatter_priv_probe(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo,
                                             const float* grad_out, float* d_vol, void* stream) {
    VolArgs v; ViewArgs a;
    int rc = make_args(vol, views, Do, Ho, Wo, v, a);
    if (rc != AFB_OK) return rc;
    if (!grad_out || !d_vol || vol->C != 8 || vol->sC != 1) return AFB_EINVAL;
    const OutGeom g = make_geom(Do, Ho, Wo);
    slice_scatter_priv_probe_kernel<<<slice_grid(g, v.B, a.V), NTHREADS, 0, (cudaStream_t)stream>>>(v, a, g, grad_out, d_vol);
    return (int)cudaGetLastError();
}
