// afb_slice.cu - slice / volume extraction (F.affine_grid + F.grid_sample of the reference,
// utils/nifti_utils.py:182-203) with the view-affine prologue fused in, forward and backward.
//
// One launch covers all S = B*V slices: grid = (tiles per slice, S), 256 threads per CTA, one
// 16x16 tile of output locations per CTA.  Lanes of a warp cover an 8x4 patch (not 32x1) so that an
// oblique plane touches few distinct 128-byte lines per load instruction.  The grid is never
// materialised; the 8 MiB (128^3 fp32) channel volume stays L2 resident across the views of a volume.
#include <type_traits>

#include "afb_device.cuh"

namespace afb {

constexpr int TILE = 16;        // 16 x 16 output locations per CTA
constexpr int NTHREADS = 256;

struct VolArgs {
    const void* data;
    int B, C, D, H, W;
    long long sB, sC, sD, sH, sW;
};

struct OutGeom {
    AxisConst ax, ay, az;       // W(x), H(y), D(z) output axes
    int Do, Ho, Wo;
    int rows, cols;             // 2-D view of the output index space: slices (Wo==1): Do x Ho, else (Do*Ho) x Wo
    int tiles_c;
};

struct Pix {
    int i, j, k;                // output indices (Do, Ho, Wo)
    bool valid;
};

__device__ __forceinline__ Pix pixel_of_thread(const OutGeom& g) {
    const int tid = threadIdx.x;
    const int w = tid >> 5, lane = tid & 31;
    const int lc = ((w & 1) << 3) + (lane & 7);
    const int lr = ((w >> 1) << 2) + (lane >> 3);
    const int tr = blockIdx.x / g.tiles_c, tc = blockIdx.x % g.tiles_c;
    const int row = tr * TILE + lr, col = tc * TILE + lc;
    Pix p;
    p.valid = row < g.rows && col < g.cols;
    if (g.Wo == 1) {
        p.i = row; p.j = col; p.k = 0;
    } else {
        p.i = row / g.Ho; p.j = row % g.Ho; p.k = col;
    }
    return p;
}

struct Sample {                 // un-normalised source coordinates of one output location
    float ix, iy, iz;
    float bx, by, bz;           // normalised base coordinates (x_k, y_j, z_i)
};

__device__ __forceinline__ Sample sample_coords(const OutGeom& g, const Pix& p, const float* G, const VolArgs& vol) {
    Sample s;
    s.bx = base_coord(p.k, g.ax);
    s.by = base_coord(p.j, g.ay);
    s.bz = base_coord(p.i, g.az);
    s.ix = unnormalize(grid_coord(G + 0, s.bx, s.by, s.bz), (float)vol.W);
    s.iy = unnormalize(grid_coord(G + 4, s.bx, s.by, s.bz), (float)vol.H);
    s.iz = unnormalize(grid_coord(G + 8, s.bx, s.by, s.bz), (float)vol.D);
    return s;
}

template <typename OffT>
struct Corners {
    float w[8];                 // ATen order tnw,tne,tsw,tse,bnw,bne,bsw,bse
    OffT off[8];                // element offsets (without batch/channel)
    unsigned inb;               // bit k set <=> corner k inside the volume
    float wx[2], wy[2], wz[2];
};

template <typename OffT>
__device__ __forceinline__ Corners<OffT> corners_of(const Sample& s, const VolArgs& vol) {
    Corners<OffT> c;
    const float x0f = floorf(s.ix), y0f = floorf(s.iy), z0f = floorf(s.iz);
    const int x0 = __float2int_rd(s.ix), y0 = __float2int_rd(s.iy), z0 = __float2int_rd(s.iz);
    c.wx[0] = __fsub_rn(__fadd_rn(x0f, 1.0f), s.ix); c.wx[1] = __fsub_rn(s.ix, x0f);
    c.wy[0] = __fsub_rn(__fadd_rn(y0f, 1.0f), s.iy); c.wy[1] = __fsub_rn(s.iy, y0f);
    c.wz[0] = __fsub_rn(__fadd_rn(z0f, 1.0f), s.iz); c.wz[1] = __fsub_rn(s.iz, z0f);
    const bool xin[2] = {x0 >= 0 && x0 < vol.W, x0 + 1 >= 0 && x0 + 1 < vol.W};
    const bool yin[2] = {y0 >= 0 && y0 < vol.H, y0 + 1 >= 0 && y0 + 1 < vol.H};
    const bool zin[2] = {z0 >= 0 && z0 < vol.D, z0 + 1 >= 0 && z0 + 1 < vol.D};
    c.inb = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
        c.w[k] = __fmul_rn(__fmul_rn(c.wx[dx], c.wy[dy]), c.wz[dz]);
        const bool in = xin[dx] && yin[dy] && zin[dz];
        c.inb |= in ? (1u << k) : 0u;
        c.off[k] = in ? (OffT)((OffT)(z0 + dz) * (OffT)vol.sD + (OffT)(y0 + dy) * (OffT)vol.sH + (OffT)(x0 + dx) * (OffT)vol.sW) : (OffT)0;
    }
    return c;
}

__device__ __forceinline__ float pad_of(int pad_mode, float pad_value, const float* pad_device) {
    if (pad_mode == AFB_PAD_DEVICE) return __ldg(pad_device);
    if (pad_mode == AFB_PAD_VALUE) return pad_value;
    return 0.0f;
}

// ------------------------------------------------------------------------------------------------
// per-CTA view state: either copied from the prologue kernel's output (views.state) or computed
// in place by warp 0 (single-call use of the ABI).  Ends with a __syncthreads().
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_view_state(const ViewArgs& va, int s, ViewState& st, bool full) {
    if (va.state) {
        const unsigned* __restrict__ src = reinterpret_cast<const unsigned*>(
            reinterpret_cast<const char*>(va.state) + (size_t)s * sizeof(ViewState));
        unsigned* dst = reinterpret_cast<unsigned*>(&st);
        const int words = full ? (int)(sizeof(ViewState) / 4) : 16;       // forward only needs g[16]
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = __ldg(src + i);
    } else if (threadIdx.x < 32) {
        view_prologue_warp0(va, s, st);
    }
    __syncthreads();
}

// One warp per slice: raw view input -> ViewState (+ the three small outputs of the forward call).
__global__ void __launch_bounds__(128)
view_prologue_kernel(ViewArgs va, int S, ViewState* __restrict__ states, float* __restrict__ grid_affine_out,
                     double* __restrict__ nii_affine_out, float* __restrict__ theta_out) {
    __shared__ ViewState sh[4];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * 4 + w;
    if (s >= S) return;
    ViewState& st = sh[w];
    view_prologue_warp0(va, s, st);
    __syncwarp();
    if (lane < 16) {
        if (grid_affine_out) grid_affine_out[(size_t)s * 16 + lane] = st.g[lane];
        if (theta_out && va.kind == AFB_AFFINE_PARAMS) theta_out[(size_t)s * 16 + lane] = st.theta[lane];
    }
    if (lane == 16 && nii_affine_out && va.kind != AFB_AFFINE_GRID) {
        double na[16];
        nii_affine_of_result(va, s / va.V, st, na);
        for (int i = 0; i < 16; ++i) nii_affine_out[(size_t)s * 16 + i] = na[i];
    }
    if (states) {
        const unsigned* src = reinterpret_cast<const unsigned*>(&st);
        unsigned* dst = reinterpret_cast<unsigned*>(states + s);
        for (int i = lane; i < (int)(sizeof(ViewState) / 4); i += 32) dst[i] = src[i];
    }
}

// ------------------------------------------------------------------------------------------------
// forward, generic strides (any layout, any dtype): one thread per output location, loop over C
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(NTHREADS)
slice_fwd_kernel(VolArgs vol, ViewArgs va, OutGeom g, int pad_mode, float pad_value, const float* pad_device,
                 T* __restrict__ out, float* __restrict__ grid_affine_out, double* __restrict__ nii_affine_out,
                 float* __restrict__ theta_out) {
    __shared__ ViewState st;
    const int s = blockIdx.y;
    load_view_state(va, s, st, false);
    if (blockIdx.x == 0 && !va.state) {
        if (grid_affine_out && threadIdx.x < 16) grid_affine_out[(size_t)s * 16 + threadIdx.x] = st.g[threadIdx.x];
        if (theta_out && va.kind == AFB_AFFINE_PARAMS && threadIdx.x >= 32 && threadIdx.x < 48)
            theta_out[(size_t)s * 16 + threadIdx.x - 32] = st.theta[threadIdx.x - 32];
        if (nii_affine_out && va.kind != AFB_AFFINE_GRID && threadIdx.x == 64) {
            double na[16];
            nii_affine_of_result(va, s / va.V, st, na);
            for (int i = 0; i < 16; ++i) nii_affine_out[(size_t)s * 16 + i] = na[i];
        }
    }
    const Pix p = pixel_of_thread(g);
    if (!p.valid) return;
    const Sample sm = sample_coords(g, p, st.g, vol);
    const int b = s / va.V;
    const T* __restrict__ src = (const T*)vol.data + (long long)b * vol.sB;
    const size_t plane = (size_t)g.Do * g.Ho * g.Wo;
    T* __restrict__ dst = out + (size_t)s * vol.C * plane + ((size_t)p.i * g.Ho + p.j) * g.Wo + p.k;

    if (MODE == AFB_NEAREST) {
        // nearbyint = round half to even (cvt.rni), ATen grid_sampler_3d nearest
        const int xn = __float2int_rn(sm.ix), yn = __float2int_rn(sm.iy), zn = __float2int_rn(sm.iz);
        const bool in = xn >= 0 && xn < vol.W && yn >= 0 && yn < vol.H && zn >= 0 && zn < vol.D;
        const long long off = in ? ((long long)zn * vol.sD + (long long)yn * vol.sH + (long long)xn * vol.sW) : 0ll;
#pragma unroll 4
        for (int c = 0; c < vol.C; ++c) {
            T v = T(0);
            if (in) v = __ldg(src + (long long)c * vol.sC + off);
            dst[(size_t)c * plane] = v;
        }
        return;
    }

    const Corners<long long> cn = corners_of<long long>(sm, vol);
    const float pad = pad_of(pad_mode, pad_value, pad_device);
#pragma unroll 2
    for (int c = 0; c < vol.C; ++c) {
        const T* __restrict__ sc = src + (long long)c * vol.sC;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (cn.inb >> k) & 1u ? Store<T>::load(sc + cn.off[k]) : pad;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if ((cn.inb >> k) & 1u) acc = __fadd_rn(acc, __fmul_rn(__fsub_rn(v[k], pad), cn.w[k]));
        dst[(size_t)c * plane] = Store<T>::from_float(__fadd_rn(acc, pad));
    }
}

// ------------------------------------------------------------------------------------------------
// forward, channels-last volumes (sC == 1: what one_hot(...).permute(...) of running/run_dl.py:261
// hands the sampler): the C channels of a voxel are contiguous, so every corner is ONE 16-byte
// vector load per 4 fp32 channels (2 loads = one full 32-byte sector for C = 8) instead of C scalar
// loads.  Arithmetic per channel is identical to the generic kernel (bitwise same results).
// ------------------------------------------------------------------------------------------------
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T, int MODE>
__global__ void __launch_bounds__(NTHREADS)
slice_fwd_cl_kernel(VolArgs vol, ViewArgs va, OutGeom g, int pad_mode, float pad_value, const float* pad_device,
                    T* __restrict__ out) {
    __shared__ ViewState st;
    const int s = blockIdx.y;
    load_view_state(va, s, st, false);
    const Pix p = pixel_of_thread(g);
    if (!p.valid) return;
    const Sample sm = sample_coords(g, p, st.g, vol);
    const int b = s / va.V;
    const T* __restrict__ src = (const T*)vol.data + (long long)b * vol.sB;
    const size_t plane = (size_t)g.Do * g.Ho * g.Wo;
    T* __restrict__ dst = out + (size_t)s * vol.C * plane + ((size_t)p.i * g.Ho + p.j) * g.Wo + p.k;
    constexpr int N = Vec16<T>::N;

    if (MODE == AFB_NEAREST) {
        const int xn = __float2int_rn(sm.ix), yn = __float2int_rn(sm.iy), zn = __float2int_rn(sm.iz);
        const bool in = xn >= 0 && xn < vol.W && yn >= 0 && yn < vol.H && zn >= 0 && zn < vol.D;
        const int off = in ? (zn * (int)vol.sD + yn * (int)vol.sH + xn * (int)vol.sW) : 0;
        for (int c0 = 0; c0 < vol.C; c0 += N) {
            uint4 raw = make_uint4(0u, 0u, 0u, 0u);
            if (in) raw = __ldg(reinterpret_cast<const uint4*>(src + off + c0));
            const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int q = 0; q < N; ++q) dst[(size_t)(c0 + q) * plane] = e[q];
        }
        return;
    }
    if constexpr (std::is_same<T, float>::value && MODE == AFB_BILINEAR) {
        const Corners<int> cn = corners_of<int>(sm, vol);
        const float pad = pad_of(pad_mode, pad_value, pad_device);
        for (int c0 = 0; c0 < vol.C; c0 += 4) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                v[k] = (cn.inb >> k) & 1u ? __ldg(reinterpret_cast<const float4*>(src + cn.off[k] + c0)) : make_float4(pad, pad, pad, pad);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((cn.inb >> k) & 1u) {
                    a0 = __fadd_rn(a0, __fmul_rn(__fsub_rn(v[k].x, pad), cn.w[k]));
                    a1 = __fadd_rn(a1, __fmul_rn(__fsub_rn(v[k].y, pad), cn.w[k]));
                    a2 = __fadd_rn(a2, __fmul_rn(__fsub_rn(v[k].z, pad), cn.w[k]));
                    a3 = __fadd_rn(a3, __fmul_rn(__fsub_rn(v[k].w, pad), cn.w[k]));
                }
            dst[(size_t)(c0 + 0) * plane] = Store<T>::from_float(__fadd_rn(a0, pad));
            dst[(size_t)(c0 + 1) * plane] = Store<T>::from_float(__fadd_rn(a1, pad));
            dst[(size_t)(c0 + 2) * plane] = Store<T>::from_float(__fadd_rn(a2, pad));
            dst[(size_t)(c0 + 3) * plane] = Store<T>::from_float(__fadd_rn(a3, pad));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward of the view prologue (analytic chain), run by warp 0 of the last CTA of a slice
// ------------------------------------------------------------------------------------------------
__device__ inline void cross3d(const double* u, const double* v, double* o) {
    o[0] = u[1] * v[2] - u[2] * v[1];
    o[1] = u[2] * v[0] - u[0] * v[2];
    o[2] = u[0] * v[1] - u[1] * v[0];
}

struct ChainScratch {
    double dG[16];
    double dpos[3];
};

__device__ inline void view_backward_warp0(const ViewArgs& va, int s, const ViewState& st, ChainScratch& cs,
                                           float* __restrict__ d_affine, float* __restrict__ d_gpre) {
    const int lane = threadIdx.x & 31;
    if (va.kind == AFB_AFFINE_GRID) {
        if (lane < 12) d_affine[(size_t)s * 12 + lane] = (float)cs.dG[lane];
        return;
    }
    const int NP = 6 + 3 * va.R + 1;
    if (lane == 0) {
        // ---- G' = P diag(s,1), s_j = rho_j / n_{2-j}  (nifti_utils.py:55-58 with the flip) ----
        double dP[16];
        for (int r = 0; r < 4; ++r) {
            for (int j = 0; j < 3; ++j) dP[r * 4 + j] = cs.dG[r * 4 + j] * st.s[j];
            dP[r * 4 + 3] = cs.dG[r * 4 + 3];
        }
        for (int j = 0; j < 3; ++j) {
            double ds = 0.0;
            for (int r = 0; r < 4; ++r) ds += cs.dG[r * 4 + j] * st.P[r * 4 + j];
            const int k = 2 - j;
            const double dn = -st.s[j] / st.n[k] * ds;
            for (int r = 0; r < 3; ++r) dP[r * 4 + k] += dn * st.P[r * 4 + k] / st.n[k];
        }
        if (va.kind == AFB_AFFINE_PRE) {
            for (int i = 0; i < 16; ++i) d_affine[(size_t)s * 16 + i] = (float)dP[i];
        } else {
            // ---- P = Gpre @ theta ----
            double dth[16];
            for (int k = 0; k < 4; ++k)
                for (int j = 0; j < 4; ++j) {
                    double acc = 0.0;
                    for (int i = 0; i < 4; ++i) acc += (double)st.gpre[i * 4 + k] * dP[i * 4 + j];
                    dth[k * 4 + j] = acc;
                }
            if (d_gpre) {
                for (int i = 0; i < 4; ++i)
                    for (int k = 0; k < 4; ++k) {
                        double acc = 0.0;
                        for (int j = 0; j < 4; ++j) acc += dP[i * 4 + j] * (double)st.theta[k * 4 + j];
                        d_gpre[(size_t)s * 16 + i * 4 + k] = (float)acc;
                    }
            }
            // ---- theta = [[zm * Rm, t]] ----
            double dzm = 0.0, dRm[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    dzm += dth[r * 4 + c] * (double)st.Rm[r * 3 + c];
                    dRm[r * 3 + c] = (double)st.zm * dth[r * 4 + c];
                }
            // ---- Rm = R0 @ Rb ----
            double dRb[9];
            for (int k = 0; k < 3; ++k)
                for (int c = 0; c < 3; ++c) {
                    double acc = 0.0;
                    for (int r = 0; r < 3; ++r) acc += (double)st.R0[r * 3 + k] * dRm[r * 3 + c];
                    dRb[k * 3 + c] = acc;
                }
            // ---- Gram-Schmidt backward (transform_utils.py:29-35) ----
            double x[3], y[3], z[3], bb[3], dx[3], dy[3], dz[3], t1[3], t2[3];
            for (int r = 0; r < 3; ++r) {
                x[r] = st.Rb[r * 3 + 0]; y[r] = st.Rb[r * 3 + 1]; z[r] = st.Rb[r * 3 + 2];
                dx[r] = dRb[r * 3 + 0]; dy[r] = dRb[r * 3 + 1]; dz[r] = dRb[r * 3 + 2];
                bb[r] = st.b[r];
            }
            (void)y;
            cross3d(x, dy, t1);                 // y = z cross x : dz += x cross dy
            cross3d(dy, z, t2);                 //                 dx += dy cross z
            for (int r = 0; r < 3; ++r) { dz[r] += t1[r]; dx[r] += t2[r]; }
            double zdz = z[0] * dz[0] + z[1] * dz[1] + z[2] * dz[2];
            double dzr[3];
            for (int r = 0; r < 3; ++r) dzr[r] = (dz[r] - z[r] * zdz) / (double)st.nz;
            cross3d(bb, dzr, t1);               // z' = x cross b : dx += b cross dz'
            double db[3];
            cross3d(dzr, x, db);                //                  db  = dz' cross x
            for (int r = 0; r < 3; ++r) dx[r] += t1[r];
            double xdx = x[0] * dx[0] + x[1] * dx[1] + x[2] * dx[2];
            float* dp = d_affine + (size_t)s * NP;
            for (int r = 0; r < 3; ++r) {
                dp[r] = (float)((dx[r] - x[r] * xdx) / (double)st.na);
                dp[3 + r] = (float)db[r];
            }
            // ---- zoom: zm = init_zp * (1 - clip*tanh(zp)) ----
            const double dzb = (double)st.init_zp * dzm;
            dp[NP - 1] = (float)(-(double)va.zoom_clip * (1.0 - (double)st.tanh_z * (double)st.tanh_z) * dzb);
            // ---- offsets: t = init_t + offs, offs = (2 pos + 1)/spat - 1 ----
            for (int c = 0; c < 3; ++c)
                cs.dpos[c] = (va.offset_clip == 0.0f) ? 0.0 : dth[c * 4 + 3] * 2.0 / (double)va.spat;
        }
    }
    __syncwarp();
    if (va.kind == AFB_AFFINE_PARAMS) {
        // soft-argmax backward: dlogit_i = p_i (arra_i - pos) dpos
        const float* prm = va.params + (size_t)s * NP;
        float* dp = d_affine + (size_t)s * NP;
        const int arra0 = (va.spat - va.R) / 2;
        for (int c = 0; c < 3; ++c) {
            const float* lg = prm + 6 + c * va.R;
            float m = -INFINITY;
            for (int i = lane; i < va.R; i += 32) m = fmaxf(m, lg[i]);
            m = warp_max(m);
            float se = 0.0f;
            for (int i = lane; i < va.R; i += 32) se += expf(lg[i] - m);
            se = warp_sum(se);
            for (int i = lane; i < va.R; i += 32) {
                const double pr = (double)expf(lg[i] - m) / (double)se;
                dp[6 + c * va.R + i] = (float)(pr * ((double)(arra0 + i) - (double)st.pos[c]) * cs.dpos[c]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward: re-gather, dVolume scatter (RED), dGrid -> dG' CTA reduction, last-CTA chain
// ------------------------------------------------------------------------------------------------
struct BwdShared {
    ViewState st;
    float red[NTHREADS / 32][13];
    ChainScratch cs;
    bool is_last;
};

__device__ __forceinline__ void grid_grad_parts(const float* dot, const float (&wx)[2], const float (&wy)[2], const float (&wz)[2],
                                                const float* w, unsigned inb, const Sample& sm, const VolArgs& vol, float gsum,
                                                float* part) {
    // d out / d (ix,iy,iz): sign pattern of ATen grid_sampler_3d_backward
    float gix = 0.0f, giy = 0.0f, giz = 0.0f, wsum = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
        const float d = ((inb >> k) & 1u) ? dot[k] : 0.0f;
        gix += (dx ? d : -d) * wy[dy] * wz[dz];
        giy += (dy ? d : -d) * wx[dx] * wz[dz];
        giz += (dz ? d : -d) * wx[dx] * wy[dy];
        wsum += ((inb >> k) & 1u) ? w[k] : 0.0f;
    }
    const float ggx = gix * (0.5f * (float)vol.W), ggy = giy * (0.5f * (float)vol.H), ggz = giz * (0.5f * (float)vol.D);
    part[0] = ggx * sm.bx; part[1] = ggx * sm.by; part[2] = ggx * sm.bz; part[3] = ggx;
    part[4] = ggy * sm.bx; part[5] = ggy * sm.by; part[6] = ggy * sm.bz; part[7] = ggy;
    part[8] = ggz * sm.bx; part[9] = ggz * sm.by; part[10] = ggz * sm.bz; part[11] = ggz;
    part[12] = gsum * (1.0f - wsum);
}

// CTA reduction of the 13 partial sums (shuffle -> smem -> one fp64 atomic per sum per CTA); the last CTA of the
// slice (atomic ticket) adds the upstream gradient of grid_affine and runs the analytic parameter chain.
__device__ __forceinline__ void bwd_epilogue(const ViewArgs& va, int s, BwdShared& sh, const float* part, int pad_mode,
                                             const float* __restrict__ grad_grid_affine, float* __restrict__ d_affine,
                                             float* __restrict__ d_gpre, float* __restrict__ d_pad,
                                             double* __restrict__ ws_acc, unsigned* __restrict__ ws_counter) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 13; ++q) {
        const float r = warp_sum(part[q]);
        if (lane == 0) sh.red[w][q] = r;
    }
    __syncthreads();
    if (threadIdx.x < 13) {
        double t = 0.0;
#pragma unroll
        for (int ww = 0; ww < NTHREADS / 32; ++ww) t += (double)sh.red[ww][threadIdx.x];
        if (threadIdx.x < 12) {
            atomicAdd(ws_acc + (size_t)s * 16 + threadIdx.x, t);
        } else if (d_pad && pad_mode != AFB_PAD_ZERO) {
            atomicAdd(d_pad, (float)t);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sh.is_last = (atomicAdd(ws_counter + s, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!sh.is_last) return;
    __threadfence();
    if (threadIdx.x < 16) {
        double t = threadIdx.x < 12 ? __ldcg(ws_acc + (size_t)s * 16 + threadIdx.x) : 0.0;
        if (grad_grid_affine) t += (double)grad_grid_affine[(size_t)s * 16 + threadIdx.x];
        sh.cs.dG[threadIdx.x] = t;
        ws_acc[(size_t)s * 16 + threadIdx.x] = 0.0;      // leave the workspace zeroed
    }
    if (threadIdx.x == 0) ws_counter[s] = 0u;
    __syncthreads();
    if (threadIdx.x < 32 && d_affine) view_backward_warp0(va, s, sh.st, sh.cs, d_affine, d_gpre);
}

template <typename T>
__global__ void __launch_bounds__(NTHREADS)
slice_bwd_kernel(VolArgs vol, ViewArgs va, OutGeom g, int pad_mode, float pad_value, const float* pad_device,
                 const float* __restrict__ grad_out, const float* __restrict__ grad_grid_affine,
                 float* __restrict__ d_vol, float* __restrict__ d_affine, float* __restrict__ d_gpre,
                 float* __restrict__ d_pad, double* __restrict__ ws_acc, unsigned* __restrict__ ws_counter) {
    __shared__ BwdShared sh;
    const int s = blockIdx.y;
    load_view_state(va, s, sh.st, true);
    float part[13];
#pragma unroll
    for (int q = 0; q < 13; ++q) part[q] = 0.0f;
    const Pix p = pixel_of_thread(g);
    if (p.valid && grad_out != nullptr) {     // grad_out == NULL: chain-only launch (grid.x == 1)
        const Sample sm = sample_coords(g, p, sh.st.g, vol);
        const Corners<long long> cn = corners_of<long long>(sm, vol);
        const float pad = pad_of(pad_mode, pad_value, pad_device);
        const int b = s / va.V;
        const T* __restrict__ src = (const T*)vol.data + (long long)b * vol.sB;
        float* __restrict__ dv = d_vol ? d_vol + (long long)b * vol.sB : nullptr;
        const size_t plane = (size_t)g.Do * g.Ho * g.Wo;
        const float* __restrict__ go_p = grad_out + (size_t)s * vol.C * plane + ((size_t)p.i * g.Ho + p.j) * g.Wo + p.k;
        float dot[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) dot[k] = 0.0f;
        float gsum = 0.0f;
#pragma unroll 2
        for (int c = 0; c < vol.C; ++c) {
            const float go = __ldg(go_p + (size_t)c * plane);
            const long long coff = (long long)c * vol.sC;
            gsum += go;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if ((cn.inb >> k) & 1u) {
                    const float v = Store<T>::load(src + coff + cn.off[k]) - pad;
                    dot[k] = fmaf(v, go, dot[k]);
                    if (dv) atomicAdd(dv + coff + cn.off[k], cn.w[k] * go);
                }
            }
        }
        grid_grad_parts(dot, cn.wx, cn.wy, cn.wz, cn.w, cn.inb, sm, vol, gsum, part);
    }
    bwd_epilogue(va, s, sh, part, pad_mode, grad_grid_affine, d_affine, d_gpre, d_pad, ws_acc, ws_counter);
}

// channels-last fp32 volumes: 16-byte gathers and 16-byte vector reductions (red.global.add.v4.f32)
__global__ void __launch_bounds__(NTHREADS)
slice_bwd_cl_kernel(VolArgs vol, ViewArgs va, OutGeom g, int pad_mode, float pad_value, const float* pad_device,
                    const float* __restrict__ grad_out, const float* __restrict__ grad_grid_affine,
                    float* __restrict__ d_vol, float* __restrict__ d_affine, float* __restrict__ d_gpre,
                    float* __restrict__ d_pad, double* __restrict__ ws_acc, unsigned* __restrict__ ws_counter) {
    __shared__ BwdShared sh;
    const int s = blockIdx.y;
    load_view_state(va, s, sh.st, true);
    float part[13];
#pragma unroll
    for (int q = 0; q < 13; ++q) part[q] = 0.0f;
    const Pix p = pixel_of_thread(g);
    if (p.valid) {
        const Sample sm = sample_coords(g, p, sh.st.g, vol);
        const Corners<int> cn = corners_of<int>(sm, vol);
        const float pad = pad_of(pad_mode, pad_value, pad_device);
        const int b = s / va.V;
        const float* __restrict__ src = (const float*)vol.data + (long long)b * vol.sB;
        float* __restrict__ dv = d_vol ? d_vol + (long long)b * vol.sB : nullptr;
        const size_t plane = (size_t)g.Do * g.Ho * g.Wo;
        const float* __restrict__ go_p = grad_out + (size_t)s * vol.C * plane + ((size_t)p.i * g.Ho + p.j) * g.Wo + p.k;
        float dot[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) dot[k] = 0.0f;
        float gsum = 0.0f;
        for (int c0 = 0; c0 < vol.C; c0 += 4) {
            const float g0 = __ldg(go_p + (size_t)(c0 + 0) * plane), g1 = __ldg(go_p + (size_t)(c0 + 1) * plane);
            const float g2 = __ldg(go_p + (size_t)(c0 + 2) * plane), g3 = __ldg(go_p + (size_t)(c0 + 3) * plane);
            gsum += (g0 + g1) + (g2 + g3);
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                v[k] = (cn.inb >> k) & 1u ? __ldg(reinterpret_cast<const float4*>(src + cn.off[k] + c0)) : make_float4(pad, pad, pad, pad);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if ((cn.inb >> k) & 1u) {
                    dot[k] = fmaf(v[k].x - pad, g0, fmaf(v[k].y - pad, g1, fmaf(v[k].z - pad, g2, fmaf(v[k].w - pad, g3, dot[k]))));
                    if (dv) {
                        const float wk = cn.w[k];
                        atomicAdd(reinterpret_cast<float4*>(dv + cn.off[k] + c0), make_float4(wk * g0, wk * g1, wk * g2, wk * g3));
                    }
                }
            }
        }
        grid_grad_parts(dot, cn.wx, cn.wy, cn.wz, cn.w, cn.inb, sm, vol, gsum, part);
    }
    bwd_epilogue(va, s, sh, part, pad_mode, grad_grid_affine, d_affine, d_gpre, d_pad, ws_acc, ws_counter);
}

// d(out)/d(pad) = sum go * (1 - sum of in-bounds weights): depends on geometry and grad_out only, so it can run
// BEFORE the dVolume fill, which lets MinBackward be fused with the zero-fill (afb_min_grad_fill).
__global__ void __launch_bounds__(NTHREADS)
slice_pad_grad_kernel(VolArgs vol, ViewArgs va, OutGeom g, const float* __restrict__ grad_out, float* __restrict__ d_pad) {
    __shared__ ViewState st;
    __shared__ float red[NTHREADS / 32];
    const int s = blockIdx.y;
    load_view_state(va, s, st, false);
    const Pix p = pixel_of_thread(g);
    float part = 0.0f;
    if (p.valid) {
        const Sample sm = sample_coords(g, p, st.g, vol);
        const Corners<int> cn = corners_of<int>(sm, vol);
        float wsum = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) wsum += ((cn.inb >> k) & 1u) ? cn.w[k] : 0.0f;
        const size_t plane = (size_t)g.Do * g.Ho * g.Wo;
        const float* __restrict__ go_p = grad_out + (size_t)s * vol.C * plane + ((size_t)p.i * g.Ho + p.j) * g.Wo + p.k;
        float gsum = 0.0f;
        for (int c = 0; c < vol.C; ++c) gsum += __ldg(go_p + (size_t)c * plane);
        part = gsum * (1.0f - wsum);
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
static OutGeom make_geom(int Do, int Ho, int Wo) {
    OutGeom g;
    g.ax = make_axis(Wo); g.ay = make_axis(Ho); g.az = make_axis(Do);
    g.Do = Do; g.Ho = Ho; g.Wo = Wo;
    if (Wo == 1) { g.rows = Do; g.cols = Ho; } else { g.rows = Do * Ho; g.cols = Wo; }
    g.tiles_c = (g.cols + TILE - 1) / TILE;
    return g;
}

static int make_view_args(const afb_views* views, int B, int D, int H, int W, int Do, int Ho, int Wo, ViewArgs& a) {
    if (!views) return AFB_EINVAL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || Do <= 0 || Ho <= 0 || Wo <= 0 || views->V <= 0) return AFB_ESHAPE;
    if ((long long)B * views->V > 65535) return AFB_ESHAPE;
    a.kind = views->kind; a.V = views->V; a.theta = views->theta; a.pre = views->pre; a.pre_is_f64 = views->pre_is_f64;
    a.params = views->params; a.gpre = views->gpre; a.init = views->init; a.R = views->R; a.spat = views->spat;
    a.offset_clip = views->offset_clip; a.zoom_clip = views->zoom_clip; a.nii_affine = views->nii_affine;
    for (int k = 0; k < 3; ++k) a.fov_mm[k] = views->fov_mm[k];
    a.D = D; a.H = H; a.W = W; a.Do = Do; a.Ho = Ho; a.Wo = Wo;
    a.state = views->state;
    switch (views->kind) {
        case AFB_AFFINE_GRID: if (!views->theta && !views->state) return AFB_EINVAL; break;
        case AFB_AFFINE_PRE: if (!views->pre && !views->state) return AFB_EINVAL; break;
        case AFB_AFFINE_PARAMS:
            if (!views->params || ((!views->gpre || !views->init) && !views->state)) return AFB_EINVAL;
            if (views->R < 0 || views->spat <= 0) return AFB_ESHAPE;
            break;
        default: return AFB_EINVAL;
    }
    return AFB_OK;
}

static int make_args(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo, VolArgs& v, ViewArgs& a) {
    if (!vol || !vol->data) return AFB_EINVAL;
    if (vol->C <= 0) return AFB_ESHAPE;
    v.data = vol->data; v.B = vol->B; v.C = vol->C; v.D = vol->D; v.H = vol->H; v.W = vol->W;
    v.sB = vol->sB; v.sC = vol->sC; v.sD = vol->sD; v.sH = vol->sH; v.sW = vol->sW;
    return make_view_args(views, vol->B, vol->D, vol->H, vol->W, Do, Ho, Wo, a);
}

// channels-last fast path: channels contiguous, everything 16-byte aligned, offsets fit in 32 bits
static bool channels_last_ok(const afb_volume* vol, int n, const void* extra_ptr) {
    if (vol->sC != 1 || vol->C % n != 0) return false;
    if (((uintptr_t)vol->data & 15u) || ((uintptr_t)extra_ptr & 15u)) return false;
    if (vol->sW % n || vol->sH % n || vol->sD % n || vol->sB % n) return false;
    if (vol->sW < 0 || vol->sH < 0 || vol->sD < 0) return false;
    const long long extent = (long long)(vol->D - 1) * vol->sD + (long long)(vol->H - 1) * vol->sH + (long long)(vol->W - 1) * vol->sW + vol->C;
    return extent < 2147483647ll;
}

static dim3 slice_grid(const OutGeom& g, int S) { return dim3(((g.rows + TILE - 1) / TILE) * g.tiles_c, S); }

template <typename T>
static int launch_fwd(const afb_volume* vol, const VolArgs& v, const ViewArgs& a, const OutGeom& g, int mode, int pad_mode,
                      float pad_value, const float* pad_device, void* out, float* ga, double* na, float* th, cudaStream_t st) {
    const int S = v.B * a.V;
    const dim3 grid = slice_grid(g, S);
    const bool cl = a.state != nullptr && channels_last_ok(vol, 16 / (int)sizeof(T), nullptr) &&
                    (mode == AFB_NEAREST || std::is_same<T, float>::value);
    if (cl) {
        if (mode == AFB_NEAREST)
            slice_fwd_cl_kernel<T, AFB_NEAREST><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out);
        else
            slice_fwd_cl_kernel<T, AFB_BILINEAR><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out);
    } else if (mode == AFB_NEAREST) {
        slice_fwd_kernel<T, AFB_NEAREST><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out, ga, na, th);
    } else {
        slice_fwd_kernel<T, AFB_BILINEAR><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out, ga, na, th);
    }
    return (int)cudaGetLastError();
}

}  // namespace afb

using namespace afb;

extern "C" int64_t afb_view_state_bytes(void) { return (int64_t)sizeof(ViewState); }

extern "C" int afb_view_prologue(const afb_views* views, int B, int D, int H, int W, int Do, int Ho, int Wo, void* state,
                                 float* grid_affine_out, double* nii_affine_out, float* theta_out, void* stream) {
    ViewArgs a;
    int rc = make_view_args(views, B, D, H, W, Do, Ho, Wo, a);
    if (rc != AFB_OK) return rc;
    a.state = nullptr;
    if (views->kind == AFB_AFFINE_GRID && !views->theta) return AFB_EINVAL;
    if (views->kind == AFB_AFFINE_PRE && !views->pre) return AFB_EINVAL;
    if (views->kind == AFB_AFFINE_PARAMS && (!views->gpre || !views->init)) return AFB_EINVAL;
    const int S = B * views->V;
    view_prologue_kernel<<<(S + 3) / 4, 128, 0, (cudaStream_t)stream>>>(a, S, (ViewState*)state, grid_affine_out, nii_affine_out, theta_out);
    return (int)cudaGetLastError();
}

extern "C" int afb_slice_fwd(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo, int mode,
                             int pad_mode, float pad_value, const float* pad_device, void* out,
                             float* grid_affine_out, double* nii_affine_out, float* theta_out, void* stream) {
    VolArgs v; ViewArgs a;
    int rc = make_args(vol, views, Do, Ho, Wo, v, a);
    if (rc != AFB_OK) return rc;
    if (!out) return AFB_EINVAL;
    if (mode != AFB_BILINEAR && mode != AFB_NEAREST) return AFB_EINVAL;
    if (pad_mode < AFB_PAD_ZERO || pad_mode > AFB_PAD_DEVICE) return AFB_EINVAL;
    if (pad_mode == AFB_PAD_DEVICE && !pad_device) return AFB_EINVAL;
    const OutGeom g = make_geom(Do, Ho, Wo);
    cudaStream_t st = (cudaStream_t)stream;
#define AFB_FWD(T) return launch_fwd<T>(vol, v, a, g, mode, pad_mode, pad_value, pad_device, out, grid_affine_out, nii_affine_out, theta_out, st)
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

extern "C" int64_t afb_slice_bwd_workspace_bytes(int S) {
    return (int64_t)S * (16 * sizeof(double) + sizeof(unsigned) * 2);
}

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
    const int S = v.B * a.V;
    const OutGeom g = make_geom(Do, Ho, Wo);
    dim3 grid = slice_grid(g, S);
    if (!grad_out) grid.x = 1;
    double* acc = (double*)workspace;
    unsigned* counter = (unsigned*)(acc + (size_t)S * 16);
    cudaStream_t st = (cudaStream_t)stream;
#define AFB_BWD(T) slice_bwd_kernel<T><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, grad_out, grad_grid_affine, d_vol, d_affine, d_gpre, d_pad, acc, counter)
    switch (vol->dtype) {
        case AFB_F32:
            if (grad_out && a.state && channels_last_ok(vol, 4, d_vol))
                slice_bwd_cl_kernel<<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, grad_out, grad_grid_affine, d_vol, d_affine, d_gpre, d_pad, acc, counter);
            else
                AFB_BWD(float);
            break;
        case AFB_BF16: AFB_BWD(__nv_bfloat16); break;
        case AFB_F16: AFB_BWD(__half); break;
        default:
            if (grad_out) return AFB_EDTYPE;
            // chain-only (integer / nearest volumes): the volume is never read
            slice_bwd_kernel<float><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, nullptr, grad_grid_affine, nullptr, d_affine, d_gpre, nullptr, acc, counter);
            break;
    }
#undef AFB_BWD
    return (int)cudaGetLastError();
}

extern "C" int afb_slice_pad_grad(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo,
                                  const float* grad_out, float* d_pad, void* stream) {
    VolArgs v; ViewArgs a;
    int rc = make_args(vol, views, Do, Ho, Wo, v, a);
    if (rc != AFB_OK) return rc;
    if (!grad_out || !d_pad) return AFB_EINVAL;
    const OutGeom g = make_geom(Do, Ho, Wo);
    slice_pad_grad_kernel<<<slice_grid(g, v.B * a.V), NTHREADS, 0, (cudaStream_t)stream>>>(v, a, g, grad_out, d_pad);
    return (int)cudaGetLastError();
}
