// afb_slice.cu - slice / volume extraction (F.affine_grid + F.grid_sample of the reference,
// utils/nifti_utils.py:182-203) with the view-affine prologue fused in, forward and backward.
//
// One launch covers all S = B*V slices: grid = (tiles per slice, S), 256 threads per CTA, one
// 16x16 tile of output locations per CTA.  Lanes of a warp cover an 8x4 patch (not 32x1) so that an
// oblique plane touches few distinct 128-byte lines per load instruction.  The grid is never
// materialised; the 8 MiB (128^3 fp32) channel volume stays L2 resident across the views of a volume.
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

struct Corners {
    float w[8];                 // ATen order tnw,tne,tsw,tse,bnw,bne,bsw,bse
    long long off[8];           // element offsets (without batch/channel)
    unsigned inb;               // bit k set <=> corner k inside the volume
    float wx[2], wy[2], wz[2];
};

__device__ __forceinline__ Corners corners_of(const Sample& s, const VolArgs& vol) {
    Corners c;
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
        c.off[k] = in ? ((long long)(z0 + dz) * vol.sD + (long long)(y0 + dy) * vol.sH + (long long)(x0 + dx) * vol.sW) : 0ll;
    }
    return c;
}

__device__ __forceinline__ float pad_of(int pad_mode, float pad_value, const float* pad_device) {
    if (pad_mode == AFB_PAD_DEVICE) return __ldg(pad_device);
    if (pad_mode == AFB_PAD_VALUE) return pad_value;
    return 0.0f;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(NTHREADS)
slice_fwd_kernel(VolArgs vol, ViewArgs va, OutGeom g, int pad_mode, float pad_value, const float* pad_device,
                 T* __restrict__ out, float* __restrict__ grid_affine_out, double* __restrict__ nii_affine_out,
                 float* __restrict__ theta_out) {
    __shared__ ViewState st;
    const int s = blockIdx.y;
    if (threadIdx.x < 32) view_prologue_warp0(va, s, st);
    __syncthreads();
    if (blockIdx.x == 0) {
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

    const Corners cn = corners_of(sm, vol);
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
// backward: re-gather, dVolume scatter (RED), dGrid -> dG' block reduction, last-CTA chain
// ------------------------------------------------------------------------------------------------
struct BwdWorkspace {           // layout inside the caller's zeroed workspace
    double* acc;                // [S][16]  (12 sums of dgrid (x) base, 1 spare)
    unsigned* counter;          // [S]
};

template <typename T>
__global__ void __launch_bounds__(NTHREADS)
slice_bwd_kernel(VolArgs vol, ViewArgs va, OutGeom g, int pad_mode, float pad_value, const float* pad_device,
                 const float* __restrict__ grad_out, const float* __restrict__ grad_grid_affine,
                 float* __restrict__ d_vol, float* __restrict__ d_affine, float* __restrict__ d_gpre,
                 float* __restrict__ d_pad, double* __restrict__ ws_acc, unsigned* __restrict__ ws_counter) {
    __shared__ ViewState st;
    __shared__ float red[NTHREADS / 32][13];
    __shared__ ChainScratch cs;
    __shared__ bool is_last;
    const int s = blockIdx.y;
    if (threadIdx.x < 32) view_prologue_warp0(va, s, st);
    __syncthreads();

    float part[13];
#pragma unroll
    for (int q = 0; q < 13; ++q) part[q] = 0.0f;

    const Pix p = pixel_of_thread(g);
    if (p.valid && grad_out != nullptr) {     // grad_out == NULL: chain-only launch (grid.x == 1)
        const Sample sm = sample_coords(g, p, st.g, vol);
        const Corners cn = corners_of(sm, vol);
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
        // d out / d (ix,iy,iz): sign pattern of ATen grid_sampler_3d_backward
        float gix = 0.0f, giy = 0.0f, giz = 0.0f, wsum = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
            const float d = ((cn.inb >> k) & 1u) ? dot[k] : 0.0f;
            gix += (dx ? d : -d) * cn.wy[dy] * cn.wz[dz];
            giy += (dy ? d : -d) * cn.wx[dx] * cn.wz[dz];
            giz += (dz ? d : -d) * cn.wx[dx] * cn.wy[dy];
            wsum += ((cn.inb >> k) & 1u) ? cn.w[k] : 0.0f;
        }
        const float ggx = gix * (0.5f * (float)vol.W), ggy = giy * (0.5f * (float)vol.H), ggz = giz * (0.5f * (float)vol.D);
        part[0] = ggx * sm.bx; part[1] = ggx * sm.by; part[2] = ggx * sm.bz; part[3] = ggx;
        part[4] = ggy * sm.bx; part[5] = ggy * sm.by; part[6] = ggy * sm.bz; part[7] = ggy;
        part[8] = ggz * sm.bx; part[9] = ggz * sm.by; part[10] = ggz * sm.bz; part[11] = ggz;
        part[12] = gsum * (1.0f - wsum);
    }
    // CTA reduction: shuffle, then one fp64 atomic per sum per CTA
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 13; ++q) {
        const float r = warp_sum(part[q]);
        if (lane == 0) red[w][q] = r;
    }
    __syncthreads();
    if (threadIdx.x < 13) {
        double t = 0.0;
#pragma unroll
        for (int ww = 0; ww < NTHREADS / 32; ++ww) t += (double)red[ww][threadIdx.x];
        if (threadIdx.x < 12) {
            atomicAdd(ws_acc + (size_t)s * 16 + threadIdx.x, t);
        } else if (d_pad && pad_mode != AFB_PAD_ZERO) {
            atomicAdd(d_pad, (float)t);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(ws_counter + s, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last CTA of this slice: total dG' (rows 0-2 from the sampler, all rows from upstream)
    if (threadIdx.x < 16) {
        double t = threadIdx.x < 12 ? __ldcg(ws_acc + (size_t)s * 16 + threadIdx.x) : 0.0;
        if (grad_grid_affine) t += (double)grad_grid_affine[(size_t)s * 16 + threadIdx.x];
        cs.dG[threadIdx.x] = t;
        ws_acc[(size_t)s * 16 + threadIdx.x] = 0.0;      // leave the workspace zeroed
    }
    if (threadIdx.x == 0) ws_counter[s] = 0u;
    __syncthreads();
    if (threadIdx.x < 32 && d_affine) view_backward_warp0(va, s, st, cs, d_affine, d_gpre);
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

static int make_args(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo, VolArgs& v, ViewArgs& a) {
    if (!vol || !views || !vol->data) return AFB_EINVAL;
    if (vol->B <= 0 || vol->C <= 0 || vol->D <= 0 || vol->H <= 0 || vol->W <= 0) return AFB_ESHAPE;
    if (Do <= 0 || Ho <= 0 || Wo <= 0 || views->V <= 0) return AFB_ESHAPE;
    v.data = vol->data; v.B = vol->B; v.C = vol->C; v.D = vol->D; v.H = vol->H; v.W = vol->W;
    v.sB = vol->sB; v.sC = vol->sC; v.sD = vol->sD; v.sH = vol->sH; v.sW = vol->sW;
    a.kind = views->kind; a.V = views->V; a.theta = views->theta; a.pre = views->pre; a.pre_is_f64 = views->pre_is_f64;
    a.params = views->params; a.gpre = views->gpre; a.init = views->init; a.R = views->R; a.spat = views->spat;
    a.offset_clip = views->offset_clip; a.zoom_clip = views->zoom_clip; a.nii_affine = views->nii_affine;
    for (int k = 0; k < 3; ++k) a.fov_mm[k] = views->fov_mm[k];
    a.D = vol->D; a.H = vol->H; a.W = vol->W; a.Do = Do; a.Ho = Ho; a.Wo = Wo;
    switch (views->kind) {
        case AFB_AFFINE_GRID: if (!views->theta) return AFB_EINVAL; break;
        case AFB_AFFINE_PRE: if (!views->pre) return AFB_EINVAL; break;
        case AFB_AFFINE_PARAMS:
            if (!views->params || !views->gpre || !views->init) return AFB_EINVAL;
            if (views->R < 0 || views->spat <= 0) return AFB_ESHAPE;
            break;
        default: return AFB_EINVAL;
    }
    return AFB_OK;
}

template <typename T>
static int launch_fwd(const VolArgs& v, const ViewArgs& a, const OutGeom& g, int mode, int pad_mode, float pad_value,
                      const float* pad_device, void* out, float* ga, double* na, float* th, cudaStream_t st) {
    const int S = v.B * a.V;
    dim3 grid(((g.rows + TILE - 1) / TILE) * g.tiles_c, S);
    if (mode == AFB_NEAREST)
        slice_fwd_kernel<T, AFB_NEAREST><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out, ga, na, th);
    else
        slice_fwd_kernel<T, AFB_BILINEAR><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, (T*)out, ga, na, th);
    return (int)cudaGetLastError();
}

}  // namespace afb

using namespace afb;

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
    if ((long long)v.B * a.V > 65535) return AFB_ESHAPE;
    const OutGeom g = make_geom(Do, Ho, Wo);
    cudaStream_t st = (cudaStream_t)stream;
    switch (vol->dtype) {
        case AFB_F32: return launch_fwd<float>(v, a, g, mode, pad_mode, pad_value, pad_device, out, grid_affine_out, nii_affine_out, theta_out, st);
        case AFB_BF16: return launch_fwd<__nv_bfloat16>(v, a, g, mode, pad_mode, pad_value, pad_device, out, grid_affine_out, nii_affine_out, theta_out, st);
        case AFB_F16: return launch_fwd<__half>(v, a, g, mode, pad_mode, pad_value, pad_device, out, grid_affine_out, nii_affine_out, theta_out, st);
        case AFB_I64: return launch_fwd<int64_t>(v, a, g, mode, pad_mode, pad_value, pad_device, out, grid_affine_out, nii_affine_out, theta_out, st);
        case AFB_I32: return launch_fwd<int32_t>(v, a, g, mode, pad_mode, pad_value, pad_device, out, grid_affine_out, nii_affine_out, theta_out, st);
        case AFB_I16: return launch_fwd<int16_t>(v, a, g, mode, pad_mode, pad_value, pad_device, out, grid_affine_out, nii_affine_out, theta_out, st);
        case AFB_U8: return launch_fwd<uint8_t>(v, a, g, mode, pad_mode, pad_value, pad_device, out, grid_affine_out, nii_affine_out, theta_out, st);
        default: return AFB_EDTYPE;
    }
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
    if (S > 65535) return AFB_ESHAPE;
    const OutGeom g = make_geom(Do, Ho, Wo);
    dim3 grid(grad_out ? ((g.rows + TILE - 1) / TILE) * g.tiles_c : 1, S);
    double* acc = (double*)workspace;
    unsigned* counter = (unsigned*)(acc + (size_t)S * 16);
    cudaStream_t st = (cudaStream_t)stream;
    switch (vol->dtype) {
        case AFB_F32:
            slice_bwd_kernel<float><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, grad_out, grad_grid_affine, d_vol, d_affine, d_gpre, d_pad, acc, counter);
            break;
        case AFB_BF16:
            slice_bwd_kernel<__nv_bfloat16><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, grad_out, grad_grid_affine, d_vol, d_affine, d_gpre, d_pad, acc, counter);
            break;
        case AFB_F16:
            slice_bwd_kernel<__half><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, grad_out, grad_grid_affine, d_vol, d_affine, d_gpre, d_pad, acc, counter);
            break;
        default:
            if (grad_out) return AFB_EDTYPE;
            // chain-only (integer / nearest volumes): the volume is never read
            slice_bwd_kernel<float><<<grid, NTHREADS, 0, st>>>(v, a, g, pad_mode, pad_value, pad_device, nullptr, grad_grid_affine, nullptr, d_affine, d_gpre, nullptr, acc, counter);
            break;
    }
    return (int)cudaGetLastError();
}
