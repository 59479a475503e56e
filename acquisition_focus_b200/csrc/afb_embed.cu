// afb_embed.cu - slice -> 3-D embedding of the hybrid U-Net skip connections
// (models/hybrid_unet.py:71-94, SkipConnector.forward) forward and backward.
//
// The reference builds a zero S^3 volume per channel, writes the 2-D feature map onto the plane
// W = S/2 (x_mid, :75-76), materialises a [B,S,S,S,3] grid from inverse(normalised slicing affine)
// (:83-87) and calls grid_sample (:88-90).  Here neither x_mid nor the grid exists: for an output
// voxel only the corners with x index == S/2 can be non-zero, so
//     out = wx(ix) * bilinear2D_zeros(x[b,ch], row = iz, col = iy)
// with the identical coordinate / weight arithmetic (afb_device.cuh), i.e. bitwise the same value
// for the same inverse affine.
//
// forward  : pure HBM write stream (B*V*c*S^3*4 bytes).  One thread = 4 consecutive w (one 16-byte
//            streaming store per channel).  A cheap conservative slab test (FMA, base coordinates from a
//            shared-memory table) rejects ~97% of the voxels; only near-slab voxels run the exact
//            bit-level coordinate code.  Coordinates are computed once and reused across the c channels.
// backward : GATHER over the 2-D feature-map pixels, no atomics on dX (deterministic): the voxels that
//            sample pixel (r,q) are the lattice points of A*([mid-1,mid+1) x [q-1,q+1) x [r-1,r+1)), a
//            parallelepiped whose bounding box (<= ~5^3 candidates) is enumerated and tested with the
//            exact forward tap code.  One thread = one pixel x one chunk of <=16 channels.  The 12 sums
//            of d(theta) are CTA-reduced and chained through inverse() and the column normalisation by the
//            last CTA of each (b,v).
#include "afb_device.cuh"

namespace afb {

constexpr int ETHREADS = 256;
constexpr int ECH = 16;             // channels per thread in the backward

struct EmbedView {                // per (b, v), shared memory
    float t[12];                  // inverse(normalised affine)[:3,:] fp32 = affine_grid theta
    float fwd[12];                // normalised affine A[:3,:] (fp32 copy) for the backward's candidate boxes
    double ga[16], n[3], Ainv[16];
};

// ga -> A = ga diag(1/|col|) -> A^-1 (affine: last row 0 0 0 1), hybrid_unet.py:83-86. One thread.
__device__ inline void embed_prologue(const float* __restrict__ ga_in, EmbedView& ev) {
    for (int i = 0; i < 16; ++i) ev.ga[i] = (double)ga_in[i];
    float nf[3];
    for (int j = 0; j < 3; ++j) {
        // fp32 column norm and reciprocal like get_zooms / 1/get_zooms on the fp32 affine
        float a0 = ga_in[0 * 4 + j], a1 = ga_in[1 * 4 + j], a2 = ga_in[2 * 4 + j];
        nf[j] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2)));
        ev.n[j] = (double)nf[j];
    }
    double A[9], t[3];
    for (int r = 0; r < 3; ++r) {
        for (int j = 0; j < 3; ++j) {
            const float a = __fmul_rn(ga_in[r * 4 + j], __fdiv_rn(1.0f, nf[j]));
            A[r * 3 + j] = (double)a;
            ev.fwd[r * 4 + j] = a;
        }
        t[r] = (double)ga_in[r * 4 + 3];
        ev.fwd[r * 4 + 3] = ga_in[r * 4 + 3];
    }
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double id = 1.0 / det;
    double inv[9];
    inv[0] = c00 * id; inv[1] = (A[2] * A[7] - A[1] * A[8]) * id; inv[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    inv[3] = c01 * id; inv[4] = (A[0] * A[8] - A[2] * A[6]) * id; inv[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    inv[6] = c02 * id; inv[7] = (A[1] * A[6] - A[0] * A[7]) * id; inv[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    for (int r = 0; r < 3; ++r) {
        for (int j = 0; j < 3; ++j) ev.Ainv[r * 4 + j] = inv[r * 3 + j];
        ev.Ainv[r * 4 + 3] = -(inv[r * 3 + 0] * t[0] + inv[r * 3 + 1] * t[1] + inv[r * 3 + 2] * t[2]);
    }
    ev.Ainv[12] = 0.0; ev.Ainv[13] = 0.0; ev.Ainv[14] = 0.0; ev.Ainv[15] = 1.0;
    for (int i = 0; i < 12; ++i) ev.t[i] = (float)ev.Ainv[i];
}

struct Tap {                  // the (up to) 4 contributing samples of one output voxel
    float w[4];               // (wx*wy)*wz in ATen order (dy,dz) = (0,0),(1,0),(0,1),(1,1)
    int off[4];               // row*S + col into x[b,ch]
    unsigned inb;             // bit k: tap k inside the slice;  0 => voxel is zero
    float sx;                 // +-1: d wx / d ix
    float wx, wy[2], wz[2];
};

__device__ __forceinline__ Tap taps_of(const float* __restrict__ t, float bx, float by, float bz, int S) {
    Tap tp;
    tp.inb = 0u;
    const float Sf = (float)S;
    const float ix = unnormalize(grid_coord(t + 0, bx, by, bz), Sf);
    const int x0 = __float2int_rd(ix);
    const int mid = S >> 1;
    if (x0 != mid && x0 + 1 != mid) return tp;
    const float x0f = floorf(ix);
    if (x0 == mid) { tp.wx = __fsub_rn(__fadd_rn(x0f, 1.0f), ix); tp.sx = -1.0f; }
    else { tp.wx = __fsub_rn(ix, x0f); tp.sx = 1.0f; }
    const float iy = unnormalize(grid_coord(t + 4, bx, by, bz), Sf);
    const float iz = unnormalize(grid_coord(t + 8, bx, by, bz), Sf);
    const float y0f = floorf(iy), z0f = floorf(iz);
    const int y0 = __float2int_rd(iy), z0 = __float2int_rd(iz);
    tp.wy[0] = __fsub_rn(__fadd_rn(y0f, 1.0f), iy); tp.wy[1] = __fsub_rn(iy, y0f);
    tp.wz[0] = __fsub_rn(__fadd_rn(z0f, 1.0f), iz); tp.wz[1] = __fsub_rn(iz, z0f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int dy = k & 1, dz = k >> 1;
        const int q = y0 + dy, r = z0 + dz;
        const bool in = q >= 0 && q < S && r >= 0 && r < S;
        tp.w[k] = __fmul_rn(__fmul_rn(tp.wx, tp.wy[dy]), tp.wz[dz]);
        tp.off[k] = in ? r * S + q : 0;
        tp.inb |= in ? (1u << k) : 0u;
    }
    return tp;
}

// ------------------------------------------------------------------------------------------------
// forward: grid = (ceil(S^3/VEC / 256), B*V)
// ------------------------------------------------------------------------------------------------
// one thread per (b,v): the 4x4 algebra runs ONCE per call, not once per CTA (a per-CTA fp64 inverse in front of
// the stores cost 2x in write bandwidth: 2.7 TB/s instead of >5, see profiles/embed_fwd_chunk_sweep.py)
__global__ void embed_prologue_kernel(const float* __restrict__ affines, int B, int V, EmbedView* __restrict__ views) {
    const int bv = blockIdx.x * blockDim.x + threadIdx.x;
    if (bv >= B * V) return;
    const int b = bv / V, v = bv % V;
    EmbedView ev;
    embed_prologue(affines + ((size_t)v * B + b) * 16, ev);
    views[bv] = ev;
}

template <int VEC>
__global__ void __launch_bounds__(ETHREADS)
embed_fwd_kernel(const float* __restrict__ x, const EmbedView* __restrict__ views, int B, int V, int c, int S,
                 AxisConst ax, float* __restrict__ out) {
    const int bv = blockIdx.y, b = bv / V, v = bv % V;
    float t[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) t[q] = __ldg(views[bv].t + q);
    const int wv = S / VEC;                                   // vectors per row
    const long long nvec = (long long)S * S * wv;
    const size_t S2 = (size_t)S * S, S3 = S2 * S;
    const float Sf = (float)S, mid = (float)(S >> 1);
    const float a1 = 2.0f / Sf, a0 = 1.0f / Sf - 1.0f;       // closed-form base coordinate (2k+1)/S-1, for the reject test only
    const float* __restrict__ xs = x + ((size_t)b * V + v) * c * S2;
    for (long long e = (long long)blockIdx.x * ETHREADS + threadIdx.x; e < nvec; e += (long long)gridDim.x * ETHREADS) {
        const int w0 = (int)(e % wv) * VEC;
        const int h = (int)((e / wv) % S), d = (int)(e / ((long long)wv * S));
        float* __restrict__ o = out + ((size_t)b * V + v) * c * S3 + ((size_t)d * S + h) * S + w0;
        // conservative slab test: approximate ix of the first and last voxel of this vector; ix is affine in w, so if
        // both ends are farther than 1.5 voxels on the same side of the plane, all VEC voxels are exactly zero
        const float rest = t[1] * (a1 * h + a0) + t[2] * (a1 * d + a0) + t[3];
        const float ia = ((t[0] * (a1 * w0 + a0) + rest + 1.0f) * Sf - 1.0f) * 0.5f - mid;
        const float ib = ((t[0] * (a1 * (w0 + VEC - 1) + a0) + rest + 1.0f) * Sf - 1.0f) * 0.5f - mid;
        const bool far = (ia > 1.5f && ib > 1.5f) || (ia < -1.5f && ib < -1.5f);
        unsigned any = 0u;
        Tap tp[VEC];
        if (!far) {
            const float by = base_coord(h, ax), bz = base_coord(d, ax);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                tp[k] = taps_of(t, base_coord(w0 + k, ax), by, bz, S);
                any |= tp[k].inb;
            }
        }
        if (any == 0u) {
            if (VEC == 4) {
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
                for (int ch = 0; ch < c; ++ch) __stcs(reinterpret_cast<float4*>(o + (size_t)ch * S3), z4);
            } else {
                for (int ch = 0; ch < c; ++ch) __stcs(o + (size_t)ch * S3, 0.0f);
            }
            continue;
        }
        for (int ch = 0; ch < c; ++ch) {
            const float* __restrict__ xc = xs + (size_t)ch * S2;
            float r[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float acc = 0.0f;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if ((tp[k].inb >> q) & 1u) acc = __fadd_rn(acc, __fmul_rn(__ldg(xc + tp[k].off[q]), tp[k].w[q]));
                r[k] = acc;
            }
            if (VEC == 4) __stcs(reinterpret_cast<float4*>(o + (size_t)ch * S3), make_float4(r[0], r[1], r[2], r[3]));
            else __stcs(o + (size_t)ch * S3, r[0]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__device__ inline void embed_chain(const EmbedView& ev, const double* __restrict__ dT /*12*/, float* __restrict__ d_ga /*16*/) {
    // d(Ainv) = [dT; 0] ; dA = -Ainv^T d(Ainv) Ainv^T
    double G[16], M[16], dA[16];
    for (int i = 0; i < 12; ++i) G[i] = dT[i];
    for (int i = 12; i < 16; ++i) G[i] = 0.0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += ev.Ainv[k * 4 + i] * G[k * 4 + j];
            M[i * 4 + j] = acc;
        }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += M[i * 4 + k] * ev.Ainv[j * 4 + k];
            dA[i * 4 + j] = -acc;
        }
    // A[:, j] = ga[:, j] / n_j  (j < 3), n_j = |ga[:3, j]|
    for (int j = 0; j < 3; ++j) {
        double dot = 0.0;
        for (int r = 0; r < 4; ++r) dot += dA[r * 4 + j] * ev.ga[r * 4 + j];
        const double n = ev.n[j];
        for (int r = 0; r < 4; ++r) {
            double g = dA[r * 4 + j] / n;
            if (r < 3) g -= ev.ga[r * 4 + j] * dot / (n * n * n);
            d_ga[r * 4 + j] = (float)g;
        }
    }
    for (int r = 0; r < 4; ++r) d_ga[r * 4 + 3] = (float)dA[r * 4 + 3];
}

// grid = (pixel tiles, channel chunks, B*V); one thread = one feature-map pixel (r,q) x <=ECH channels
__global__ void __launch_bounds__(ETHREADS)
embed_bwd_kernel(const float* __restrict__ go, const float* __restrict__ x, const EmbedView* __restrict__ views,
                 int B, int V, int c, int S, AxisConst ax, float* __restrict__ d_x, float* __restrict__ d_aff,
                 double* __restrict__ ws_acc, unsigned* __restrict__ ws_counter) {
    __shared__ EmbedView ev;
    __shared__ float base[256];
    __shared__ float red[ETHREADS / 32][12];
    __shared__ double dT[12];
    __shared__ bool is_last;
    const int bv = blockIdx.z, b = bv / V, v = bv % V;
    {
        const unsigned* __restrict__ src = reinterpret_cast<const unsigned*>(views + bv);
        unsigned* dst = reinterpret_cast<unsigned*>(&ev);
        for (int i = threadIdx.x; i < (int)(sizeof(EmbedView) / 4); i += ETHREADS) dst[i] = __ldg(src + i);
    }
    const bool use_tab = S <= 256;
    if (use_tab) for (int i = threadIdx.x; i < S; i += ETHREADS) base[i] = base_coord(i, ax);
    __syncthreads();
    float part[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) part[q] = 0.0f;

    const int pix = blockIdx.x * ETHREADS + threadIdx.x;
    const int c0 = blockIdx.y * ECH;
    const int nch = min(ECH, c - c0);
    const size_t S2 = (size_t)S * S, S3 = S2 * S;
    if (pix < S * S) {
        const int r = pix / S, q = pix % S;                     // r: row (D index), q: column (H index)
        const int mid = S >> 1;
        const float Sf = (float)S;
        // centre of the pixel's influence box in volume index space: A * normalised(mid, q, r)
        const float pn[3] = {(2.0f * mid + 1.0f) / Sf - 1.0f, (2.0f * q + 1.0f) / Sf - 1.0f, (2.0f * r + 1.0f) / Sf - 1.0f};
        int lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float g = ev.fwd[k * 4 + 0] * pn[0] + ev.fwd[k * 4 + 1] * pn[1] + ev.fwd[k * 4 + 2] * pn[2] + ev.fwd[k * 4 + 3];
            const float vc = ((g + 1.0f) * Sf - 1.0f) * 0.5f;
            const float ext = fabsf(ev.fwd[k * 4 + 0]) + fabsf(ev.fwd[k * 4 + 1]) + fabsf(ev.fwd[k * 4 + 2]) + 0.05f;
            lo[k] = max(0, (int)ceilf(vc - ext));
            hi[k] = min(S - 1, (int)floorf(vc + ext));
        }
        const float* __restrict__ gbase = go + (((size_t)b * V + v) * c + c0) * S3;
        const float* __restrict__ xp = x + (((size_t)b * V + v) * c + c0) * S2 + pix;
        float acc[ECH], xv[ECH];
#pragma unroll
        for (int ch = 0; ch < ECH; ++ch) { acc[ch] = 0.0f; xv[ch] = (d_aff && ch < nch) ? __ldg(xp + (size_t)ch * S2) : 0.0f; }
        // (lo/hi index order: k = 0 -> w (x), 1 -> h (y), 2 -> d (z))
        for (int d = lo[2]; d <= hi[2]; ++d) {
            const float bz = use_tab ? base[d] : base_coord(d, ax);
            for (int h = lo[1]; h <= hi[1]; ++h) {
                const float by = use_tab ? base[h] : base_coord(h, ax);
                for (int w = lo[0]; w <= hi[0]; ++w) {
                    const float bx = use_tab ? base[w] : base_coord(w, ax);
                    const Tap tp = taps_of(ev.t, bx, by, bz, S);
                    if (tp.inb == 0u) continue;
                    int t = -1;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (((tp.inb >> k) & 1u) && tp.off[k] == pix) t = k;
                    if (t < 0) continue;
                    const float wt = tp.w[t];
                    const float* __restrict__ gp = gbase + ((size_t)d * S + h) * S + w;
                    float s = 0.0f;
#pragma unroll
                    for (int ch = 0; ch < ECH; ++ch) {
                        if (ch < nch) {
                            const float gv = __ldg(gp + (size_t)ch * S3);
                            acc[ch] = fmaf(wt, gv, acc[ch]);
                            s = fmaf(gv, xv[ch], s);
                        }
                    }
                    if (d_aff) {
                        const int dy = t & 1, dz = t >> 1;
                        const float hs = 0.5f * Sf;
                        const float ggx = tp.sx * s * tp.wy[dy] * tp.wz[dz] * hs;
                        const float ggy = (dy ? s : -s) * tp.wx * tp.wz[dz] * hs;
                        const float ggz = (dz ? s : -s) * tp.wx * tp.wy[dy] * hs;
                        part[0] += ggx * bx; part[1] += ggx * by; part[2] += ggx * bz; part[3] += ggx;
                        part[4] += ggy * bx; part[5] += ggy * by; part[6] += ggy * bz; part[7] += ggy;
                        part[8] += ggz * bx; part[9] += ggz * by; part[10] += ggz * bz; part[11] += ggz;
                    }
                }
            }
        }
        if (d_x) {
            float* __restrict__ dxp = d_x + (((size_t)b * V + v) * c + c0) * S2 + pix;
#pragma unroll
            for (int ch = 0; ch < ECH; ++ch)
                if (ch < nch) dxp[(size_t)ch * S2] = acc[ch];
        }
    }
    if (!d_aff) return;
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        const float rr = warp_sum(part[q]);
        if (lane == 0) red[wi][q] = rr;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        double t = 0.0;
#pragma unroll
        for (int ww = 0; ww < ETHREADS / 32; ++ww) t += (double)red[ww][threadIdx.x];
        if (t != 0.0) atomicAdd(ws_acc + (size_t)bv * 16 + threadIdx.x, t);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ws_counter + bv, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < 12) {
        dT[threadIdx.x] = __ldcg(ws_acc + (size_t)bv * 16 + threadIdx.x);
        ws_acc[(size_t)bv * 16 + threadIdx.x] = 0.0;
    }
    if (threadIdx.x == 0) ws_counter[bv] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) embed_chain(ev, dT, d_aff + ((size_t)v * B + b) * 16);
}

}  // namespace afb

using namespace afb;

extern "C" int64_t afb_embed_workspace_bytes(int n_slices) {
    // [acc: n x 16 double][counter: n x 2 u32][EmbedView x n]
    return (int64_t)n_slices * (16 * sizeof(double) + 2 * sizeof(unsigned) + sizeof(EmbedView));
}

static EmbedView* views_of(void* workspace, int n) {
    return reinterpret_cast<EmbedView*>((char*)workspace + (size_t)n * (16 * sizeof(double) + 2 * sizeof(unsigned)));
}

extern "C" int afb_embed_fwd(const float* x, const float* affines, int B, int V, int c, int S, float* out, void* workspace,
                             void* stream) {
    if (!x || !affines || !out || !workspace) return AFB_EINVAL;
    if (B <= 0 || V <= 0 || c <= 0 || S <= 0 || (long long)B * V > 65535) return AFB_ESHAPE;
    const AxisConst ax = make_axis(S);
    cudaStream_t st = (cudaStream_t)stream;
    EmbedView* views = views_of(workspace, B * V);
    embed_prologue_kernel<<<(B * V + 31) / 32, 32, 0, st>>>(affines, B, V, views);
    const bool vec = (S % 4 == 0) && (((uintptr_t)out & 15u) == 0);
    const long long nvec = (long long)S * S * (vec ? S / 4 : S);
    // grid-stride: ~8 CTAs per SM in total so that each CTA streams many rows
    long long gx = (nvec + ETHREADS - 1) / ETHREADS;
    const long long cap = (148 * 8 + B * V - 1) / (B * V);
    if (gx > cap) gx = cap < 1 ? 1 : cap;
    dim3 grid((unsigned)gx, B * V);
    if (vec) embed_fwd_kernel<4><<<grid, ETHREADS, 0, st>>>(x, views, B, V, c, S, ax, out);
    else embed_fwd_kernel<1><<<grid, ETHREADS, 0, st>>>(x, views, B, V, c, S, ax, out);
    return (int)cudaGetLastError();
}

extern "C" int afb_embed_bwd(const float* grad_out, const float* x, const float* affines, int B, int V, int c, int S,
                             float* d_x, float* d_affines, void* workspace, void* stream) {
    if (!grad_out || !x || !affines || !workspace) return AFB_EINVAL;
    if (!d_x && !d_affines) return AFB_EINVAL;
    if (B <= 0 || V <= 0 || c <= 0 || S <= 0 || (long long)B * V > 65535) return AFB_ESHAPE;
    const int chunks = (c + ECH - 1) / ECH;
    if (chunks > 65535) return AFB_ESHAPE;
    const AxisConst ax = make_axis(S);
    cudaStream_t st = (cudaStream_t)stream;
    double* acc = (double*)workspace;
    unsigned* counter = (unsigned*)(acc + (size_t)B * V * 16);
    EmbedView* views = views_of(workspace, B * V);
    embed_prologue_kernel<<<(B * V + 31) / 32, 32, 0, st>>>(affines, B, V, views);
    dim3 grid((unsigned)((S * S + ETHREADS - 1) / ETHREADS), chunks, B * V);
    embed_bwd_kernel<<<grid, ETHREADS, 0, st>>>(grad_out, x, views, B, V, c, S, ax, d_x, d_affines, acc, counter);
    return (int)cudaGetLastError();
}
