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
#include <cstdlib>

#include "afb_device.cuh"

namespace afb {

constexpr int ETHREADS = 256;
constexpr int ECH = 16;             // channels per thread in the backward

__device__ __forceinline__ AxisConst make_axis_dev(int K) {   // device twin of make_axis (same IEEE fp32 results)
    AxisConst a;
    a.K = K;
    a.km1 = (float)(K - 1);
    a.kf = (float)K;
    a.step = K > 1 ? __fdiv_rn(2.0f, (float)(K - 1)) : 0.0f;
    a.inv_k = (K & (K - 1)) == 0 ? __fdiv_rn(1.0f, (float)K) : 0.0f;
    return a;
}

struct EmbedView {                // per (b, v), shared memory
    float t[12];                  // inverse(normalised affine)[:3,:] fp32 = affine_grid theta
    float fwd[12];                // normalised affine A[:3,:] (fp32 copy) for the backward's candidate boxes
    int axis;                     // forward slab scan: volume axis (0=w,1=h,2=d) along which ix changes fastest
    int K;                        // ... and the max number of slab voxels on one line along that axis
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
    // d ix / d (w,h,d) = t[0..2] in index units: the slab |ix - S/2| < 1 is thinnest along the largest component
    const float m0 = fabsf(ev.t[0]), m1 = fabsf(ev.t[1]), m2 = fabsf(ev.t[2]);
    ev.axis = (m0 >= m1 && m0 >= m2) ? 0 : (m1 >= m2 ? 1 : 2);
    const float ma = fmaxf(fmaxf(m0, m1), fmaxf(m2, 1e-6f));
    ev.K = (int)floorf(2.2f / ma) + 2;
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

// phase A of the forward: the output is ~97% zeros, written as one sequential stream (16-byte streaming stores)
__global__ void __launch_bounds__(ETHREADS)
embed_zero_kernel(float4* __restrict__ out4, size_t n4, float* __restrict__ out, size_t n) {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t stride = (size_t)gridDim.x * ETHREADS;
    for (size_t i = (size_t)blockIdx.x * ETHREADS + threadIdx.x; i < n4; i += stride) __stcs(out4 + i, z4);
    if (blockIdx.x == 0)
        for (size_t i = n4 * 4 + threadIdx.x; i < n; i += ETHREADS) out[i] = 0.0f;
}

// phase B: only the slab |ix - S/2| < 1 is non-zero.  It is thinnest along the volume axis with the largest
// |d ix / d axis| (EmbedView::axis); every line along that axis crosses it in <= K voxels, found analytically and
// then evaluated with the exact (bit-level) tap code.  One thread = one candidate voxel, all c channels
// (channel loop unrolled so that the 4-tap gathers of several channels are in flight together).
__global__ void __launch_bounds__(ETHREADS)
embed_slab_kernel(const float* __restrict__ x, const EmbedView* __restrict__ views, int B, int V, int c, int S,
                  int Kgrid, AxisConst ax, float* __restrict__ out, int ch_per_cta) {
    const int bv = blockIdx.y, b = bv / V, v = bv % V;
    // blockIdx.z: channel chunk.  Measured on the B200 (stage 0, profiles/r2_ab_embed_slab.json): spreading the channels of a
    // candidate voxel over CTAs LOSES (1 / 4 / 16 channels per CTA: 0.58 / 0.42 / 0.39 ms for zero + slab) - the exact tap
    // computation per candidate dominates, not the gather chain - so the host passes all channels (one chunk)
    const int ch0 = blockIdx.z * ch_per_cta, ch1 = min(c, ch0 + ch_per_cta);
    float t[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) t[q] = __ldg(views[bv].t + q);
    const int axis = __ldg(&views[bv].axis), K = min(S, __ldg(&views[bv].K));
    const unsigned idx = blockIdx.x * ETHREADS + threadIdx.x;
    const unsigned line = idx / (unsigned)Kgrid;                // Kgrid candidates per line are laid out in the grid;
    const int k0 = (int)(idx - line * (unsigned)Kgrid);         // a view that needs more (K > Kgrid) loops
    if (line >= (unsigned)S * (unsigned)S) return;
    const int u2 = (int)(line / (unsigned)S), u1 = (int)(line - (unsigned)u2 * (unsigned)S);
    const float Sf = (float)S, mid = (float)(S >> 1);
    const float a1 = 2.0f / Sf, a0 = 1.0f / Sf - 1.0f;       // closed-form base coordinate (2k+1)/S-1 for the line solve
    // position p along the scan axis: ix(p) ~= ix0 + ta * p   (ta = t[axis], index units)
    int w = 0, h = 0, d = 0;
    if (axis == 0) { h = u1; d = u2; } else if (axis == 1) { w = u1; d = u2; } else { w = u1; h = u2; }
    const float ta = t[axis];
    const float g0 = t[0] * (a1 * w + a0) + t[1] * (a1 * h + a0) + t[2] * (a1 * d + a0) + t[3];
    const float ix0 = ((g0 + 1.0f) * Sf - 1.0f) * 0.5f;
    // |ta| ~ 0 on the steepest axis means ix is (numerically) constant over the whole volume - a degenerate / strongly
    // anisotropic inverse affine: every voxel of the line is a candidate or none is (no division by ~0)
    const bool flat = fabsf(ta) < 1e-6f;
    if (flat && fabsf(ix0 - mid) >= 1.05f) return;
    const float pc = flat ? 0.0f : (mid - ix0) / ta, half = flat ? 3.0e38f : 1.05f / fabsf(ta);
    const int plo = flat ? 0 : (int)ceilf(pc - half);
    const size_t S2 = (size_t)S * S, S3 = S2 * S;
    const float* __restrict__ xs = x + ((size_t)b * V + v) * c * S2;
    for (int k = k0; k < K; k += Kgrid) {
        const int p = plo + k;
        if (p < 0 || p >= S || (float)p > pc + half) continue;
        if (axis == 0) w = p; else if (axis == 1) h = p; else d = p;
        const Tap tp = taps_of(t, base_coord(w, ax), base_coord(h, ax), base_coord(d, ax), S);
        if (tp.inb == 0u) continue;
        float* __restrict__ o = out + ((size_t)b * V + v) * c * S3 + ((size_t)d * S + h) * S + w;
#pragma unroll 4
        for (int ch = ch0; ch < ch1; ++ch) {
            const float* __restrict__ xc = xs + (size_t)ch * S2;
            float acc = 0.0f;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if ((tp.inb >> q) & 1u) acc = __fadd_rn(acc, __fmul_rn(__ldg(xc + tp.off[q]), tp.w[q]));
            o[(size_t)ch * S3] = acc;
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

// ================================================================================================
// Batched, single-pass forward/backward over ALL stages of one U-Net pass (HybridUnet.forward embeds the six encoder skips
// with the same affines, models/hybrid_unet.py:40-43).
//
// forward: ONE launch.  A CTA owns a chunk of consecutive output rows (d,h) of one (b,v) and ALL its c channels.  Phase 1
// streams zeros over the chunk (16-byte stores, the ~97 % of the output that is zero); after a CTA barrier phase 2 patches
// the slab voxels of the SAME rows (exact bit-level tap code).  The patch stores land microseconds after the zero stores of
// the same sectors, i.e. while those lines are still dirty in L2, so every sector reaches HBM once - the earlier
// zero-kernel-then-slab-kernel pair re-fetched each slab sector from HBM (1.6 GB of zeros had long been evicted).
// The slab is found per row along w (the contiguous axis): ix is affine in w with slope t[0] (index units), so the row
// meets |ix - S/2| < 1 in one interval; |t0| ~ 0 (slab parallel to the rows) degenerates to "whole row or nothing".
// ================================================================================================
constexpr int EMAX_STAGES = 8;

struct EmbedStage {
    const float* x;            // [B, V*c, S, S]
    float* out;                // [B, V*c, S, S, S]              (forward)
    const float* go;           // grad_out, same shape as out    (backward)
    float* dx;                 // [B, V*c, S, S] | NULL          (backward)
    int c, S;
    int rows_per_cta;          // forward: rows (d,h) per CTA
    unsigned chunks;           // forward: row chunks per (b,v);  backward: pixel tiles x channel chunks per (b,v)
    unsigned cta_begin;        // first CTA of this stage in the flattened grid
    int ch_chunks;             // backward: channel chunks
};

struct EmbedBatch {
    int n, B, V;
    EmbedStage st[EMAX_STAGES];
};

__device__ __forceinline__ int stage_of_cta(const EmbedBatch& eb, unsigned cta) {
    int s = 0;
#pragma unroll
    for (int q = 1; q < EMAX_STAGES; ++q)
        if (q < eb.n && cta >= eb.st[q].cta_begin) s = q;
    return s;
}

__global__ void __launch_bounds__(ETHREADS)
embed_fwd_fused_kernel(const __grid_constant__ EmbedBatch eb, const EmbedView* __restrict__ views) {
    const int si = stage_of_cta(eb, blockIdx.x);
    const EmbedStage& sg = eb.st[si];
    const unsigned local = blockIdx.x - sg.cta_begin;
    const int bv = (int)(local / sg.chunks), chunk = (int)(local - (unsigned)bv * sg.chunks);
    const int S = sg.S, c = sg.c;
    const int nrows = S * S;
    const int r0 = chunk * sg.rows_per_cta, r1 = min(nrows, r0 + sg.rows_per_cta);
    const size_t S2 = (size_t)S * S, S3 = S2 * S;
    float* __restrict__ ob = sg.out + (size_t)bv * c * S3;           // [b, v*c .. v*c+c) == (b*V + v) * c channels
    // ---- phase 1: zeros over rows [r0, r1) of every channel (contiguous span per channel) ----
    const size_t span = (size_t)(r1 - r0) * S;                        // floats per channel
    const size_t off0 = (size_t)r0 * S;
    if ((S & 3) == 0 && (((uintptr_t)sg.out) & 15u) == 0) {
        // (channel, 16-byte column) flattened over the threads: at S <= 32 a channel's span is shorter than the CTA
        const int span4 = (int)(span >> 2);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int step_ch = ETHREADS / span4, step_i = ETHREADS - step_ch * span4;
        int ch = (int)threadIdx.x / span4, i = (int)threadIdx.x - ch * span4;
        while (ch < c) {
            __stcs(reinterpret_cast<float4*>(ob + (size_t)ch * S3 + off0) + i, z4);
            ch += step_ch; i += step_i;
            if (i >= span4) { i -= span4; ++ch; }
        }
    } else {
        for (int ch = 0; ch < c; ++ch) {
            float* __restrict__ p = ob + (size_t)ch * S3 + off0;
            for (int i = threadIdx.x; i < (int)span; i += ETHREADS) p[i] = 0.0f;
        }
    }
    __syncthreads();              // orders the zero stores before the patch stores of the same addresses (CTA scope)
    // ---- phase 2: patch the slab voxels of these rows ----
    // work item = (candidate k along w, row, group of ECG channels), k fastest: neighbouring lanes store neighbouring voxels, and
    // the c channels of a candidate are spread over threads (a chunk has only rows x ~3 candidates: one thread per candidate
    // with a serial channel loop left 9 of 10 threads idle behind a 4-round dependent gather chain; the tap arithmetic is
    // recomputed per channel group instead).  Threads without an item leave before the per-thread set-up.
    // (Writing zeros and slab values in ONE full-width store per 16 bytes instead was built and measured: same time -
    // profiles/experiments/embed_fwd_merged_store.cu.txt.)
    constexpr int ECG = 4;
    const float Sf = (float)S, mid = (float)(S >> 1);
    const float t0 = __ldg(views[bv].t + 0);                          // d ix / d w in index units
    const bool flat = fabsf(t0) * Sf < 0.05f;                        // ix changes by < 0.05 voxel along the whole row
    const int Kc = flat ? S : min(S, (int)(2.1f / fabsf(t0)) + 2);
    const int ngroups = (c + ECG - 1) / ECG;
    const int per_group = (r1 - r0) * Kc;
    const int items = per_group * ngroups;
    if ((int)threadIdx.x >= items) return;
    float t[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) t[q] = __ldg(views[bv].t + q);
    const AxisConst ax = make_axis_dev(S);
    const float a1 = 2.0f / Sf, a0 = 1.0f / Sf - 1.0f;              // closed-form base coordinate (2k+1)/S-1 for the row solve
    const float* __restrict__ xs = sg.x + (size_t)bv * c * S2;
    for (int it = threadIdx.x; it < items; it += ETHREADS) {
        const int cg = it / per_group, rem = it - cg * per_group;
        const int rr = rem / Kc, k = rem - rr * Kc;
        const int row = r0 + rr;
        const int d = row / S, h = row - d * S;
        const float g0 = t[0] * a0 + t[1] * (a1 * h + a0) + t[2] * (a1 * d + a0) + t[3];      // grid x at w = 0 (closed form)
        const float ix0 = ((g0 + 1.0f) * Sf - 1.0f) * 0.5f;
        int lo, hi;                                                   // candidate interval [lo, hi] along w
        if (flat) {
            if (fabsf(ix0 + 0.5f * t0 * Sf - mid) < 1.1f) { lo = 0; hi = S - 1; } else { lo = 1; hi = 0; }
        } else {
            const float wc = (mid - ix0) / t0, half = 1.05f / fabsf(t0);
            lo = max(0, (int)ceilf(wc - half));
            hi = min(S - 1, (int)floorf(wc + half));
        }
        const int w = lo + k;
        if (w > hi) continue;
        const Tap tp = taps_of(t, base_coord(w, ax), base_coord(h, ax), base_coord(d, ax), S);
        if (tp.inb == 0u) continue;
        const int ch0 = cg * ECG;
        float* __restrict__ o = ob + (size_t)ch0 * S3 + (size_t)row * S + w;
        const float* __restrict__ xc = xs + (size_t)ch0 * S2;
        float acc[ECG];
#pragma unroll
        for (int j = 0; j < ECG; ++j) {
            acc[j] = 0.0f;
            if (ch0 + j < c) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if ((tp.inb >> q) & 1u) acc[j] = __fadd_rn(acc[j], __fmul_rn(__ldg(xc + (size_t)j * S2 + tp.off[q]), tp.w[q]));
            }
        }
#pragma unroll
        for (int j = 0; j < ECG; ++j)
            if (ch0 + j < c) o[(size_t)j * S3] = acc[j];
    }
}

// backward, all stages in one launch: GATHER over the 2-D feature-map pixels per (stage, pixel tile, channel chunk,
// b*V + v); the 12 sums of d(theta) are accumulated over ALL stages (the chain through inverse() and the column normalisation
// is linear in them) and chained once by the last CTA of each (b,v).
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
embed_bwd_fused_kernel(const __grid_constant__ EmbedBatch eb, const EmbedView* __restrict__ views, float* __restrict__ d_aff,
                       double* __restrict__ ws_acc, unsigned* __restrict__ ws_counter, unsigned ctas_per_bv) {
    __shared__ EmbedView ev;
    __shared__ float base[256];
    __shared__ float red[NT / 32][12];
    __shared__ double dT[12];
    __shared__ bool is_last;
    const int si = stage_of_cta(eb, blockIdx.x);
    const EmbedStage& sg = eb.st[si];
    const unsigned local = blockIdx.x - sg.cta_begin;
    const int bv = (int)(local / sg.chunks);
    const unsigned rest = local - (unsigned)bv * sg.chunks;
    const int tile = (int)(rest / (unsigned)sg.ch_chunks), cchunk = (int)(rest - (unsigned)tile * (unsigned)sg.ch_chunks);
    const int S = sg.S, c = sg.c, B = eb.B, V = eb.V;
    const int b = bv / V, v = bv % V;
    {
        const unsigned* __restrict__ src = reinterpret_cast<const unsigned*>(views + bv);
        unsigned* dst = reinterpret_cast<unsigned*>(&ev);
        for (int i = threadIdx.x; i < (int)(sizeof(EmbedView) / 4); i += NT) dst[i] = __ldg(src + i);
    }
    const AxisConst ax = make_axis_dev(S);
    const bool use_tab = S <= 256;
    if (use_tab) for (int i = threadIdx.x; i < S; i += NT) base[i] = base_coord(i, ax);
    __syncthreads();
    float part[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) part[q] = 0.0f;
    const int pix = tile * NT + threadIdx.x;
    const int c0 = cchunk * ECH;
    const int nch = min(ECH, c - c0);
    const size_t S2 = (size_t)S * S, S3 = S2 * S;
    if (pix < S * S) {
        const int r = pix / S, q = pix % S;
        const int mid = S >> 1;
        const float Sf = (float)S;
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
        const float* __restrict__ gbase = sg.go + ((size_t)bv * c + c0) * S3;
        const float* __restrict__ xp = sg.x + ((size_t)bv * c + c0) * S2 + pix;
        float acc[ECH], xv[ECH];
#pragma unroll
        for (int ch = 0; ch < ECH; ++ch) { acc[ch] = 0.0f; xv[ch] = (d_aff && ch < nch) ? __ldg(xp + (size_t)ch * S2) : 0.0f; }
        // Along a line (d,h) the three source coordinates are affine in w with slopes t[0], t[4], t[8] (index units), so the
        // voxels that can sample this pixel - ix in [mid-1, mid+1), iy in (q-1, q+1), iz in (r-1, r+1) - form ONE interval of w:
        // solved in closed form (0.05-voxel margin), only its few members run the exact tap code.  (The first version tested
        // every voxel of the bounding box.)  Per line: three fma pairs; the reciprocal slopes are per-thread constants.
        const float cen[3] = {(float)mid, (float)q, (float)r};
        float c000[3], inv[3];                 // coordinate k (index units) of voxel (0,0,0), relative to the pixel; 1/slope or 0 (flat)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float* __restrict__ tk = ev.t + 4 * k;
            const float a0 = 1.0f / Sf - 1.0f;
            const float g0 = (tk[0] + tk[1] + tk[2]) * a0 + tk[3];
            c000[k] = ((g0 + 1.0f) * Sf - 1.0f) * 0.5f - cen[k];
            inv[k] = fabsf(tk[0]) * Sf < 0.02f ? 0.0f : 1.0f / tk[0];
        }
        for (int d = lo[2]; d <= hi[2]; ++d) {
            const float bz = use_tab ? base[d] : base_coord(d, ax);
            for (int h = lo[1]; h <= hi[1]; ++h) {
                const float by = use_tab ? base[h] : base_coord(h, ax);
                float wlo = (float)lo[0], whi = (float)hi[0];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float rel = fmaf(ev.t[4 * k + 1], (float)h, fmaf(ev.t[4 * k + 2], (float)d, c000[k]));   // at w = 0
                    if (inv[k] == 0.0f) {
                        if (fabsf(rel) > 1.07f) { wlo = 1.0f; whi = 0.0f; }               // constant along the line and out of reach
                    } else {
                        const float u = (-1.05f - rel) * inv[k], v2 = (1.05f - rel) * inv[k];
                        wlo = fmaxf(wlo, fminf(u, v2));
                        whi = fminf(whi, fmaxf(u, v2));
                    }
                }
                const int w0 = (int)ceilf(wlo), w1 = (int)floorf(whi);
                for (int w = w0; w <= w1; ++w) {
                    const float bx = use_tab ? base[w] : base_coord(w, ax);
                    // does voxel (w,h,d) sample pixel (q,r)?  Same arithmetic as taps_of (the forward's), reduced to the one
                    // tap (dy,dz) = (q - floor(iy), r - floor(iz)) that can hit this pixel.
                    const float ix = unnormalize(grid_coord(ev.t + 0, bx, by, bz), Sf);
                    const float x0f = floorf(ix);
                    const int x0 = (int)x0f;
                    if (x0 != mid && x0 + 1 != mid) continue;
                    const float iy = unnormalize(grid_coord(ev.t + 4, bx, by, bz), Sf);
                    const float y0f = floorf(iy);
                    const int dy = q - (int)y0f;
                    if ((unsigned)dy > 1u) continue;
                    const float iz = unnormalize(grid_coord(ev.t + 8, bx, by, bz), Sf);
                    const float z0f = floorf(iz);
                    const int dz = r - (int)z0f;
                    if ((unsigned)dz > 1u) continue;
                    const float wx = x0 == mid ? __fsub_rn(__fadd_rn(x0f, 1.0f), ix) : __fsub_rn(ix, x0f);
                    const float sx = x0 == mid ? -1.0f : 1.0f;
                    const float wy = dy ? __fsub_rn(iy, y0f) : __fsub_rn(__fadd_rn(y0f, 1.0f), iy);
                    const float wz = dz ? __fsub_rn(iz, z0f) : __fsub_rn(__fadd_rn(z0f, 1.0f), iz);
                    const float wt = __fmul_rn(__fmul_rn(wx, wy), wz);
                    const float* __restrict__ gp = gbase + ((size_t)d * S + h) * S + w;
                    float ssum = 0.0f;
#pragma unroll
                    for (int ch = 0; ch < ECH; ++ch) {
                        if (ch < nch) {
                            const float gv = __ldg(gp + (size_t)ch * S3);
                            acc[ch] = fmaf(wt, gv, acc[ch]);
                            ssum = fmaf(gv, xv[ch], ssum);
                        }
                    }
                    if (d_aff) {
                        const float hs = 0.5f * Sf;
                        const float ggx = sx * ssum * wy * wz * hs;
                        const float ggy = (dy ? ssum : -ssum) * wx * wz * hs;
                        const float ggz = (dz ? ssum : -ssum) * wx * wy * hs;
                        part[0] += ggx * bx; part[1] += ggx * by; part[2] += ggx * bz; part[3] += ggx;
                        part[4] += ggy * bx; part[5] += ggy * by; part[6] += ggy * bz; part[7] += ggy;
                        part[8] += ggz * bx; part[9] += ggz * by; part[10] += ggz * bz; part[11] += ggz;
                    }
                }
            }
        }
        if (sg.dx) {
            float* __restrict__ dxp = sg.dx + ((size_t)bv * c + c0) * S2 + pix;
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
        double tsum = 0.0;
#pragma unroll
        for (int ww = 0; ww < NT / 32; ++ww) tsum += (double)red[ww][threadIdx.x];
        if (tsum != 0.0) atomicAdd(ws_acc + (size_t)bv * 16 + threadIdx.x, tsum);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ws_counter + bv, 1u) == ctas_per_bv - 1;
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
    // phase A: zero fill (one sequential stream); phase B: the slab
    const size_t n = (size_t)B * V * c * S * S * S;
    const bool al = (((uintptr_t)out & 15u) == 0);
    const size_t n4 = al ? n / 4 : 0;
    size_t want = (n4 + ETHREADS - 1) / ETHREADS;
    const unsigned zgrid = (unsigned)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    embed_zero_kernel<<<zgrid, ETHREADS, 0, st>>>((float4*)out, n4, out, n);
    // |t_axis| >= 0.58 for a rotation => K <= 5 slab voxels per line; views that need more loop inside the kernel
    if ((long long)S * S * 8 >= 2147483647ll) return AFB_ESHAPE;
    const int Kgrid = S < 6 ? S : 6;
    const char* es = getenv("AFB_EMBED_SLAB_CH");          // A/B knob (profiles/ab_embed_roles.py): channels per slab CTA
    int chp = es ? atoi(es) : c;
    if (chp < 1) chp = 1;
    if ((c + chp - 1) / chp > 65535) chp = (c + 65534) / 65535;
    dim3 grid((unsigned)(((long long)S * S * Kgrid + ETHREADS - 1) / ETHREADS), B * V, (unsigned)((c + chp - 1) / chp));
    // (a register budget for 3 or 4 resident CTAs per SM instead of 2 changes nothing: 0.412 / 0.421 / 0.404 ms at stage 0)
    embed_slab_kernel<<<grid, ETHREADS, 0, st>>>(x, views, B, V, c, S, Kgrid, ax, out, chp);
    return (int)cudaGetLastError();
}

extern "C" int afb_embed_multi_bwd(int n_stages, const float* const* grad_out, const float* const* x, const int* c, const int* S,
                                   float* const* d_x, const float* affines, int B, int V, float* d_affines, void* workspace,
                                   void* stream);

extern "C" int afb_embed_bwd(const float* grad_out, const float* x, const float* affines, int B, int V, int c, int S,
                             float* d_x, float* d_affines, void* workspace, void* stream) {
    if (!grad_out || !x || !affines || !workspace) return AFB_EINVAL;
    if (!d_x && !d_affines) return AFB_EINVAL;
    const float* go1[1] = {grad_out};
    const float* x1[1] = {x};
    float* dx1[1] = {d_x};
    return afb_embed_multi_bwd(1, go1, x1, &c, &S, d_x ? dx1 : nullptr, affines, B, V, d_affines, workspace, stream);
}


// ------------------------------------------------------------------------------------------------
// batched entry points: all stages of one pass in one launch each
// ------------------------------------------------------------------------------------------------
static int check_batch(int n_stages, const int* c, const int* S, int B, int V) {
    if (n_stages <= 0 || n_stages > EMAX_STAGES) return AFB_ESHAPE;
    if (B <= 0 || V <= 0 || (long long)B * V > 65535) return AFB_ESHAPE;
    for (int i = 0; i < n_stages; ++i)
        if (c[i] <= 0 || S[i] <= 0 || (long long)S[i] * S[i] * 8 >= 2147483647ll) return AFB_ESHAPE;
    return AFB_OK;
}

extern "C" int afb_embed_multi_fwd(int n_stages, const float* const* x, const int* c, const int* S, float* const* out,
                                   const float* affines, int B, int V, void* workspace, void* stream) {
    if (!x || !c || !S || !out || !affines || !workspace) return AFB_EINVAL;
    int rc = check_batch(n_stages, c, S, B, V);
    if (rc != AFB_OK) return rc;
    const char* ck = getenv("AFB_EMBED_FWD_CHUNK_KB");      // A/B knob (profiles/ab_embed_chunk.py)
    const long long chunk_bytes = (ck && atoi(ck) > 0 ? atoi(ck) : 32) * 1024ll;
    EmbedBatch eb;
    eb.n = n_stages; eb.B = B; eb.V = V;
    unsigned long long cta = 0;
    for (int i = 0; i < n_stages; ++i) {
        if (!x[i] || !out[i]) return AFB_EINVAL;
        EmbedStage& sg = eb.st[i];
        sg.x = x[i]; sg.out = out[i]; sg.go = nullptr; sg.dx = nullptr; sg.c = c[i]; sg.S = S[i]; sg.ch_chunks = 1;
        // ~32 KB of output per CTA (all channels of its rows; swept in profiles/r2_ab_embed_chunk.json), at least one row, at most all rows
        const long long row_bytes = (long long)c[i] * S[i] * 4;
        long long rpc = (chunk_bytes + row_bytes - 1) / row_bytes;
        const long long nrows = (long long)S[i] * S[i];
        if (rpc < 1) rpc = 1;
        if (rpc > nrows) rpc = nrows;
        sg.rows_per_cta = (int)rpc;
        sg.chunks = (unsigned)((nrows + rpc - 1) / rpc);
        sg.cta_begin = (unsigned)cta;
        cta += (unsigned long long)sg.chunks * B * V;
        if (cta >= 2147483647ull) return AFB_ESHAPE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    EmbedView* views = views_of(workspace, B * V);
    embed_prologue_kernel<<<(B * V + 31) / 32, 32, 0, st>>>(affines, B, V, views);
    embed_fwd_fused_kernel<<<(unsigned)cta, ETHREADS, 0, st>>>(eb, views);
    return (int)cudaGetLastError();
}

extern "C" int afb_embed_multi_bwd(int n_stages, const float* const* grad_out, const float* const* x, const int* c, const int* S,
                                   float* const* d_x, const float* affines, int B, int V, float* d_affines, void* workspace,
                                   void* stream) {
    if (!grad_out || !x || !c || !S || !affines || !workspace) return AFB_EINVAL;
    if (!d_x && !d_affines) return AFB_EINVAL;
    int rc = check_batch(n_stages, c, S, B, V);
    if (rc != AFB_OK) return rc;
    static const int variant = [] { const char* e = getenv("AFB_EMBED_BWD_VARIANT"); return e ? atoi(e) : 0; }();
    // measured (profiles/r2_ab_embed_bwd.jsonl): 256 x 2 and 128 x 4 CTAs/SM tie (0.388 ms, 128 registers); forcing 3 x 256 or
    // 5 x 128 CTAs/SM (80 / 96 registers, spills inside the match loop) is 1.8x slower
    const int nt = variant == 1 ? 128 : 256;
    EmbedBatch eb;
    eb.n = 0; eb.B = B; eb.V = V;
    unsigned long long cta = 0, per_bv = 0;
    for (int i = 0; i < n_stages; ++i) {
        if (!grad_out[i]) continue;               // this stage's output took no part in the loss
        if (!x[i]) return AFB_EINVAL;
        EmbedStage& sg = eb.st[eb.n++];
        sg.x = x[i]; sg.out = nullptr; sg.go = grad_out[i]; sg.dx = d_x ? d_x[i] : nullptr; sg.c = c[i]; sg.S = S[i];
        sg.rows_per_cta = 0;
        sg.ch_chunks = (c[i] + ECH - 1) / ECH;
        const unsigned tiles = (unsigned)((S[i] * S[i] + nt - 1) / nt);
        sg.chunks = tiles * (unsigned)sg.ch_chunks;
        sg.cta_begin = (unsigned)cta;
        cta += (unsigned long long)sg.chunks * B * V;
        per_bv += sg.chunks;
        if (cta >= 2147483647ull) return AFB_ESHAPE;
    }
    if (eb.n == 0) return AFB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    double* acc = (double*)workspace;
    unsigned* counter = (unsigned*)(acc + (size_t)B * V * 16);
    EmbedView* views = views_of(workspace, B * V);
    embed_prologue_kernel<<<(B * V + 31) / 32, 32, 0, st>>>(affines, B, V, views);
    if (variant == 1) embed_bwd_fused_kernel<128, 4><<<(unsigned)cta, 128, 0, st>>>(eb, views, d_affines, acc, counter, (unsigned)per_bv);
    else embed_bwd_fused_kernel<256, 2><<<(unsigned)cta, 256, 0, st>>>(eb, views, d_affines, acc, counter, (unsigned)per_bv);
    return (int)cudaGetLastError();
}
