// afb_sampler.cuh - building blocks shared by the samplers (afb_slice.cu, afb_onehot.cu): thread -> output location
// mapping, bit-exact coordinates, the 8 trilinear corners, the dgrid (x) base reduction of the backward.
#pragma once
#include <cstdlib>

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
    int tiles_c_shift;          // log2(tiles_c) when it is a power of two, else -1
    int wide;                   // 1: 3-D outputs with Wo >= 32: a warp covers 32 consecutive w of one row (8 x 32 tile per CTA) so
                                // that each channel's store is one full 128-byte line; 0: 16 x 16 tile, 8 x 4 warp patch (slices:
                                // an oblique plane then touches the fewest 128-byte lines per gather instruction)
    int tile_r, tile_c;
};

struct Pix {
    int i, j, k;                // output indices (Do, Ho, Wo)
    bool valid;
};

__device__ __forceinline__ Pix pixel_of_tile(const OutGeom& g, int tile) {
    const int tid = threadIdx.x;
    const int w = tid >> 5, lane = tid & 31;
    const int lc = g.wide ? lane : ((w & 1) << 3) + (lane & 7);
    const int lr = g.wide ? w : ((w >> 1) << 2) + (lane >> 3);
    int tr, tc;
    if (g.tiles_c_shift >= 0) { tr = tile >> g.tiles_c_shift; tc = tile & (g.tiles_c - 1); }
    else { tr = tile / g.tiles_c; tc = tile - tr * g.tiles_c; }
    const int row = tr * g.tile_r + lr, col = tc * g.tile_c + lc;
    Pix p;
    p.valid = row < g.rows && col < g.cols;
    if (g.Wo == 1) {
        p.i = row; p.j = col; p.k = 0;
    } else {
        p.i = row / g.Ho; p.j = row % g.Ho; p.k = col;
    }
    return p;
}

__device__ __forceinline__ Pix pixel_of_thread(const OutGeom& g) { return pixel_of_tile(g, blockIdx.x); }

struct Sample {                 // un-normalised source coordinates of one output location
    float ix, iy, iz;
    float bx, by, bz;           // normalised base coordinates (x_k, y_j, z_i)
};

// grid affine of slice s: the first 12 floats of its ViewState (uniform address -> one broadcast per warp)
__device__ __forceinline__ Sample sample_coords(const OutGeom& g, const Pix& p, const ViewArgs& va, int s, const VolArgs& vol) {
    // three 16-byte broadcast loads (ViewState is 16-byte aligned and starts with G')
    const float4* __restrict__ G = reinterpret_cast<const float4*>(reinterpret_cast<const ViewState*>(va.state) + s);
    float t[12];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const float4 r = __ldg(G + q);
        t[4 * q] = r.x; t[4 * q + 1] = r.y; t[4 * q + 2] = r.z; t[4 * q + 3] = r.w;
    }
    Sample sm;
    sm.bx = base_coord(p.k, g.ax);
    sm.by = base_coord(p.j, g.ay);
    sm.bz = base_coord(p.i, g.az);
    sm.ix = unnormalize(grid_coord(t + 0, sm.bx, sm.by, sm.bz), (float)vol.W);
    sm.iy = unnormalize(grid_coord(t + 4, sm.bx, sm.by, sm.bz), (float)vol.H);
    sm.iz = unnormalize(grid_coord(t + 8, sm.bx, sm.by, sm.bz), (float)vol.D);
    return sm;
}

// Channel vectors of the channels-last kernels.  One gather moves VB = 16 or 32 bytes of a voxel's contiguous channels:
// sm_100 has 256-bit global loads (ld.global.nc.v8.b32 -> LDG.E.256), so the 8 fp32 channels of a one-hot corner
// (exactly one 32-byte L2 sector) are ONE load instruction instead of two - half the LSU instructions and half the
// per-instruction line replays of an L1-wavefront-bound gather.
template <int VB> struct Raw { unsigned w[VB / 4]; };

template <int VB> __device__ __forceinline__ Raw<VB> gather_nc(const void* p);
template <> __device__ __forceinline__ Raw<16> gather_nc<16>(const void* p) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    Raw<16> r;
    r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
    return r;
}
template <> __device__ __forceinline__ Raw<32> gather_nc<32>(const void* p) {     // p must be 32-byte aligned
    Raw<32> r;
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
        : "l"(p));
    return r;
}
template <int VB> __device__ __forceinline__ Raw<VB> raw_zero() {
    Raw<VB> r;
#pragma unroll
    for (int i = 0; i < VB / 4; ++i) r.w[i] = 0u;
    return r;
}

// how the storage type widens to fp32, per 32-bit word of a gathered vector
template <typename T> struct Widen { static constexpr bool is_float = false; };
template <> struct Widen<float> {
    static constexpr bool is_float = true;
    template <int VB> static __device__ __forceinline__ void decode(const Raw<VB>& r, float* o) {
#pragma unroll
        for (int i = 0; i < VB / 4; ++i) o[i] = __uint_as_float(r.w[i]);
    }
};
template <> struct Widen<__nv_bfloat16> {
    static constexpr bool is_float = true;
    template <int VB> static __device__ __forceinline__ void decode(const Raw<VB>& r, float* o) {
#pragma unroll
        for (int i = 0; i < VB / 4; ++i) {      // bf16 -> fp32 is a 16-bit shift
            o[2 * i] = __uint_as_float(r.w[i] << 16);
            o[2 * i + 1] = __uint_as_float(r.w[i] & 0xffff0000u);
        }
    }
};
template <> struct Widen<__half> {
    static constexpr bool is_float = true;
    template <int VB> static __device__ __forceinline__ void decode(const Raw<VB>& r, float* o) {
#pragma unroll
        for (int i = 0; i < VB / 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r.w[i]));
            o[2 * i] = f.x; o[2 * i + 1] = f.y;
        }
    }
};

// The 8 trilinear corners of one sample, kept lean (one base offset + 6 axis weights + in-bounds mask);
// corner k = (dx,dy,dz) = (k&1, (k>>1)&1, k>>2) is ATen's order tnw,tne,tsw,tse,bnw,bne,bsw,bse.
// Weights are formed exactly as ATen does: (wx*wy)*wz.
struct Corners {
    int base;                   // element offset of corner (x0,y0,z0); only dereferenced where `inb` allows
    unsigned inb;               // bit k set <=> corner k inside the volume
    float wx[2], wy[2], wz[2];
    __device__ __forceinline__ float w(int k) const {
        return __fmul_rn(__fmul_rn(wx[k & 1], wy[(k >> 1) & 1]), wz[k >> 2]);
    }
    __device__ __forceinline__ int off(int k, const VolArgs& vol) const {
        return base + ((k & 1) ? (int)vol.sW : 0) + (((k >> 1) & 1) ? (int)vol.sH : 0) + ((k >> 2) ? (int)vol.sD : 0);
    }
    __device__ __forceinline__ bool in(int k) const { return (inb >> k) & 1u; }
};

__device__ __forceinline__ Corners corners_of(const Sample& s, const VolArgs& vol) {
    Corners c;
    const float x0f = floorf(s.ix), y0f = floorf(s.iy), z0f = floorf(s.iz);
    // clamp far-out-of-field samples so that the base offset cannot overflow (all their corners are masked)
    const int x0 = max(-2, min(__float2int_rd(s.ix), vol.W + 1));
    const int y0 = max(-2, min(__float2int_rd(s.iy), vol.H + 1));
    const int z0 = max(-2, min(__float2int_rd(s.iz), vol.D + 1));
    c.wx[0] = __fsub_rn(__fadd_rn(x0f, 1.0f), s.ix); c.wx[1] = __fsub_rn(s.ix, x0f);
    c.wy[0] = __fsub_rn(__fadd_rn(y0f, 1.0f), s.iy); c.wy[1] = __fsub_rn(s.iy, y0f);
    c.wz[0] = __fsub_rn(__fadd_rn(z0f, 1.0f), s.iz); c.wz[1] = __fsub_rn(s.iz, z0f);
    const bool xin[2] = {x0 >= 0 && x0 < vol.W, x0 + 1 >= 0 && x0 + 1 < vol.W};
    const bool yin[2] = {y0 >= 0 && y0 < vol.H, y0 + 1 >= 0 && y0 + 1 < vol.H};
    const bool zin[2] = {z0 >= 0 && z0 < vol.D, z0 + 1 >= 0 && z0 + 1 < vol.D};
    c.inb = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) c.inb |= (xin[k & 1] && yin[(k >> 1) & 1] && zin[k >> 2]) ? (1u << k) : 0u;
    c.base = z0 * (int)vol.sD + y0 * (int)vol.sH + x0 * (int)vol.sW;
    return c;
}

__device__ __forceinline__ float pad_of(int pad_mode, float pad_value, const float* pad_device) {
    if (pad_mode == AFB_PAD_DEVICE) return __ldg(pad_device);
    if (pad_mode == AFB_PAD_VALUE) return pad_value;
    return 0.0f;
}

__device__ __forceinline__ void grid_grad_parts(const float* dot, const Corners& cn, const Sample& sm, const VolArgs& vol,
                                                float gsum, float* part) {
    // d out / d (ix,iy,iz): sign pattern of ATen grid_sampler_3d_backward
    float gix = 0.0f, giy = 0.0f, giz = 0.0f, wsum = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
        const float d = cn.in(k) ? dot[k] : 0.0f;
        gix += (dx ? d : -d) * cn.wy[dy] * cn.wz[dz];
        giy += (dy ? d : -d) * cn.wx[dx] * cn.wz[dz];
        giz += (dz ? d : -d) * cn.wx[dx] * cn.wy[dy];
        wsum += cn.in(k) ? cn.w(k) : 0.0f;
    }
    const float ggx = gix * (0.5f * (float)vol.W), ggy = giy * (0.5f * (float)vol.H), ggz = giz * (0.5f * (float)vol.D);
    part[0] = ggx * sm.bx; part[1] = ggx * sm.by; part[2] = ggx * sm.bz; part[3] = ggx;
    part[4] = ggy * sm.bx; part[5] = ggy * sm.by; part[6] = ggy * sm.bz; part[7] = ggy;
    part[8] = ggz * sm.bx; part[9] = ggz * sm.by; part[10] = ggz * sm.bz; part[11] = ggz;
    part[12] = gsum * (1.0f - wsum);
}

__device__ __forceinline__ void bwd_reduce(int s, const float* part, int pad_mode, float* __restrict__ d_pad,
                                           double* __restrict__ ws_acc) {
    __shared__ float red[NTHREADS / 32][13];
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
}

// allow_wide: 32 x 1 warp rows when Wo >= 32 (full-line stores of the float samplers' 3-D outputs).  The one-hot-from-index
// samplers pass false: their gathers are single bytes, the 8 x 4 patch keeps more of them in one sector (f1 from uint8
// labels, 128^3 -> 128^3, B = 8, three calls: 1.47 ms against 1.58 ms with rows)
inline OutGeom make_geom(int Do, int Ho, int Wo, bool allow_wide = true) {
    OutGeom g;
    g.ax = make_axis(Wo); g.ay = make_axis(Ho); g.az = make_axis(Do);
    g.Do = Do; g.Ho = Ho; g.Wo = Wo;
    if (Wo == 1) { g.rows = Do; g.cols = Ho; } else { g.rows = Do * Ho; g.cols = Wo; }
    g.wide = (allow_wide && Wo >= 32 && !getenv("AFB_NO_WIDE_PATCH")) ? 1 : 0;
    g.tile_r = g.wide ? NTHREADS / 32 : TILE;
    g.tile_c = g.wide ? 32 : TILE;
    g.tiles_c = (g.cols + g.tile_c - 1) / g.tile_c;
    g.tiles_c_shift = -1;
    for (int sh = 0; sh < 31; ++sh)
        if ((1 << sh) == g.tiles_c) g.tiles_c_shift = sh;
    return g;
}

// (tiles of one slice, views, volumes): slice s = blockIdx.z * V + blockIdx.y, volume b = blockIdx.z
inline dim3 slice_grid(const OutGeom& g, int B, int V) { return dim3(((g.rows + g.tile_r - 1) / g.tile_r) * g.tiles_c, V, B); }


}  // namespace afb
