// afb_onehot.cu - slicing one-hot label volumes straight from the INTEGER label map.
//
// In the reference the soft label handed to the slicer is always one_hot(label).float()
// (running/run_dl.py:261-264), an 8-channel fp32 volume (64 MiB at 128^3) next to the same one-hot as int64
// (128 MiB).  Both slicings can be synthesised from the index map (2 MiB as uint8):
//     y_soft[c]  = sum_k w_k * [label_k == c]      (bilinear of the one-hot; pad value = min = 0)
//     y_label[c] = [label_nearest == c]             (nearest of the int64 one-hot)
// The sums are accumulated in ATen's corner order with unfused adds and the skipped terms are exact zeros, so
// y_soft is BITWISE what afb_slice_fwd / the reference produce on the materialised one-hot volume; out-of-field
// samples give the all-zero one-hot (argmax 0 = background) exactly as there.  The backward needs no volume
// gradient (integer input - this is the reference's training case: only dTheta is consumed) and the corner
// "dot" of the dGrid formula collapses to a register lookup grad_out[label_k].
// Gather traffic drops from 8 corners x 32 B (fp32 one-hot) to 8 corners x 1 B per output location.
#include "afb_sampler.cuh"

namespace afb {

constexpr int MAXC = 16;          // channels held in registers

struct LabArgs {
    const void* data;
    int B, D, H, W;
    long long sB, sD, sH, sW;
    int C;
};

template <typename L>
__device__ __forceinline__ int load_label(const L* p) { return (int)__ldg(p); }
template <>
__device__ __forceinline__ int load_label<int64_t>(const int64_t* p) { return (int)__ldg((const long long*)p); }

// labels of the 8 corners (-1 where the corner is outside the volume)
template <typename L>
__device__ __forceinline__ void corner_labels(const L* __restrict__ src, const Corners& cn, const VolArgs& vol, int* lab) {
#pragma unroll
    for (int k = 0; k < 8; ++k) lab[k] = cn.in(k) ? load_label<L>(src + cn.off(k, vol)) : -1;
}

template <typename L, int LABEL_OUT>      // LABEL_OUT: 0 none, 1 one-hot int64 [S,C,N], 2 index uint8 [S,N]
__global__ void __launch_bounds__(NTHREADS, 3)
onehot_fwd_kernel(LabArgs la, ViewArgs va, OutGeom g, float* __restrict__ y_soft, void* __restrict__ y_label) {
    const int s = blockIdx.z * va.V + blockIdx.y;      // grid = (tiles, V, B): no integer division for the batch index
    const Pix p = pixel_of_thread(g);
    if (!p.valid) return;
    VolArgs vol;
    vol.data = la.data; vol.B = la.B; vol.C = 1; vol.D = la.D; vol.H = la.H; vol.W = la.W;
    vol.sB = la.sB; vol.sC = 0; vol.sD = la.sD; vol.sH = la.sH; vol.sW = la.sW;
    const Sample sm = sample_coords(g, p, va, s, vol);
    const int b = blockIdx.z;
    const L* __restrict__ src = (const L*)la.data + (long long)b * la.sB;
    const int plane = g.Do * g.Ho * g.Wo;      // C * plane < 2^31 (checked on the host): 32-bit channel offsets
    const size_t pix = ((size_t)p.i * g.Ho + p.j) * g.Wo + p.k;
    if (y_soft) {
        const Corners cn = corners_of(sm, vol);
        int lab[8];
        corner_labels<L>(src, cn, vol, lab);
        float w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = cn.w(k);
        float* __restrict__ dst = y_soft + (size_t)s * la.C * plane + pix;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (c < la.C) {
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = __fadd_rn(acc, lab[k] == c ? w[k] : 0.0f);
                dst[c * plane] = acc;
            }
        }
    }
    if (LABEL_OUT != 0) {
        const int xn = __float2int_rn(sm.ix), yn = __float2int_rn(sm.iy), zn = __float2int_rn(sm.iz);
        const bool in = xn >= 0 && xn < la.W && yn >= 0 && yn < la.H && zn >= 0 && zn < la.D;
        const int ln = in ? load_label<L>(src + (long long)zn * la.sD + (long long)yn * la.sH + (long long)xn * la.sW) : -1;
        if (LABEL_OUT == 1) {
            int64_t* __restrict__ dl = (int64_t*)y_label + (size_t)s * la.C * plane + pix;
            for (int c = 0; c < la.C; ++c) dl[c * plane] = (ln == c) ? 1 : 0;
        } else {
            ((uint8_t*)y_label)[(size_t)s * plane + pix] = in ? (uint8_t)ln : (uint8_t)0;
        }
    }
}

template <typename L>
__global__ void __launch_bounds__(NTHREADS, 3)
onehot_bwd_kernel(LabArgs la, ViewArgs va, OutGeom g, const float* __restrict__ grad_out, double* __restrict__ ws_acc) {
    const int s = blockIdx.z * va.V + blockIdx.y;      // grid = (tiles, V, B): no integer division for the batch index
    float part[13];
#pragma unroll
    for (int q = 0; q < 13; ++q) part[q] = 0.0f;
    const Pix p = pixel_of_thread(g);
    if (p.valid) {
        VolArgs vol;
        vol.data = la.data; vol.B = la.B; vol.C = 1; vol.D = la.D; vol.H = la.H; vol.W = la.W;
        vol.sB = la.sB; vol.sC = 0; vol.sD = la.sD; vol.sH = la.sH; vol.sW = la.sW;
        const Sample sm = sample_coords(g, p, va, s, vol);
        const Corners cn = corners_of(sm, vol);
        const int b = blockIdx.z;
        const L* __restrict__ src = (const L*)la.data + (long long)b * la.sB;
        int lab[8];
        corner_labels<L>(src, cn, vol, lab);
        const int plane = g.Do * g.Ho * g.Wo;      // C * plane < 2^31 (checked on the host): 32-bit channel offsets
        const float* __restrict__ go_p = grad_out + (size_t)s * la.C * plane + ((size_t)p.i * g.Ho + p.j) * g.Wo + p.k;
        float dot[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) dot[k] = 0.0f;
        float gsum = 0.0f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (c < la.C) {
                const float go = __ldg(go_p + c * plane);
                gsum += go;
#pragma unroll
                for (int k = 0; k < 8; ++k) dot[k] += (lab[k] == c) ? go : 0.0f;      // dot_k = grad_out[label_k]
            }
        }
        grid_grad_parts(dot, cn, sm, vol, gsum, part);
    }
    bwd_reduce(s, part, AFB_PAD_ZERO, nullptr, ws_acc);
}

static int make_lab_args(const afb_volume* lab, int num_classes, LabArgs& la) {
    if (!lab || !lab->data) return AFB_EINVAL;
    if (lab->C != 1 || num_classes <= 0 || num_classes > MAXC) return AFB_ESHAPE;
    if (lab->sD < 0 || lab->sH < 0 || lab->sW < 0) return AFB_EUNSUPPORTED;
    if ((long long)(lab->D + 2) * lab->sD + (long long)(lab->H + 2) * lab->sH + (long long)(lab->W + 2) * lab->sW >= 2147483647ll)
        return AFB_EUNSUPPORTED;
    la.data = lab->data; la.B = lab->B; la.D = lab->D; la.H = lab->H; la.W = lab->W;
    la.sB = lab->sB; la.sD = lab->sD; la.sH = lab->sH; la.sW = lab->sW; la.C = num_classes;
    return AFB_OK;
}

template <typename L>
static int launch_onehot_fwd(const LabArgs& la, const ViewArgs& a, const OutGeom& g, float* y_soft, void* y_label,
                             int label_out, cudaStream_t st) {
    const dim3 grid = slice_grid(g, la.B, a.V);
    switch (label_out) {
        case 0: onehot_fwd_kernel<L, 0><<<grid, NTHREADS, 0, st>>>(la, a, g, y_soft, y_label); break;
        case 1: onehot_fwd_kernel<L, 1><<<grid, NTHREADS, 0, st>>>(la, a, g, y_soft, y_label); break;
        case 2: onehot_fwd_kernel<L, 2><<<grid, NTHREADS, 0, st>>>(la, a, g, y_soft, y_label); break;
        default: return AFB_EINVAL;
    }
    return (int)cudaGetLastError();
}

}  // namespace afb

using namespace afb;

extern "C" int afb_slice_onehot_fwd(const afb_volume* labels, int num_classes, const afb_views* views, int Do, int Ho, int Wo,
                                    float* y_soft, void* y_label, int label_out, void* stream) {
    LabArgs la; ViewArgs a;
    int rc = make_lab_args(labels, num_classes, la);
    if (rc != AFB_OK) return rc;
    rc = make_view_args(views, la.B, la.D, la.H, la.W, Do, Ho, Wo, /*need_state=*/true, a);
    if (rc != AFB_OK) return rc;
    if (!y_soft && label_out == 0) return AFB_EINVAL;
    if (label_out != 0 && !y_label) return AFB_EINVAL;
    if ((long long)num_classes * Do * Ho * Wo >= 2147483647ll) return AFB_EUNSUPPORTED;     // 32-bit channel offsets
    const OutGeom g = make_geom(Do, Ho, Wo, /*allow_wide=*/false);
    cudaStream_t st = (cudaStream_t)stream;
    switch (labels->dtype) {
        case AFB_U8: return launch_onehot_fwd<uint8_t>(la, a, g, y_soft, y_label, label_out, st);
        case AFB_I16: return launch_onehot_fwd<int16_t>(la, a, g, y_soft, y_label, label_out, st);
        case AFB_I32: return launch_onehot_fwd<int32_t>(la, a, g, y_soft, y_label, label_out, st);
        case AFB_I64: return launch_onehot_fwd<int64_t>(la, a, g, y_soft, y_label, label_out, st);
        default: return AFB_EDTYPE;
    }
}

extern "C" int afb_slice_onehot_bwd(const afb_volume* labels, int num_classes, const afb_views* views, int Do, int Ho, int Wo,
                                    const float* grad_y_soft, const float* grad_grid_affine, float* d_affine, float* d_gpre,
                                    void* workspace, void* stream) {
    LabArgs la; ViewArgs a;
    int rc = make_lab_args(labels, num_classes, la);
    if (rc != AFB_OK) return rc;
    rc = make_view_args(views, la.B, la.D, la.H, la.W, Do, Ho, Wo, /*need_state=*/true, a);
    if (rc != AFB_OK) return rc;
    if (!workspace || (!grad_y_soft && !grad_grid_affine) || !d_affine) return AFB_EINVAL;
    if (a.kind == AFB_AFFINE_PARAMS && !views->params) return AFB_EINVAL;
    if ((long long)num_classes * Do * Ho * Wo >= 2147483647ll) return AFB_EUNSUPPORTED;     // 32-bit channel offsets
    const OutGeom g = make_geom(Do, Ho, Wo, /*allow_wide=*/false);
    const int S = la.B * a.V;
    const dim3 grid = slice_grid(g, la.B, a.V);
    double* acc = (double*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_y_soft) {
        switch (labels->dtype) {
            case AFB_U8: onehot_bwd_kernel<uint8_t><<<grid, NTHREADS, 0, st>>>(la, a, g, grad_y_soft, acc); break;
            case AFB_I16: onehot_bwd_kernel<int16_t><<<grid, NTHREADS, 0, st>>>(la, a, g, grad_y_soft, acc); break;
            case AFB_I32: onehot_bwd_kernel<int32_t><<<grid, NTHREADS, 0, st>>>(la, a, g, grad_y_soft, acc); break;
            case AFB_I64: onehot_bwd_kernel<int64_t><<<grid, NTHREADS, 0, st>>>(la, a, g, grad_y_soft, acc); break;
            default: return AFB_EDTYPE;
        }
        rc = (int)cudaGetLastError();
        if (rc != 0) return rc;
    }
    return launch_view_chain(a, S, acc, grad_grid_affine, d_affine, d_gpre, st);
}
