// afb_views.cu - the view-parameter side of the path, kept out of the samplers so that those stay lean
// (the fp64 4x4 algebra would otherwise set the register count of every sampler thread):
//   view_prologue_kernel : raw view input -> ViewState per slice (one warp per slice, ONE launch per
//                          acquisition, shared by the soft-label / label / image samplers and the backward)
//                          = utils/nifti_utils.py:36-71 + models/learnable_transform.py:144-230,262-289
//   view_chain_kernel    : total dG' (sampler reduction + upstream grad of grid_affine) -> gradient of the
//                          view input (theta | P | raw parameters): analytic chain, SURVEY 3.5
#include "afb_device.cuh"

namespace afb {

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
    const unsigned* src = reinterpret_cast<const unsigned*>(&st);
    unsigned* dst = reinterpret_cast<unsigned*>(states + s);
    for (int i = lane; i < (int)(sizeof(ViewState) / 4); i += 32) dst[i] = src[i];
}

// One warp per slice: dG' = sampler sums (fp64 workspace, re-zeroed here) + upstream grad -> d(view input).
__global__ void __launch_bounds__(128)
view_chain_kernel(ViewArgs va, int S, double* __restrict__ ws_acc, const float* __restrict__ grad_grid_affine,
                  float* __restrict__ d_affine, float* __restrict__ d_gpre) {
    __shared__ ViewState sh[4];
    __shared__ ChainScratch cs[4];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * 4 + w;
    if (s >= S) return;
    const unsigned* src = reinterpret_cast<const unsigned*>(reinterpret_cast<const ViewState*>(va.state) + s);
    unsigned* dst = reinterpret_cast<unsigned*>(&sh[w]);
    for (int i = lane; i < (int)(sizeof(ViewState) / 4); i += 32) dst[i] = src[i];
    if (lane < 16) {
        double t = lane < 12 ? ws_acc[(size_t)s * 16 + lane] : 0.0;
        if (grad_grid_affine) t += (double)grad_grid_affine[(size_t)s * 16 + lane];
        cs[w].dG[lane] = t;
        ws_acc[(size_t)s * 16 + lane] = 0.0;              // leave the workspace zeroed for the next call
    }
    __syncwarp();
    if (d_affine) view_backward_warp0(va, s, sh[w], cs[w], d_affine, d_gpre);
}

int launch_view_chain(const ViewArgs& a, int S, double* ws_acc, const float* grad_grid_affine, float* d_affine,
                      float* d_gpre, cudaStream_t st) {
    view_chain_kernel<<<(S + 3) / 4, 128, 0, st>>>(a, S, ws_acc, grad_grid_affine, d_affine, d_gpre);
    return (int)cudaGetLastError();
}

}  // namespace afb

using namespace afb;

extern "C" int64_t afb_view_state_bytes(void) { return (int64_t)sizeof(ViewState); }

extern "C" int afb_view_prologue(const afb_views* views, int B, int D, int H, int W, int Do, int Ho, int Wo, void* state,
                                 float* grid_affine_out, double* nii_affine_out, float* theta_out, void* stream) {
    ViewArgs a;
    int rc = make_view_args(views, B, D, H, W, Do, Ho, Wo, /*need_state=*/false, a);
    if (rc != AFB_OK) return rc;
    if (!state) return AFB_EINVAL;
    if (views->kind == AFB_AFFINE_GRID && !views->theta) return AFB_EINVAL;
    if (views->kind == AFB_AFFINE_PRE && !views->pre) return AFB_EINVAL;
    if (views->kind == AFB_AFFINE_PARAMS && (!views->params || !views->gpre || !views->init)) return AFB_EINVAL;
    a.state = nullptr;
    const int S = B * views->V;
    view_prologue_kernel<<<(S + 3) / 4, 128, 0, (cudaStream_t)stream>>>(a, S, (ViewState*)state, grid_affine_out, nii_affine_out, theta_out);
    return (int)cudaGetLastError();
}
