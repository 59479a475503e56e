/* afb200.h - C ABI of libafb200.so: B200 (sm_100a) kernels for the differentiable
 * view-acquisition hot path of multimodallearning/acquisition-focus.
 *
 * The reference has no FFI layer of its own: the path is pure Python over torch ATen.  The
 * boundary this library replaces is therefore the set of ATen calls the reference makes
 * (paths relative to /root/reference/acquisition_focus):
 *
 *   utils/nifti_utils.py:36-71    fp64 affine bookkeeping (grid affine + NIfTI affine)
 *   utils/nifti_utils.py:182-184  F.affine_grid
 *   utils/nifti_utils.py:87-94    F.grid_sample (checkpointed), bilinear / nearest
 *   utils/nifti_utils.py:200-203  min-shift  (volume - min, + min)
 *   utils/transform_utils.py:27-58            R6 -> rotation matrix
 *   models/learnable_transform.py:144-230     init / batch affines (R6, soft-argmax offset, tanh zoom)
 *   models/learnable_transform.py:262-289     theta = T@R@Z, pre = Gpre @ theta
 *   models/hybrid_unet.py:71-94               SkipConnector: slice -> 3-D embedding
 *   + the autograd of all of the above.
 *
 * Conventions
 *   - every entry point returns int: 0 = ok, >0 = cudaError_t of the launch, <0 = AFB_E* argument error
 *   - all data pointers are CALLER-OWNED DEVICE pointers unless marked "host"; the library
 *     allocates nothing and keeps no state; work is enqueued on `stream` and is asynchronous
 *   - one device per call (the caller sets the current device); re-entrant, thread-safe
 *   - tensors are row-major; volume strides are in ELEMENTS; (D,H,W) index order throughout,
 *     torch grid convention for affines (x->W, y->H, z->D, normalised [-1,1], align_corners=False)
 *   - slices are ordered [b][v] (batch-major): slice s = b*V + v
 */
#ifndef AFB200_H
#define AFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFB_VERSION 100

/* error codes (negative) */
#define AFB_OK 0
#define AFB_EINVAL (-1)      /* bad argument (null pointer, non-positive size, bad enum) */
#define AFB_EDTYPE (-2)      /* dtype/mode combination not supported */
#define AFB_ESHAPE (-3)      /* inconsistent shapes */
#define AFB_EUNSUPPORTED (-4)

/* storage dtypes */
#define AFB_F32 0
#define AFB_BF16 1
#define AFB_F16 2
#define AFB_I64 3
#define AFB_I32 4
#define AFB_I16 5
#define AFB_U8 6

/* interpolation (F.grid_sample mode; padding_mode='zeros', align_corners=False always) */
#define AFB_BILINEAR 0
#define AFB_NEAREST 1

/* out-of-bounds value of the bilinear path.  The reference shifts by the global minimum of the
 * tensor passed in one call (nifti_utils.py:200-203), so out-of-bounds evaluates to min(volume). */
#define AFB_PAD_ZERO 0      /* plain grid_sample zeros padding */
#define AFB_PAD_VALUE 1     /* host-supplied value (e.g. 0 for one-hot volumes)            */
#define AFB_PAD_DEVICE 2    /* value read from a device float (output of afb_volume_min)   */

/* how the per-slice grid affine is obtained (fused prologue of the sampler) */
#define AFB_AFFINE_GRID 0   /* theta[S,3,4] fp32 used as is (F.affine_grid input)                         */
#define AFB_AFFINE_PRE 1    /* pre_grid_sample_affine P[S,4,4] -> nifti_utils.py:36-58                    */
#define AFB_AFFINE_PARAMS 2 /* raw view parameters -> learnable_transform.py:144-230,262-289 -> PRE path  */

typedef struct afb_volume {
    const void* data;       /* [B,C,D,H,W] with arbitrary element strides */
    int dtype;              /* AFB_F32 ... */
    int B, C, D, H, W;
    int64_t sB, sC, sD, sH, sW;
} afb_volume;

typedef struct afb_views {
    int kind;               /* AFB_AFFINE_* */
    int V;                  /* views per volume; S = B*V slices, slice s = b*V + v */
    /* AFB_AFFINE_GRID */
    const float* theta;     /* [S,3,4] */
    /* AFB_AFFINE_PRE */
    const void* pre;        /* [S,4,4] fp32 or fp64 */
    int pre_is_f64;
    /* AFB_AFFINE_PARAMS: per slice [R6(6) | offset logits (3*R) | zoom logit (1)] (MLP-head output) */
    const float* params;    /* [S, 6+3R+1] */
    const float* gpre;      /* [S,4,4] fp32 grid_affine_pre_mlp (clinical view affine incl. augmentation) */
    const float* init;      /* [V,10]: init_theta_ap[6], init_theta_t_offsets[3], init_theta_zp[1]        */
    int R;                  /* vox_range = round(offset_clip*spat), learnable_transform.py:112-115        */
    int spat;               /* volume_fov_vox[0], learnable_transform.py:110                              */
    float offset_clip;      /* 0 => offsets forced to 0 (learnable_transform.py:211-212)                  */
    float zoom_clip;
    /* PRE and PARAMS: NIfTI bookkeeping inputs */
    const double* nii_affine;   /* [B,4,4] fp64 (NULL => identity; only ratios/zooms of it are used)      */
    double fov_mm[3];           /* host: target_fov_mm in (D,H,W) order; <=0 => input FOV (nifti_utils.py:140) */
    /* per-slice view state, [S * afb_view_state_bytes()] bytes of caller-owned device memory: WRITTEN by
     * afb_view_prologue, READ by afb_slice_fwd / afb_slice_bwd / afb_slice_pad_grad (required there).  The
     * soft-label, label and image slicings of one acquisition and their backward share ONE prologue launch;
     * keeping the fp64 4x4 algebra out of the samplers is what keeps those at <= 85 registers per thread. */
    void* state;
} afb_views;

/* ---- library info -------------------------------------------------------------------------- */
int afb_version(void);
const char* afb_error_string(int code);

/* measurement helper (bench.py): `passes` read sweeps over buf[n_bytes] with 16-byte ld.global.cg loads in one
 * launch; with n_bytes <= ~32 MiB the buffer is L2 resident and n_bytes*passes/time is the L2 read bandwidth. */
int afb_probe_read(const void* buf, int64_t n_bytes, int passes, float* sink, void* stream);

/* ---- host-side stage of the upload path (runs on the host cores, no CUDA call) ------------------------------
 * Packs an integer label map (AFB_I64 | AFB_I32 | AFB_I16; run_dl.py:261 uploads torch.long) to uint8 with n_threads host
 * threads, so that 1 instead of 8 bytes per voxel cross PCIe; afb_onehot_expand takes the uint8 map.  *out_of_range = 1 when a
 * value lies outside [0, 255] (the output is then unusable).  src and dst are HOST pointers (pinned or pageable). */
int afb_host_narrow_labels(const void* src, int src_dtype, int64_t n, uint8_t* dst, int n_threads, int* out_of_range);

/* ---- min pre-pass of the bilinear path (nifti_utils.py:200) -------------------------------- */
/* Scans n_elements of a dense tensor; writes out_min_count[0] = min as float32,
 * out_min_count[1] = number of elements equal to the min (as float32, exact < 2^24, else rounded).
 * workspace: >= afb_volume_min_workspace_bytes() bytes, contents arbitrary. */
int64_t afb_volume_min_workspace_bytes(void);
int afb_volume_min(const void* data, int dtype, int64_t n_elements, float* out_min_count,
                   void* workspace, void* stream);

/* fp32 variant that additionally leaves a 1-bit-per-voxel record (per 512-voxel chunk: its minimum + the bitmask
 * "== chunk minimum") in `mask` (>= afb_min_mask_bytes(n) bytes, 16-byte aligned).  afb_min_grad_fill_mask rebuilds
 * MinBackward from it WITHOUT re-reading the volume: 4.1 instead of 8 bytes of HBM traffic per voxel.        */
int64_t afb_min_mask_bytes(int64_t n_elements);
int afb_volume_min_mask(const float* data, int64_t n_elements, float* out_min_count, void* mask,
                        void* workspace, void* stream);
int afb_min_grad_fill_mask(const void* mask, int64_t n_elements, const float* min_count, const float* d_pad,
                           float* d_vol, void* stream);

/* The same record for bf16 / fp16 volumes (dtype AFB_BF16 | AFB_F16): identical size (afb_min_mask_bytes) and meaning,
 * lane vectors of 8 elements; the fill still writes fp32 dVolume (32-byte aligned).  afb_min_count_from_mask serves both. */
int afb_volume_min_mask_half(const void* data, int dtype, int64_t n_elements, float* out_min_count, void* mask,
                             void* workspace, void* stream);
int afb_min_grad_fill_mask_half(const void* mask, int64_t n_elements, const float* min_count, const float* d_pad,
                                float* d_vol, void* stream);

/* fp32 -> bf16 / fp16 (round to nearest even) over n contiguous elements: dVolume of a half-precision volume is
 * accumulated in fp32 (afb_slice_bwd) and handed back in the volume's own dtype, as the reference's autograd does. */
int afb_cast_from_f32(const float* src, void* dst, int dst_dtype, int64_t n_elements, void* stream);

/* ---- one-hot materialisation (running/run_dl.py:261-264) fused with the min record ---------- */
/* labels: n_voxels integers (label_dtype: AFB_U8/I16/I32/I64).  Writes, channels-last ([voxel][class], i.e. the strides
 * of `one_hot(label, C).permute(0,4,1,2,3)` and of its `.float()`):
 *   onehot_i64 (nullable) : int64 one-hot, what run_dl.py:261-262 hands the nearest-neighbour label slicing
 *   soft_f32   (nullable) : fp32 one-hot, what run_dl.py:263-264 hands the bilinear soft-label slicing
 * and, when `mask` is given, the soft volume's chunk record at its place inside the record of a tensor of
 * total_elements fp32 values that this range starts at element elem_offset of (a multiple of 512; ranges other than
 * the last must also hold a multiple of 512 elements): a host batch can be expanded range by range while later
 * ranges are still crossing PCIe.  afb_min_count_from_mask then gives [min, multiplicity] of the whole tensor from
 * the record alone - no 4 B/voxel min pass over a volume this library produced itself.
 * Labels outside [0, num_classes) give an all-zero voxel (torch's one_hot raises instead).                       */
int afb_onehot_expand(const void* labels, int label_dtype, int64_t n_voxels, int num_classes, int64_t* onehot_i64,
                      float* soft_f32, void* mask, int64_t total_elements, int64_t elem_offset, void* stream);
int afb_min_count_from_mask(const void* mask, int64_t n_elements, float* out_min_count, void* workspace, void* stream);

/* ---- view prologue: raw view input -> grid affine, once per acquisition -------------------------
 * Computes for all S = B*V slices what nifti_utils.py:36-71 and learnable_transform.py:144-230,262-289
 * compute on the host in ~100 tiny fp32/fp64 torch ops: state (for the samplers), grid_affine_out
 * [S,4,4] fp32, nii_affine_out [S,4,4] fp64 (PRE/PARAMS), theta_out [S,4,4] fp32 (PARAMS). Any output
 * may be NULL.  (D,H,W) = input volume size, (Do,Ho,Wo) = output size.                              */
int64_t afb_view_state_bytes(void);
int afb_view_prologue(const afb_views* views, int B, int D, int H, int W, int Do, int Ho, int Wo,
                      void* state, float* grid_affine_out, double* nii_affine_out, float* theta_out,
                      void* stream);

/* ---- slice / volume extraction, forward ------------------------------------------------------
 * out[b, v, c, i, j, k] for (i,j,k) in (Do,Ho,Wo); out dtype = volume dtype.  views->state must have been
 * filled by afb_view_prologue for the same (B, V, sizes).  One launch for all S slices.             */
int afb_slice_fwd(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo, int mode,
                  int pad_mode, float pad_value, const float* pad_device, void* out, void* stream);

/* The three slicings of ONE acquisition in one launch (models/learnable_transform.py:287-306: soft label bilinear, one-hot
 * label nearest, image bilinear with the same pre-affine): coordinates, corners and weights are computed once per output
 * location.  soft: fp32, channels-last (sC == 1, C % 4 == 0, 16-byte aligned); label (may be NULL): u8/i16/i32/i64
 * channels-last with 16-byte channel vectors; image (may be NULL): fp32, any strides.  All three share B, D, H, W.
 * Results are bitwise those of three afb_slice_fwd calls.  Returns AFB_EUNSUPPORTED when a layout does not qualify (call
 * afb_slice_fwd per volume instead). */
int afb_slice_fwd3(const afb_volume* soft, const afb_volume* label, const afb_volume* image, const afb_views* views,
                   int Do, int Ho, int Wo, int pad_mode_soft, float pad_value_soft, const float* pad_device_soft,
                   int pad_mode_image, float pad_value_image, const float* pad_device_image,
                   float* y_soft, void* y_label, float* y_image, void* stream);

/* ---- slice extraction, backward (bilinear only) ----------------------------------------------
 * grad_out         [S,C,Do,Ho,Wo] fp32 contiguous; NULL => chain-only (only grad_grid_affine is
 *                  propagated to d_affine; used for nearest / integer volumes)
 * grad_grid_affine [S,4,4] fp32 or NULL: upstream gradient w.r.t. the returned grid_affine_out
 * d_vol            fp32, SAME element strides as the volume, pre-zeroed by the caller, or NULL
 *                  (training case: the volume never requires grad)
 * d_affine         gradient w.r.t. the view input of `views->kind`:
 *                    GRID   -> [S,3,4]   PRE -> [S,4,4]   PARAMS -> [S, 6+3R+1]        (fp32)
 * d_gpre           PARAMS only, [S,4,4] fp32 or NULL
 * d_pad            device float accumulator (+=) of d(out)/d(pad value) = sum go*(1-sum w_inbounds),
 *                  or NULL; the caller zeroes it.  Feeds afb_min_grad (MinBackward of the reference).
 * workspace        >= afb_slice_bwd_workspace_bytes(S) bytes, ZEROED by the caller before the first
 *                  use; the call leaves it zeroed again on completion.
 * Two launches: the sampler (re-gather, dVolume RED, CTA-reduced fp64 sums of dgrid (x) base per slice) and a
 * one-warp-per-slice chain kernel that turns the sums + upstream gradient into d_affine / d_gpre.       */
int64_t afb_slice_bwd_workspace_bytes(int S);
int afb_slice_bwd(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo,
                  int pad_mode, float pad_value, const float* pad_device,
                  const float* grad_out, const float* grad_grid_affine,
                  float* d_vol, float* d_affine, float* d_gpre, float* d_pad,
                  void* workspace, void* stream);

/* d_pad += sum go * (1 - sum of in-bounds weights): the pad-value gradient alone (reads grad_out and the
 * geometry only).  Lets the caller run MinBackward fused with the dVolume zero-fill BEFORE the scatter:
 *   afb_slice_pad_grad -> afb_min_grad_fill -> afb_slice_bwd(d_pad = NULL).                         */
int afb_slice_pad_grad(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo,
                       const float* grad_out, float* d_pad, void* stream);

/* Scatter-only half of the backward: d_vol (fp32, strides of `vol`, already filled) += w_k * grad_out at the 8 corners of
 * every output location.  With it the backward splits into  afb_slice_bwd(d_vol = NULL)  (re-gather, dTheta reduction and
 * chain: independent of the MinBackward fill, can run on another stream under it)  and  afb_slice_scatter  after the fill;
 * together they give exactly what afb_slice_bwd with d_vol does. */
int afb_slice_scatter(const afb_volume* vol, const afb_views* views, int Do, int Ho, int Wo, const float* grad_out,
                      float* d_vol, void* stream);

/* d_vol[i] = (vol[i] == min) ? d_pad / count : 0   for all i (initialises d_vol; replaces memset + afb_min_grad) */
int afb_min_grad_fill(const void* vol, int dtype, int64_t n_elements, const float* min_count,
                      const float* d_pad, float* d_vol, void* stream);

/* ---- one-hot label slicing straight from the integer label map (running/run_dl.py:261-264 + the two label
 * slicings of learnable_transform.py:287-298, without materialising the fp32 / int64 one-hot volumes) ----------
 * labels      [B,1,D,H,W] integer index map (AFB_U8 / I16 / I32 / I64), C field = 1, any strides
 * num_classes <= 16
 * y_soft      [S,num_classes,Do,Ho,Wo] fp32 = bilinear slice of one_hot(labels).float(), bitwise what afb_slice_fwd
 *             gives on the materialised volume (pad = min = 0); NULL to skip
 * y_label     label_out 1: [S,num_classes,Do,Ho,Wo] int64 = nearest slice of the int64 one-hot;
 *             label_out 2: [S,Do,Ho,Wo] uint8 nearest label index (0 out of field); label_out 0: none
 * Backward: gradient w.r.t. the view input only (an integer volume has none): the reference's training case.
 * workspace as for afb_slice_bwd.                                                                       */
int afb_slice_onehot_fwd(const afb_volume* labels, int num_classes, const afb_views* views, int Do, int Ho,
                         int Wo, float* y_soft, void* y_label, int label_out, void* stream);
int afb_slice_onehot_bwd(const afb_volume* labels, int num_classes, const afb_views* views, int Do, int Ho,
                         int Wo, const float* grad_y_soft, const float* grad_grid_affine, float* d_affine,
                         float* d_gpre, void* workspace, void* stream);

/* MinBackward of `volume.min()` (evenly distributed over all elements equal to the min):
 * d_vol[i] += (vol[i] == min) * d_pad / count.  vol dense (any permutation), d_vol same layout. */
int afb_min_grad(const void* vol, int dtype, int64_t n_elements, const float* min_count,
                 const float* d_pad, float* d_vol, void* stream);

/* ---- R6 -> rotation (utils/transform_utils.py:27-58) ------------------------------------------ */
int afb_r6_fwd(const float* ortho /*[N,6]*/, int N, float* mat /*[N,4,4]*/, void* stream);
int afb_r6_bwd(const float* ortho, const float* grad_mat /*[N,4,4]*/, int N, float* d_ortho /*[N,6]*/,
               void* stream);

/* ---- slice -> 3-D embedding (models/hybrid_unet.py:71-94) -------------------------------------
 * x        [B, V*c, S, S] fp32 contiguous (view-major channels, as torch.chunk(dim=1))
 * affines  [V, B, 4, 4] fp32: the slicing grid affines (b_grid_affines stacked)
 * out      [B, V*c, S, S, S] fp32
 * Backward: d_x [B,V*c,S,S] (fully overwritten: gather formulation, no atomics) and d_affines [V,B,4,4]
 * (either may be NULL).
 * workspace (both calls): >= afb_embed_workspace_bytes(B*V) bytes; the first B*V*136 bytes must be zero before the
 * first backward call and are left zeroed; the rest is scratch for the per-(b,v) inverse affines.   */
int64_t afb_embed_workspace_bytes(int n_slices /* B*V */);
int afb_embed_fwd(const float* x, const float* affines, int B, int V, int c, int S, float* out,
                  void* workspace, void* stream);
int afb_embed_bwd(const float* grad_out, const float* x, const float* affines, int B, int V, int c,
                  int S, float* d_x, float* d_affines, void* workspace, void* stream);


/* ---- callers either side of the samplers ------------------------------------------------------
 * afb_compose_pre_affine: Gpre[b] = base[b]^-1 @ view[b] (@ aug[b]) - running/run_dl.py:227-234 (+ the augmentation product of
 *   :208-223), computed in fp64 like the reference (base_affine carries the NIfTI affine's dtype) and rounded once to fp32.
 *   base [B,4,4] fp64, view [B,4,4] fp32 or fp64, aug [B,4,4] fp32 or NULL, out [B,4,4] fp32; *singular_flag (device int, may be
 *   NULL) is set to 1 if a base matrix is singular (its output is NaN).
 * afb_upsample2d_fwd/bwd: F.interpolate(x[N,C,h,w,1], size=[H,W,1], mode='trilinear', align_corners=False) of running/
 *   run_dl.py:193-197 on n_planes = N*C contiguous [h,w] planes; d_x is fully overwritten (gather form, deterministic).
 * afb_rot3_fwd/bwd: utils/transform_utils.py:62-178, params [N,3] -> homogeneous [N,4,4]. */
#define AFB_ROT_ANGLE_AXIS 0
#define AFB_ROT_NORMAL 1
int afb_compose_pre_affine(const double* base, const void* view, int view_is_f64, const float* aug, int B, float* out,
                           int* singular_flag, void* stream);
int afb_upsample2d_fwd(const float* x, int64_t n_planes, int h, int w, int H, int W, float* out, void* stream);
int afb_upsample2d_bwd(const float* grad_out, int64_t n_planes, int h, int w, int H, int W, float* d_x, void* stream);
int afb_rot3_fwd(int kind, const float* params, int N, float* mat, void* stream);
int afb_rot3_bwd(int kind, const float* params, const float* grad_mat, int N, float* d_params, void* stream);

/* The voxel passes of get_clinical_cardiac_view_affines (functional/clinical_cardiac_views.py:223-364; sparse CPU tensors in
 * the reference).  labels: contiguous [D,H,W] integer label map (values 1..31 take part).  A GROUP is a bit mask over label
 * values (bit l set <=> label l belongs to it).
 * afb_label_group_moments: for each of n_groups (<= 8) groups the exact integer sums {count, sum d, sum h, sum w, sum dd, dh,
 *   dw, hh, hw, ww} of the voxel indices (-> centre and inertia tensor, utils/torch_sparse_tensor_utils.py:34-56), one pass;
 *   out [n_groups][10] uint64, zeroed by the caller.
 * afb_label_extent_search: get_min_max_extent_along_axis (:51-62): bisection for the extent of the group along +dir and -dir
 *   from `center` (fp64 factors, fp32 distances like the reference); out[0], out[1] = the two factors. */
int afb_label_group_moments(const void* labels, int dtype, int D, int H, int W, const unsigned* group_masks_dev, int n_groups,
                            unsigned long long* out_dev, void* stream);
int afb_label_extent_search(const void* labels, int dtype, int D, int H, int W, unsigned group_mask, const float* center_dev,
                            const float* dir_dev, double init_end, double* out_dev, void* stream);

/* ---- the sharded path's collectives over NVLink peer memory -------------------------
 * One single-CTA kernel: push the local contribution (optionally summed over `pre_sum` rows of `in`) into this rank's slot
 * of EVERY rank's symmetric buffer, publish the epoch to every peer, wait (bounded, ~2 s, then *err != 0) until every peer has
 * published, reduce the slots of the own buffer in rank order.  op 0: out[n] = sum over ranks; op 1: out[world*n] = all-gather (rank-major);
 * op 2: n/2 (min, multiplicity) pairs -> the pairs of the whole batch (minimum over ranks, multiplicities of its holders summed).
 * bufs_dev: DEVICE array of `world` pointers, entry r = this process' mapping of rank r's buffer of
 * afb_peer_buffer_floats(n_channels, n_max, world) floats (symmetric / peer-mapped memory, zeroed before the first call; the host -
 * e.g. torch.distributed._symmetric_memory - allocates and exchanges the mappings).  epoch: device uint32[n_channels],
 * zeroed, private to the rank.  Independent exchanges use different channels.  Stream-ordered, CUDA-graph capturable. */
int64_t afb_peer_buffer_floats(int n_channels, int n_max, int world);
int afb_peer_collective(void* const* bufs_dev, int rank, int world, int op, int channel, int n_channels, int n, int n_max,
                        int pre_sum, const float* in, float* out, void* epoch, int* err, void* stream);

/* All stages of one U-Net pass in one launch each: HybridUnet.forward embeds every encoder skip with the same affines
 * (models/hybrid_unet.py:40-43: `[self.skip_connector(s, b_grid_affines) for s in skips]`).
 * x[i] [B, V*c[i], S[i], S[i]], out[i] / grad_out[i] [B, V*c[i], S[i]^3]; n_stages <= 8.  Backward: grad_out[i] == NULL
 * skips stage i; d_x (array, may be NULL) and its entries may be NULL; d_affines [V,B,4,4] receives the SUM over the stages
 * (may be NULL).  Same workspace contract as afb_embed_fwd / afb_embed_bwd.  Forward: a CTA streams zeros over a chunk of
 * rows of all channels and then patches the slab voxels of the same rows (the patch stores hit lines still dirty in L2). */
int afb_embed_multi_fwd(int n_stages, const float* const* x, const int* c, const int* S, float* const* out,
                        const float* affines, int B, int V, void* workspace, void* stream);
int afb_embed_multi_bwd(int n_stages, const float* const* grad_out, const float* const* x, const int* c, const int* S,
                        float* const* d_x, const float* affines, int B, int V, float* d_affines, void* workspace,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AFB200_H */
