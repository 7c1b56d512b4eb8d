/*
 * diffrender.h -- C ABI of libdiffrender.so, the sm_100a implementation of Differender's differentiable
 * ray-march hot path.
 *
 * Drop-in boundary.  The reference has no FFI: its "operator API" is the torch.autograd.Function
 * `RaycastFunction` (reference differender/volume_raycaster.py:392-476), which drives the Taichi object
 * `VolumeRaycaster` (:56-389).  Each entry point below names the reference calls it replaces.  The Python host
 * (differender_b200/volume_raycaster.py) binds these symbols with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller (PyTorch);
 *     the library never allocates, frees or retains memory and has no global mutable state
 *     (contrast: Taichi's process-global runtime :486 and forward state held in the object :429-430).
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*); no internal syncs.
 *   - return 0 on success, a negative DR_E* code otherwise; dr_last_error() gives a thread-local message.
 *   - volume axes use the reference's Taichi order (:481): X = torch W (unit stride), Y = torch D
 *     (slowest), Z = torch H.  A "linear" volume is the contiguous torch tensor [B][Y][Z][X].
 *   - images: DR_F_OUT_IMAGE set  -> [BS][4][H][W], row H-1-j holds raw row j (the flip/permute of :538-548
 *     fused into the kernel); clear -> the reference's raw `output_rgba.to_torch()` order [BS][W][H][4].
 *     grad_out uses the same layout as out_rgba.
 *   - jitter, out_K, out_Tprev: [BS][H][W] in image orientation: pixel (i,j) -> [H-1-j][i].
 */
#ifndef DIFFRENDER_H_
#define DIFFRENDER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DR_VERSION 104

/* error codes */
#define DR_OK 0
#define DR_EINVAL (-1)    /* bad shape / flag / null pointer            */
#define DR_EDTYPE (-2)    /* unsupported voxel dtype                    */
#define DR_EALIGN (-3)    /* pointer not aligned as required            */
#define DR_ECUDA (-4)     /* a CUDA call / launch failed                */
#define DR_EWORKSPACE (-5) /* workspace too small                        */

/* voxel storage types */
#define DR_VOX_F32 0
#define DR_VOX_F16 1
#define DR_VOX_U8 2       /* uint8-stored volume, voxel value = u8 / 255 in fp32 exactly as dr_ingest_u8 rounds it (the reference's skull.raw,
                             examples/taichi_volume_raycaster.py:548-550); marched from the cell-major copy only (DR_F_LAYOUT_CELL8: 8-byte records) */

/* DrDesc.flags */
#define DR_F_NONDIFF 1u        /* raycast_nondiff + get_final_image_nondiff (:308-361) instead of raycast (:261-306) */
#define DR_F_NEEDS_VOL_GRAD 2u /* ctx.needs_input_grad[volume]                                       */
#define DR_F_NEEDS_TF_GRAD 4u  /* ctx.needs_input_grad[tf]                                           */
#define DR_F_HAS_JITTER 8u     /* `jitter` flag of compute_entry_exit (:222, 254); jitter pointer must be non-null */
#define DR_F_OUT_IMAGE 16u     /* out_rgba / grad_out are [BS][4][H][W] flipped (else raw [BS][W][H][4])   */
#define DR_F_TF_4R 32u         /* tf / grad_tf are [Btf][4][R] (torch layout; else the reference's [Btf][R][4], :567,571) */
#define DR_F_GENERIC_TAPS 64u  /* force the 7x8-load tap path (always used when a normal tap can skip a whole cell) */
#define DR_F_LAYOUT_BRICK8 256u /* `vol` is the 8x8x8-bricked copy made by dr_brick_volume; clear: `vol` is the caller's linear [Bvol][Y][Z][X] tensor, read in place */
#define DR_F_LAYOUT_CELL8 2048u /* `vol` is the cell-major copy made by dr_expand_cells: [Bvol][X*Y*Z][8], record c = the 8 corners of the cell whose low corner has torch-linear index c (8x the volume's bytes; one address and two 16-byte loads per cell) */
#define DR_F_FUSED_MSE 1024u    /* internal: set by dr_backward_mse (grad_out slot holds the target image) */
#define DR_F_COUNT_SHADED 512u  /* diagnostic: out_K counts only samples with non-zero opacity (do not feed such a K to dr_backward) */
/* 128u: reserved (was DR_F_NO_REG_ACCUM, a tuning flag of early builds; ignored) */

/* Plain-data description of one call.  Fill it with dr_desc_init(); do not hand-edit derived fields. */
typedef struct DrDesc {
    int32_t X, Y, Z;        /* volume resolution, Taichi order                (:58-60, :481)         */
    int32_t W, H;           /* render resolution (w, h)                        (:74)                  */
    int32_t R;              /* transfer-function resolution                    (:113)                 */
    int32_t M;              /* max_samples                                     (:90)                  */
    int32_t BS;             /* number of views (camera positions) in this call                        */
    int32_t Bvol, Btf;      /* 1 (shared by all views, no clones - contrast :566-567) or BS           */
    int32_t vox_dtype;      /* DR_VOX_*                                                               */
    uint32_t flags;         /* DR_F_*                                                                 */
    /* derived (folded on the host in double, rounded to fp32 like Python scalars captured by the Taichi kernels) */
    float sr, inv_sr;       /* sampling_rate, 1/sampling_rate                  (:222, :285)           */
    float near_, near_w, near_h; /* :146-148                                                          */
    float scale[3];         /* (X,Y,Z) - 1 - 1e-4                              (:165)                 */
    float vol_diag;         /* ||(X,Y,Z) - 1||                                 (:248-249)             */
    float tf_len;           /* R - 1                                           (:215)                 */
    float ambient, diffuse, specular; /* 0.4, 0.8, 0.3                         (:91-93); shininess is fixed at 32 (:94) */
    float ert;              /* early-ray-termination threshold 0.99            (:267, :318)           */
    float delta;            /* normal tap offset 1e-3                          (:193)                 */
    float alpha_skip;       /* nondiff alpha skip 1e-3                         (:334)                 */
    int32_t nbx, nby, nbz;  /* bricks (8x8x8 voxels) per axis                                         */
    int32_t tap_generic;    /* 1 if delta spans >= 1 voxel on some axis (dims > ~2000): generic taps  */
} DrDesc;

/* Library version (DR_VERSION of the build). */
int dr_version(void);

/* Debug aid: out-of-range volume loads / gradient reductions counted by a -DDR_BOUNDS_CHECK build (synchronises the
 * device); always 0 in a normal build. */
long long dr_debug_oob_count(void);

/* Thread-local message of the last failure on this thread ("" if none). */
const char* dr_last_error(void);

/*
 * Fills *d.  Replaces VolumeRaycaster.__init__ (:58-116): resolution bookkeeping, constants, layout choice.
 * fov_deg/near as in Raycaster.__init__ (:479); sampling_rate as in RaycastFunction.forward (:395).
 */
int dr_desc_init(DrDesc* d, int32_t X, int32_t Y, int32_t Z, int32_t W, int32_t H, int32_t R, int32_t M,
                 int32_t BS, int32_t Bvol, int32_t Btf, int32_t vox_dtype, uint32_t flags,
                 double sampling_rate, double fov_deg, double near_plane);

/* Elements (not bytes) of ONE bricked volume: nbx*nby*nbz*512. */
size_t dr_bricked_elems(const DrDesc* d);

/*
 * Re-lays Bvol linear volumes [Bvol][Y][Z][X] (fp32 or fp16 per vox_dtype) into 8x8x8 bricks
 * [Bvol][nbz][nby][nbx][8][8][8], x fastest, same dtype.  Replaces set_volume / field.from_torch (:118-119),
 * which copied into Taichi's 4x4x4-bricked SNode (:97-101).  Needed once per DISTINCT volume, not per view, and only
 * for calls that set DR_F_LAYOUT_BRICK8 (the default layout reads the linear tensor in place).
 */
int dr_brick_volume(const DrDesc* d, const void* vol_linear, void* vol_bricked, void* stream);

/*
 * Cell-major copy of the volume for DR_F_LAYOUT_CELL8 (same role as dr_brick_volume; replaces set_volume's from_torch
 * into the reference's 4x4x4-blocked field, :97-101, :118-119): vol_cells is [Bvol][X*Y*Z][8] of the volume's dtype
 * (dr_grad_cells_elems(d) elements per volume, 32-byte aligned); record c holds the 8 corners of the cell whose low corner
 * has torch-linear index c, slot c + 2a + 4b = voxel (min(x+a,X-1), min(y+b,Y-1), min(z+c,Z-1)) -- z pairs adjacent, so the two
 * 16-byte loads of a record land in the register pairs the packed (f32x2) trilinear mixes work on.
 */
int dr_expand_cells(const DrDesc* d, const void* vol_linear, void* vol_cells, void* stream);

/*
 * Forward march of BS views.  Replaces, per view: set_cam_pos/set_tf_tex (:121-125), clear_framebuffer
 * (:374-382), compute_entry_exit (:221-259), raycast (:261-306) or raycast_nondiff (:308-351), and
 * get_final_image (:363-372) or get_final_image_nondiff (:353-361), i.e. the body of
 * RaycastFunction.forward (:418-438) and Raycaster.raycast_nondiff (:502-523).
 *   vol [Bvol] volumes: linear [Y][Z][X] (default, zero copy), bricked (DR_F_LAYOUT_BRICK8) or cell-major (DR_F_LAYOUT_CELL8), fp32/fp16
 *   tf [Btf][R][4] or [Btf][4][R]     cam [BS][3]
 *   jitter [BS][H][W] uniform [0,1) or NULL (replaces ti.random, :255)
 *   out_rgba  see layout note            out_K [BS][H][W] active samples per ray (valid_sample_step_count-1,
 *   :303,367) or NULL      out_Tprev [BS][H][W] transmittance before the last active sample, or NULL
 *   (out_K and out_Tprev are what the backward needs instead of the O(W*H*M) render_tape, :82,102-103).
 */
int dr_forward(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
               float* out_rgba, int32_t* out_K, float* out_Tprev, void* stream);

/* Bytes of scratch dr_backward needs for this descriptor (privatised TF-gradient copies). */
size_t dr_workspace_bytes(const DrDesc* d);

/* Floats (not bytes) of ONE cell-major volume-gradient buffer: X*Y*Z*8 (one 32-byte cell per voxel position). */
size_t dr_grad_cells_elems(const DrDesc* d);

/*
 * Backward of BS views.  Replaces, per view: clear_grad (:384-389), output_rgba.grad.from_torch (:459,469),
 * get_final_image.grad() (:460,470) and raycast.grad() (:461,471), i.e. RaycastFunction.backward (:440-476).
 * Each ray is re-marched in reverse from (out_rgba, K, Tprev); no per-sample tape exists.
 *   grad_vol_cells [Bvol][Y*Z*X][8] fp32 cell-major gradient (slot a+2b+4c of cell (x,y,z) belongs to voxel
 *     (x+a, y+b, z+c)), ACCUMULATED into (caller zeroes), 32-byte aligned; may be NULL without NEEDS_VOL_GRAD.
 *     It is shared by all views of the call, so one shared volume gets ONE summed gradient (contrast :447, :463).
 *   grad_tf [Btf] in the tf layout, fp32, ACCUMULATED into (caller zeroes); may be NULL without NEEDS_TF_GRAD
 *   workspace: dr_workspace_bytes(d) bytes, 16-byte aligned, contents undefined on entry and exit.
 */
int dr_backward(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
                const float* grad_out, const float* out_rgba, const int32_t* K, const float* Tprev,
                float* grad_vol_cells, float* grad_tf, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Cell-major fp32 gradient -> linear [Bvol][Y][Z][X] fp32 (sums the up-to-8 cell slots that alias each voxel) with
 * nan_to_num applied (NaN -> 0, +-inf -> +-FLT_MAX), as volume.grad.to_torch + torch.nan_to_num (:463, :474).
 * accumulate != 0 adds into grad_linear.
 */
int dr_gather_grad(const DrDesc* d, const float* grad_vol_cells, float* grad_linear, int accumulate, void* stream);

/* ---- caller-side steps either side of the march ------------------------------------------------ */

/*
 * dr_forward with the mean-squared-error loss of the reference's optimisation loops fused into the epilogue
 * (torch.nn.functional.mse_loss(output_rgba, reference), examples/taichi_volume_raycaster.py:425-447;
 * F.mse_loss(pred, targ), examples/test_opt_tf.py:70-72): besides everything dr_forward writes, loss_sum[b] +=
 * sum over the view's pixels and 4 channels of (out - target)^2.  target has the layout of out_rgba; loss_sum is
 * [BS] fp32, caller-zeroed.  The mean is the caller's division.
 */
int dr_forward_mse(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
                   const float* target, float* out_rgba, int32_t* out_K, float* out_Tprev, float* loss_sum, void* stream);

/*
 * Exact empty-space skipping for the forward march (no counterpart in the reference, which marches every sample of every ray,
 * :264-306).  dr_build_skip_grid classifies every MACRO-CELL (8x8x8 cells; nbx*nby*nbz of them per volume) as "exactly
 * transparent under this call's transfer function": every TF bin that a trilinear value of the macro-cell's voxels can select
 * has alpha 0.  A sample there has opacity exactly 0 and leaves the accumulated colour bit-identical, so dr_forward_ex replaces
 * runs of such samples by their count: images, out_K and out_Tprev are bit-identical with and without the grid.
 *   minmax     [Bvol][nbz*nby*nbx] float2 workspace (dr_skip_minmax_bytes): per-macro-cell voxel min / max.  It depends on the
 *              volume only: pass minmax_valid = 1 to reuse it while only the transfer function changes (vol_linear may be NULL).
 *   skip_grid  dr_skip_grid_bytes(d) bytes, 4-byte aligned: a 16-byte header (uint32 number of empty macro-cells; 0 makes
 *              dr_forward_ex march as if no grid were given) followed by [1 or BS][nbz*nby*nbx] bytes, one grid per view unless
 *              volume and TF are shared.
 * dr_forward_ex is dr_forward (target == NULL) or dr_forward_mse (target, loss_sum != NULL) with an optional skip_grid
 * (NULL = march every sample).  The grid must have been built from the same volume, transfer function and descriptor.
 */
size_t dr_skip_minmax_bytes(const DrDesc* d);
size_t dr_skip_grid_bytes(const DrDesc* d);
int dr_build_skip_grid(const DrDesc* d, const void* vol_linear, const float* tf, void* minmax, int minmax_valid, uint8_t* skip_grid,
                       void* stream);
int dr_forward_ex(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter, const float* target,
                  const uint8_t* skip_grid, float* out_rgba, int32_t* out_K, float* out_Tprev, float* loss_sum, void* stream);

/*
 * dr_backward for that loss: dL/d(out) = scale * (out - target) is formed inside the kernel from out_rgba and target,
 * so the gradient image never exists in HBM (replaces output_rgba.grad.from_torch, :436).  For the mean over
 * BS*4*H*W elements pass scale = 2 / (BS*4*H*W) times the upstream gradient of the loss.
 */
int dr_backward_mse(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
                    const float* target, float scale, const float* out_rgba, const int32_t* K, const float* Tprev,
                    float* grad_vol_cells, float* grad_tf, void* workspace, size_t workspace_bytes, void* stream);

/*
 * dr_backward (grad_out != NULL, target == NULL) or dr_backward_mse (target != NULL, grad_out == NULL, `scale` as there; an optional
 * DEVICE scalar `scale_dev` is multiplied into it inside the kernel -- the upstream gradient of the loss, which the host then never
 * has to read back) with the forward's skip grid (dr_build_skip_grid; NULL = march every sample).  The grid is used by the VOLUME-ONLY backward (flags without
 * DR_F_NEEDS_TF_GRAD): a sample in a macro-cell that is exactly transparent under the transfer function has zero opacity and two
 * transparent TF bins, so it contributes nothing to the volume gradient and leaves the transmittance and dL/dA unchanged; runs of
 * such samples are jumped over.  With a TF gradient every transparent sample contributes d(alpha) and the grid is ignored.
 * This is the reference's own optimisation case (examples/test_opt_tf.py optimises the VOLUME, :49-55, :86-88).
 */
int dr_backward_ex(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter, const float* grad_out,
                   const float* target, float scale, const float* scale_dev, const uint8_t* skip_grid, const float* out_rgba, const int32_t* K,
                   const float* Tprev, float* grad_vol_cells, float* grad_tf, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Momentum-SGD step with gradient clipping and projection, the reference's `apply_grad`
 * (examples/taichi_volume_raycaster.py:375-381: tf_momentum = gamma*tf_momentum + lr*clamp(grad, +-max_grad);
 * tf -= tf_momentum; tf = max(tf, 0)) generalised with an upper clamp so that it also covers the volume projection
 * vol.clamp_(0, 1) of examples/test_opt_tf.py:86-88:
 *   m = gamma*m + lr*clamp(g, -max_grad, max_grad);  p = clamp(p - m, lo, hi)      over n fp32 elements, in place.
 */
int dr_momentum_step(float* param, const float* grad, float* momentum, size_t n, float lr, float gamma, float max_grad,
                     float lo, float hi, void* stream);

/*
 * The whole parameter update of a volume-optimisation step in ONE kernel: gather of the cell-major
 * gradient with nan_to_num (what dr_gather_grad does; volume.grad.to_torch + torch.nan_to_num, :463, :474), the momentum-SGD
 * step with clipping and projection of dr_momentum_step (examples/taichi_volume_raycaster.py:375-381; vol.clamp_(0, 1) of
 * examples/test_opt_tf.py:86-88), and the refresh of the cell-major VOLUME copy that the next forward reads (so the next
 * step needs no dr_expand_cells; replaces set_volume's from_torch, :118-119).
 *   grad_vol_cells  [X*Y*Z][8] fp32 cell-major gradient as scattered by dr_backward          } exactly one of the two;
 *   grad_linear     [Y][Z][X] fp32, already gathered (e.g. all-reduced across GPUs)          } the other is NULL
 *   param, momentum [Y][Z][X] fp32, updated in place
 *   vol_cells       cell-major volume copy [X*Y*Z][8] of d->vox_dtype (fp32 / fp16) to refresh from the new param, or NULL
 *   grad_out        optional [Y][Z][X] fp32: receives the gathered gradient (for logging), or NULL
 * Results are bit-identical to dr_gather_grad followed by dr_momentum_step followed by dr_expand_cells.  d->Bvol must be 1.
 */
int dr_gather_step(const DrDesc* d, const float* grad_vol_cells, const float* grad_linear, float* param, float* momentum,
                   void* vol_cells, float* grad_out, float lr, float gamma, float max_grad, float lo, float hi, void* stream);

/*
 * Diagnostic for the benchmark's L2 roofline (the L2 -> SM bandwidth is not in MEASURED_PEAKS.json and has to
 * be measured on the box): 2 CTAs per SM each read the whole buffer `reps` times with 16-byte ld.global.cg loads.  With a
 * buffer that stays L2-resident (32-64 MiB) the bytes moved / elapsed time is the L2 read bandwidth.  Returns the number of
 * bytes the launch reads (>= 0) or a negative DR_E* code; `sink` is 4 bytes of scratch.  Time it with events on `stream`.
 */
long long dr_probe_l2_read(const void* buf, size_t bytes, int reps, void* sink, void* stream);

/*
 * Raw uint8 volume -> linear voxel volume [Y][Z][X] of d->vox_dtype with value = u8 / 255, as the reference ingests
 * skull.raw (np.fromfile(uint8).reshape(256,256,256), swapaxes(0,1), / 255.0: examples/taichi_volume_raycaster.py:548-550).
 * swap_axes01 != 0: src is [Z][Y][X] and the two slow axes are swapped on the fly; else src is already [Y][Z][X].
 */
int dr_ingest_u8(const DrDesc* d, const uint8_t* src, void* vol_linear, int swap_axes01, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFRENDER_H_ */
