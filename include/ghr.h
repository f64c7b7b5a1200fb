/*
 * ghr.h -- C ABI of libghr.so, the B200-native (sm_100a) Gaussian-splatting rasterizer.
 *
 * Drop-in boundary.  GuassianHand calls the rasterizer only through the Python package
 * `diff_gaussian_rasterization` (/root/reference/tgs/models/renderer_one_shot.py:3, :281-296,
 * :338-346, :355-379; same lines in renderer_one_shot_edit.py).  That package's native surface is
 * a pybind module with three entry points (upstream rasterize_points.cu):
 *     rasterize_gaussians(...)            -> ghr_forward
 *     rasterize_gaussians_backward(...)   -> ghr_backward
 *     mark_visible(...)                   -> ghr_mark_visible
 * The functions below are what an FFI binding of that path binds instead.  Conventions:
 *   - extern "C", plain pointers and sizes, no C++/torch types, no exceptions across the ABI.
 *   - every pointer in the argument structs is a DEVICE pointer unless its comment says "host";
 *     NULL means "absent" for the optional inputs.  All arrays are contiguous fp32 / int32.
 *   - the caller owns all memory.  The library never allocates device memory and never
 *     synchronises unless GHR_FLAG_DEBUG is set; all work is enqueued on the given stream
 *     (a cudaStream_t passed as void*).
 *   - return value: 0 on success, negative GHR_E* on failure; ghr_last_error() gives the text
 *     (thread-local).  Capacity overflow of the binning buffers is reported asynchronously in
 *     GhrStatus (device) because detecting it on the host would need the device sync this
 *     design removes; see ghr_read_status().
 *   - V >= 1 cameras ("views") may be rendered by one call: Gaussians are shared, every
 *     per-view array is laid out view-major.  V == 1 is exactly the upstream call.
 */
#ifndef GHR_H_
#define GHR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GHR_OK 0
#define GHR_EINVAL (-1)     /* bad argument (NULL required pointer, negative size, ...) */
#define GHR_ENOSPC (-2)     /* state/temp buffer smaller than ghr_layout() demands */
#define GHR_ECUDA (-3)      /* a CUDA call failed; text in ghr_last_error() */
#define GHR_EOVERFLOW (-4)  /* ghr_read_status: R exceeded R_cap in the last forward */

#define GHR_STATUS_OVERFLOW 1u
#define GHR_STATUS_PREFILTER 2u
#define GHR_FLAG_PREFILTERED 1u /* settings.prefiltered (renderer_one_shot.py:292) */
#define GHR_FLAG_DEBUG 2u       /* settings.debug (:293): sync + check after every launch */

#define GHR_ABI_VERSION 11
#define GHR_SEGMENT 128 /* instances per backward work unit / forward checkpoint interval */

/* stage ids for the optional stage_events arrays */
#define GHR_NSTAGES_FWD 5 /* 0 preprocess (+ per-tile counts), 1 tile scan + schedule, 2 duplicate, 3 per-tile sort + gather, 4 blend */
#define GHR_NSTAGES_BWD 2 /* 0 blend backward, 1 preprocess backward */

/* Problem dimensions.  T = ceil(W/16)*ceil(H/16) tiles per view, N = H*W pixels per view. */
typedef struct GhrDims {
  int32_t P;         /* Gaussians */
  int32_t V;         /* views (cameras) in this call */
  int32_t H, W;      /* image size, identical for all views of a call */
  int32_t M;         /* SH coefficients per Gaussian (shs.shape[1]); 0 with colors_precomp */
  int32_t sh_degree; /* active SH degree 0..3 (settings.sh_degree) */
  int64_t R_cap;     /* capacity (entries) of the (tile, Gaussian) instance buffers, all views */
} GhrDims;

/* Byte offsets inside the caller-provided `state` buffer (kept from forward to backward) and
 * sizes of both buffers.  Filled by ghr_layout(); also lets tests read intermediates. */
typedef struct GhrLayout {
  size_t state_bytes, temp_bytes, temp_bwd_bytes;
  size_t off_status;   /* GhrStatus */
  size_t off_geom;     /* float4[4] per (view, Gaussian): {x,y,conic.x,conic.y} {conic.z,opacity,thr,0}
                          {r,g,b,depth} {radius(int bits), tiles_touched(uint bits),0,0}; thr = 2 ln(255
                          opacity) + margin: bound on d^T Q d for alpha >= 1/255 (culling only) */
  size_t off_clamped;  /* uint8 per (view, Gaussian): bit c set if SH colour channel c was clamped */
  size_t off_ranges;   /* uint32[2] per (view, tile): [start,end) into the sorted instances */
  size_t off_tilemax;  /* uint32[2] per (view, tile): {max n_contrib in the tile, forward CTAs of the tile that have finished} */
  size_t off_records;  /* 48-byte instance records in sorted order:
                          {x,y,conic.x,conic.z} {conic.y,opacity,thr,0} {r,g,b,id(uint bits)} */
  size_t off_final_T;  /* float per (view, pixel) */
  size_t off_ncontrib; /* uint32 per (view, pixel) */
  size_t off_order;    /* uint32 per (view, tile): blend launch order, longest instance list first */
  size_t off_masks;    /* uint8 per sorted instance: bit w set if the instance can reach alpha >= 1/255
                          inside the 8x4-pixel sub-block w = 2*(row/4) + (col/8) of its tile (culling only) */
  size_t off_tilefinal; /* float4 per (view, tile, pixel-in-tile): {C.r, C.g, C.b, T_final} of the forward
                          blend before the background term (the backward's suffix sums start from it) */
  size_t off_ckpt;     /* float4 {T, C.r, C.g, C.b} per (slot, pixel-in-tile): the forward's running state at
                          every GHR_SEGMENT-instance boundary of a tile list; slot(tile, s) =
                          ranges[tile].start / GHR_SEGMENT + tile + s, s >= 1 */
  size_t off_units;    /* uint32[4] (view*T + tile, segment, slab start, instances up to the tile's last contributor)
                          per backward work unit, GhrStatus.reserved[1] of them */
} GhrLayout;

typedef struct GhrStatus {
  uint64_t R;          /* number of (tile, Gaussian) instances the forward produced (all views) */
  uint32_t overflow;   /* bit 0 (GHR_STATUS_OVERFLOW): R > R_cap, the forward output is invalid, retry with a larger
                          R_cap; bit 1 (GHR_STATUS_PREFILTER): GHR_FLAG_PREFILTERED was set and a Gaussian failed the
                          near-plane test (upstream's in_frustum traps there); the point was culled */
  uint32_t n_visible;  /* Gaussians (summed over views) with radius > 0 */
  uint64_t reserved[2]; /* [0] = GhrForwardArgs.seq of the forward that wrote this status;
                           [1] = number of backward work units (tile, segment) the forward blend emitted */
} GhrStatus;

typedef struct GhrForwardArgs {
  GhrDims dims;
  uint32_t flags;
  float scale_modifier;
  float tanfovx, tanfovy;      /* host scalars; used when `tanfov` is NULL */
  /* cameras */
  const float *viewmatrix;     /* [V,16]  settings.viewmatrix  (= w2c^T, renderer_one_shot.py:96) */
  const float *projmatrix;     /* [V,16]  settings.projmatrix  (:104-106) */
  const float *campos;         /* [V,3]   settings.campos      (:107) */
  const float *tanfov;         /* [V,2] (tanfovx, tanfovy) per view, or NULL */
  const float *bg;             /* [3] (bg_stride 0) or [V,3] (bg_stride 3) */
  int32_t bg_stride;
  /* Gaussians, shared by all views */
  const float *means3D;        /* [P,3] */
  const float *opacities;      /* [P]   */
  const float *scales;         /* [P,3] or NULL */
  const float *rotations;      /* [P,4] or NULL */
  const float *cov3D_precomp;  /* [P,6] or NULL */
  const float *shs;            /* [P,M,3] or NULL */
  const float *colors_precomp; /* [P,3] or NULL */
  /* outputs */
  float *out_color;            /* [V,3,H,W] */
  int32_t *radii;              /* [V,P] */
  float *out_mask;             /* [V,H,W] or NULL: coverage 1 - T_final of the same pass.  Equals the
                                  reference's second "mask" render (colors = 1, bg = 0,
                                  renderer_one_shot.py:353-380) without rendering twice */
  /* workspaces */
  void *state; size_t state_bytes;
  void *temp;  size_t temp_bytes;
  /* optional parity/debug exports (device, may be NULL): the upstream intermediates */
  uint64_t *dbg_keys_sorted;   /* [R_cap] (tile<<32 | depth bits), tile local to its view */
  uint32_t *dbg_point_list;    /* [R_cap] Gaussian index within its view */
  /* optional early status report: if non-NULL (PINNED HOST memory), GhrStatus is copied there
   * as soon as R is known (after the tile scan, before duplication, the per-tile sort and the blend are enqueued), with
   * reserved[0] == seq.  The host can poll it while the GPU keeps working -- this replaces the
   * blocking D2H read of num_rendered in upstream's forward without stalling the pipeline. */
  GhrStatus *host_status;
  uint64_t seq;
  /* optional per-stage timing: HOST array of 2*GHR_NSTAGES_FWD cudaEvent_t (start,stop per stage,
   * created with ghr_event_create), recorded on the stream around each stage; NULL = off. */
  void **stage_events;
  /* optional geometry reuse (V == 1): the state of an earlier ghr_forward with IDENTICAL geometry inputs
   * (means3D, opacities, scales/rotations or cov3D_precomp, camera, image size, scale_modifier, R_cap) whose
   * colours / background / SH inputs may differ -- the reference renders every view twice this way (RGB, then
   * an all-ones "mask" render, /root/reference/tgs/models/renderer_one_shot.py:338-346, :372-379).  The
   * projected geometry, tile ranges, depth order and cull masks are copied, only the colours are recomputed
   * and the blend runs: preprocess, scan, duplication and sort are skipped.  reuse_M = the M of that call. */
  const void *reuse_state;
  int32_t reuse_M;
} GhrForwardArgs;

typedef struct GhrBackwardArgs {
  GhrDims dims;
  uint32_t flags;
  float scale_modifier;
  float tanfovx, tanfovy;
  const float *viewmatrix, *projmatrix, *campos, *tanfov, *bg;
  int32_t bg_stride;
  const float *means3D, *opacities, *scales, *rotations, *cov3D_precomp, *shs, *colors_precomp;
  const float *dL_dout_color;  /* [V,3,H,W] */
  const float *dL_dout_mask;   /* [V,H,W] or NULL: gradient w.r.t. out_mask, folded into the same pass */
  const void *state; size_t state_bytes;   /* as written by ghr_forward */
  void *temp; size_t temp_bytes;           /* >= layout.temp_bwd_bytes */
  /* Gradient outputs.  Summed over the V views.  accumulate != 0: "+=" into the buffers
   * (multi-call accumulation before an all-reduce), else overwritten.  NULL = not wanted. */
  int32_t accumulate;
  float *dL_dmeans3D;   /* [P,3] */
  float *dL_dmeans2D;   /* [V,P,3] per view, NDC units (x 0.5W, 0.5H), z = 0; never accumulated */
  float *dL_dcolors;    /* [P,3]   (colors_precomp path) */
  float *dL_dopacity;   /* [P]     */
  float *dL_dcov3D;     /* [P,6]   (always available; the grad when cov3D_precomp is used) */
  float *dL_dsh;        /* [P,M,3] (SH path) */
  float *dL_dscales;    /* [P,3] */
  float *dL_drotations; /* [P,4] */
  float *dL_dconic;     /* [V,P,4] optional debug export (xx, xy, 0, yy) */
  void **stage_events;  /* HOST array of 2*GHR_NSTAGES_BWD cudaEvent_t or NULL */
} GhrBackwardArgs;

int ghr_abi_version(void);
const char *ghr_last_error(void);
/* sizeof() of an ABI struct by name ("GhrForwardArgs", ...), 0 if unknown: lets a foreign-language
 * binding verify its mirror of the structs at load time. */
size_t ghr_struct_size(const char *name);

/* Sizes/offsets of the caller-owned buffers for the given dimensions. */
int ghr_layout(const GhrDims *dims, GhrLayout *out);

/* Replaces upstream rasterize_gaussians(): preprocess -> bin -> sort -> blend, V views. */
int ghr_forward(const GhrForwardArgs *args, void *cuda_stream);

/* Replaces upstream rasterize_gaussians_backward(). */
int ghr_backward(const GhrBackwardArgs *args, void *cuda_stream);

/* Replaces upstream mark_visible(): present[i] = (view-space z > 0.2). */
int ghr_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                     uint8_t *present, void *cuda_stream);

/* Cameras of V views on the device, from world-to-camera matrices and intrinsics.
 * Replaces the per-view host work of /root/reference/tgs/models/renderer_one_shot.py:61-112 (Camera.from_w2c,
 * getProjectionMatrix_refine, intrinsic_to_fov) and the two math.tan(cuda scalar) host syncs of :278-279:
 * one launch, no synchronisation.  w2c [V,4,4] and K [V,3,3] row-major device fp32; outputs are exactly the
 * per-view arrays GhrForwardArgs takes: viewmatrix [V,16] (= w2c^T), projmatrix [V,16] (= viewmatrix @ P^T),
 * campos [V,3] (= inverse(viewmatrix)[3,:3]), tanfov [V,2].  The reference forces znear = 0.01, zfar = 1000. */
int ghr_cameras_from_w2c(int32_t V, const float *w2c, const float *K, int32_t H, int32_t W, float znear, float zfar,
                         float *viewmatrix, float *projmatrix, float *campos, float *tanfov, void *cuda_stream);

/* Thin event helpers so a host without a CUDA binding can time stages (cudaEvent_t as void*). */
int ghr_event_create(void **event_out);
int ghr_event_destroy(void *event);
int ghr_event_record(void *event, void *cuda_stream);
int ghr_event_elapsed_ms(void *start, void *stop, float *ms_out); /* both must have completed */

/* FP32 roofline denominator: launches a dependent-FMA kernel (8 chains/thread, 128 FMA instructions per
 * loop trip, all three operands in registers) on `cuda_stream`; packed != 0 issues fp32 pairs (FFMA2, the
 * form the blend kernels use).  `in` = 2 device floats (multiplier, addend), `sink` = 4 device bytes;
 * *flops_out = floating-point operations the launch executes.  Time it with events. */
int ghr_fp32_probe(int32_t iters, int32_t packed, const float *in, float *sink, double *flops_out, void *cuda_stream);

/* ---- fused attribute head ----
 * Activations of GSLayer.forward (/root/reference/tgs/models/renderer_one_shot.py:191-214: normalize,
 * trunc_exp (+clamp), sigmoid, xyz offset) and the interaction-aware attribute blending of
 * forward_single_view (:298-334: + xyz_b, + opacity_b, colour*w0 + w1 - 1 + b0), use_rgb path, in one
 * launch; outputs are the tensors ghr_forward takes.  All pointers device fp32, contiguous. */
#define GHR_ATTR_XYZ_OFFSET 1u      /* GSLayer.Config.xyz_offset (:161) */
#define GHR_ATTR_RESTRICT_OFFSET 2u /* GSLayer.Config.restrict_offset (:162) */
#define GHR_ATTR_CLIP_SCALING 4u    /* clip_scaling is set (:164) */

typedef struct GhrAttributeArgs {
  int32_t P;
  uint32_t flags;
  float clip_scaling;
  /* head outputs (after the Linear layers) and the points they offset */
  const float *xyz_raw;      /* [P,3] */
  const float *pts;          /* [P,3] */
  const float *scaling_raw;  /* [P,3] */
  const float *rotation_raw; /* [P,4] */
  const float *opacity_raw;  /* [P]   */
  const float *rgb_raw;      /* [P,3] */
  /* blending terms, each may be NULL */
  const float *xyz_b;        /* [P,3] */
  const float *opacity_b;    /* [P]   */
  const float *color_w0;     /* [P,3] = color_w.view(-1,16,3)[:,0,:] */
  const float *color_w1;     /* [P,3] = color_w.view(-1,16,3)[:,1,:] */
  const float *color_b0;     /* [P,3] = color_b.view(-1,16,3)[:,0,:] */
  /* outputs */
  float *means3D;            /* [P,3] */
  float *scales;             /* [P,3] */
  float *rotations;          /* [P,4] */
  float *opacities;          /* [P]   */
  float *colors;             /* [P,3] */
} GhrAttributeArgs;

typedef struct GhrAttributeGrads {
  /* incoming: gradients w.r.t. the five outputs (NULL = zero) */
  const float *dL_dmeans3D, *dL_dscales, *dL_drotations, *dL_dopacity, *dL_dcolors;
  /* outgoing (NULL = not wanted) */
  float *d_xyz_raw, *d_pts, *d_scaling_raw, *d_rotation_raw, *d_opacity_raw, *d_rgb_raw;
  float *d_xyz_b, *d_opacity_b, *d_color_w0, *d_color_w1, *d_color_b0;
} GhrAttributeGrads;

int ghr_attributes_forward(const GhrAttributeArgs *args, void *cuda_stream);
int ghr_attributes_backward(const GhrAttributeArgs *args, const GhrAttributeGrads *grads, void *cuda_stream);

/* SH path of the attribute blending (use_rgb = false; /root/reference/tgs/models/renderer_one_shot.py:329-334), n =
 * P * 16 * 3 coefficients, elementwise:   w == NULL: out = x;   w only: out = x * w;   w and b: out = (x * w) * w + b
 * (the reference multiplies by color_w a second time when color_b is given, :333-334 -- kept).  b without w is
 * an error, as in the reference (it dereferences color_w).  The geometry attributes of that path go through
 * ghr_attributes_forward with rgb_raw == colors == NULL. */
int ghr_sh_blend_forward(int64_t n, const float *x, const float *w, const float *b, float *out, void *cuda_stream);
/* d_x / d_w / d_b: NULL = not wanted */
int ghr_sh_blend_backward(int64_t n, const float *x, const float *w, const float *b, const float *dL_dout, float *d_x,
                          float *d_w, float *d_b, void *cuda_stream);

/* ---- gradient all-reduce over NVLink peer memory (SURVEY.md §8(e) stage 2) ----
 * Replaces the NCCL all-reduce Lightning DDP performs for the reference (/root/reference/infer_one_shot.py:631,638)
 * for ranks of ONE node (one process per GPU): a two-shot, in-place, deterministic sum of a float buffer that
 * lives in an allocation of this library (the one exception to "the library never allocates": the buffer must
 * be exportable through CUDA IPC).  Set-up, once: every rank creates its communicator, the host side
 * all-gathers the GHR_COMM_HANDLE_BYTES-byte handles (any transport: torch.distributed, MPI, a file) and
 * connects.  Per step: the backward writes its packed gradients into ghr_comm_buffer() and
 * ghr_comm_allreduce() enqueues ONE kernel on the stream (CUDA-graph capturable, no host argument changes
 * between calls); every rank must enqueue the same sequence of all-reduces.  A peer that never arrives makes
 * the kernel give up after ~2 s and sets the error word (ghr_comm_status) instead of hanging the GPU. */
#define GHR_COMM_MAX_RANKS 8
#define GHR_COMM_HANDLE_BYTES 128
typedef struct GhrComm GhrComm;
int ghr_comm_create(int32_t rank, int32_t world, size_t bytes, GhrComm **out);   /* current device */
int ghr_comm_handle(GhrComm *comm, void *handle_out /* GHR_COMM_HANDLE_BYTES, host */);
int ghr_comm_connect(GhrComm *comm, const void *all_handles /* world handles in rank order, host */);
void *ghr_comm_buffer(GhrComm *comm);                                           /* device pointer, `bytes` long */
int ghr_comm_allreduce(GhrComm *comm, size_t nfloats /* multiple of 4 */, void *cuda_stream);
int ghr_comm_status(GhrComm *comm, uint32_t *epochs_done, uint32_t *error /* 0 ok, 1/2 = timed out in barrier A/B */);
int ghr_comm_destroy(GhrComm *comm);

/* Enqueue an async copy of GhrStatus from `state` into pinned host memory `host_status`. */
int ghr_read_status_async(const void *state, GhrStatus *host_status, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* GHR_H_ */
