// C ABI of libghr.so (include/ghr.h): argument checks, buffer layout, kernel sequencing.
// No allocation, no synchronisation (unless GHR_FLAG_DEBUG), everything on the caller's stream.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ghr_internal.cuh"

namespace ghr {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

int compute_layout(const GhrDims &d, Layout *L) {
  if (d.P < 0 || d.V < 1 || d.H < 1 || d.W < 1 || d.M < 0 || d.R_cap < 0 || d.sh_degree < 0 || d.sh_degree > 3) {
    set_error("ghr_layout: bad dims P=%d V=%d H=%d W=%d M=%d deg=%d R_cap=%lld", d.P, d.V, d.H, d.W, d.M,
              d.sh_degree, (long long)d.R_cap);
    return GHR_EINVAL;
  }
  memset(L, 0, sizeof(*L));
  L->gx = (d.W + kTile - 1) / kTile;
  L->gy = (d.H + kTile - 1) / kTile;
  L->T = L->gx * L->gy;
  const uint64_t VP = (uint64_t)d.V * d.P, VT = (uint64_t)d.V * L->T, VN = (uint64_t)d.V * d.H * d.W;
  if (VP >= (1ull << 30) || VT >= (1ull << 31) || (uint64_t)d.R_cap >= (1ull << 30)) {
    set_error("ghr_layout: V*P, V*T and R_cap must stay below 2^30 (got %llu, %llu, %lld)",
              (unsigned long long)VP, (unsigned long long)VT, (long long)d.R_cap);
    return GHR_EINVAL;
  }
  // ---- state ----
  size_t o = 0;
  L->pub.off_status = o;   o = align_up(o + sizeof(GhrStatus));
  L->pub.off_geom = o;     o = align_up(o + VP * 64);
  L->pub.off_clamped = o;  o = align_up(o + (d.M > 0 ? VP : 0));
  L->pub.off_ranges = o;   o = align_up(o + VT * 8);
  L->pub.off_tilemax = o;  o = align_up(o + VT * 8);
  L->pub.off_records = o;  o = align_up(o + (size_t)d.R_cap * kRecBytes);
  L->pub.off_final_T = o;  o = align_up(o + VN * 4);
  L->pub.off_ncontrib = o; o = align_up(o + VN * 4);
  L->pub.off_order = o;    o = align_up(o + VT * 4);
  L->pub.off_masks = o;    o = align_up(o + (size_t)d.R_cap + 64);
  L->n_slots = (size_t)d.R_cap / GHR_SEGMENT + VT + 1;
  L->pub.off_tilefinal = o; o = align_up(o + VT * 256 * 16);
  L->pub.off_ckpt = o;     o = align_up(o + L->n_slots * 256 * 16);
  L->pub.off_units = o;    o = align_up(o + L->n_slots * 16);
  L->pub.state_bytes = o;

  // ---- temp (forward): zeroed prefix first ----
  L->n_heavy = (size_t)d.R_cap / kChunk + 1;          // tiles with more than kChunk instances
  L->n_hchunks = 2 * ((size_t)d.R_cap / kChunk) + 1;  // their kChunk-sized pieces
  L->n_chunks = 3 * ((size_t)d.R_cap / kChunk) + VT + 8;   // sort items: light tiles + depth buckets
  size_t t = 0;
  L->t_tile_count = t;   t = align_up(t + VT * 4);
  L->t_cursor = t;       t = align_up(t + VT * 4);
  L->t_misc = t;         t = align_up(t + 64);
  L->t_dmm = t;          t = align_up(t + L->n_heavy * 8);
  L->t_slab_count = t;   t = align_up(t + L->n_heavy * 256 * 4);
  L->t_zero_bytes = t;
  L->t_slab_off = t;     t = align_up(t + L->n_heavy * 256 * 4);
  L->t_heavy_flag = t;   t = align_up(t + L->n_heavy * 4);
  L->t_heavy = t;        t = align_up(t + L->n_heavy * 4);
  L->t_heavy_id = t;     t = align_up(t + VT * 4);
  L->t_hchunks = t;      t = align_up(t + L->n_hchunks * 8);
  L->t_inst = t;         t = align_up(t + (size_t)d.R_cap * 8);
  L->t_inst_b = t;       t = align_up(t + (size_t)d.R_cap * 8);
  L->t_chunks = t;       t = align_up(t + L->n_chunks * 16);
  L->pub.temp_bytes = t;
  L->pub.temp_bwd_bytes = align_up(VP * kAccStride * 4);
  return GHR_OK;
}

static int check_cuda(cudaError_t e, const char *what, bool debug, cudaStream_t s) {
  if (e == cudaSuccess && debug) {
    e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return GHR_ECUDA;
  }
  return GHR_OK;
}

#define GHR_TRY(call, what)                                  \
  do {                                                       \
    int _rc = check_cuda((call), what, debug, s);            \
    if (_rc != GHR_OK) return _rc;                           \
  } while (0)

struct StageTimer {
  void **ev; cudaStream_t s;
  void start(int i) { if (ev) cudaEventRecord((cudaEvent_t)ev[2 * i], s); }
  void stop(int i) { if (ev) cudaEventRecord((cudaEvent_t)ev[2 * i + 1], s); }
};

// FP32 roofline probes: 8 independent dependent-FMA chains per thread, 16x unrolled (128 FMA instructions
// per loop trip), multiplier and addend in registers (the 3-register form the blend kernels issue).
// kPacked: the chains are fp32 pairs (FFMA2), two FMAs per instruction.
template <bool kPacked>
__global__ void __launch_bounds__(256) fp32_probe_kernel(int iters, const float *__restrict__ in, float *sink) {
  const float m = in[0], c = in[1];
  if (kPacked) {
    f32x2 a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = pk2(threadIdx.x * 1e-3f + j, threadIdx.x * 2e-3f + j);
    const f32x2 m2 = bc2(m), c2 = bc2(c);
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 16; u++)
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = fma2(a[j], m2, c2);
    }
    float r = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) r += lo2(a[j]) + hi2(a[j]);
    if (r == 123.456f) *sink = r;
  } else {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = threadIdx.x * 1e-3f + j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 16; u++)
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = __fmaf_rn(a[j], m, c);
    }
    float r = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) r += a[j];
    if (r == 123.456f) *sink = r;
  }
}

}  // namespace ghr

using namespace ghr;

extern "C" {

int ghr_abi_version(void) { return GHR_ABI_VERSION; }
const char *ghr_last_error(void) { return g_err; }

size_t ghr_struct_size(const char *name) {
  if (!name) return 0;
  if (!strcmp(name, "GhrDims")) return sizeof(GhrDims);
  if (!strcmp(name, "GhrLayout")) return sizeof(GhrLayout);
  if (!strcmp(name, "GhrStatus")) return sizeof(GhrStatus);
  if (!strcmp(name, "GhrForwardArgs")) return sizeof(GhrForwardArgs);
  if (!strcmp(name, "GhrBackwardArgs")) return sizeof(GhrBackwardArgs);
  if (!strcmp(name, "GhrAttributeArgs")) return sizeof(GhrAttributeArgs);
  if (!strcmp(name, "GhrAttributeGrads")) return sizeof(GhrAttributeGrads);
  return 0;
}

int ghr_layout(const GhrDims *dims, GhrLayout *out) {
  if (!dims || !out) { set_error("ghr_layout: NULL argument"); return GHR_EINVAL; }
  Layout L;
  int rc = compute_layout(*dims, &L);
  if (rc != GHR_OK) return rc;
  *out = L.pub;
  return GHR_OK;
}

static int check_inputs(const char *fn, const GhrDims &d, const float *means3D, const float *opac, const float *scales,
                        const float *rots, const float *cov, const float *shs, const float *colors,
                        const float *view, const float *proj, const float *campos, const float *bg) {
  if (d.P > 0 && (!means3D || !opac)) { set_error("%s: means3D/opacities are required", fn); return GHR_EINVAL; }
  if (!view || !proj || !campos || !bg) { set_error("%s: viewmatrix/projmatrix/campos/bg are required", fn); return GHR_EINVAL; }
  if (d.P > 0) {
    if ((shs != nullptr) == (colors != nullptr)) { set_error("%s: provide exactly one of shs / colors_precomp", fn); return GHR_EINVAL; }
    bool sr = scales != nullptr && rots != nullptr;
    if (sr == (cov != nullptr) || ((scales != nullptr) != (rots != nullptr))) {
      set_error("%s: provide exactly one of (scales, rotations) / cov3D_precomp", fn);
      return GHR_EINVAL;
    }
    if (shs && d.M < (d.sh_degree + 1) * (d.sh_degree + 1)) {
      set_error("%s: M=%d SH coefficients cannot hold degree %d", fn, d.M, d.sh_degree);
      return GHR_EINVAL;
    }
  }
  return GHR_OK;
}

int ghr_forward(const GhrForwardArgs *a, void *cuda_stream) {
  if (!a) { set_error("ghr_forward: NULL args"); return GHR_EINVAL; }
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const GhrDims &d = a->dims;
  Layout L;
  int rc = compute_layout(d, &L);
  if (rc != GHR_OK) return rc;
  rc = check_inputs("ghr_forward", d, a->means3D, a->opacities, a->scales, a->rotations, a->cov3D_precomp, a->shs,
                    a->colors_precomp, a->viewmatrix, a->projmatrix, a->campos, a->bg);
  if (rc != GHR_OK) return rc;
  if (!a->out_color || (d.P > 0 && !a->radii) || !a->state || !a->temp) {
    set_error("ghr_forward: out_color/radii/state/temp are required");
    return GHR_EINVAL;
  }
  if (a->state_bytes < L.pub.state_bytes || a->temp_bytes < L.pub.temp_bytes) {
    set_error("ghr_forward: workspace too small (state %zu < %zu or temp %zu < %zu)", a->state_bytes,
              L.pub.state_bytes, a->temp_bytes, L.pub.temp_bytes);
    return GHR_ENOSPC;
  }
  const bool debug = a->flags & GHR_FLAG_DEBUG;
  char *state = (char *)a->state, *temp = (char *)a->temp;
  Cameras cam{a->viewmatrix, a->projmatrix, a->campos, a->tanfov, a->bg, a->tanfovx, a->tanfovy, a->bg_stride};
  Gaussians g{a->means3D, a->opacities, a->scales, a->rotations, a->cov3D_precomp, a->shs, a->colors_precomp};

  StageTimer tm{a->stage_events, s};
  if (a->reuse_state) {
    // geometry reuse: copy what depends on geometry only, recompute the colours, blend
    if (d.V != 1) { set_error("ghr_forward: reuse_state needs V == 1"); return GHR_EINVAL; }
    GhrDims dold = d;
    dold.M = a->reuse_M;
    Layout Lold;
    rc = compute_layout(dold, &Lold);
    if (rc != GHR_OK) return rc;
    const char *old_state = (const char *)a->reuse_state;
    tm.start(0);
    GHR_TRY(launch_recolor_geom(d, L, Lold, cam, g, old_state, state, a->radii, s), "ghr_forward: recolour geometry");
    tm.stop(0);
    tm.start(1);
    tm.stop(1);
    tm.start(2);
    tm.stop(2);
    tm.start(3);
    GHR_TRY(launch_reuse_binning(d, L, Lold, old_state, state, a->seq, s), "ghr_forward: reuse binning");
    tm.stop(3);
    if (a->host_status)
      GHR_TRY(cudaMemcpyAsync(a->host_status, state + L.pub.off_status, sizeof(GhrStatus), cudaMemcpyDeviceToHost, s),
              "ghr_forward: status copy");
    tm.start(4);
    GHR_TRY(launch_blend_forward(d, L, cam, state, a->out_color, a->out_mask, s), "ghr_forward: blend");
    tm.stop(4);
    return GHR_OK;
  }
  tm.start(0);
  GHR_TRY(cudaMemsetAsync(temp, 0, L.t_zero_bytes, s), "ghr_forward: memset(temp)");
  GHR_TRY(launch_preprocess(d, L, cam, g, a->scale_modifier, a->flags, state, temp, a->radii, s),
          "ghr_forward: preprocess");
  tm.stop(0);
  tm.start(1);
  GHR_TRY(launch_tile_scan_schedule(d, L, state, temp, a->seq, s), "ghr_forward: tile scan+schedule");
  tm.stop(1);
  if (a->host_status)
    GHR_TRY(cudaMemcpyAsync(a->host_status, state + L.pub.off_status, sizeof(GhrStatus), cudaMemcpyDeviceToHost, s),
            "ghr_forward: status copy");
  tm.start(2);
  GHR_TRY(launch_duplicate(d, L, state, temp, s), "ghr_forward: duplicate");
  tm.stop(2);
  tm.start(3);
  GHR_TRY(launch_sort_gather(d, L, state, temp, a->dbg_keys_sorted, a->dbg_point_list, s),
          "ghr_forward: sort+gather");
  tm.stop(3);
  tm.start(4);
  GHR_TRY(launch_blend_forward(d, L, cam, state, a->out_color, a->out_mask, s), "ghr_forward: blend");
  tm.stop(4);
  return GHR_OK;
}

int ghr_backward(const GhrBackwardArgs *a, void *cuda_stream) {
  if (!a) { set_error("ghr_backward: NULL args"); return GHR_EINVAL; }
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const GhrDims &d = a->dims;
  Layout L;
  int rc = compute_layout(d, &L);
  if (rc != GHR_OK) return rc;
  rc = check_inputs("ghr_backward", d, a->means3D, a->opacities, a->scales, a->rotations, a->cov3D_precomp, a->shs,
                    a->colors_precomp, a->viewmatrix, a->projmatrix, a->campos, a->bg);
  if (rc != GHR_OK) return rc;
  if (!a->dL_dout_color || !a->state || !a->temp) {
    set_error("ghr_backward: dL_dout_color/state/temp are required");
    return GHR_EINVAL;
  }
  if (a->state_bytes < L.pub.state_bytes || a->temp_bytes < L.pub.temp_bwd_bytes) {
    set_error("ghr_backward: workspace too small (state %zu < %zu or temp %zu < %zu)", a->state_bytes,
              L.pub.state_bytes, a->temp_bytes, L.pub.temp_bwd_bytes);
    return GHR_ENOSPC;
  }
  if (d.P == 0) return GHR_OK;
  const bool debug = a->flags & GHR_FLAG_DEBUG;
  const char *state = (const char *)a->state;
  float *acc = (float *)a->temp;
  Cameras cam{a->viewmatrix, a->projmatrix, a->campos, a->tanfov, a->bg, a->tanfovx, a->tanfovy, a->bg_stride};
  Gaussians g{a->means3D, a->opacities, a->scales, a->rotations, a->cov3D_precomp, a->shs, a->colors_precomp};
  GradOut go{a->accumulate, a->dL_dmeans3D, a->dL_dmeans2D, a->dL_dcolors, a->dL_dopacity, a->dL_dcov3D,
             a->dL_dsh, a->dL_dscales, a->dL_drotations, a->dL_dconic};

  StageTimer tm{a->stage_events, s};
  tm.start(0);
  GHR_TRY(cudaMemsetAsync(acc, 0, (size_t)d.V * d.P * kAccStride * sizeof(float), s), "ghr_backward: memset(acc)");
  GHR_TRY(launch_blend_backward(d, L, cam, state, a->dL_dout_color, a->dL_dout_mask, acc, s), "ghr_backward: blend");
  tm.stop(0);
  tm.start(1);
  GHR_TRY(launch_preprocess_backward(d, L, cam, g, a->scale_modifier, state, acc, go, s),
          "ghr_backward: preprocess");
  tm.stop(1);
  return GHR_OK;
}

int ghr_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                     uint8_t *present, void *cuda_stream) {
  (void)projmatrix;
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
    set_error("ghr_mark_visible: bad arguments");
    return GHR_EINVAL;
  }
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const bool debug = false;
  GHR_TRY(launch_mark_visible(P, means3D, viewmatrix, present, s), "ghr_mark_visible");
  return GHR_OK;
}

int ghr_cameras_from_w2c(int32_t V, const float *w2c, const float *K, int32_t H, int32_t W, float znear, float zfar,
                         float *viewmatrix, float *projmatrix, float *campos, float *tanfov, void *cuda_stream) {
  if (V < 0 || H < 1 || W < 1 || !(zfar > znear) ||
      (V > 0 && (!w2c || !K || !viewmatrix || !projmatrix || !campos || !tanfov))) {
    set_error("ghr_cameras_from_w2c: bad arguments");
    return GHR_EINVAL;
  }
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const bool debug = false;
  GHR_TRY(launch_cameras_from_w2c(V, w2c, K, H, W, znear, zfar, viewmatrix, projmatrix, campos, tanfov, s),
          "ghr_cameras_from_w2c");
  return GHR_OK;
}

int ghr_event_create(void **event_out) {
  if (!event_out) { set_error("ghr_event_create: NULL"); return GHR_EINVAL; }
  cudaEvent_t e;
  cudaError_t rc = cudaEventCreate(&e);
  if (rc != cudaSuccess) { set_error("ghr_event_create: %s", cudaGetErrorString(rc)); return GHR_ECUDA; }
  *event_out = (void *)e;
  return GHR_OK;
}
int ghr_event_destroy(void *event) {
  if (event) cudaEventDestroy((cudaEvent_t)event);
  return GHR_OK;
}
int ghr_event_record(void *event, void *cuda_stream) {
  cudaError_t rc = cudaEventRecord((cudaEvent_t)event, (cudaStream_t)cuda_stream);
  if (rc != cudaSuccess) { set_error("ghr_event_record: %s", cudaGetErrorString(rc)); return GHR_ECUDA; }
  return GHR_OK;
}
int ghr_event_elapsed_ms(void *start, void *stop, float *ms_out) {
  if (!ms_out) { set_error("ghr_event_elapsed_ms: NULL"); return GHR_EINVAL; }
  cudaError_t rc = cudaEventElapsedTime(ms_out, (cudaEvent_t)start, (cudaEvent_t)stop);
  if (rc != cudaSuccess) { set_error("ghr_event_elapsed_ms: %s", cudaGetErrorString(rc)); return GHR_ECUDA; }
  return GHR_OK;
}

int ghr_fp32_probe(int32_t iters, int32_t packed, const float *in, float *sink, double *flops_out, void *cuda_stream) {
  if (iters <= 0 || !sink || !in) { set_error("ghr_fp32_probe: bad arguments"); return GHR_EINVAL; }
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const bool debug = false;
  const int blocks = 148 * 8, threads = 256;
  if (packed) fp32_probe_kernel<true><<<blocks, threads, 0, s>>>(iters, in, sink);
  else fp32_probe_kernel<false><<<blocks, threads, 0, s>>>(iters, in, sink);
  if (flops_out) *flops_out = 2.0 * 128.0 * (packed ? 2.0 : 1.0) * (double)iters * blocks * threads;
  GHR_TRY(cudaGetLastError(), "ghr_fp32_probe");
  return GHR_OK;
}

int ghr_read_status_async(const void *state, GhrStatus *host_status, void *cuda_stream) {
  if (!state || !host_status) { set_error("ghr_read_status_async: NULL argument"); return GHR_EINVAL; }
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const bool debug = false;
  GHR_TRY(cudaMemcpyAsync(host_status, state, sizeof(GhrStatus), cudaMemcpyDeviceToHost, s),
          "ghr_read_status_async");
  return GHR_OK;
}

}  // extern "C"
