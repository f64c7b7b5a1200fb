// Upstream-STRUCTURED GPU stand-in for the binning + blend stages (bench-only; never part of the product).
//
// The reference's rasterizer (pip diff-gaussian-rasterization, /root/reference/environment.yml:129) cannot be
// installed here, so the "reference CUDA rasterizer on the same B200" of BASELINE.json's north_star has no
// measurable value.  This file restates the STRUCTURE of its stages K2-K7 from the behavioural specification in
// SURVEY.md Appendix A.4-A.6 / §2a -- not from its source, which is not in /root/reference -- so that libghr's
// numbers get a GPU-class denominator and every redesign (no global sort, no host sync, segment-parallel
// backward, warp reductions) can be priced:
//   K2  cub::DeviceScan::InclusiveSum over tiles_touched + BLOCKING device->host read of the instance count
//   K3  duplicateWithKeys: one (tile << 32 | depth bits) key + Gaussian id per touched tile
//   K4  cub::DeviceRadixSort::SortPairs on bits [0, 32 + msb(tiles))
//   K5  identifyTileRanges
//   K6  render forward: one 16x16 CTA per tile, rounds of 256 entries staged in shared memory
//   K7  render backward: same CTA shape, back to front, NINE per-thread atomicAdd per contributing pair
// It is reported as gpu_baseline.kind = "restatement": it can NOT earn the "x upstream" claim.  Per-Gaussian
// inputs (xy, depth, radius, conic, opacity, colour, tiles_touched) come from libghr's own preprocess (the
// geometry block of a ghr_forward state), unpacked into upstream's structure-of-arrays layout first.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

namespace {

constexpr int kBlockX = 16, kBlockY = 16, kBlock = kBlockX * kBlockY;

struct Buf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t need(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes + bytes / 4);
    if (e == cudaSuccess) cap = bytes + bytes / 4;
    return e;
  }
  ~Buf() { if (p) cudaFree(p); }
};

}  // namespace

struct Sgs {
  // geometry state (upstream's geomBuffer), structure of arrays
  Buf xy, conic_opacity, rgb, depths, radii, tiles_touched, point_offsets, scan_tmp;
  // binning state
  Buf keys_unsorted, keys, vals_unsorted, vals, sort_tmp, ranges;
  // image state
  Buf final_T, n_contrib;
  uint32_t *host_R = nullptr;   // pinned
  int P = 0;
  uint32_t R = 0;
};

namespace {

__global__ void unpack_kernel(int P, const float4 *__restrict__ geom, float2 *xy, float4 *conic_opacity, float *rgb,
                              float *depths, int *radii, uint32_t *tiles_touched) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  // libghr geometry record: {x, y, A, B} {C, opacity, thr, 0} {r, g, b, depth} {radius, tiles, 0, 0}
  const float4 q0 = geom[4 * (size_t)i], q1 = geom[4 * (size_t)i + 1], q2 = geom[4 * (size_t)i + 2],
               q3 = geom[4 * (size_t)i + 3];
  xy[i] = make_float2(q0.x, q0.y);
  conic_opacity[i] = make_float4(q0.z, q0.w, q1.x, q1.y);
  rgb[3 * i] = q2.x;
  rgb[3 * i + 1] = q2.y;
  rgb[3 * i + 2] = q2.z;
  depths[i] = q2.w;
  radii[i] = __float_as_int(q3.x);
  tiles_touched[i] = __float_as_uint(q3.y);
}

__device__ __forceinline__ void get_rect(float2 p, int r, int gx, int gy, int &x0, int &y0, int &x1, int &y1) {
  x0 = min(gx, max(0, (int)((p.x - r) / kBlockX)));
  y0 = min(gy, max(0, (int)((p.y - r) / kBlockY)));
  x1 = min(gx, max(0, (int)((p.x + r + kBlockX - 1) / kBlockX)));
  y1 = min(gy, max(0, (int)((p.y + r + kBlockY - 1) / kBlockY)));
}

__global__ void duplicate_with_keys(int P, const float2 *xy, const float *depths, const uint32_t *offsets, const int *radii,
                                    int gx, int gy, uint64_t *keys, uint32_t *vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P || radii[i] <= 0) return;
  uint32_t off = i == 0 ? 0u : offsets[i - 1];
  int x0, y0, x1, y1;
  get_rect(xy[i], radii[i], gx, gy, x0, y0, x1, y1);
  for (int y = y0; y < y1; y++)
    for (int x = x0; x < x1; x++) {
      keys[off] = ((uint64_t)(y * gx + x) << 32) | (uint64_t)__float_as_uint(depths[i]);
      vals[off] = (uint32_t)i;
      off++;
    }
}

__global__ void identify_tile_ranges(uint32_t R, const uint64_t *keys, uint2 *ranges) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const uint32_t t = (uint32_t)(keys[i] >> 32);
  if (i == 0) {
    ranges[t].x = 0;
  } else {
    const uint32_t tp = (uint32_t)(keys[i - 1] >> 32);
    if (t != tp) {
      ranges[tp].y = i;
      ranges[t].x = i;
    }
  }
  if (i == R - 1) ranges[t].y = R;
}

__global__ void __launch_bounds__(kBlock)
render_forward(const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, int W, int H,
               const float2 *__restrict__ xy, const float *__restrict__ rgb, const float4 *__restrict__ conic_opacity,
               float *__restrict__ final_T, uint32_t *__restrict__ n_contrib, const float *__restrict__ bg,
               float *__restrict__ out) {
  __shared__ int s_id[kBlock];
  __shared__ float2 s_xy[kBlock];
  __shared__ float4 s_co[kBlock];
  const int gx = (W + kBlockX - 1) / kBlockX;
  const int px = blockIdx.x * kBlockX + threadIdx.x, py = blockIdx.y * kBlockY + threadIdx.y;
  const int tid = threadIdx.y * kBlockX + threadIdx.x;
  const bool inside = px < W && py < H;
  const float2 pixf = make_float2((float)px, (float)py);
  const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
  const int rounds = ((int)(range.y - range.x) + kBlock - 1) / kBlock;
  int todo = (int)(range.y - range.x);
  bool done = !inside;
  float T = 1.0f, C[3] = {0.f, 0.f, 0.f};
  uint32_t contributor = 0, last = 0;
  for (int r = 0; r < rounds; r++, todo -= kBlock) {
    if (__syncthreads_count(done) == kBlock) break;
    const int progress = r * kBlock + tid;
    if ((int)range.x + progress < (int)range.y) {
      const int id = (int)point_list[range.x + progress];
      s_id[tid] = id;
      s_xy[tid] = xy[id];
      s_co[tid] = conic_opacity[id];
    }
    __syncthreads();
    for (int j = 0; !done && j < min(kBlock, todo); j++) {
      contributor++;
      const float2 p = s_xy[j];
      const float dx = p.x - pixf.x, dy = p.y - pixf.y;
      const float4 co = s_co[j];
      const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
      if (power > 0.0f) continue;
      const float alpha = min(0.99f, co.w * __expf(power));
      if (alpha < 1.0f / 255.0f) continue;
      const float test_T = T * (1.f - alpha);
      if (test_T < 0.0001f) {
        done = true;
        continue;
      }
      const float *c = rgb + 3 * (size_t)s_id[j];
      for (int ch = 0; ch < 3; ch++) C[ch] += c[ch] * alpha * T;
      T = test_T;
      last = contributor;
    }
  }
  if (inside) {
    const size_t pix = (size_t)py * W + px;
    final_T[pix] = T;
    n_contrib[pix] = last;
    for (int ch = 0; ch < 3; ch++) out[(size_t)ch * H * W + pix] = C[ch] + T * bg[ch];
  }
}

__global__ void __launch_bounds__(kBlock)
render_backward(const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, int W, int H,
                const float *__restrict__ bg, const float2 *__restrict__ xy, const float4 *__restrict__ conic_opacity,
                const float *__restrict__ rgb, const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
                const float *__restrict__ dL_dpixels, float *__restrict__ dL_dmean2D, float *__restrict__ dL_dconic,
                float *__restrict__ dL_dopacity, float *__restrict__ dL_dcolors) {
  __shared__ int s_id[kBlock];
  __shared__ float2 s_xy[kBlock];
  __shared__ float4 s_co[kBlock];
  __shared__ float s_c[3 * kBlock];
  const int gx = (W + kBlockX - 1) / kBlockX;
  const int px = blockIdx.x * kBlockX + threadIdx.x, py = blockIdx.y * kBlockY + threadIdx.y;
  const int tid = threadIdx.y * kBlockX + threadIdx.x;
  const bool inside = px < W && py < H;
  const float2 pixf = make_float2((float)px, (float)py);
  const size_t pix = (size_t)py * W + px;
  const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
  const int rounds = ((int)(range.y - range.x) + kBlock - 1) / kBlock;
  int todo = (int)(range.y - range.x);
  bool done = !inside;
  const float T_final = inside ? final_T[pix] : 0.f;
  float T = T_final;
  uint32_t contributor = (uint32_t)todo;
  const uint32_t last = inside ? n_contrib[pix] : 0u;
  float accum[3] = {0.f, 0.f, 0.f}, dLp[3] = {0.f, 0.f, 0.f};
  if (inside)
    for (int ch = 0; ch < 3; ch++) dLp[ch] = dL_dpixels[(size_t)ch * H * W + pix];
  float last_alpha = 0.f, last_c[3] = {0.f, 0.f, 0.f};
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  for (int r = 0; r < rounds; r++, todo -= kBlock) {
    __syncthreads();
    const int progress = r * kBlock + tid;
    if ((int)range.x + progress < (int)range.y) {
      const int id = (int)point_list[range.y - progress - 1];
      s_id[tid] = id;
      s_xy[tid] = xy[id];
      s_co[tid] = conic_opacity[id];
      for (int ch = 0; ch < 3; ch++) s_c[ch * kBlock + tid] = rgb[3 * (size_t)id + ch];
    }
    __syncthreads();
    for (int j = 0; !done && j < min(kBlock, todo); j++) {
      contributor--;
      if (contributor >= last) continue;
      const float2 p = s_xy[j];
      const float dx = p.x - pixf.x, dy = p.y - pixf.y;
      const float4 co = s_co[j];
      const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
      if (power > 0.0f) continue;
      const float G = __expf(power);
      const float alpha = min(0.99f, co.w * G);
      if (alpha < 1.0f / 255.0f) continue;
      T = T / (1.f - alpha);
      const float dchannel_dcolor = alpha * T;
      float dL_dalpha = 0.f;
      const int id = s_id[j];
      for (int ch = 0; ch < 3; ch++) {
        const float c = s_c[ch * kBlock + j];
        accum[ch] = last_alpha * last_c[ch] + (1.f - last_alpha) * accum[ch];
        last_c[ch] = c;
        dL_dalpha += (c - accum[ch]) * dLp[ch];
        atomicAdd(&dL_dcolors[3 * (size_t)id + ch], dchannel_dcolor * dLp[ch]);
      }
      dL_dalpha *= T;
      last_alpha = alpha;
      float bg_dot = 0.f;
      for (int ch = 0; ch < 3; ch++) bg_dot += bg[ch] * dLp[ch];
      dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
      const float dL_dG = co.w * dL_dalpha;
      const float gdx = G * dx, gdy = G * dy;
      const float dG_ddelx = -gdx * co.x - gdy * co.y, dG_ddely = -gdy * co.z - gdx * co.y;
      atomicAdd(&dL_dmean2D[2 * (size_t)id], dL_dG * dG_ddelx * ddelx_dx);
      atomicAdd(&dL_dmean2D[2 * (size_t)id + 1], dL_dG * dG_ddely * ddely_dy);
      atomicAdd(&dL_dconic[3 * (size_t)id], -0.5f * gdx * dx * dL_dG);
      atomicAdd(&dL_dconic[3 * (size_t)id + 1], -0.5f * gdx * dy * dL_dG);
      atomicAdd(&dL_dconic[3 * (size_t)id + 2], -0.5f * gdy * dy * dL_dG);
      atomicAdd(&dL_dopacity[id], G * dL_dalpha);
    }
  }
}

int fail(cudaError_t e, const char *what) {
  fprintf(stderr, "standin: %s: %s\n", what, cudaGetErrorString(e));
  return 1;
}
#define SGS_TRY(x, what)                      \
  do {                                        \
    cudaError_t _e = (x);                     \
    if (_e != cudaSuccess) return fail(_e, what); \
  } while (0)

// position of the highest set bit + 1, found by bisection from 16
uint32_t higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4, step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step;
    else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

}  // namespace

extern "C" {

int sgs_create(Sgs **out) {
  Sgs *s = new Sgs();
  if (cudaMallocHost((void **)&s->host_R, sizeof(uint32_t)) != cudaSuccess) {
    delete s;
    return 1;
  }
  *out = s;
  return 0;
}

int sgs_destroy(Sgs *s) {
  if (!s) return 0;
  if (s->host_R) cudaFreeHost(s->host_R);
  delete s;
  return 0;
}

// untimed: libghr geometry block (P x 4 float4) -> upstream's structure-of-arrays geometry state
int sgs_unpack(Sgs *s, int P, const void *geom, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  s->P = P;
  SGS_TRY(s->xy.need((size_t)P * 8), "alloc");
  SGS_TRY(s->conic_opacity.need((size_t)P * 16), "alloc");
  SGS_TRY(s->rgb.need((size_t)P * 12), "alloc");
  SGS_TRY(s->depths.need((size_t)P * 4), "alloc");
  SGS_TRY(s->radii.need((size_t)P * 4), "alloc");
  SGS_TRY(s->tiles_touched.need((size_t)P * 4), "alloc");
  SGS_TRY(s->point_offsets.need((size_t)P * 4), "alloc");
  unpack_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, (const float4 *)geom, (float2 *)s->xy.p, (float4 *)s->conic_opacity.p,
                                                 (float *)s->rgb.p, (float *)s->depths.p, (int *)s->radii.p,
                                                 (uint32_t *)s->tiles_touched.p);
  SGS_TRY(cudaGetLastError(), "unpack");
  return 0;
}

// K2-K6.  Blocks the host once (the instance count), like upstream.
int sgs_forward(Sgs *s, int H, int W, const float *bg, float *out_color, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int P = s->P, gx = (W + kBlockX - 1) / kBlockX, gy = (H + kBlockY - 1) / kBlockY;
  size_t scan_bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (uint32_t *)s->tiles_touched.p, (uint32_t *)s->point_offsets.p, P, st);
  SGS_TRY(s->scan_tmp.need(scan_bytes), "alloc");
  SGS_TRY(cub::DeviceScan::InclusiveSum(s->scan_tmp.p, scan_bytes, (uint32_t *)s->tiles_touched.p,
                                        (uint32_t *)s->point_offsets.p, P, st), "scan");
  SGS_TRY(cudaMemcpyAsync(s->host_R, (uint32_t *)s->point_offsets.p + (P - 1), 4, cudaMemcpyDeviceToHost, st), "D2H");
  SGS_TRY(cudaStreamSynchronize(st), "sync");                      // upstream's blocking read of num_rendered
  const uint32_t R = *s->host_R;
  s->R = R;
  SGS_TRY(s->keys_unsorted.need((size_t)R * 8 + 8), "alloc");
  SGS_TRY(s->keys.need((size_t)R * 8 + 8), "alloc");
  SGS_TRY(s->vals_unsorted.need((size_t)R * 4 + 4), "alloc");
  SGS_TRY(s->vals.need((size_t)R * 4 + 4), "alloc");
  SGS_TRY(s->ranges.need((size_t)gx * gy * 8), "alloc");
  SGS_TRY(s->final_T.need((size_t)H * W * 4), "alloc");
  SGS_TRY(s->n_contrib.need((size_t)H * W * 4), "alloc");
  duplicate_with_keys<<<(P + 255) / 256, 256, 0, st>>>(P, (const float2 *)s->xy.p, (const float *)s->depths.p,
                                                       (const uint32_t *)s->point_offsets.p, (const int *)s->radii.p, gx, gy,
                                                       (uint64_t *)s->keys_unsorted.p, (uint32_t *)s->vals_unsorted.p);
  const int bit = (int)higher_msb((uint32_t)(gx * gy));
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint64_t *)s->keys_unsorted.p, (uint64_t *)s->keys.p,
                                  (uint32_t *)s->vals_unsorted.p, (uint32_t *)s->vals.p, (int)R, 0, 32 + bit, st);
  SGS_TRY(s->sort_tmp.need(sort_bytes), "alloc");
  SGS_TRY(cub::DeviceRadixSort::SortPairs(s->sort_tmp.p, sort_bytes, (uint64_t *)s->keys_unsorted.p, (uint64_t *)s->keys.p,
                                          (uint32_t *)s->vals_unsorted.p, (uint32_t *)s->vals.p, (int)R, 0, 32 + bit, st),
          "sort");
  SGS_TRY(cudaMemsetAsync(s->ranges.p, 0, (size_t)gx * gy * 8, st), "memset");
  if (R) identify_tile_ranges<<<(R + 255) / 256, 256, 0, st>>>(R, (const uint64_t *)s->keys.p, (uint2 *)s->ranges.p);
  render_forward<<<dim3(gx, gy), dim3(kBlockX, kBlockY), 0, st>>>(
      (const uint2 *)s->ranges.p, (const uint32_t *)s->vals.p, W, H, (const float2 *)s->xy.p, (const float *)s->rgb.p,
      (const float4 *)s->conic_opacity.p, (float *)s->final_T.p, (uint32_t *)s->n_contrib.p, bg, out_color);
  SGS_TRY(cudaGetLastError(), "forward launches");
  return 0;
}

// K7 (+ the memsets of its four accumulators).  dL_dmean2D [P,2], dL_dconic [P,3] (xx, xy, yy), dL_dopacity [P],
// dL_dcolors [P,3]
int sgs_backward(Sgs *s, int H, int W, const float *bg, const float *dL_dout, float *dL_dmean2D, float *dL_dconic,
                 float *dL_dopacity, float *dL_dcolors, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int P = s->P, gx = (W + kBlockX - 1) / kBlockX, gy = (H + kBlockY - 1) / kBlockY;
  SGS_TRY(cudaMemsetAsync(dL_dmean2D, 0, (size_t)P * 8, st), "memset");
  SGS_TRY(cudaMemsetAsync(dL_dconic, 0, (size_t)P * 12, st), "memset");
  SGS_TRY(cudaMemsetAsync(dL_dopacity, 0, (size_t)P * 4, st), "memset");
  SGS_TRY(cudaMemsetAsync(dL_dcolors, 0, (size_t)P * 12, st), "memset");
  render_backward<<<dim3(gx, gy), dim3(kBlockX, kBlockY), 0, st>>>(
      (const uint2 *)s->ranges.p, (const uint32_t *)s->vals.p, W, H, bg, (const float2 *)s->xy.p,
      (const float4 *)s->conic_opacity.p, (const float *)s->rgb.p, (const float *)s->final_T.p,
      (const uint32_t *)s->n_contrib.p, dL_dout, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors);
  SGS_TRY(cudaGetLastError(), "backward launch");
  return 0;
}

uint32_t sgs_num_rendered(Sgs *s) { return s->R; }
const void *sgs_n_contrib(Sgs *s) { return s->n_contrib.p; }
const void *sgs_final_T(Sgs *s) { return s->final_T.p; }
const void *sgs_ranges(Sgs *s) { return s->ranges.p; }
const void *sgs_sorted_keys(Sgs *s) { return s->keys.p; }
int sgs_copy(void *dst, const void *src, size_t n, void *stream) {
  SGS_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "copy");
  return 0;
}

}  // extern "C"
