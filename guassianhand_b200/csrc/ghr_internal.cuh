// Internal declarations shared by the translation units of libghr.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ghr.h"

namespace ghr {

constexpr int kTile = 16;            // 16x16 pixel tiles
constexpr int kRecBytes = 48;        // sorted instance record: 3 x float4
constexpr int kChunk = 2048;         // instances per work item of the per-tile sort
// words of the zero-initialised temp block `misc` (16 words): [0..4] binning work counts (binning.cu), then
constexpr int kMiscVisible = 5;     // Gaussians with radius > 0, summed over views (preprocess)
constexpr int kMiscPrefilter = 6;   // GHR_FLAG_PREFILTERED violated (preprocess)
constexpr int kMaxSmemTiles = 16384; // tiles per view whose per-block counters fit in shared memory
constexpr int kSeg = GHR_SEGMENT;     // instances per backward work unit
constexpr int kAccStride = 12;       // floats per (view,Gaussian) backward accumulator

// ---- canonical fp32 order: explicit rn intrinsics are never re-contracted ----
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
// a0*b0 + a1*b1  ->  fma(a0,b0, a1*b1)
__device__ __forceinline__ float dot2(float a0, float b0, float a1, float b1) {
  return ffma(a0, b0, fmul(a1, b1));
}
// a0*b0 + a1*b1 + a2*b2  ->  fma(a2,b2, fma(a0,b0, a1*b1))
__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  return ffma(a2, b2, ffma(a0, b0, fmul(a1, b1)));
}
__device__ __forceinline__ float dot3a(float a0, float b0, float a1, float b1, float a2, float b2, float c) {
  return fadd(dot3(a0, b0, a1, b1, a2, b2), c);
}

#ifdef GHR_EXACT_EXP
// Test-only build variant (libghr_exact.so, tests/test_gpu_exact_variant.py): libdevice expf and an IEEE
// divide in the blend kernels instead of ex2.approx / rcp.approx, to COUNT the pixels whose threshold
// decisions the fast functions change.  Never the shipped library.
__device__ __forceinline__ float rcp_fast(float x) { return __fdiv_rn(1.0f, x); }
#else
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#endif

// ---- packed fp32 pairs (sm_100a FFMA2 / FMUL2 / FADD2) ----
// Two IEEE round-to-nearest fp32 operations per instruction, one per half: the bits of every half are
// those of the scalar __fmaf_rn / __fmul_rn / __fadd_rn, so the canonical arithmetic does not change.
// ptxas folds a {x, x} pair into a scalar-broadcast operand and immediates likewise (no extra moves).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 bc2(float x) { return pk2(x, x); }
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float lo2(f32x2 v) { float a, b; upk2(v, a, b); return a; }
__device__ __forceinline__ float hi2(f32x2 v) { float a, b; upk2(v, a, b); return b; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Sub-block culling (never part of the canonical arithmetic).
// Exact box-constrained minimum of q(d) = A dx^2 + 2 B dx dy + C dy^2 over the pixel block
// [bx0,bx1] x [by0,by1] (d = pixel - centre) against the instance's threshold.  The minimiser of a
// convex quadratic over a box is the centre itself (q = 0) or lies on an edge facing the centre; the
// two candidates below cover every case.  Returns true if the instance may contribute to the block.
__device__ __forceinline__ bool cull_hit(const float4 a, const float4 b, float bx0, float bx1, float by0, float by1) {
  const float A = a.z, B = a.w, C = b.x, thr = b.z;
  const float u0 = bx0 - a.x, u1 = bx1 - a.x, v0 = by0 - a.y, v1 = by1 - a.y;
  const float uc = fminf(fmaxf(0.f, u0), u1), vc = fminf(fmaxf(0.f, v0), v1);
  const float v_e = fminf(fmaxf(-B * uc * rcp_fast(C), v0), v1);   // best v on the edge u = uc
  const float u_e = fminf(fmaxf(-B * vc * rcp_fast(A), u0), u1);   // best u on the edge v = vc
  const float q1 = A * uc * uc + 2.f * B * uc * v_e + C * v_e * v_e;
  const float q2 = A * u_e * u_e + 2.f * B * u_e * vc + C * vc * vc;
  return !(fminf(q1, q2) > thr);   // NaN-safe: anything unordered counts as a hit
}


// 8-bit mask of the 8x4-pixel sub-blocks of a 16x16 tile an instance may contribute to (bit w =
// sub-block (w&1, w>>1), the pixel block blend warp w owns).  Computed once per instance by
// gather_ranges and stored in record[1].w; the blend kernels only test a bit.  Same test as
// cull_hit for each of the 8 sub-blocks, with the per-column / per-row parts shared (2 columns x 4
// rows): ~11 instructions per sub-block.
__device__ __forceinline__ uint32_t subblock_mask(const float4 a, const float4 b, int tile_x0, int tile_y0) {
  const float A = a.z, B = a.w, C = b.x, thr = b.z;
  if (!(thr >= 0.f)) return 0u;          // opacity < 1/255: never contributes
  if (thr > 1e37f) return 0xFFu;         // ill-conditioned conic: never culled
  const float k1 = -B * rcp_fast(C), k2 = -B * rcp_fast(A), B2 = 2.f * B;
  float u0[2], u1[2], Auc2[2], Buc[2], vun[2];
#pragma unroll
  for (int c = 0; c < 2; c++) {
    u0[c] = (float)(tile_x0 + 8 * c) - a.x;
    u1[c] = u0[c] + 7.f;
    const float uc = fminf(fmaxf(0.f, u0[c]), u1[c]);
    Auc2[c] = A * uc * uc;
    Buc[c] = B2 * uc;
    vun[c] = k1 * uc;                    // unconstrained best v on the edge u = uc
  }
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float v0 = (float)(tile_y0 + 4 * j) - a.y, v1 = v0 + 3.f;
    const float vc = fminf(fmaxf(0.f, v0), v1);
    const float Cvc2 = C * vc * vc, Bvc = B2 * vc, uun = k2 * vc;
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const float v_e = fminf(fmaxf(vun[c], v0), v1);
      const float u_e = fminf(fmaxf(uun, u0[c]), u1[c]);
      const float q1 = fmaf(fmaf(C, v_e, Buc[c]), v_e, Auc2[c]);
      const float q2 = fmaf(fmaf(A, u_e, Bvc), u_e, Cvc2);
      if (!(fminf(q1, q2) > thr)) m |= 1u << (2 * j + c);
    }
  }
  return m;
}

// Exact division of n < 2^31 by a run-time constant: q = (n * m) >> s with a host-made magic
// (m = floor(2^(31+k)/d) + 1, k = ceil(log2 d), s = 31 + k); two integer instructions instead of
// the ~25-instruction emulated 32-bit divide.
struct FastDiv {
  uint32_t m, s, d;
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
    return (uint32_t)(((uint64_t)n * m) >> s);
  }
  __host__ __device__ __forceinline__ uint32_t mod(uint32_t n) const { return n - div(n) * d; }
};
inline FastDiv make_fastdiv(uint32_t d) {
  if (d == 0) d = 1;
  uint32_t k = 0;
  while ((1ull << k) < d) k++;
  FastDiv f;
  f.m = (uint32_t)(((1ull << (31 + k)) / d) + 1ull);
  f.s = 31 + k;
  f.d = d;
  return f;
}

struct Cameras {
  const float *view, *proj, *campos, *tanfov, *bg;
  float tanfovx, tanfovy;
  int bg_stride;
};

struct Gaussians {
  const float *means3D, *opacities, *scales, *rotations, *cov3D_precomp, *shs, *colors_precomp;
};

// Host-side layout of the two caller-owned buffers.
struct Layout {
  GhrLayout pub;
  int gx, gy, T;               // tiles per view
  size_t n_slots;              // checkpoint slots = upper bound of backward work units
  // temp (forward)
  size_t t_zero_bytes;         // prefix of temp that must be zeroed before a forward
  size_t t_tile_count, t_cursor, t_misc;   // uint32 per (view, tile) x 2; misc words
  size_t t_inst;               // uint2 (depth bits, view*P + id) per instance, unordered inside a tile
  size_t t_inst_b;             // the heavy tiles' instances again, partitioned into depth buckets
  size_t t_chunks;             // uint4 sort items (view*T + tile, offset in the tile list, count, flags)
  size_t n_chunks;             // their upper bound: 3 R_cap / kChunk + V*T
  // heavy tiles (more than kChunk instances): depth partition scratch
  size_t n_heavy, n_hchunks;   // upper bounds: heavy tiles, their kChunk-sized pieces
  size_t t_dmm;                // uint2 per heavy tile: (max of ~depth bits, max of depth bits), zeroed
  size_t t_slab_count;         // uint32[256] per heavy tile: slab histogram, then the scatter's cursors, zeroed
  size_t t_slab_off;           // uint32[256] per heavy tile: slab offsets inside the tile list
  size_t t_heavy_flag;         // uint32 per heavy tile: 1 = not partitioned (plain chunks + merge)
  size_t t_heavy, t_heavy_id;  // heavy tile list; heavy index per (view, tile)
  size_t t_hchunks;            // uint2 (view*T + tile, chunk index) pieces of the heavy tiles
};

int compute_layout(const GhrDims &d, Layout *L);
void set_error(const char *fmt, ...);

// ---- kernel launchers (each returns cudaGetLastError()) ----
cudaError_t launch_preprocess(const GhrDims &d, const Layout &L, const Cameras &cam, const Gaussians &g,
                              float scale_modifier, uint32_t flags, char *state, char *temp, int32_t *radii,
                              cudaStream_t s);
cudaError_t launch_recolor_geom(const GhrDims &d, const Layout &L, const Layout &Lold, const Cameras &cam,
                                const Gaussians &g, const char *old_state, char *state, int32_t *radii, cudaStream_t s);
cudaError_t launch_reuse_binning(const GhrDims &d, const Layout &L, const Layout &Lold, const char *old_state,
                                 char *state, uint64_t seq, cudaStream_t s);
cudaError_t launch_tile_scan_schedule(const GhrDims &d, const Layout &L, char *state, char *temp, uint64_t seq,
                                      cudaStream_t s);
cudaError_t launch_duplicate(const GhrDims &d, const Layout &L, char *state, char *temp, cudaStream_t s);
cudaError_t launch_sort_gather(const GhrDims &d, const Layout &L, char *state, char *temp, uint64_t *dbg_keys,
                               uint32_t *dbg_plist, cudaStream_t s);
cudaError_t launch_blend_forward(const GhrDims &d, const Layout &L, const Cameras &cam, char *state,
                                 float *out_color, float *out_mask, cudaStream_t s);
cudaError_t launch_blend_backward(const GhrDims &d, const Layout &L, const Cameras &cam, const char *state,
                                  const float *dL_dout, const float *dL_dmask, float *acc, cudaStream_t s);
struct GradOut {
  int accumulate;
  float *dmeans3D, *dmeans2D, *dcolors, *dopacity, *dcov3D, *dsh, *dscales, *drots, *dconic;
};
cudaError_t launch_preprocess_backward(const GhrDims &d, const Layout &L, const Cameras &cam,
                                       const Gaussians &g, float scale_modifier, const char *state,
                                       const float *acc, const GradOut &go, cudaStream_t s);
cudaError_t launch_cameras_from_w2c(int V, const float *w2c, const float *K, int H, int W, float znear, float zfar,
                                    float *view, float *proj, float *campos, float *tanfov, cudaStream_t s);
cudaError_t launch_mark_visible(int P, const float *means3D, const float *view, uint8_t *present,
                                cudaStream_t s);

}  // namespace ghr
