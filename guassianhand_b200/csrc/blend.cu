// Per-tile alpha blending, forward and backward.
//
// Forward replaces upstream renderCUDA<3> forward; backward replaces
// renderCUDA<3> backward (K7, A.6).  One CTA per (tile, view), one thread per pixel, a warp covers
// an 8x4 pixel block.  The tile's instance list is a contiguous slab of 48-byte records (written
// by gather_ranges), streamed into shared memory by 1-D bulk async copies (cp.async.bulk ->
// UBLKCP) that complete on mbarriers, three stages deep, issued by one elected thread.  Inside a
// stage each warp first culls: lane l tests instance (base+l)'s conservative alpha>=1/255 box against
// the warp's 8x4 pixel block, a ballot compacts the survivors, and only those are blended.
// Backward: per-instance partial gradients are reduced across the warp with a reduce-scatter
// butterfly (16 shuffles for 9 values) and leave the SM as one RED.ADD per value per warp.
#include "ghr_internal.cuh"

namespace ghr {

namespace {

constexpr int kBatch = 256;   // instances per stage
constexpr int kStages = 3;
constexpr float kAlphaMin = 1.0f / 255.0f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

struct __align__(128) StageBuf {
  float4 rec[kStages][kBatch * 3];
  uint64_t bar[kStages];
};

// thread -> pixel inside the tile: warp w covers the 8x4 block (w&1, w>>1)
__device__ __forceinline__ void pixel_of_thread(int tid, int &lx, int &ly) {
  int w = tid >> 5, l = tid & 31;
  lx = ((w & 1) << 3) + (l & 7);
  ly = ((w >> 1) << 2) + (l >> 3);
}

__global__ void __launch_bounds__(256)
blend_forward_kernel(int H, int W, int gx, int T, Cameras cam, const uint2 *__restrict__ ranges,
                     const float4 *__restrict__ records, float *__restrict__ final_T,
                     uint32_t *__restrict__ n_contrib, uint32_t *__restrict__ tilemax,
                     float *__restrict__ out_color) {
  __shared__ StageBuf sb;
  const int tile = blockIdx.x, v = blockIdx.y;
  const int tid = threadIdx.x;
  int lx, ly;
  pixel_of_thread(tid, lx, ly);
  const int px = (tile % gx) * kTile + lx, py = (tile / gx) * kTile + ly;
  const bool inside = px < W && py < H;
  const size_t N = (size_t)H * W;
  const float pxf = (float)px, pyf = (float)py;
  const int lane = tid & 31;
  // pixel block of this warp (all lanes): x in [bx0,bx1], y in [by0,by1]
  const float bx0 = (float)((tile % gx) * kTile + (((tid >> 5) & 1) << 3)), bx1 = bx0 + 7.f;
  const float by0 = (float)((tile / gx) * kTile + ((tid >> 6) << 2)), by1 = by0 + 3.f;

  const uint2 range = ranges[(size_t)v * T + tile];
  const uint32_t n = range.y - range.x;
  const uint32_t rounds = (n + kBatch - 1) / kBatch;
  const float4 *src = records + 3 * (size_t)range.x;

  if (tid == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(&sb.bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t issued = 0;
  auto issue = [&](uint32_t r) {
    uint32_t cnt = min((uint32_t)kBatch, n - r * kBatch);
    uint32_t bytes = cnt * kRecBytes;
    int s = r % kStages;
    mbar_expect_tx(&sb.bar[s], bytes);
    bulk_g2s(&sb.rec[s][0], src + 3 * (size_t)r * kBatch, bytes, &sb.bar[s]);
  };
  if (tid == 0)
    for (; issued < rounds && issued < kStages; issued++) issue(issued);

  bool done = !inside;
  float Tr = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last = 0;
  uint32_t r = 0;
  for (; r < rounds; r++) {
    const int s = r % kStages;
    mbar_wait(&sb.bar[s], (r / kStages) & 1);
    const uint32_t cnt = min((uint32_t)kBatch, n - r * kBatch);
    const float4 *rec = &sb.rec[s][0];
    if (!__all_sync(0xFFFFFFFFu, done)) {
      for (uint32_t base = 0; base < cnt; base += 32) {
        // each lane tests ONE instance's cull box against this warp's 8x4 pixel block
        const uint32_t e = base + lane;
        bool hit = false;
        if (e < cnt) {
          const float2 c = *reinterpret_cast<const float2 *>(&rec[3 * e]);
          const float2 w = *reinterpret_cast<const float2 *>(&rec[3 * e + 1].z);
          hit = (c.x + w.x >= bx0) && (c.x - w.x <= bx1) && (c.y + w.y >= by0) && (c.y - w.y <= by1);
        }
        uint32_t mask = __ballot_sync(0xFFFFFFFFu, hit);
        while (mask) {
          const uint32_t j = base + (uint32_t)__ffs(mask) - 1u;
          mask &= mask - 1u;
          if (done) continue;
          float4 a = rec[3 * j], b = rec[3 * j + 1];
          float dx = fsub(a.x, pxf), dy = fsub(a.y, pyf);
          float q = ffma(fmul(a.z, dx), dx, fmul(fmul(b.x, dy), dy));
          float power = ffma(-0.5f, q, -fmul(fmul(a.w, dx), dy));
          if (power > 0.0f) continue;
          float alpha = fminf(0.99f, fmul(b.y, expf(power)));
          if (alpha < kAlphaMin) continue;
          float test_T = fmul(Tr, fsub(1.f, alpha));
          if (test_T < 0.0001f) {
            done = true;
            continue;
          }
          float4 c = rec[3 * j + 2];
          C0 = ffma(fmul(c.x, alpha), Tr, C0);
          C1 = ffma(fmul(c.y, alpha), Tr, C1);
          C2 = ffma(fmul(c.z, alpha), Tr, C2);
          Tr = test_T;
          last = r * kBatch + j + 1;
        }
        if (__all_sync(0xFFFFFFFFu, done)) break;
      }
    }
    int ndone = __syncthreads_count(done);
    if (ndone == 256) { r++; break; }
    if (tid == 0 && issued < rounds) { issue(issued); issued++; }
  }
  // never leave the CTA with bulk copies in flight into its shared memory
  if (tid == 0)
    for (uint32_t q = r; q < issued; q++) mbar_wait(&sb.bar[q % kStages], (q / kStages) & 1);

  if (inside) {
    const float *bg = cam.bg + (size_t)cam.bg_stride * v;
    size_t pix = (size_t)py * W + px;
    final_T[(size_t)v * N + pix] = Tr;
    n_contrib[(size_t)v * N + pix] = last;
    float *o = out_color + (size_t)v * 3 * N + pix;
    o[0] = ffma(Tr, bg[0], C0);
    o[N] = ffma(Tr, bg[1], C1);
    o[2 * N] = ffma(Tr, bg[2], C2);
  }
  uint32_t wmax = __reduce_max_sync(0xFFFFFFFFu, last);
  if ((tid & 31) == 0 && wmax) atomicMax(&tilemax[(size_t)v * T + tile], wmax);
}

// Reduce-scatter butterfly over the warp for 9 values held in v[0..8]; on return lane L with
// (L & 1) == 0 and slot(L) < 9 holds the warp total of value slot(L) in the return value,
// slot(L) = ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1).
__device__ __forceinline__ float warp_reduce_scatter9(const float (&v)[9], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  float a[8];
  // 16 -> 8 : lower half keeps v[0..7], upper half keeps v[8..15] (v[9..15] = 0)
#pragma unroll
  for (int k = 0; k < 8; k++) {
    float mine_hi = (k == 0) ? v[8] : 0.f;
    float send = b4 ? v[k] : mine_hi;
    float recv = __shfl_xor_sync(0xFFFFFFFFu, send, 16);
    a[k] = (b4 ? mine_hi : v[k]) + recv;
  }
  float c[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float send = b3 ? a[k] : a[k + 4];
    float recv = __shfl_xor_sync(0xFFFFFFFFu, send, 8);
    c[k] = (b3 ? a[k + 4] : a[k]) + recv;
  }
  float e[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    float send = b2 ? c[k] : c[k + 2];
    float recv = __shfl_xor_sync(0xFFFFFFFFu, send, 4);
    e[k] = (b2 ? c[k + 2] : c[k]) + recv;
  }
  float send = b1 ? e[0] : e[1];
  float recv = __shfl_xor_sync(0xFFFFFFFFu, send, 2);
  float f = (b1 ? e[1] : e[0]) + recv;
  f += __shfl_xor_sync(0xFFFFFFFFu, f, 1);
  return f;
}

__global__ void __launch_bounds__(256)
blend_backward_kernel(int H, int W, int gx, int T, int P, Cameras cam, const uint2 *__restrict__ ranges,
                      const float4 *__restrict__ records, const float *__restrict__ final_T,
                      const uint32_t *__restrict__ n_contrib, const uint32_t *__restrict__ tilemax,
                      const float *__restrict__ dL_dout, float *__restrict__ acc) {
  __shared__ StageBuf sb;
  const int tile = blockIdx.x, v = blockIdx.y;
  const uint32_t maxc = tilemax[(size_t)v * T + tile];
  if (maxc == 0) return;
  const int tid = threadIdx.x, lane = tid & 31;
  int lx, ly;
  pixel_of_thread(tid, lx, ly);
  const int px = (tile % gx) * kTile + lx, py = (tile / gx) * kTile + ly;
  const bool inside = px < W && py < H;
  const size_t N = (size_t)H * W;
  const float pxf = (float)px, pyf = (float)py;
  const float bx0 = (float)((tile % gx) * kTile + (((tid >> 5) & 1) << 3)), bx1 = bx0 + 7.f;
  const float by0 = (float)((tile / gx) * kTile + ((tid >> 6) << 2)), by1 = by0 + 3.f;

  const uint2 range = ranges[(size_t)v * T + tile];
  const uint32_t n = min(range.y - range.x, maxc);     // instances past the last contributor never matter
  const uint32_t rounds = (n + kBatch - 1) / kBatch;
  const float4 *src = records + 3 * (size_t)range.x;

  if (tid == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(&sb.bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  // rounds are consumed back to front: step k handles round (rounds-1-k)
  uint32_t issued = 0;
  auto issue = [&](uint32_t k) {
    uint32_t rr = rounds - 1 - k;
    uint32_t cnt = min((uint32_t)kBatch, n - rr * kBatch);
    uint32_t bytes = cnt * kRecBytes;
    int s = k % kStages;
    mbar_expect_tx(&sb.bar[s], bytes);
    bulk_g2s(&sb.rec[s][0], src + 3 * (size_t)rr * kBatch, bytes, &sb.bar[s]);
  };
  if (tid == 0)
    for (; issued < rounds && issued < kStages; issued++) issue(issued);

  float T_final = 0.f, dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f;
  uint32_t last = 0;
  if (inside) {
    size_t pix = (size_t)py * W + px;
    T_final = final_T[(size_t)v * N + pix];
    last = n_contrib[(size_t)v * N + pix];
    const float *g = dL_dout + (size_t)v * 3 * N + pix;
    dLp0 = g[0];
    dLp1 = g[N];
    dLp2 = g[2 * N];
  }
  const float *bg = cam.bg + (size_t)cam.bg_stride * v;
  const float bgdot = bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2;
  float Tr = T_final;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;

  // which slot of the butterfly this lane ends up owning
  const int slot = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  const bool owner = ((lane & 1) == 0) && slot < 9;
  float *accv = acc + (size_t)v * P * kAccStride + slot;

  for (uint32_t k = 0; k < rounds; k++) {
    const int s = k % kStages;
    mbar_wait(&sb.bar[s], (k / kStages) & 1);
    const uint32_t rr = rounds - 1 - k;
    const uint32_t cnt = min((uint32_t)kBatch, n - rr * kBatch);
    const float4 *rec = &sb.rec[s][0];
    const uint32_t wlast = __reduce_max_sync(0xFFFFFFFFu, last);
    for (int base = (int)((cnt - 1) & ~31u); base >= 0; base -= 32) {
      const uint32_t el = (uint32_t)base + lane;          // index inside the stage
      bool hit = false;
      if (el < cnt && rr * kBatch + el < wlast) {
        const float2 c = *reinterpret_cast<const float2 *>(&rec[3 * el]);
        const float2 w = *reinterpret_cast<const float2 *>(&rec[3 * el + 1].z);
        hit = (c.x + w.x >= bx0) && (c.x - w.x <= bx1) && (c.y + w.y >= by0) && (c.y - w.y <= by1);
      }
      uint32_t mask = __ballot_sync(0xFFFFFFFFu, hit);
      while (mask) {
        const int bit = 31 - __clz(mask);                 // back to front
        mask &= ~(1u << bit);
        const int j = base + bit;
        const uint32_t e = rr * kBatch + (uint32_t)j;     // position in the tile list
        float4 a = rec[3 * j], b = rec[3 * j + 1];
        float vals[9];
#pragma unroll
        for (int t = 0; t < 9; t++) vals[t] = 0.f;
        bool contrib = false;
        if (e < last) {
          float dx = fsub(a.x, pxf), dy = fsub(a.y, pyf);
          float q = ffma(fmul(a.z, dx), dx, fmul(fmul(b.x, dy), dy));
          float power = ffma(-0.5f, q, -fmul(fmul(a.w, dx), dy));
          if (power <= 0.0f) {
            float G = expf(power);
            float alpha = fminf(0.99f, fmul(b.y, G));
            if (alpha >= kAlphaMin) {
              contrib = true;
              float4 c = rec[3 * j + 2];
              float rc = __frcp_rn(1.f - alpha);
              Tr = Tr * rc;
              float wgt = alpha * Tr;
              acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0;
              acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1;
              acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2;
              lc0 = c.x; lc1 = c.y; lc2 = c.z;
              float dL_dalpha = (c.x - acc0) * dLp0 + (c.y - acc1) * dLp1 + (c.z - acc2) * dLp2;
              dL_dalpha *= Tr;
              last_alpha = alpha;
              dL_dalpha += (-T_final * rc) * bgdot;
              float wG = G * dL_dalpha;
              float m10 = wG * dx, m01 = wG * dy;
              vals[0] = wgt * dLp0; vals[1] = wgt * dLp1; vals[2] = wgt * dLp2;
              vals[3] = wG; vals[4] = m10; vals[5] = m01;
              vals[6] = m10 * dx; vals[7] = m10 * dy; vals[8] = m01 * dy;
            }
          }
        }
        if (!__any_sync(0xFFFFFFFFu, contrib)) continue;
        float tot = warp_reduce_scatter9(vals, lane);
        if (owner) {
          uint32_t id = __float_as_uint(rec[3 * j + 2].w);
          atomicAdd(accv + (size_t)id * kAccStride, tot);
        }
      }
    }
    __syncthreads();
    if (tid == 0 && issued < rounds) { issue(issued); issued++; }
  }
}

}  // namespace

cudaError_t launch_blend_forward(const GhrDims &d, const Layout &L, const Cameras &cam, char *state,
                                 float *out_color, cudaStream_t s) {
  if (L.T == 0 || d.V == 0) return cudaSuccess;
  dim3 grid(L.T, d.V), block(256);
  blend_forward_kernel<<<grid, block, 0, s>>>(d.H, d.W, L.gx, L.T, cam, (const uint2 *)(state + L.pub.off_ranges),
                                              (const float4 *)(state + L.pub.off_records),
                                              (float *)(state + L.pub.off_final_T),
                                              (uint32_t *)(state + L.pub.off_ncontrib),
                                              (uint32_t *)(state + L.pub.off_tilemax), out_color);
  return cudaGetLastError();
}

cudaError_t launch_blend_backward(const GhrDims &d, const Layout &L, const Cameras &cam, const char *state,
                                  const float *dL_dout, float *acc, cudaStream_t s) {
  if (L.T == 0 || d.V == 0) return cudaSuccess;
  dim3 grid(L.T, d.V), block(256);
  blend_backward_kernel<<<grid, block, 0, s>>>(
      d.H, d.W, L.gx, L.T, d.P, cam, (const uint2 *)(state + L.pub.off_ranges),
      (const float4 *)(state + L.pub.off_records), (const float *)(state + L.pub.off_final_T),
      (const uint32_t *)(state + L.pub.off_ncontrib), (const uint32_t *)(state + L.pub.off_tilemax), dL_dout, acc);
  return cudaGetLastError();
}

}  // namespace ghr
