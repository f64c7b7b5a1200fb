// Per-tile alpha blending, forward and backward.
//
// Forward replaces upstream renderCUDA<3> forward; backward replaces
// renderCUDA<3> backward (K7, A.6).  One CTA per (view, tile), taken heaviest-first from the tile
// schedule.  8 consumer warps (one thread per pixel, a warp covers an 8x4 pixel block) + 1 producer
// warp.  The tile's instance list is a contiguous slab of 48-byte records (written by the chunk
// sort / merge of binning.cu); the producer streams it into a ring of shared-memory stages with 1-D bulk async
// copies (cp.async.bulk -> UBLKCP) completing on "full" mbarriers; consumer warps release a stage on
// its "empty" mbarrier, so warps drift apart by up to kStages-1 stages instead of meeting at a CTA
// barrier every round.  Inside a stage each warp first culls: an instance's precomputed 8-bit mask says
// which 8x4 sub-blocks it can reach with alpha >= 1/255, a ballot compacts the survivors into a queue,
// and only those are blended.  Backward: one (tile, 128-instance segment) work unit per two 4-warp CTAs,
// restarted from the forward's checkpoints; per-instance partial gradients are parked in shared memory
// three instances at a time, reduced across the warp with conflict-free 128-bit loads and leave the SM
// as one RED.ADD per value per warp.
#include <cstdlib>

#include "ghr_internal.cuh"

namespace ghr {

namespace {

constexpr int kStageN = 128;  // instances per stage (6 KB)
constexpr int kStages = 4;
constexpr int kConsumerWarps = 8;
constexpr int kBlendThreads = (kConsumerWarps + 1) * 32;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr int kMaxIlpB = 2;
constexpr int kDirectMax = 4;    // <= this many contributing lanes: no warp reduction, direct REDs
constexpr int kQPad = 8;         // padding entries on both sides of a survivor queue

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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Blocking wait: try_wait with a suspend-time hint parks the warp in hardware; between retries the
// warp sleeps `backoff_ns` (warps that only keep the ring turning -- all their pixels terminated --
// pass a long backoff so their polling does not take issue slots from warps that still blend).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t backoff_ns = 32u) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (!ok) __nanosleep(backoff_ns);
  } while (!ok);
}

constexpr int kMaskBytes = kStageN + 16;   // a stage's mask bytes start at a 16-byte boundary at or before it
struct __align__(128) StageBuf {
  float4 rec[kStages][kStageN * 3 + 3];   // + one all-zero record (index kStageN): queue padding, alpha = 0
  uint8_t msk[kStages][kMaskBytes];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint32_t done_warps;
  uint32_t stop_round;
  uint32_t tmax;          // max n_contrib of the tile (forward)
};

// Producer side of one stage: the contiguous record slab plus the 16-byte-aligned window of the
// mask byte array that covers the same instances, both completing on the stage's full barrier.
__device__ __forceinline__ void stage_load(StageBuf &sb, int s, const float4 *records, const uint8_t *masks,
                                           size_t first, uint32_t cnt) {
  const size_t m0 = first & ~(size_t)15;
  const uint32_t mbytes = (uint32_t)(((first + cnt + 15) & ~(size_t)15) - m0);
  mbar_expect_tx(&sb.full[s], cnt * kRecBytes + mbytes);
  bulk_g2s(&sb.rec[s][0], records + 3 * first, cnt * kRecBytes, &sb.full[s]);
  bulk_g2s(&sb.msk[s][0], masks + m0, mbytes, &sb.full[s]);
}

__device__ __forceinline__ void stage_init(StageBuf &sb, int tid) {
  if (tid == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(&sb.full[s], 1);
      mbar_init(&sb.empty[s], kConsumerWarps);
    }
    sb.done_warps = 0;
    sb.stop_round = 0xFFFFFFFFu;
    sb.tmax = 0;
    mbar_fence_init();
  }
  if (tid < 3 * kStages) sb.rec[tid / 3][kStageN * 3 + tid % 3] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
}

// exp(x) for x <= 0 as one FMUL + MUFU.EX2 (ex2.approx: <= 2 ulp; argument rounding adds <= 3.3e-7
// relative at |x| <= 5.6, the largest exponent that can still pass alpha >= 1/255).  libdevice expf
// costs 9 more issue slots per pair in kernels that are issue-bound.  Forward and backward use the
// same function, so every alpha / skip decision of the backward replays the forward's exactly.
__device__ __forceinline__ float exp_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

// thread -> pixel inside the tile: warp w covers the 8x4 block (w&1, w>>1)
__device__ __forceinline__ void pixel_of_thread(int tid, int &lx, int &ly) {
  int w = tid >> 5, l = tid & 31;
  lx = ((w & 1) << 3) + (l & 7);
  ly = ((w >> 1) << 2) + (l >> 3);
}

// Survivor queue of one warp for one stage: the indices (inside the stage, ascending) of the
// instances whose sub-block mask (one byte per instance, written by gather_ranges) has this warp's
// bit set.  Each lane tests 4 instances (one byte LDS each), 4 ballots compact them; the blend loop then
// takes kIlp indices per iteration from one broadcast LDS instead of peeling bits off a mask.
// Survivor i is q[kQPad + i]; kQPad entries pad both ends (front: index 0, back: `pad` -- the forward
// passes kStageN, its all-zero record whose alpha is 0).  Returns the survivor count.
__device__ __forceinline__ uint32_t build_queue(const uint8_t *msk, uint32_t cnt, uint32_t limit, int warp, int lane,
                                                uint8_t *q, uint32_t pad) {
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t total = 0;
  __syncwarp();   // every lane has finished reading the previous stage's queue
#pragma unroll
  for (int w = 0; w < kStageN / 32; w++) {
    const uint32_t e = w * 32 + lane;
    bool hit = false;
    if (e < cnt && e < limit) hit = (msk[e] >> warp) & 1u;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
    if (hit) q[kQPad + total + __popc(m & lt)] = (uint8_t)e;
    total += __popc(m);
  }
  if (lane < kQPad) q[kQPad + total + lane] = (uint8_t)pad;
  __syncwarp();
  return total;
}

// Forward variant of the survivor queue: entries are the 16-bit shared-memory ADDRESSES of the
// survivors' records (the kernel's shared window is far below 64 KB), eight of them per LDS.128, so an
// instance costs one extract and no address arithmetic before its three record loads.
__device__ __forceinline__ uint32_t build_queue_addr(const uint8_t *msk, uint32_t cnt, int warp, int lane, uint16_t *q,
                                                     uint32_t rec_base, uint32_t pad_addr) {
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t total = 0;
  __syncwarp();   // every lane has finished reading the previous stage's queue
#pragma unroll
  for (int w = 0; w < kStageN / 32; w++) {
    const uint32_t e = w * 32 + lane;
    bool hit = false;
    if (e < cnt) hit = (msk[e] >> warp) & 1u;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
    if (hit) q[kQPad + total + __popc(m & lt)] = (uint16_t)(rec_base + e * kRecBytes);
    total += __popc(m);
  }
  if (lane < kQPad) q[kQPad + total + lane] = (uint16_t)pad_addr;
  __syncwarp();
  return total;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// kIlpF = instances blended per inner iteration (ILP): 4 or 8
template <int kIlpF>
__global__ void __launch_bounds__(kBlendThreads, kIlpF <= 4 ? 4 : 3)
blend_forward_kernel(int H, int W, int gx, int T, Cameras cam, const uint32_t *__restrict__ order,
                     const uint2 *__restrict__ ranges, const float4 *__restrict__ records,
                     const uint8_t *__restrict__ masks, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
                     uint32_t *__restrict__ tilemax, float4 *__restrict__ tilefinal, float4 *__restrict__ ckpt,
                     uint4 *__restrict__ units, GhrStatus *__restrict__ status, float *__restrict__ out_color,
                     float *__restrict__ out_mask, uint32_t bo_active, uint32_t bo_done, uint32_t bo_prod) {
  __shared__ StageBuf sb;
  __shared__ __align__(16) uint16_t s_q[kConsumerWarps][kStageN + 2 * kQPad];
  const uint32_t vt = order[blockIdx.x];
  const int v = vt / (uint32_t)T, tile = vt % (uint32_t)T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  const uint2 range = ranges[vt];
  const uint32_t n = range.y - range.x;
  if (n == 0) {
    // empty tile (most of the frame): background only, no barriers, no staging
    if (warp < kConsumerWarps) {
      int lx, ly;
      pixel_of_thread(tid, lx, ly);
      const int px = (tile % gx) * kTile + lx, py = (tile / gx) * kTile + ly;
      if (px < W && py < H) {
        const size_t N = (size_t)H * W, pix = (size_t)py * W + px;
        const float *bg = cam.bg + (size_t)cam.bg_stride * v;
        final_T[(size_t)v * N + pix] = 1.0f;
        n_contrib[(size_t)v * N + pix] = 0u;
        float *o = out_color + (size_t)v * 3 * N + pix;
        o[0] = ffma(1.0f, bg[0], 0.f);
        o[N] = ffma(1.0f, bg[1], 0.f);
        o[2 * N] = ffma(1.0f, bg[2], 0.f);
        if (out_mask) out_mask[(size_t)v * N + pix] = 0.f;
      }
    }
    return;
  }
  const uint32_t rounds = (n + kStageN - 1) / kStageN;
  stage_init(sb, tid);

  if (warp == kConsumerWarps) {
    // ---------------- producer ----------------
    if (lane == 0) {
      for (uint32_t r = 0; r < rounds; r++) {
        const int s = r % kStages;
        if (r >= kStages) mbar_wait(&sb.empty[s], ((r / kStages) - 1) & 1, bo_prod);
        if (*(volatile uint32_t *)&sb.done_warps == kConsumerWarps) {
          // every pixel of the tile has terminated: complete the phase without data ("poison")
          *(volatile uint32_t *)&sb.stop_round = r;
          mbar_arrive(&sb.full[s]);
          break;
        }
        const uint32_t cnt = min((uint32_t)kStageN, n - r * kStageN);
        stage_load(sb, s, records, masks, (size_t)range.x + (size_t)r * kStageN, cnt);
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  int lx, ly;
  pixel_of_thread(tid, lx, ly);
  const int px = (tile % gx) * kTile + lx, py = (tile / gx) * kTile + ly;
  const bool inside = px < W && py < H;
  const size_t N = (size_t)H * W;
  const float pxf = (float)px, pyf = (float)py;
  uint16_t *q = &s_q[warp][0];

  // Tw: transmittance while the pixel is live, 0 once it has terminated (or lies outside the image) --
  // then test_T = 0 keeps `term` set and every later weight is 0; Tr: the last live value (the output)
  float Tw = inside ? 1.0f : 0.f, Tr = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  bool wdone = __all_sync(0xFFFFFFFFu, !inside);
  if (wdone && lane == 0) atomicAdd(&sb.done_warps, 1u);
  uint32_t last = 0;
  for (uint32_t r = 0; r < rounds; r++) {
    const int s = r % kStages;
    // a consumer that waits here is ahead of the tile's slowest warp by the whole ring: it can sleep long
    // between polls (its wake-up is not on the critical path), the producer polls `empty` tightly
    mbar_wait(&sb.full[s], (r / kStages) & 1, wdone ? bo_done : bo_active);
    if (r >= *(volatile uint32_t *)&sb.stop_round) break;
    if (!wdone) {
      // running state at every kSeg-instance boundary: the backward restarts from it (one work unit per
      // segment).  A warp whose pixels have all terminated writes nothing: no later unit reads it.
      if (r > 0 && (r * kStageN) % kSeg == 0)
        ckpt[((size_t)(range.x / kSeg) + vt + (r * kStageN) / kSeg) * 256 + tid] = make_float4(Tr, C0, C1, C2);
      const uint32_t cnt = min((uint32_t)kStageN, n - r * kStageN);
      const uint32_t rec_base = smem_u32(&sb.rec[s][0]);
      const uint32_t total = build_queue_addr(&sb.msk[s][(range.x + r * kStageN) & 15u], cnt, warp, lane, q, rec_base,
                                              rec_base + kStageN * kRecBytes);
      uint32_t lastq = 0;                                     // 1 + queue index of the last blended survivor
      for (uint32_t b = 0; b < total; b += kIlpF) {
        // kIlpF survivors at a time: their alphas do not depend on the running transmittance, so
        // the long chains (LDS -> quadratic form -> exp) of several instances overlap; only the
        // short T / colour update is serial (and branch-free: a rejected pair blends alpha = 0).
        float al[kIlpF];
        float4 col[kIlpF];
        uint32_t packed[kIlpF / 2];
        if constexpr (kIlpF == 8) {
          const uint4 p4 = *reinterpret_cast<const uint4 *>(q + kQPad + b);
          packed[0] = p4.x; packed[1] = p4.y; packed[2] = p4.z; packed[3] = p4.w;
        } else {
          const uint2 p2 = *reinterpret_cast<const uint2 *>(q + kQPad + b);
          packed[0] = p2.x; packed[1] = p2.y;
        }
#pragma unroll
        for (int k = 0; k < kIlpF; k++) {
          const uint32_t addr = (k & 1) ? packed[k >> 1] >> 16 : packed[k >> 1] & 0xFFFFu;
          const float4 a = lds128(addr), bq = lds128(addr + 16);
          col[k] = lds128(addr + 32);
          const float dx = fsub(a.x, pxf), dy = fsub(a.y, pyf);
          const float qf = ffma(fmul(a.z, dx), dx, fmul(fmul(bq.x, dy), dy));
          const float power = ffma(-0.5f, qf, -fmul(fmul(a.w, dx), dy));
          const float alpha = fminf(0.99f, fmul(bq.y, exp_fast(power)));
          al[k] = (power <= 0.0f && alpha >= kAlphaMin) ? alpha : 0.f;   // padding: opacity 0 -> alpha 0
        }
#pragma unroll
        for (int k = 0; k < kIlpF; k++) {
          const float test_T = fmul(Tw, fsub(1.f, al[k]));      // alpha == 0: test_T == Tw exactly
          // a live Tw is >= 1e-4 (the stopping Gaussian is never applied), so a live pixel terminates only
          // on alpha != 0; a terminated one (Tw == 0) stays terminated
          const bool term = test_T < 0.0001f;
          const float Tm = term ? 0.f : Tw;                     // the stopping Gaussian is not blended
          C0 = ffma(fmul(col[k].x, al[k]), Tm, C0);             // upstream's order: (c * alpha) * T
          C1 = ffma(fmul(col[k].y, al[k]), Tm, C1);
          C2 = ffma(fmul(col[k].z, al[k]), Tm, C2);
          Tw = term ? 0.f : test_T;
          Tr = term ? Tr : test_T;
          lastq = (!term && al[k] != 0.f) ? b + k + 1 : lastq;
        }
        if (__all_sync(0xFFFFFFFFu, Tw == 0.f)) {
          wdone = true;
          if (lane == 0) atomicAdd(&sb.done_warps, 1u);
          break;
        }
      }
      if (lastq) last = r * kStageN + ((uint32_t)q[kQPad + lastq - 1] - rec_base) / kRecBytes + 1;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sb.empty[s]);
  }

  if (inside) {
    const float *bg = cam.bg + (size_t)cam.bg_stride * v;
    size_t pix = (size_t)py * W + px;
    final_T[(size_t)v * N + pix] = Tr;
    n_contrib[(size_t)v * N + pix] = last;
    float *o = out_color + (size_t)v * 3 * N + pix;
    o[0] = ffma(Tr, bg[0], C0);
    o[N] = ffma(Tr, bg[1], C1);
    o[2 * N] = ffma(Tr, bg[2], C2);
    if (out_mask) out_mask[(size_t)v * N + pix] = 1.0f - Tr;   // = sum_j alpha_j T_j
  }
  tilefinal[(size_t)vt * 256 + tid] = make_float4(C0, C1, C2, Tr);
  // backward work units of this tile: one per kSeg instances up to the tile's last contributor
  const uint32_t wmax = __reduce_max_sync(0xFFFFFFFFu, last);
  if (lane == 0 && wmax) atomicMax(&sb.tmax, wmax);
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
  if (warp == 0) {
    const uint32_t tmax = sb.tmax, nseg = (tmax + kSeg - 1) / kSeg;
    if (nseg) {
      uint32_t base = 0;
      if (lane == 0) {
        tilemax[vt] = tmax;
        base = (uint32_t)atomicAdd((unsigned long long *)&status->reserved[1], (unsigned long long)nseg);
      }
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      for (uint32_t i = lane; i < nseg; i += 32) units[base + i] = make_uint4(vt, i, range.x, tmax);
    }
  }
}

// Warp reduction of 9 values per lane through shared memory: lane L stores its 9 partials as row L
// of a 32x9 tile (stride 9 is odd: conflict-free), then lane (g,c) = (L/9, L%9), L < 27, sums column
// c over 12 (8 for g=2) rows and the three row groups are combined with two shuffles.  36 issue slots
// per instance instead of 62 for a register reduce-scatter butterfly (16 SHFL + 16 FADD + 30 FSEL) or
// 90 for five plain xor steps.  On return lanes 0..8 hold the warp totals of values 0..8.
__device__ __forceinline__ void warp_store9(float *buf, const float (&v)[9], int lane) {
#pragma unroll
  for (int t = 0; t < 9; t++) buf[lane * 9 + t] = v[t];
}
__device__ __forceinline__ float warp_colsum9(const float *buf, int lane) {
  const int g = lane / 9, c = lane - 9 * g;
  float sum = 0.f;
  if (lane < 27) {
    const float *col = buf + (g * 12) * 9 + c;
#pragma unroll
    for (int i = 0; i < 8; i++) sum += col[i * 9];
    if (g < 2) {
#pragma unroll
      for (int i = 8; i < 12; i++) sum += col[i * 9];
    }
  }
  const float s1 = __shfl_down_sync(0xFFFFFFFFu, sum, 9);
  const float s2 = __shfl_down_sync(0xFFFFFFFFu, sum, 18);
  return sum + s1 + s2;
}

// Ring reduction (kRing): instead of reducing every instance on its own, a warp parks the 9 partials
// of up to kRingSlots instances as rows of 32 floats (row stride 36 floats: lane l writes column l,
// conflict-free) and reduces the slots together: the 27 rows are cut into 54 half-rows, lane l sums
// half-row l (then 32 + l) with four LDS.128 + 15 FADD, one xor-shuffle joins the halves, and the even
// lanes send the totals as REDs -- two rounds for three instances (~16 issue slots per instance
// instead of ~40 for the per-instance column sum).  Eight consecutive half-rows start in eight
// different 4-bank groups, so the 128-bit loads are conflict-free as well.
constexpr int kRingSlots = 3;
constexpr int kRowStride = 36;
constexpr int kSlotFloats = 9 * kRowStride;
__device__ __forceinline__ void ring_flush(const float *ring, const uint32_t *ids, uint32_t nslots, float *accb,
                                           int lane) {
  const uint32_t ntask = nslots * 18u;
#pragma unroll
  for (int round = 0; round < 2; round++) {
    if ((uint32_t)round * 32u >= ntask) break;          // warp-uniform
    const uint32_t t = round * 32 + lane, r = t >> 1;
    float sum = 0.f;
    if (t < ntask) {
      const float4 *p = reinterpret_cast<const float4 *>(ring + r * kRowStride + (t & 1u) * 16u);
      const float4 a = p[0], b = p[1], c = p[2], d = p[3];
      sum = (((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) +
            (((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w)));
    }
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, 1);
    if (t < ntask && !(t & 1u)) {
      const uint32_t slot = (r * 57u) >> 9, val = r - 9u * slot;      // r / 9 for r < 27
      atomicAdd(accb + (size_t)ids[slot] * kAccStride + val, sum);
    }
  }
}

// Backward blend: one CTA per work unit = (view, tile, segment of kSeg instances), emitted by the
// forward.  Each unit restarts the per-pixel recursion from the forward's checkpoint at the segment
// start and walks the segment FRONT TO BACK, so units of one tile are independent: the launch has
// R/kSeg uniform units instead of one CTA per tile (no heavy-tile tail, and a single view fills the
// GPU).  With T_j the transmittance in front of contributor j, w_j = alpha_j T_j and
//   D_j = sum_{k>j} (c_k . dL/dpix) w_k          (colour still to come behind j, from C_total - C_prefix)
// upstream's  dL/dalpha_j = T_j (c_j - A_j).dL/dpix - T_final/(1-alpha_j) bg.dL/dpix  (A_j = suffix colour
// normalised by T_{j+1}) becomes  T_j (c_j.dL/dpix) - (D_j + T_final bg.dL/dpix) / (1-alpha_j):  two scalar
// recurrences (T, D) instead of the back-to-front vector one, and T replays the forward's products exactly.
// kW warps per CTA: 8 (the whole tile) or 4 (half a tile: two CTAs per unit, each staging the slab).
// Warps of a CTA finish at very different times (a sub-block outside the hands has few survivors) and the
// CTA keeps its registers and shared memory until the slowest one is done; smaller CTAs give those
// resources back sooner.
template <int kIlpB, bool kRing, int kMinCtas, int kW>
__global__ void __launch_bounds__(kW * 32, kMinCtas)
blend_backward_kernel(int H, int W, int gx, int T, int P, Cameras cam, const GhrStatus *__restrict__ status,
                      const uint4 *__restrict__ units, const float4 *__restrict__ records,
                      const uint8_t *__restrict__ masks, const float4 *__restrict__ tilefinal,
                      const float4 *__restrict__ ckpt, const uint32_t *__restrict__ n_contrib,
                      const float *__restrict__ dL_dout, const float *__restrict__ dL_dmask,
                      float *__restrict__ acc, int direct_max, int reverse) {
  constexpr int kRedBufs = kIlpB < kMaxIlpB ? kIlpB : kMaxIlpB;
  __shared__ __align__(128) float4 s_rec[kSeg * 3];
  __shared__ __align__(16) uint8_t s_msk[kSeg + 16];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ __align__(16) float s_red[kW][kRing ? kRingSlots * kSlotFloats : kRedBufs * 32 * 9];
  __shared__ uint32_t s_ids[kW][4];
  __shared__ __align__(16) uint8_t s_q[kW][kStageN + 2 * kQPad];
  constexpr uint32_t kParts = kConsumerWarps / kW;       // CTAs per unit
  const uint32_t bid = blockIdx.x / kParts, part = blockIdx.x % kParts;
  // The unit record {view*T + tile, segment, start of the tile's slab, instances up to the tile's last
  // contributor} carries everything the copy needs, and it is read together with the unit count (the
  // list is allocated to its upper bound): one memory round trip between CTA start and the bulk copy.
  // `reverse`: the forward appends a tile's units when the tile completes, so the heaviest tiles sit at
  // the end of the list; walking it backwards starts their (long, dense) units first and leaves the
  // short ones to fill the tail of the launch.
  const uint32_t n_units = (uint32_t)status->reserved[1];
  uint32_t uidx = bid;
  if (reverse) {
    if (bid >= n_units) return;
    uidx = n_units - 1u - bid;
  }
  const uint4 unit = units[uidx];
  if (bid >= n_units) return;
  const uint32_t vt = unit.x, first = unit.y * kSeg;      // first = position of the segment in the tile list
  const int v = vt / (uint32_t)T, tile = vt % (uint32_t)T;
  const int lane = threadIdx.x & 31, wloc = threadIdx.x >> 5;      // warp inside the CTA
  const int warp = (int)part * kW + wloc, tid = warp * 32 + lane;   // warp / thread inside the tile
  uint8_t *q = &s_q[wloc][0];
  const uint32_t cnt = min((uint32_t)kSeg, unit.w - first);   // instances past the last contributor never matter
  const size_t g0 = (size_t)unit.z + first;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    const size_t m0 = g0 & ~(size_t)15;
    const uint32_t mbytes = (uint32_t)(((g0 + cnt + 15) & ~(size_t)15) - m0);
    mbar_expect_tx(&s_bar, cnt * kRecBytes + mbytes);
    bulk_g2s(s_rec, records + 3 * g0, cnt * kRecBytes, &s_bar);
    bulk_g2s(s_msk, masks + m0, mbytes, &s_bar);
  }
  if (lane < kQPad) q[lane] = 0;

  // per-pixel state while the segment is in flight (all loads independent of each other)
  int lx, ly;
  pixel_of_thread(tid, lx, ly);
  const int px = (tile % gx) * kTile + lx, py = (tile / gx) * kTile + ly;
  const bool inside = px < W && py < H;
  const size_t N = (size_t)H * W;
  const float pxf = (float)px, pyf = (float)py;
  float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f, dLm = 0.f;
  uint32_t last = 0;
  const float4 fin = tilefinal[(size_t)vt * 256 + tid];
  float4 c = make_float4(1.f, 0.f, 0.f, 0.f);
  // (a checkpoint no later unit needs was never written: whatever is read there is not used)
  if (first) c = ckpt[((size_t)(unit.z / kSeg) + vt + unit.y) * 256 + tid];
  if (inside) {
    const size_t pix = (size_t)py * W + px;
    last = n_contrib[(size_t)v * N + pix];
    const float *g = dL_dout + (size_t)v * 3 * N + pix;
    dLp0 = g[0];
    dLp1 = g[N];
    dLp2 = g[2 * N];
    if (dL_dmask) dLm = dL_dmask[(size_t)v * N + pix];
  }
  float Tr = 0.f, Drem = 0.f, Tb = 0.f;
  if (last > first) {
    Tr = c.x;
    Drem = (fin.x - c.y) * dLp0 + (fin.y - c.z) * dLp1 + (fin.z - c.w) * dLp2;
    const float *bg = cam.bg + (size_t)cam.bg_stride * v;
    // coverage output m = 1 - T_final: dm/dalpha_j = +T_final/(1-alpha_j), the background term with
    // the opposite sign, so its gradient folds into the same product
    Tb = fin.w * (bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2 - dLm);
  }
  const bool owner = lane < 9;      // lanes 0..8 own the warp totals of the 9 accumulated values
  float *accb = acc + (size_t)v * P * kAccStride;
  const uint32_t wlast = __reduce_max_sync(0xFFFFFFFFu, last);
  __syncthreads();                  // barrier initialised before anyone polls it
  if (wlast <= first) return;
  mbar_wait(&s_bar, 0);
  float *ring = &s_red[wloc][0];
  uint32_t *ids = &s_ids[wloc][0];
  uint32_t pend = 0;                // instances parked in the ring (warp-uniform)
  const float4 *s_rec_cur = s_rec;
  const uint8_t *s_msk_cur = s_msk;
  const uint32_t g0lo = (uint32_t)g0 & 15u;

  for (uint32_t sub = 0; sub * kStageN < cnt; sub++) {
    const uint32_t pos0 = first + sub * kStageN;
    if (pos0 >= wlast) break;
    const uint32_t scnt = min((uint32_t)kStageN, cnt - sub * kStageN);
    const float4 *rec = s_rec_cur + 3 * sub * kStageN;
    // survivors of this warp's sub-block among the instances that precede the warp's last contributor
    const uint32_t total = build_queue(&s_msk_cur[g0lo + sub * kStageN], scnt, wlast - pos0, warp, lane, q, 0u);
    for (uint32_t b = 0; b < total; b += kIlpB) {
      // phase 1 (independent per instance): alpha, G, 1/(1-alpha), offsets, colour . dL/dpix
      float al[kIlpB], Gk[kIlpB], omk[kIlpB], rck[kIlpB], dxk[kIlpB], dyk[kIlpB], cdk[kIlpB];
      uint32_t idk[kIlpB];
#pragma unroll
      for (int k = 0; k < kIlpB; k++) {
        const uint32_t jj = q[kQPad + b + k];                // tail pad (index 0) when b + k >= total
        const float4 a = rec[3 * jj], bq = rec[3 * jj + 1], col = rec[3 * jj + 2];
        const float dx = fsub(a.x, pxf), dy = fsub(a.y, pyf);
        const float qf = ffma(fmul(a.z, dx), dx, fmul(fmul(bq.x, dy), dy));
        const float power = ffma(-0.5f, qf, -fmul(fmul(a.w, dx), dy));
        const float G = exp_fast(power);
        const float alpha = fminf(0.99f, fmul(bq.y, G));
        const bool c = b + k < total && pos0 + jj < last && power <= 0.0f && alpha >= kAlphaMin;
        al[k] = c ? alpha : 0.f;
        Gk[k] = c ? G : 0.f;
        omk[k] = fsub(1.f, al[k]);
        rck[k] = rcp_fast(omk[k]);
        dxk[k] = dx;
        dyk[k] = dy;
        cdk[k] = col.x * dLp0 + col.y * dLp1 + col.z * dLp2;
        idk[k] = __float_as_uint(col.w);
      }
      // phase 2 (serial, short, branch-free: a rejected pair has alpha = G = 0 and adds zeros)
      float vals[kIlpB][9];
      bool contrib[kIlpB];
#pragma unroll
      for (int k = 0; k < kIlpB; k++) {
        contrib[k] = al[k] != 0.f;
        const float wgt = al[k] * Tr;
        Drem = fmaf(-cdk[k], wgt, Drem);
        const float dL_dalpha = fmaf(Tr, cdk[k], -(Drem + Tb) * rck[k]);
        Tr = fmul(Tr, omk[k]);                                 // the forward's own product
        const float wG = Gk[k] * dL_dalpha;
        const float m10 = wG * dxk[k], m01 = wG * dyk[k];
        vals[k][0] = wgt * dLp0; vals[k][1] = wgt * dLp1; vals[k][2] = wgt * dLp2;
        vals[k][3] = wG; vals[k][4] = m10; vals[k][5] = m01;
        vals[k][6] = m10 * dxk[k]; vals[k][7] = m10 * dyk[k]; vals[k][8] = m01 * dyk[k];
      }
      // phase 3: slots with few contributing lanes send their partials straight to L2 (9 REDs for the
      // warp); the others are reduced through shared memory first and leave as one RED per value
      if constexpr (kRing) {
#pragma unroll
        for (int k = 0; k < kIlpB; k++) {
          const uint32_t cm = __ballot_sync(0xFFFFFFFFu, contrib[k]);
          if (cm == 0u) continue;
          if (__popc(cm) <= direct_max) {
            if (contrib[k]) {
              float *dst = accb + (size_t)idk[k] * kAccStride;
#pragma unroll
              for (int t = 0; t < 9; t++) atomicAdd(dst + t, vals[k][t]);
            }
          } else {
            float *row = ring + pend * kSlotFloats + lane;
#pragma unroll
            for (int t = 0; t < 9; t++) row[t * kRowStride] = vals[k][t];
            if (lane == 0) ids[pend] = idk[k];
            if (++pend == kRingSlots) {
              __syncwarp();
              ring_flush(ring, ids, kRingSlots, accb, lane);
              __syncwarp();
              pend = 0;
            }
          }
        }
      } else {
#pragma unroll
        for (int k0 = 0; k0 < kIlpB; k0 += kRedBufs) {
          int mode[kRedBufs];     // 0 nothing, 1 direct, 2 reduce
#pragma unroll
          for (int qi = 0; qi < kRedBufs; qi++) {
            const int k = k0 + qi < kIlpB ? k0 + qi : 0;
            const uint32_t cm = k0 + qi < kIlpB ? __ballot_sync(0xFFFFFFFFu, contrib[k]) : 0u;
            mode[qi] = cm == 0u ? 0 : (__popc(cm) <= direct_max ? 1 : 2);
            if (mode[qi] == 2) warp_store9(&s_red[wloc][qi * 32 * 9], vals[k], lane);
            if (mode[qi] == 1 && contrib[k]) {
              float *dst = accb + (size_t)idk[k] * kAccStride;
#pragma unroll
              for (int t = 0; t < 9; t++) atomicAdd(dst + t, vals[k][t]);
            }
          }
          __syncwarp();
#pragma unroll
          for (int qi = 0; qi < kRedBufs; qi++) {
            const int k = k0 + qi < kIlpB ? k0 + qi : 0;
            if (mode[qi] != 2) continue;
            float tot = warp_colsum9(&s_red[wloc][qi * 32 * 9], lane);
            if (owner) atomicAdd(accb + (size_t)idk[k] * kAccStride + lane, tot);
          }
          __syncwarp();
        }
      }
    }
  }
  if constexpr (kRing) {
    if (pend) {
      __syncwarp();
      ring_flush(ring, ids, pend, accb, lane);
    }
  }
}

int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

cudaError_t launch_blend_forward(const GhrDims &d, const Layout &L, const Cameras &cam, char *state,
                                 float *out_color, float *out_mask, cudaStream_t s) {
  if (L.T == 0 || d.V == 0) return cudaSuccess;
  dim3 grid(L.T * d.V), block(kBlendThreads);
  static const int ilp = env_int("GHR_ILPF", 8);
  static const int bo_active = env_int("GHR_BO_ACTIVE", 256), bo_done = env_int("GHR_BO_DONE", 1024);
  static const int bo_prod = env_int("GHR_BO_PROD", 256);
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<grid, block, 0, s>>>(d.H, d.W, L.gx, L.T, cam, (const uint32_t *)(state + L.pub.off_order),
                                (const uint2 *)(state + L.pub.off_ranges),
                                (const float4 *)(state + L.pub.off_records),
                                (const uint8_t *)(state + L.pub.off_masks), (float *)(state + L.pub.off_final_T),
                                (uint32_t *)(state + L.pub.off_ncontrib), (uint32_t *)(state + L.pub.off_tilemax),
                                (float4 *)(state + L.pub.off_tilefinal), (float4 *)(state + L.pub.off_ckpt),
                                (uint4 *)(state + L.pub.off_units), (GhrStatus *)(state + L.pub.off_status),
                                out_color, out_mask, (uint32_t)bo_active, (uint32_t)bo_done, (uint32_t)bo_prod);
  };
  if (ilp <= 4) launch(blend_forward_kernel<4>);
  else launch(blend_forward_kernel<8>);
  return cudaGetLastError();
}

cudaError_t launch_blend_backward(const GhrDims &d, const Layout &L, const Cameras &cam, const char *state,
                                  const float *dL_dout, const float *dL_dmask, float *acc, cudaStream_t s) {
  if (L.T == 0 || d.V == 0 || d.R_cap <= 0) return cudaSuccess;
  // upper bound of the unit count (the forward wrote the exact one to GhrStatus.reserved[1]); surplus
  // CTAs exit on their first instructions
  static const int ilp = env_int("GHR_ILPB", 2);
  static const int direct = env_int("GHR_DIRECT", kDirectMax);
  static const int ring = env_int("GHR_RING", 1);
  static const int reverse = env_int("GHR_BWD_REV", 1);
  static const int warps = env_int("GHR_BWD_WARPS", 4);
  auto launch = [&](auto kern, int kw) {
    dim3 grid((unsigned)((L.n_slots - 1) * (kConsumerWarps / kw))), block(kw * 32);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<grid, block, 0, s>>>(d.H, d.W, L.gx, L.T, d.P, cam, (const GhrStatus *)(state + L.pub.off_status),
                                (const uint4 *)(state + L.pub.off_units),
                                (const float4 *)(state + L.pub.off_records),
                                (const uint8_t *)(state + L.pub.off_masks),
                                (const float4 *)(state + L.pub.off_tilefinal),
                                (const float4 *)(state + L.pub.off_ckpt),
                                (const uint32_t *)(state + L.pub.off_ncontrib), dL_dout, dL_dmask, acc, direct,
                                reverse);
  };
  if (warps <= 4) {
    if (!ring) launch(blend_backward_kernel<2, false, 8, 4>, 4);
    else launch(blend_backward_kernel<2, true, 8, 4>, 4);
  } else if (ring) {
    if (ilp <= 2) launch(blend_backward_kernel<2, true, 4, 8>, 8);
    else launch(blend_backward_kernel<3, true, 4, 8>, 8);
  } else {
    if (ilp <= 1) launch(blend_backward_kernel<1, false, 4, 8>, 8);
    else if (ilp <= 2) launch(blend_backward_kernel<2, false, 4, 8>, 8);
    else launch(blend_backward_kernel<3, false, 4, 8>, 8);
  }
  return cudaGetLastError();
}

}  // namespace ghr
