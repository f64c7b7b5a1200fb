// Per-tile alpha blending, forward and backward.
//
// Forward replaces upstream renderCUDA<3> forward; backward replaces
// renderCUDA<3> backward (K7, A.6).
//
// Both kernels are bound by instruction issue, not by FP32 throughput or memory, so both evaluate the
// per-(pixel, instance) math as packed fp32 pairs (fma/mul/add.rn.f32x2 -> FFMA2/FMUL2/FADD2; IEEE-rn per
// half: the canonical bits are those of the scalar formulation) -- but they pair different things:
//
// Forward: one CTA (8 warps) per (view, tile), taken heaviest-first from the tile schedule; thread = one
// pixel, warp = one 8x4 sub-block, the pair = TWO INSTANCES (see the forward section for why).  The tile's
// instance list is a contiguous slab of 48-byte records (written by the chunk sort / merge of
// binning.cu); it is streamed through a ring of shared-memory stages with 1-D bulk async copies
// (cp.async.bulk -> UBLKCP) completing on "full" mbarriers.  There is no producer warp: the LAST warp to
// release a stage refills it (one shared-memory atomic per warp and stage), so warps drift apart by up to
// kStages-1 stages instead of meeting at a CTA barrier every round, and no warp slot or register budget
// is spent on polling.  Inside a stage each warp first culls: an instance's precomputed 8-bit mask says
// which 8x4 sub-blocks it can reach with alpha >= 1/255; only the survivors are blended.
// Backward: one (tile, 128-instance segment) work unit per CTA, restarted from the forward's checkpoints;
// thread = TWO PIXELS of one column, 4 rows apart (a warp covers an 8x8 block, four warps the tile): the
// two pixels share the instance's record loads, its queue entry and dx, and the thread sums its two
// pixels' partial gradients before the warp reduction.  Per-instance partials are parked in shared
// memory three instances at a time, reduced across the warp with conflict-free 128-bit loads and leave
// the SM as one RED.ADD per value per warp.
#include "ghr_internal.cuh"

namespace ghr {

namespace {

constexpr int kStageN = 128;  // instances per stage (6 KB)
#ifndef GHR_FWD_STAGES
#define GHR_FWD_STAGES 4
#endif
constexpr int kStages = GHR_FWD_STAGES;      // forward stage ring (build-time A/B)
constexpr int kBlendWarps = 4;               // backward: 8x8-pixel blocks of a tile
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kLog2e = 1.4426950408889634f;
#ifndef GHR_BWD_ILP
#define GHR_BWD_ILP 2
#endif
#ifndef GHR_BWD_MINCTAS
#define GHR_BWD_MINCTAS 7       // 72 registers (A/B: 6 CTAs at 76 registers 191 us, 7 at 72: 188.5, 8 at 64: 197.6)
#endif
constexpr int kIlpB = GHR_BWD_ILP;   // instances per backward iteration (build-time A/B)
#ifndef GHR_BWD_DIRECTMAX
#define GHR_BWD_DIRECTMAX 4
#endif
constexpr int kDirectMax = GHR_BWD_DIRECTMAX;    // <= this many contributing lanes: no warp reduction, direct REDs
constexpr int kQPad = 8;         // padding entries on both sides of a survivor queue
#ifndef GHR_BWD_WARPS
#define GHR_BWD_WARPS 4          // build-time A/B: 4 = one CTA per unit, 2 = two half-tile CTAs per unit
#endif
// Backward survivor queues per HALF-warp (one 8x4 sub-block each) instead of per warp (an 8x8 block): 12 % fewer
// warp iterations on the two-hand scene (68.9M -> 60.5M evaluated pairs), kernel 189 -> 183.5 us.
// -DGHR_BWD_NO_HALFQ builds the per-warp queues (A/B); the two-CTA variant keeps them.
#if !defined(GHR_BWD_NO_HALFQ) && !defined(GHR_BWD_HALFQ) && GHR_BWD_WARPS == 4
#define GHR_BWD_HALFQ 1
#endif
static_assert(kStageN == kSeg, "a forward stage is one backward segment");

#ifdef GHR_COUNT
// counting variant only (libghr_count.so, bench.py cull_efficiency): (pixel, instance) pairs the blend kernels
// evaluate after culling and pairs that actually contribute -- [0] fwd evaluated, [1] fwd contributing,
// [2] bwd evaluated, [3] bwd contributing
__device__ unsigned long long g_counts[4];
__device__ __forceinline__ void count_add(int i, uint32_t v) {
  v = __reduce_add_sync(0xFFFFFFFFu, v);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(&g_counts[i], (unsigned long long)v);
}
#endif

#ifdef GHR_TIMELINE
// profiling variant only (libghr_timeline.so): {start ns, stop ns, smid, work} per CTA of the blend kernels
constexpr uint32_t kTimelineCap = 1u << 17;
__device__ ulonglong4 g_timeline[kTimelineCap];
__device__ unsigned int g_timeline_n;
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void timeline_put(unsigned long long t0, uint32_t kind, uint32_t a, uint32_t b) {
  uint32_t smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  const uint32_t i = atomicAdd(&g_timeline_n, 1u);
  if (i < kTimelineCap)
    g_timeline[i] = make_ulonglong4(t0, gtime_ns(), ((unsigned long long)kind << 32) | smid,
                                    ((unsigned long long)a << 32) | b);
}
#endif

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
// Release counter of a ring stage: every warp's add RELEASES its generic-proxy reads of the stage, the add that
// completes the count ACQUIRES them, and the refilling lane then orders the async-proxy write behind them with
// fence.proxy.async -- the WAR edge between a warp's LDS of a stage and the next bulk copy into it.
__device__ __forceinline__ uint32_t atom_add_acq_rel_cta(uint32_t *p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Blocking wait: try_wait with a suspend-time hint parks the warp in hardware; between retries the
// warp sleeps `backoff_ns` (warps whose pixels have all terminated only keep the ring turning and pass a
// long backoff so their polling does not take issue slots from warps that still blend).
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
  uint32_t released[kStages];   // warps that have released the stage's current round
  uint32_t done_warps;
  uint32_t stop_round;
  uint32_t tmax;                // max n_contrib of the tile
};

// One stage: the contiguous record slab plus the 16-byte-aligned window of the mask byte array that
// covers the same instances, both completing on the stage's full barrier.
__device__ __forceinline__ void stage_load(StageBuf &sb, int s, const float4 *records, const uint8_t *masks,
                                           size_t first, uint32_t cnt) {
  const size_t m0 = first & ~(size_t)15;
  const uint32_t mbytes = (uint32_t)(((first + cnt + 15) & ~(size_t)15) - m0);
  mbar_expect_tx(&sb.full[s], cnt * kRecBytes + mbytes);
  bulk_g2s(&sb.rec[s][0], records + 3 * first, cnt * kRecBytes, &sb.full[s]);
  bulk_g2s(&sb.msk[s][0], masks + m0, mbytes, &sb.full[s]);
}

// exp(x) for x <= 0 as one FMUL + MUFU.EX2 (ex2.approx: <= 2 ulp; argument rounding adds <= 3.3e-7
// relative at |x| <= 5.6, the largest exponent that can still pass alpha >= 1/255).  libdevice expf
// costs 9 more issue slots per pair in kernels that are issue-bound.  Forward and backward use the
// same function, so every alpha / skip decision of the backward replays the forward's exactly.
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// thread -> its two pixels inside the tile: warp w covers the 8x8 block (w&1, w>>1); lane l the column
// l&7 and the rows l>>3 and (l>>3) + 4.  slot(h) = w*64 + h*32 + l indexes the per-tile float4 arrays
// (tilefinal, checkpoints): a warp's 32 slots of one half are contiguous (coalesced 512-byte accesses).
__device__ __forceinline__ void pixels_of_thread(int warp, int lane, int &lx, int &ly0) {
  lx = ((warp & 1) << 3) + (lane & 7);
  ly0 = ((warp >> 1) << 3) + (lane >> 3);
}
// bits of the instance masks (bit = 2*(row/4) + col/8 of an 8x4 sub-block) that fall into warp w's block
__device__ __forceinline__ uint32_t hitmask_of_warp(int warp) {
  const uint32_t b0 = 4u * (uint32_t)(warp >> 1) + (uint32_t)(warp & 1);
  return (1u << b0) | (1u << (b0 + 2u));
}

// alpha of one instance at the thread's two pixels, canonical order, packed:
//   dx = x - px; dy = y - py; q = fma(A dx, dx, (C dy) dy); power = fma(-0.5, q, -(B dx) dy)
//   alpha = min(0.99, opacity * exp(power)); skipped (0) if power > 0 or alpha < 1/255
__device__ __forceinline__ void pair_alpha(const float4 a, const float4 bq, float pxf, f32x2 npy, float &dx,
                                           f32x2 &dy, float &p0, float &p1, float &G0, float &G1, float &a0,
                                           float &a1) {
  dx = fsub(a.x, pxf);
  dy = add2(bc2(a.y), npy);
  // record layout: a = {x, y, A, C}, bq = {B, opacity, thr, 0}
  const f32x2 v = mul2(mul2(bc2(a.w), dy), dy);
  const f32x2 qf = fma2(bc2(fmul(a.z, dx)), bc2(dx), v);
  const f32x2 z = mul2(bc2(fmul(-bq.x, dx)), dy);
  const f32x2 power = fma2(bc2(-0.5f), qf, z);
  upk2(power, p0, p1);
#ifdef GHR_EXACT_EXP
  G0 = expf(p0);
  G1 = expf(p1);
#else
  float e0, e1;
  upk2(mul2(power, bc2(kLog2e)), e0, e1);
  G0 = ex2_fast(e0);
  G1 = ex2_fast(e1);
#endif
  upk2(mul2(bc2(bq.y), pk2(G0, G1)), a0, a1);
  a0 = fminf(0.99f, a0);
  a1 = fminf(0.99f, a1);
}

// Survivor queue of one backward warp: 8-bit indices inside the segment (ascending) of the instances whose
// sub-block mask hits the warp's 8x8 block.  Each lane tests 4 instances (one byte LDS each), 4 ballots
// compact them.  Survivor i is q[kQPad + i]; `limit` = instances that precede the warp's last
// contributor; kQPad entries pad the end with index kSeg, the all-zero record behind the segment.
__device__ __forceinline__ uint32_t build_queue_idx(const uint8_t *msk, uint32_t cnt, uint32_t limit, uint32_t hitmask,
                                                    int lane, uint8_t *q) {
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t total = 0;
#pragma unroll
  for (int w = 0; w < kSeg / 32; w++) {
    const uint32_t e = w * 32 + lane;
    bool hit = false;
    if (e < cnt && e < limit) hit = (msk[e] & hitmask) != 0u;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
    if (hit) q[kQPad + total + __popc(m & lt)] = (uint8_t)e;
    total += __popc(m);
  }
  if (lane < kQPad) q[kQPad + total + lane] = (uint8_t)kSeg;
  __syncwarp();
  return total;
}
#ifdef GHR_BWD_HALFQ
// Build-time variant (A/B): the two halves of a backward warp own ONE 8x4 sub-block each (two pixels per thread,
// two rows apart) and walk their OWN survivor queue, so an instance that reaches only one of the warp's two
// sub-blocks costs half a warp iteration instead of a whole one.  Returns the upper half's count, totalL the lower's.
__device__ __forceinline__ uint32_t build_queue_half(const uint8_t *msk, uint32_t cnt, uint32_t limitU, uint32_t limitL,
                                                     uint32_t bitU, uint32_t bitL, int lane, uint8_t *qU, uint8_t *qL,
                                                     uint32_t &totalL) {
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t tU = 0, tL = 0;
#pragma unroll
  for (int w = 0; w < kSeg / 32; w++) {
    const uint32_t e = w * 32 + lane;
    const uint32_t m = e < cnt ? (uint32_t)msk[e] : 0u;
    const bool hitU = e < limitU && (m & bitU) != 0u, hitL = e < limitL && (m & bitL) != 0u;
    const uint32_t mU = __ballot_sync(0xFFFFFFFFu, hitU), mL = __ballot_sync(0xFFFFFFFFu, hitL);
    if (hitU) qU[kQPad + tU + __popc(mU & lt)] = (uint8_t)e;
    if (hitL) qL[kQPad + tL + __popc(mL & lt)] = (uint8_t)e;
    tU += __popc(mU);
    tL += __popc(mL);
  }
  __syncwarp();
  totalL = tL;
  return tU;
}
#endif

// ---- forward ----
// Thread = ONE pixel, a warp = one 8x4 sub-block, 8 warps per tile in two CTAs.  A tile's list is a serial
// chain per warp (the transmittance recursion), and a launch has only a few hundred non-empty tiles: its
// duration is set by the longest chains, i.e. by how fast ONE scheduler issues what a warp does per stage
// and per instance (per-warp clocks, tools/blend_timeline.py: with all 8 warps of a tile in one CTA the two
// busiest sub-blocks shared a scheduler and the tile ran at half speed while most of the GPU was idle).
// Hence:
//  - the tile is spread over as many warps as it has 32-pixel blocks, four per CTA (one per scheduler, and
//    the two CTAs of a tile usually land on different SMs); the packed pair is the pixel's
//    (dx, dy): records carry (x, y) and (A, C) as aligned pairs, so d = (x, y) - p and (A dx, C dy) are one
//    instruction each, (r, g) * alpha likewise -- every half in the canonical order;
//  - per stage a warp only builds a queue of the 16-bit shared-memory addresses of its survivors (mask
//    bit of its sub-block; 4 ballots) and reads their records in place: three LDS.128 per instance;
//  - eight instances per iteration: their alphas do not depend on the running transmittance, so the long
//    chains (LDS -> quadratic form -> exp -> thresholds, ~150 cycles) of 8 instances overlap;
//  - the serial part is ONE multiply per instance: tc <- tc * (1 - alpha) runs on as a pure product even
//    past the pixel's termination (1 - alpha <= 1, so tc < 1e-4 is sticky without a select in the chain);
//    the frozen output value, the weight mask and the last-contributor index hang off it as selects.
constexpr int kFwdWarps = 4;                  // warps per CTA: one per scheduler
constexpr int kFwdParts = 2;                  // CTAs per tile (upper / lower half), each staging the slab
constexpr int kFwdThreads = kFwdWarps * 32;
// A/B on B200 (tools/stage_times.py, 8 views, graphed step with two chains): {8 instances per iteration, 94
// registers, 6 stages: 5 CTAs per SM} 0.4165 ms; {8, 64 registers, 4 stages: 8 CTAs} 0.4109; {4, 64, 4: 8 CTAs}
// 0.4065.  The kernel's own duration barely moves (93-96 us): the smaller CTAs let the other chain's kernels in.
#ifndef GHR_FWD_ILP
#define GHR_FWD_ILP 4
#endif
#ifndef GHR_FWD_MINCTAS
#define GHR_FWD_MINCTAS 8
#endif
constexpr int kIlpF = GHR_FWD_ILP;            // instances per iteration (build-time A/B)
struct __align__(128) FwdSmem {
  StageBuf sb;
  uint16_t q[kFwdWarps][kStageN + 2 * kQPad];
};

// Survivor queue of one warp for one stage: the 16-bit shared-memory ADDRESSES of the records (the
// kernel's shared window is below 64 KB) of the instances, ascending, whose sub-block mask has the warp's
// bit.  Each lane tests 4 instances (one byte LDS each), 4 ballots compact them; the blend loop then takes
// kIlpF addresses per LDS.128, so an instance costs one extract and no address arithmetic before its
// three record loads.  Survivor i is q[kQPad + i]; kQPad entries pad the end with the address of the
// stage's all-zero record (alpha = 0).  Returns the survivor count.
__device__ __forceinline__ uint32_t build_queue_addr(const uint8_t *msk, uint32_t cnt, uint32_t hitbit, int lane,
                                                     uint16_t *q, uint32_t rec_base, uint32_t pad_addr) {
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t mb[kStageN / 32], m[kStageN / 32];
  __syncwarp();   // every lane has finished reading the previous stage's queue
#pragma unroll
  for (int w = 0; w < kStageN / 32; w++) mb[w] = msk[w * 32 + lane];     // (bytes past cnt are stale, not out of bounds)
#pragma unroll
  for (int w = 0; w < kStageN / 32; w++)
    m[w] = __ballot_sync(0xFFFFFFFFu, (uint32_t)(w * 32 + lane) < cnt && (mb[w] & hitbit) != 0u);
  uint32_t total = 0;
#pragma unroll
  for (int w = 0; w < kStageN / 32; w++) {
    if ((m[w] >> lane) & 1u) q[kQPad + total + __popc(m[w] & lt)] = (uint16_t)(rec_base + (w * 32 + lane) * kRecBytes);
    total += __popc(m[w]);
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

// alpha and (colour * alpha) of one instance at the thread's pixel, canonical order:
//   dx = x - px; dy = y - py; q = fma(A dx, dx, (C dy) dy); power = fma(-0.5, q, (-B dx) dy);
//   alpha = min(0.99, opacity * exp(power)), 0 if power > 0 or alpha < 1/255
__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float fwd_alpha(const float4 a, const float2 bo, f32x2 npxy) {
  const f32x2 d = add2(pk2(a.x, a.y), npxy);
  float dx, dy, mA, mC;
  upk2(d, dx, dy);
  upk2(mul2(pk2(a.z, a.w), d), mA, mC);
  const float qf = ffma(mA, dx, fmul(mC, dy));
  const float power = ffma(-0.5f, qf, fmul(fmul(-bo.x, dx), dy));
#ifdef GHR_EXACT_EXP
  const float G = expf(power);
#else
  const float G = ex2_fast(power * kLog2e);
#endif
  const float alpha = fminf(0.99f, fmul(bo.y, G));
  return (power <= 0.0f && alpha >= kAlphaMin) ? alpha : 0.f;   // padding: opacity 0 -> alpha 0
}

// One instance of the serial part, branch-free (written in PTX so that it stays a select chain):
//   t = tc * (1 - alpha)   test_T; alpha == 0: t == tc exactly.  A live tc is >= 1e-4 (the stopping
//                          Gaussian is never applied), so a live pixel terminates only on alpha != 0
//   term = t < 1e-4        sticky: tc keeps shrinking after the pixel's termination
//   C += (c alpha) * (term ? 0 : tc)         the stopping Gaussian is not blended
//   Tr = term ? Tr : t                       the output transmittance freezes at its last live value
//   last = (!term && alpha != 0) ? idx : last
__device__ __forceinline__ void fwd_blend_step(float &tc, float &Tr, float &Cr, float &Cg, float &Cb, uint32_t &last,
                                               float om, float a, float r, float g, float b, uint32_t idx) {
  asm("{\n\t"
      ".reg .pred pt, pc;\n\t"
      ".reg .f32 t, tm;\n\t"
      "mul.rn.f32 t, %0, %6;\n\t"
      "setp.lt.f32 pt, t, 0f38D1B717;\n\t"
      "setp.neu.and.f32 pc, %7, 0f00000000, !pt;\n\t"
      "selp.f32 tm, 0f00000000, %0, pt;\n\t"
      "fma.rn.f32 %2, %8, tm, %2;\n\t"
      "fma.rn.f32 %3, %9, tm, %3;\n\t"
      "fma.rn.f32 %4, %10, tm, %4;\n\t"
      "selp.f32 %1, %1, t, pt;\n\t"
      "selp.b32 %5, %11, %5, pc;\n\t"
      "mov.f32 %0, t;\n\t"
      "}"
      : "+f"(tc), "+f"(Tr), "+f"(Cr), "+f"(Cg), "+f"(Cb), "+r"(last)
      : "f"(om), "f"(a), "f"(r), "f"(g), "f"(b), "r"(idx));
}

__device__ __forceinline__ void pixel_of_thread(int warp, int lane, int &lx, int &ly) {
  lx = ((warp & 1) << 3) + (lane & 7);
  ly = ((warp >> 1) << 2) + (lane >> 3);
}

__global__ void __launch_bounds__(kFwdThreads, GHR_FWD_MINCTAS)
blend_forward_kernel(int H, int W, int gx, int T, Cameras cam, const uint32_t *__restrict__ order,
                     const uint2 *__restrict__ ranges, const float4 *__restrict__ records,
                     const uint8_t *__restrict__ masks, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
                     uint32_t *__restrict__ tilemax, float4 *__restrict__ tilefinal, float4 *__restrict__ ckpt,
                     uint4 *__restrict__ units, GhrStatus *__restrict__ status, float *__restrict__ out_color,
                     float *__restrict__ out_mask) {
  __shared__ FwdSmem sm;
  StageBuf &sb = sm.sb;
  const uint32_t vt = order[blockIdx.x / kFwdParts], part = blockIdx.x % kFwdParts;
  const int v = vt / (uint32_t)T, tile = vt % (uint32_t)T;
  const int lane = threadIdx.x & 31, wloc = threadIdx.x >> 5;
  const int warp = (int)part * kFwdWarps + wloc, tid = warp * 32 + lane;   // warp / thread inside the tile
  int lx, ly;
  pixel_of_thread(warp, lane, lx, ly);
  const int px = (tile % gx) * kTile + lx, py = (tile / gx) * kTile + ly;
  const bool inside = px < W && py < H;
  const size_t N = (size_t)H * W, pix = (size_t)py * W + px;
  const float *bg = cam.bg + (size_t)cam.bg_stride * v;

  const uint2 range = ranges[vt];
  const uint32_t n = range.y - range.x;
  if (n == 0) {
    // empty tile (most of the frame): background only, no barriers, no staging
    if (inside) {
      final_T[(size_t)v * N + pix] = 1.0f;
      n_contrib[(size_t)v * N + pix] = 0u;
      float *o = out_color + (size_t)v * 3 * N + pix;
      o[0] = ffma(1.0f, bg[0], 0.f);
      o[N] = ffma(1.0f, bg[1], 0.f);
      o[2 * N] = ffma(1.0f, bg[2], 0.f);
      if (out_mask) out_mask[(size_t)v * N + pix] = 0.f;
    }
    return;
  }
  const uint32_t rounds = (n + kStageN - 1) / kStageN;
#ifdef GHR_TIMELINE
  const unsigned long long tl0 = gtime_ns();
#endif
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(&sb.full[s], 1);
      sb.released[s] = 0;
    }
    sb.done_warps = 0;
    sb.stop_round = 0xFFFFFFFFu;
    sb.tmax = 0;
    mbar_fence_init();
    for (uint32_t r = 0; r < rounds && r < (uint32_t)kStages; r++)
      stage_load(sb, (int)r, records, masks, (size_t)range.x + (size_t)r * kStageN,
                 min((uint32_t)kStageN, n - r * kStageN));
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + 3 * kStages)
    sb.rec[(threadIdx.x - 32) / 3][kStageN * 3 + (threadIdx.x - 32) % 3] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  const f32x2 npxy = pk2(-(float)px, -(float)py);
  const uint32_t hitbit = 1u << warp;
  uint16_t *q = &sm.q[wloc][0];

  // tc: running product of (1 - alpha); live while >= 1e-4 (0 from the start outside the image).
  // Tr: the output transmittance, frozen at the pixel's last live value.
  float tc = inside ? 1.0f : 0.f, Tr = 1.0f, Cr = 0.f, Cg = 0.f, Cb = 0.f;
  bool wdone = __all_sync(0xFFFFFFFFu, !inside);
  if (wdone && lane == 0) atomicAdd(&sb.done_warps, 1u);
  uint32_t last = 0;
#ifdef GHR_TIMELINE
  uint32_t tl_surv = 0, tl_iter = 0, tl_pass = 0, tl_rounds = 0;
  long long tl_wait = 0, tl_comp = 0, tl_blend = 0, tl_c0 = 0;
#endif
  for (uint32_t r = 0; r < rounds; r++) {
    const int s = r % kStages;
#ifdef GHR_TIMELINE
    tl_c0 = clock64();
#endif
    mbar_wait(&sb.full[s], (r / kStages) & 1, wdone ? 256u : 32u);
#ifdef GHR_TIMELINE
    tl_wait += clock64() - tl_c0;
    if (!wdone) tl_rounds = r + 1;
    tl_c0 = clock64();
#endif
    if (r >= *(volatile uint32_t *)&sb.stop_round) break;
    if (!wdone) {
      // running state at every kSeg-instance boundary: the backward restarts from it (one work unit per
      // segment).  A warp whose pixels have all terminated writes nothing: no later unit reads it.
      if (r > 0) ckpt[((size_t)(range.x / kSeg) + vt + r) * 256 + tid] = make_float4(Tr, Cr, Cg, Cb);
      const uint32_t cnt = min((uint32_t)kStageN, n - r * kStageN);
      const uint32_t rec_base = smem_u32(&sb.rec[s][0]);
      const uint32_t total = build_queue_addr(&sb.msk[s][(range.x + r * kStageN) & 15u], cnt, hitbit, lane, q, rec_base,
                                              rec_base + kStageN * kRecBytes);
#ifdef GHR_TIMELINE
      tl_surv += total;
      tl_iter += (total + kIlpF - 1u) / kIlpF;
      tl_pass++;
      tl_comp += clock64() - tl_c0;
      tl_c0 = clock64();
#endif
      uint32_t lastq = 0;                                     // 1 + queue index of the last blended survivor
#ifdef GHR_COUNT
      uint32_t n_contributing = 0;
#endif
      for (uint32_t b = 0; b < total; b += kIlpF) {
        uint32_t packed[4];
        if (kIlpF == 8) {
          const uint4 p4 = *reinterpret_cast<const uint4 *>(q + kQPad + b);
          packed[0] = p4.x; packed[1] = p4.y; packed[2] = p4.z; packed[3] = p4.w;
        } else {
          const uint2 p2 = *reinterpret_cast<const uint2 *>(q + kQPad + b);
          packed[0] = p2.x; packed[1] = p2.y; packed[2] = 0u; packed[3] = 0u;
        }
        uint32_t addr[kIlpF];
        float4 ga[kIlpF];
        float2 gb[kIlpF];
        float al[kIlpF];
        // phase 1: the geometry of all kIlpF instances is requested before any of it is used, so the eight
        // alpha chains run side by side; phase 2 fetches the colours as the serial update reaches them
#pragma unroll
        for (int k = 0; k < kIlpF; k++) {
          addr[k] = (k & 1) ? packed[k >> 1] >> 16 : packed[k >> 1] & 0xFFFFu;
          ga[k] = lds128(addr[k]);
          gb[k] = lds64(addr[k] + 16);
        }
#pragma unroll
        for (int k = 0; k < kIlpF; k++) al[k] = fwd_alpha(ga[k], gb[k], npxy);
#pragma unroll
        for (int k = 0; k < kIlpF; k++) {
          const float4 c = lds128(addr[k] + 32);
          float r0, g0;
          upk2(mul2(pk2(c.x, c.y), bc2(al[k])), r0, g0);               // upstream's order: (c * alpha) * T
          fwd_blend_step(tc, Tr, Cr, Cg, Cb, lastq, fsub(1.0f, al[k]), al[k], r0, g0, fmul(c.z, al[k]), b + k + 1);
#ifdef GHR_COUNT
          n_contributing += lastq == b + k + 1;
#endif
        }
        if (__all_sync(0xFFFFFFFFu, tc < 0.0001f)) {
          wdone = true;
          if (lane == 0) atomicAdd(&sb.done_warps, 1u);
          break;
        }
      }
      if (lastq) last = r * kStageN + ((uint32_t)q[kQPad + lastq - 1] - rec_base) / kRecBytes + 1;
#ifdef GHR_COUNT
      count_add(0, total);          // x 32 pixels, applied on the host
      count_add(1, n_contributing);
#endif
#ifdef GHR_TIMELINE
      tl_blend += clock64() - tl_c0;
#endif
    }
    // release the stage; the last warp to do so refills it with round r + kStages (acq_rel counter: see
    // atom_add_acq_rel_cta)
    __syncwarp();
    if (lane == 0) {
      if (atom_add_acq_rel_cta(&sb.released[s], 1u) == (uint32_t)kFwdWarps - 1u) {
        *(volatile uint32_t *)&sb.released[s] = 0u;
        const uint32_t nr = r + kStages;
        if (nr < rounds) {
          if (*(volatile uint32_t *)&sb.done_warps == (uint32_t)kFwdWarps) {
            // every pixel of the tile has terminated: complete the phase without data ("poison")
            atomicMin(&sb.stop_round, nr);
            __threadfence_block();
            mbar_arrive(&sb.full[s]);
          } else {
            fence_proxy_async();
            stage_load(sb, s, records, masks, (size_t)range.x + (size_t)nr * kStageN,
                       min((uint32_t)kStageN, n - nr * kStageN));
          }
        }
      }
    }
  }

  if (inside) {
    final_T[(size_t)v * N + pix] = Tr;
    n_contrib[(size_t)v * N + pix] = last;
    float *o = out_color + (size_t)v * 3 * N + pix;
    o[0] = ffma(Tr, bg[0], Cr);
    o[N] = ffma(Tr, bg[1], Cg);
    o[2 * N] = ffma(Tr, bg[2], Cb);
    if (out_mask) out_mask[(size_t)v * N + pix] = 1.0f - Tr;   // = sum_j alpha_j T_j
  }
  tilefinal[(size_t)vt * 256 + tid] = make_float4(Cr, Cg, Cb, Tr);
  // backward work units of this tile: one per kSeg instances up to the tile's last contributor, emitted by
  // the CTA of the tile that finishes last
  const uint32_t wmax = __reduce_max_sync(0xFFFFFFFFu, last);
  if (lane == 0 && wmax) atomicMax(&sb.tmax, wmax);
  __syncthreads();
  if (wloc == 0) {
    uint32_t tmax = 0;
    if (lane == 0) {
      if (sb.tmax) atomicMax(&tilemax[2 * vt], sb.tmax);
      __threadfence();
      if (atomicAdd(&tilemax[2 * vt + 1], 1u) == (uint32_t)kFwdParts - 1u) {
        __threadfence();
        tmax = atomicMax(&tilemax[2 * vt], 0u);
      }
    }
    tmax = __shfl_sync(0xFFFFFFFFu, tmax, 0);
    const uint32_t nseg = (tmax + kSeg - 1) / kSeg;
    if (nseg) {
      uint32_t base = 0;
      if (lane == 0) base = (uint32_t)atomicAdd((unsigned long long *)&status->reserved[1], (unsigned long long)nseg);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      for (uint32_t i = lane; i < nseg; i += 32) units[base + i] = make_uint4(vt, i, range.x, tmax);
    }
  }
#ifdef GHR_TIMELINE
  if (threadIdx.x == 0) timeline_put(tl0, 0u, vt, n);
  // per warp: survivors evaluated, iterations, compaction passes, stages it was live in
  if (lane == 0) {
    timeline_put(tl0, 1u + (uint32_t)warp, (tl_surv << 12) | tl_iter, (tl_pass << 12) | tl_rounds);
    timeline_put(tl0, 16u + (uint32_t)warp, (uint32_t)(tl_wait >> 4), ((uint32_t)(tl_comp >> 4) << 16) | (uint32_t)(tl_blend >> 4));
  }
#endif
}

// Ring reduction: instead of reducing every instance on its own, a warp parks the 9 partials of up to
// kRingSlots instances as rows of 32 floats (row stride 36 floats: lane l writes column l,
// conflict-free) and reduces the slots together: the 27 rows are cut into 54 half-rows, lane l sums
// half-row l (then 32 + l) with four LDS.128 + 15 FADD, one xor-shuffle joins the halves, and the even
// lanes send the totals as REDs -- two rounds for three instances (~16 issue slots per instance
// instead of ~40 for a per-instance column sum).  Eight consecutive half-rows start in eight
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
      // the 16 partials as 8 packed pairs (the four LDS.128 deliver aligned register pairs): 7 FADD2 + 1 FADD
      // instead of 15 FADD
      const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(ring + r * kRowStride + (t & 1u) * 16u);
      const ulonglong2 a = p[0], b = p[1], c = p[2], d = p[3];
      const f32x2 s2 = add2(add2(add2(a.x, a.y), add2(b.x, b.y)), add2(add2(c.x, c.y), add2(d.x, d.y)));
      sum = lo2(s2) + hi2(s2);
    }
#ifdef GHR_BWD_HALFQ
    // the two half-rows of a row belong to different instances (ids[2 * slot + half]): no join, one RED each
    if (t < ntask && sum != 0.f) {
      const uint32_t slot = (r * 57u) >> 9, val = r - 9u * slot;      // r / 9 for r < 27
#ifdef GHR_NO_RED
      if (sum == 123.456f) accb[(size_t)ids[2u * slot + (t & 1u)] * kAccStride + val] = sum;
#else
      atomicAdd(accb + (size_t)ids[2u * slot + (t & 1u)] * kAccStride + val, sum);
#endif
    }
    continue;
#endif
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, 1);
    if (t < ntask && !(t & 1u)) {
      const uint32_t slot = (r * 57u) >> 9, val = r - 9u * slot;      // r / 9 for r < 27
#ifdef GHR_NO_RED   // experiment only: what the kernel costs without its global reductions
      if (sum == 123.456f) accb[(size_t)ids[slot] * kAccStride + val] = sum;
#else
      atomicAdd(accb + (size_t)ids[slot] * kAccStride + val, sum);
#endif
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
// The kernel tracks Dn = -D and TbN = -T_final bg.dL/dpix so that every update is a plain packed fma.
// kW warps per CTA: 4 (the whole tile) or 2 (half a tile: two CTAs per unit, each staging the slab).
template <int kW>
__global__ void __launch_bounds__(kW * 32, kW == 4 ? GHR_BWD_MINCTAS : 2 * GHR_BWD_MINCTAS)
blend_backward_kernel(int H, int W, int gx, int T, int P, Cameras cam, const GhrStatus *__restrict__ status,
                      const uint4 *__restrict__ units, const float4 *__restrict__ records,
                      const uint8_t *__restrict__ masks, const float4 *__restrict__ tilefinal,
                      const float4 *__restrict__ ckpt, const uint32_t *__restrict__ n_contrib,
                      const float *__restrict__ dL_dout, const float *__restrict__ dL_dmask,
                      float *__restrict__ acc) {
  __shared__ __align__(128) float4 s_rec[kSeg * 3 + 3];   // + the all-zero padding record (index kSeg)
  __shared__ __align__(16) uint8_t s_msk[kSeg + 16];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ __align__(16) float s_red[kW][kRingSlots * kSlotFloats];
#ifdef GHR_BWD_HALFQ
  static_assert(kW == 4, "the half-warp queues are built for one CTA per unit");
  __shared__ uint32_t s_ids[kW][8];
  __shared__ __align__(16) uint8_t s_q[kW][2][kSeg + 2 * kQPad];
#else
  __shared__ uint32_t s_ids[kW][4];
  __shared__ __align__(16) uint8_t s_q[kW][kSeg + 2 * kQPad];
#endif
  constexpr uint32_t kParts = kBlendWarps / kW;       // CTAs per unit
  const uint32_t bid = blockIdx.x / kParts, part = blockIdx.x % kParts;
  // The unit record {view*T + tile, segment, start of the tile's slab, instances up to the tile's last
  // contributor} carries everything the copy needs, and it is read together with the unit count (the
  // list is allocated to its upper bound): one memory round trip between CTA start and the bulk copy.
  // The forward appends a tile's units when the tile completes, so the heaviest tiles sit at the end of
  // the list; walking it backwards starts their (long, dense) units first and leaves the short ones to
  // fill the tail of the launch.
  const uint32_t n_units = (uint32_t)status->reserved[1];
  if (bid >= n_units) return;
  const uint4 unit = units[n_units - 1u - bid];
  const uint32_t vt = unit.x, first = unit.y * kSeg;      // first = position of the segment in the tile list
  const int v = vt / (uint32_t)T, tile = vt % (uint32_t)T;
  const int lane = threadIdx.x & 31, wloc = threadIdx.x >> 5;      // warp inside the CTA
  const int warp = (int)part * kW + wloc;                           // warp inside the tile
#ifdef GHR_BWD_HALFQ
  const int half = lane >> 4, hl = lane & 15;
  uint8_t *q = &s_q[wloc][half][0];
#else
  uint8_t *q = &s_q[wloc][0];
#endif
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
  if (threadIdx.x >= 32 && threadIdx.x < 35) s_rec[kSeg * 3 + threadIdx.x - 32] = make_float4(0.f, 0.f, 0.f, 0.f);

  // per-pixel state while the segment is in flight (all loads independent of each other)
  int lx, ly0;
#ifdef GHR_BWD_HALFQ
  // half u of warp w owns sub-block (w&1, 2*(w>>1) + u): column hl & 7, rows hl >> 3 and (hl >> 3) + 2
  lx = ((warp & 1) << 3) + (hl & 7);
  ly0 = ((warp >> 1) << 3) + 4 * half + (hl >> 3);
  const int px = (tile % gx) * kTile + lx, py0 = (tile / gx) * kTile + ly0, py1 = py0 + 2;
#else
  pixels_of_thread(warp, lane, lx, ly0);
  const int px = (tile % gx) * kTile + lx, py0 = (tile / gx) * kTile + ly0, py1 = py0 + 4;
#endif
  const bool in0 = px < W && py0 < H, in1 = px < W && py1 < H;
  const size_t N = (size_t)H * W;
  const float pxf = (float)px;
  const f32x2 npy = pk2(-(float)py0, -(float)py1);
  // per-tile float4 arrays are indexed by the FORWARD's thread id (8x4 sub-block * 32 + lane): the
  // thread's upper pixel lies in sub-block 4*(warp>>1) + (warp&1), the lower one two sub-blocks on
#ifdef GHR_BWD_HALFQ
  // (forward lane of a pixel = (row % 4) * 8 + column % 8: hl for the first pixel, hl + 16 for the second)
  constexpr size_t kSlot1 = 16;
  const size_t slot0 = (size_t)(4 * (warp >> 1) + (warp & 1) + 2 * half) * 32 + hl;
#else
  constexpr size_t kSlot1 = 64;
  const size_t slot0 = (size_t)(4 * (warp >> 1) + (warp & 1)) * 32 + lane;
#endif
  const float4 fin0 = tilefinal[(size_t)vt * 256 + slot0], fin1 = tilefinal[(size_t)vt * 256 + slot0 + kSlot1];
  float4 c0 = make_float4(1.f, 0.f, 0.f, 0.f), c1 = c0;
  // (a checkpoint no later unit needs was never written: whatever is read there is not used)
  if (first) {
    const float4 *ck = ckpt + ((size_t)(unit.z / kSeg) + vt + unit.y) * 256 + slot0;
    c0 = ck[0];
    c1 = ck[kSlot1];
  }
  float d0[3] = {0.f, 0.f, 0.f}, d1[3] = {0.f, 0.f, 0.f}, dm0 = 0.f, dm1 = 0.f;
  uint32_t last0 = 0, last1 = 0;
  if (in0) {
    const size_t pix = (size_t)py0 * W + px;
    last0 = n_contrib[(size_t)v * N + pix];
    const float *g = dL_dout + (size_t)v * 3 * N + pix;
    d0[0] = g[0]; d0[1] = g[N]; d0[2] = g[2 * N];
    if (dL_dmask) dm0 = dL_dmask[(size_t)v * N + pix];
  }
  if (in1) {
    const size_t pix = (size_t)py1 * W + px;
    last1 = n_contrib[(size_t)v * N + pix];
    const float *g = dL_dout + (size_t)v * 3 * N + pix;
    d1[0] = g[0]; d1[1] = g[N]; d1[2] = g[2 * N];
    if (dL_dmask) dm1 = dL_dmask[(size_t)v * N + pix];
  }
  const float *bg = cam.bg + (size_t)cam.bg_stride * v;
  float tr0 = 0.f, tr1 = 0.f, dn0 = 0.f, dn1 = 0.f, tb0 = 0.f, tb1 = 0.f;
  if (last0 > first) {
    tr0 = c0.x;
    dn0 = -((fin0.x - c0.y) * d0[0] + (fin0.y - c0.z) * d0[1] + (fin0.z - c0.w) * d0[2]);
    // coverage output m = 1 - T_final: dm/dalpha_j = +T_final/(1-alpha_j), the background term with
    // the opposite sign, so its gradient folds into the same product
    tb0 = -(fin0.w * (bg[0] * d0[0] + bg[1] * d0[1] + bg[2] * d0[2] - dm0));
  }
  if (last1 > first) {
    tr1 = c1.x;
    dn1 = -((fin1.x - c1.y) * d1[0] + (fin1.y - c1.z) * d1[1] + (fin1.z - c1.w) * d1[2]);
    tb1 = -(fin1.w * (bg[0] * d1[0] + bg[1] * d1[1] + bg[2] * d1[2] - dm1));
  }
  f32x2 Tr = pk2(tr0, tr1), Dn = pk2(dn0, dn1);
  const f32x2 TbN = pk2(tb0, tb1);
  const f32x2 dLr = pk2(d0[0], d1[0]), dLg = pk2(d0[1], d1[1]), dLb = pk2(d0[2], d1[2]);
  float *accb = acc + (size_t)v * P * kAccStride;
  const uint32_t wlast = __reduce_max_sync(0xFFFFFFFFu, max(last0, last1));
  __syncthreads();                  // barrier initialised before anyone polls it
  if (wlast <= first) return;
  mbar_wait(&s_bar, 0);
  float *ring = &s_red[wloc][0];
  uint32_t *ids = &s_ids[wloc][0];
  uint32_t pend = 0;                // instances parked in the ring (warp-uniform)

  // survivors of this warp's block among the instances that precede the warp's last contributor
#ifdef GHR_BWD_HALFQ
  uint32_t hlast = max(last0, last1);                        // last contributor of the thread's HALF
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) hlast = max(hlast, __shfl_xor_sync(0xFFFFFFFFu, hlast, o));
  const uint32_t lastU = __shfl_sync(0xFFFFFFFFu, hlast, 0), lastL = __shfl_sync(0xFFFFFFFFu, hlast, 16);
  const uint32_t b0 = 4u * (uint32_t)(warp >> 1) + (uint32_t)(warp & 1);
  uint32_t totalL;
  const uint32_t totalU = build_queue_half(&s_msk[(uint32_t)g0 & 15u], cnt, lastU > first ? lastU - first : 0u,
                                           lastL > first ? lastL - first : 0u, 1u << b0, 1u << (b0 + 2u), lane,
                                           &s_q[wloc][0][0], &s_q[wloc][1][0], totalL);
  const uint32_t mine = half ? totalL : totalU;             // survivors of this lane's half
  const uint32_t total = max(totalU, totalL);
#else
  const uint32_t total = build_queue_idx(&s_msk[(uint32_t)g0 & 15u], cnt, wlast - first, hitmask_of_warp(warp), lane, q);
#endif
#ifdef GHR_COUNT
  count_add(2, total);              // warp iterations: x 64 pixels, applied on the host
  uint32_t n_contributing = 0;
#endif
  for (uint32_t b = 0; b < total; b += kIlpB) {
    // phase 1 (independent per instance): alpha, G, 1/(1-alpha), offsets, colour . dL/dpix
    f32x2 al[kIlpB], Gk[kIlpB], omk[kIlpB], rck[kIlpB], dyk[kIlpB], cdk[kIlpB];
    float dxk[kIlpB];
    uint32_t idk[kIlpB];
#pragma unroll
    for (int k = 0; k < kIlpB; k++) {
#ifdef GHR_BWD_HALFQ
      const uint32_t jj = b + k < mine ? (uint32_t)q[kQPad + b + k] : (uint32_t)kSeg;   // exhausted half: zero record
#else
      const uint32_t jj = q[kQPad + b + k];                // tail pad: the all-zero record (alpha = 0)
#endif
      const float4 a = s_rec[3 * jj], bq = s_rec[3 * jj + 1], col = s_rec[3 * jj + 2];
      float p0, p1, G0, G1, a0, a1;
      pair_alpha(a, bq, pxf, npy, dxk[k], dyk[k], p0, p1, G0, G1, a0, a1);
      const uint32_t pos = first + jj;
      const bool ok0 = pos < last0 && p0 <= 0.0f && a0 >= kAlphaMin;
      const bool ok1 = pos < last1 && p1 <= 0.0f && a1 >= kAlphaMin;
      al[k] = pk2(ok0 ? a0 : 0.f, ok1 ? a1 : 0.f);
      Gk[k] = pk2(ok0 ? G0 : 0.f, ok1 ? G1 : 0.f);
      omk[k] = fma2(al[k], bc2(-1.0f), bc2(1.0f));
      float o0, o1;
      upk2(omk[k], o0, o1);
      rck[k] = pk2(rcp_fast(o0), rcp_fast(o1));
      cdk[k] = fma2(bc2(col.z), dLb, fma2(bc2(col.y), dLg, mul2(bc2(col.x), dLr)));
      idk[k] = __float_as_uint(col.w);
    }
    // phase 2 (serial, short, branch-free: a rejected pair has alpha = G = 0 and adds zeros), then the
    // thread's own two pixels are summed: per instance 9 values per thread
    float vals[kIlpB][9];
    bool contrib[kIlpB];
#pragma unroll
    for (int k = 0; k < kIlpB; k++) {
      float a0, a1;
      upk2(al[k], a0, a1);
      contrib[k] = a0 != 0.f || a1 != 0.f;
#ifdef GHR_COUNT
      n_contributing += (a0 != 0.f) + (a1 != 0.f);
#endif
      const f32x2 wgt = mul2(al[k], Tr);
      Dn = fma2(cdk[k], wgt, Dn);
      const f32x2 dL_dalpha = fma2(Tr, cdk[k], mul2(add2(Dn, TbN), rck[k]));
      Tr = mul2(Tr, omk[k]);                                 // the forward's own product
      const f32x2 wG = mul2(Gk[k], dL_dalpha);
      const f32x2 t01 = mul2(wG, dyk[k]), t02 = mul2(t01, dyk[k]);
      const f32x2 cr = mul2(wgt, dLr), cg = mul2(wgt, dLg), cb = mul2(wgt, dLb);
      const float wGs = lo2(wG) + hi2(wG), m01 = lo2(t01) + hi2(t01), m10 = wGs * dxk[k];
      vals[k][0] = lo2(cr) + hi2(cr);
      vals[k][1] = lo2(cg) + hi2(cg);
      vals[k][2] = lo2(cb) + hi2(cb);
      vals[k][3] = wGs;
      vals[k][4] = m10;
      vals[k][5] = m01;
      vals[k][6] = m10 * dxk[k];
      vals[k][7] = m01 * dxk[k];
      vals[k][8] = lo2(t02) + hi2(t02);
    }
    // phase 3: slots with few contributing lanes send their partials straight to L2 (9 REDs per lane);
    // the others are reduced through shared memory first and leave as one RED per value
#pragma unroll
    for (int k = 0; k < kIlpB; k++) {
      const uint32_t cm = __ballot_sync(0xFFFFFFFFu, contrib[k]);
      if (cm == 0u) continue;
      if (__popc(cm) <= kDirectMax) {
        if (contrib[k]) {
          float *dst = accb + (size_t)idk[k] * kAccStride;
#pragma unroll
#ifdef GHR_NO_RED
          for (int t = 0; t < 9; t++) if (vals[k][t] == 123.456f) dst[t] = vals[k][t];
#else
          for (int t = 0; t < 9; t++) atomicAdd(dst + t, vals[k][t]);
#endif
        }
      } else {
        float *row = ring + pend * kSlotFloats + lane;
#pragma unroll
        for (int t = 0; t < 9; t++) row[t * kRowStride] = vals[k][t];
#ifdef GHR_BWD_HALFQ
        if (hl == 0) ids[2u * pend + half] = idk[k];
#else
        if (lane == 0) ids[pend] = idk[k];
#endif
        if (++pend == kRingSlots) {
          __syncwarp();
          ring_flush(ring, ids, kRingSlots, accb, lane);
          __syncwarp();
          pend = 0;
        }
      }
    }
  }
  if (pend) {
    __syncwarp();
    ring_flush(ring, ids, pend, accb, lane);
  }
#ifdef GHR_COUNT
  count_add(3, n_contributing);
#endif
}

// shared-memory carveout preference, set once per device and kernel
template <typename K>
void prefer_shared(K kern) {
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || done[dev]) return;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  done[dev] = true;
}

}  // namespace

#ifdef GHR_COUNT
extern "C" int ghr_debug_counts(unsigned long long *out4, int reset) {
  cudaDeviceSynchronize();
  if (out4) cudaMemcpyFromSymbol(out4, g_counts, sizeof(unsigned long long) * 4);
  if (reset) {
    const unsigned long long z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(g_counts, z, sizeof(z));
  }
  return 0;
}
#endif

#ifdef GHR_TIMELINE
extern "C" int ghr_debug_timeline(unsigned long long *host_out, unsigned int cap, unsigned int *count) {
  unsigned int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, g_timeline_n, sizeof(n));
  if (n > kTimelineCap) n = kTimelineCap;
  if (n > cap) n = cap;
  if (n) cudaMemcpyFromSymbol(host_out, g_timeline, (size_t)n * sizeof(ulonglong4));
  const unsigned int zero = 0;
  cudaMemcpyToSymbol(g_timeline_n, &zero, sizeof(zero));
  *count = n;
  return 0;
}
#endif

cudaError_t launch_blend_forward(const GhrDims &d, const Layout &L, const Cameras &cam, char *state,
                                 float *out_color, float *out_mask, cudaStream_t s) {
  if (L.T == 0 || d.V == 0) return cudaSuccess;
  prefer_shared(blend_forward_kernel);
  blend_forward_kernel<<<L.T * d.V * kFwdParts, kFwdThreads, 0, s>>>(
      d.H, d.W, L.gx, L.T, cam, (const uint32_t *)(state + L.pub.off_order), (const uint2 *)(state + L.pub.off_ranges),
      (const float4 *)(state + L.pub.off_records), (const uint8_t *)(state + L.pub.off_masks),
      (float *)(state + L.pub.off_final_T), (uint32_t *)(state + L.pub.off_ncontrib),
      (uint32_t *)(state + L.pub.off_tilemax), (float4 *)(state + L.pub.off_tilefinal),
      (float4 *)(state + L.pub.off_ckpt), (uint4 *)(state + L.pub.off_units), (GhrStatus *)(state + L.pub.off_status),
      out_color, out_mask);
  return cudaGetLastError();
}

cudaError_t launch_blend_backward(const GhrDims &d, const Layout &L, const Cameras &cam, const char *state,
                                  const float *dL_dout, const float *dL_dmask, float *acc, cudaStream_t s) {
  if (L.T == 0 || d.V == 0 || d.R_cap <= 0) return cudaSuccess;
  // upper bound of the unit count (the forward wrote the exact one to GhrStatus.reserved[1]); surplus
  // CTAs exit on their first instructions
  constexpr int kW = GHR_BWD_WARPS;
  prefer_shared(blend_backward_kernel<kW>);
  const unsigned grid = (unsigned)((L.n_slots - 1) * (kBlendWarps / kW));
  blend_backward_kernel<kW><<<grid, kW * 32, 0, s>>>(
      d.H, d.W, L.gx, L.T, d.P, cam, (const GhrStatus *)(state + L.pub.off_status),
      (const uint4 *)(state + L.pub.off_units), (const float4 *)(state + L.pub.off_records),
      (const uint8_t *)(state + L.pub.off_masks), (const float4 *)(state + L.pub.off_tilefinal),
      (const float4 *)(state + L.pub.off_ckpt), (const uint32_t *)(state + L.pub.off_ncontrib), dL_dout, dL_dmask,
      acc);
  return cudaGetLastError();
}

}  // namespace ghr
