// Stable LSD radix sort, one kernel per 8-bit digit ("onesweep": the global digit histograms are
// known before the pass, the cross-block digit prefix is resolved by decoupled look-back).
//
// Replaces upstream's cub::DeviceRadixSort::SortPairs on 64-bit (tile<<32|depth) keys over the R
// duplicated instances.  The sort is factorised: a stable sort on the low
// word (depth, V*P elements, 4 passes -- this file, launch_depth_sort) followed, after duplication
// in depth order, by a stable sort on the high word ((view,tile), R elements, ceil(bits/8) passes
// -- launch_tile_sort).  An LSD sort is exactly that sequence, so the final order is identical to
// the upstream stable 64-bit sort while the 4 depth passes touch 8 B x V*P instead of 12 B x R.
#include "ghr_internal.cuh"

namespace ghr {

namespace {

constexpr uint32_t kFlagLocal = 1u << 30;
constexpr uint32_t kFlagIncl = 2u << 30;
constexpr uint32_t kValMask = (1u << 30) - 1;
constexpr uint32_t kGroup = 32;                 // blocks per group of the two-level prefix
constexpr uint32_t kTwoLevelMaxBlocks = 2048;   // above: classic decoupled look-back

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile(uint32_t *p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One look-back window: W predecessors b, b-1, ... loaded back to back (independent loads), then
// consumed nearest-first until an inclusive prefix is met.  Returns true when one was found.
template <int W>
__device__ __forceinline__ bool lookback_window(const uint32_t *status, int b, int d, uint32_t &excl) {
  uint32_t sv[W];
#pragma unroll
  for (int w = 0; w < W; w++) sv[w] = (b - w >= 0) ? ld_volatile(&status[(size_t)(b - w) * 256 + d]) : kFlagIncl;
  bool found = false;
#pragma unroll
  for (int w = 0; w < W; w++) {
    if (found) break;
    uint32_t x = sv[w];
    while ((x & ~kValMask) == 0) x = ld_volatile(&status[(size_t)(b - w) * 256 + d]);
    excl += x & kValMask;
    if ((x & ~kValMask) == kFlagIncl) found = true;
  }
  return found;
}

// One digit pass over one segment (blockIdx.y).  Items are taken in warp-striped order so that
// (warp, iteration, lane) is the stable order.
template <int ITEMS, bool IDENTITY_VALS, bool WRITE_KEYS>
__global__ void __launch_bounds__(kSortThreads)
radix_pass_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                  const uint64_t *__restrict__ n_ptr, uint32_t n_fixed, uint64_t n_cap, size_t seg_stride,
                  int shift, const uint32_t *__restrict__ hist, uint32_t *__restrict__ status,
                  size_t status_seg_stride, uint32_t *__restrict__ tickets, uint32_t nblk_cap) {
  constexpr int kWarps = kSortThreads / 32;
  constexpr int kBlockItems = kSortThreads * ITEMS;
  __shared__ uint32_t warp_cnt[kWarps][256];
  __shared__ uint32_t digit_base[256];
  __shared__ uint32_t s_scan[kWarps];
  __shared__ uint32_t s_blk;

  const int seg = blockIdx.y;
  uint32_t n = n_fixed;
  if (n_ptr) {
    uint64_t nn = *n_ptr;
    n = (uint32_t)(nn < n_cap ? nn : n_cap);
  }
  const uint32_t nblk = (n + kBlockItems - 1) / kBlockItems;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_blk = atomicAdd(&tickets[seg], 1u);
  for (int i = tid; i < kWarps * 256; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t blk = s_blk;
  if (blk >= nblk) return;

  keys_in += seg * seg_stride;
  keys_out += seg * seg_stride;
  vals_out += seg * seg_stride;
  if (!IDENTITY_VALS) vals_in += seg * seg_stride;
  hist += seg * 256;
  status += seg * status_seg_stride;

  const uint32_t base = blk * kBlockItems + warp * (32 * ITEMS) + lane;
  uint32_t key[ITEMS], val[ITEMS];
  uint32_t rank[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    uint32_t idx = base + i * 32;
    bool ok = idx < n;
    key[i] = ok ? keys_in[idx] : 0xFFFFFFFFu;
    val[i] = IDENTITY_VALS ? idx : (ok ? vals_in[idx] : 0u);
  }
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    uint32_t idx = base + i * 32;
    bool ok = idx < n;
    uint32_t d = ok ? ((key[i] >> shift) & 255u) : 256u;   // 256 = padding, never counted
    uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader && ok) {
      old = warp_cnt[warp][d];
      warp_cnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xFFFFFFFFu, old, leader);
    rank[i] = old + __popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();

  // thread d owns digit d: prefix over warps, publish, look back, add the global digit base
  {
    const int d = tid;
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      uint32_t c = warp_cnt[w][d];
      warp_cnt[w][d] = total;
      total += c;
    }
    uint32_t excl = 0;
    const bool two_level = nblk <= kTwoLevelMaxBlocks;
    if (blk == 0) {
      st_volatile(&status[(size_t)blk * 256 + d], kFlagIncl | total);
      if (two_level) atomicAdd(&status[(size_t)nblk_cap * 256 + d], (1u << 24) | total);
    } else {
      st_volatile(&status[(size_t)blk * 256 + d], kFlagLocal | total);
      if (two_level) {
        // At these sizes every block of the pass is resident at once, so a classic decoupled look-back
        // degenerates into a serial chain (a block cannot count on its predecessor being finished).
        // Two-level sum instead, free of any chain: every block adds its count to its group's word
        // ((arrivals << 24) | sum, one RED) right after ranking; a block needs the local counts of the
        // earlier blocks of its own group and the words of the earlier, complete groups -- all of
        // which only wait for ranking, never for another block's look-back.
        uint32_t *grp = status + (size_t)nblk_cap * 256;
        const uint32_t g = blk / kGroup;
        atomicAdd(&grp[(size_t)g * 256 + d], (1u << 24) | total);
#pragma unroll 1
        for (uint32_t b0 = g * kGroup; b0 < blk; b0 += 8) {
          uint32_t sv[8];
#pragma unroll
          for (int w = 0; w < 8; w++) sv[w] = (b0 + w < blk) ? ld_volatile(&status[(size_t)(b0 + w) * 256 + d]) : kFlagLocal;
#pragma unroll
          for (int w = 0; w < 8; w++) {
            uint32_t x = sv[w];
            while ((x & ~kValMask) == 0) x = ld_volatile(&status[(size_t)(b0 + w) * 256 + d]);
            excl += x & kValMask;
          }
        }
#pragma unroll 1
        for (uint32_t g0 = 0; g0 < g; g0 += 8) {
          uint32_t sv[8];
#pragma unroll
          for (int w = 0; w < 8; w++) sv[w] = (g0 + w < g) ? ld_volatile(&grp[(size_t)(g0 + w) * 256 + d]) : (kGroup << 24);
#pragma unroll
          for (int w = 0; w < 8; w++) {
            uint32_t x = sv[w];
            while ((x >> 24) != kGroup) x = ld_volatile(&grp[(size_t)(g0 + w) * 256 + d]);
            excl += x & 0xFFFFFFu;
          }
        }
      } else {
        // classic decoupled look-back (multi-wave grids: predecessors are mostly finished), windows
        // of 8 independent loads consumed nearest-first until an inclusive prefix is met
        int b = (int)blk - 1;
        bool found = false;
#pragma unroll 1
        while (!found) {
          found = lookback_window<8>(status, b, d, excl);
          b -= 8;
        }
        st_volatile(&status[(size_t)blk * 256 + d], kFlagIncl | (excl + total));
      }
    }
    // exclusive scan of the global histogram over digits
    uint32_t h = hist[d];
    uint32_t incl = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++)
      if (w < warp) wbase += s_scan[w];
    digit_base[d] = wbase + incl - h + excl;
  }
  __syncthreads();

#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    uint32_t idx = base + i * 32;
    if (idx < n) {
      uint32_t d = (key[i] >> shift) & 255u;
      uint32_t pos = digit_base[d] + warp_cnt[warp][d] + rank[i];
      if (WRITE_KEYS) keys_out[pos] = key[i];
      vals_out[pos] = val[i];
    }
  }
}

template <int ITEMS>
cudaError_t run_passes(int npass, int first_shift, bool identity_first, bool drop_last_keys, uint32_t *keys[2],
                       uint32_t *vals[2], const uint64_t *n_ptr, uint32_t n_fixed, uint64_t n_cap, int segs,
                       size_t seg_stride, const uint32_t *hist /*[pass][seg][256]*/, uint32_t *status,
                       size_t status_pass_stride, size_t status_seg_stride, uint32_t *tickets, int nblk,
                       cudaStream_t s) {
  dim3 grid(nblk, segs), block(kSortThreads);
  if (nblk == 0 || segs == 0) return cudaSuccess;
  for (int p = 0; p < npass; p++) {
    int in = p & 1, out = in ^ 1;
    const uint32_t *h = hist + (size_t)p * segs * 256;
    uint32_t *st = status + (size_t)p * status_pass_stride;
    uint32_t *tk = tickets + (size_t)p * segs;
    int shift = first_shift + 8 * p;
    bool ident = identity_first && p == 0;
    bool wk = !(drop_last_keys && p == npass - 1);
    if (ident)
      radix_pass_kernel<ITEMS, true, true><<<grid, block, 0, s>>>(keys[in], nullptr, keys[out], vals[out], n_ptr,
                                                                   n_fixed, n_cap, seg_stride, shift, h, st,
                                                                   status_seg_stride, tk, (uint32_t)nblk);
    else if (wk)
      radix_pass_kernel<ITEMS, false, true><<<grid, block, 0, s>>>(keys[in], vals[in], keys[out], vals[out],
                                                                    n_ptr, n_fixed, n_cap, seg_stride, shift, h, st,
                                                                    status_seg_stride, tk, (uint32_t)nblk);
    else
      radix_pass_kernel<ITEMS, false, false><<<grid, block, 0, s>>>(keys[in], vals[in], keys[out], vals[out],
                                                                     n_ptr, n_fixed, n_cap, seg_stride, shift, h,
                                                                     st, status_seg_stride, tk, (uint32_t)nblk);
  }
  return cudaGetLastError();
}

}  // namespace

// Depth sort: V segments of P (depth-bits, Gaussian index) pairs; 4 passes; result in buffer 0.
cudaError_t launch_depth_sort(const GhrDims &d, const Layout &L, char *temp, cudaStream_t s) {
  uint32_t *keys[2] = {(uint32_t *)(temp + L.t_dkeys[0]), (uint32_t *)(temp + L.t_dkeys[1])};
  uint32_t *vals[2] = {(uint32_t *)(temp + L.t_dvals[0]), (uint32_t *)(temp + L.t_dvals[1])};
  const uint32_t *hist = (const uint32_t *)(temp + L.t_dhist);
  uint32_t *status = (uint32_t *)(temp + L.t_dstatus);
  uint32_t *tickets = (uint32_t *)(temp + L.t_tickets);
  size_t seg_status = status_words(L.nblk_d);
  size_t pass_status = seg_status * d.V;
  if (L.items_d == 4)
    return run_passes<4>(4, 0, true, true, keys, vals, nullptr, (uint32_t)d.P, 0, d.V, (size_t)d.P, hist, status,
                         pass_status, seg_status, tickets, L.nblk_d, s);
  return run_passes<16>(4, 0, true, true, keys, vals, nullptr, (uint32_t)d.P, 0, d.V, (size_t)d.P, hist, status,
                        pass_status, seg_status, tickets, L.nblk_d, s);
}

// Tile sort: one segment of R ((view*T+tile), view*P+idx) pairs, R read from GhrStatus on device.
cudaError_t launch_tile_sort(const GhrDims &d, const Layout &L, char *state, char *temp, cudaStream_t s) {
  uint32_t *keys[2] = {(uint32_t *)(temp + L.t_tkeys[0]), (uint32_t *)(temp + L.t_tkeys[1])};
  uint32_t *vals[2] = {(uint32_t *)(temp + L.t_tvals[0]), (uint32_t *)(temp + L.t_tvals[1])};
  const uint32_t *hist = (const uint32_t *)(temp + L.t_thist);
  uint32_t *status = (uint32_t *)(temp + L.t_tstatus);
  uint32_t *tickets = (uint32_t *)(temp + L.t_tickets) + 4 * (size_t)d.V;
  const uint64_t *n_ptr = &((const GhrStatus *)(state + L.pub.off_status))->R;
  size_t seg_status = status_words(L.nblk_t);
  if (L.items_t == 4)
    return run_passes<4>(L.npt, 0, false, false, keys, vals, n_ptr, 0, (uint64_t)d.R_cap, 1, 0, hist, status,
                         seg_status, seg_status, tickets, L.nblk_t, s);
  return run_passes<16>(L.npt, 0, false, false, keys, vals, n_ptr, 0, (uint64_t)d.R_cap, 1, 0, hist, status,
                        seg_status, seg_status, tickets, L.nblk_t, s);
}

}  // namespace ghr
