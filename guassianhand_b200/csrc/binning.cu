// Tile binning: from per-tile instance counts (preprocess) to the per-tile, depth-sorted slabs of
// instance records the blend kernels stream.
//
// Upstream: InclusiveSum over Gaussians -> blocking D2H of num_rendered ->
// duplicateWithKeys -> one global stable radix sort of R 64-bit (tile<<32|depth) keys ->
// identifyTileRanges.  The final order is "by tile, then by depth bits, ties in Gaussian-index order"
// -- a total order on (tile, depth bits, index).  Here the sort is done most-significant part first:
//   tile_scan_schedule : exclusive scan of the per-tile counts = the tile ranges (known BEFORE any
//                        instance exists), R and the overflow flag, plus the heaviest-first launch
//                        order of the tiles.  One CTA; V*T is a few thousand.
//   duplicate          : every visible Gaussian drops one (depth bits, id) pair into each tile it
//                        touches, at a slot taken from the tile's cursor -- unordered inside the tile.
//                        Slots are reserved per (block, tile) from shared-memory counts, so the global
//                        atomics are one per touched tile per block, not one per instance.
//   sort_chunks        : one CTA sorts one work item of at most kChunk instances by (depth bits, id) in
//                        shared memory: a counting pass on the 11 leading significant bits, then every key
//                        ranks itself inside its bin.  An item that is a whole tile list (or a depth bucket,
//                        below) is finished here: its 48-byte geometry records are gathered in sorted order
//                        with their sub-block cull masks.
//   merge_gather       : a list of a few chunks: one CTA per chunk ranks its keys in the other (sorted)
//                        chunks -- final position = own index + ranks -- and gathers its records straight
//                        to their final place.  Quadratic in the chunks of a list.
//   heavy_*            : when long lists are likely, every multi-chunk list is first partitioned by depth
//                        (min/max, 256-slab histogram, bucket plan, scatter) into depth-disjoint buckets of
//                        at most kChunk instances, each an ordinary sort item that is final in place:
//                        linear in the list length; merge_gather is then only the fallback for lists whose
//                        depths are too degenerate to cut.
// No global sort, no cross-block prefix: all work items are uniform, independent after the scan.  Result:
// bit for bit the upstream order (checked against the oracle's sorted keys / point list / ranges).
#include <cstdlib>

#include <algorithm>

#include "ghr_internal.cuh"

namespace ghr {

namespace {

__device__ __forceinline__ void rect_of(float px, float py, int radius, int gx, int gy, int &minx, int &miny,
                                        int &maxx, int &maxy) {
  float rf = (float)radius;
  // x / 16 == x * 0.0625 bit for bit (power of two)
  minx = min(gx, max(0, (int)fmul(fsub(px, rf), 0.0625f)));
  miny = min(gy, max(0, (int)fmul(fsub(py, rf), 0.0625f)));
  maxx = min(gx, max(0, (int)fmul(fsub(fadd(fadd(px, rf), 16.0f), 1.0f), 0.0625f)));
  maxy = min(gy, max(0, (int)fmul(fsub(fadd(fadd(py, rf), 16.0f), 1.0f), 0.0625f)));
}

// Size class of a tile list: 0 for an empty tile, else 1 + ceil(log2 n) (n = 1 -> 1, 2 -> 2, 3..4 -> 3, ...).
// Lists of one class share the padded power-of-two length 2^(class-1) of the sorting network.
__device__ __forceinline__ int size_class(uint32_t n) { return n == 0 ? 0 : 33 - __clz(n - 1); }

constexpr int kScanThreads1 = 1024;
constexpr int kClasses = 34;

// Sort work items.  A tile list of at most kChunk instances is one item and is final after its chunk
// sort.  A list of a few chunks is sorted chunk by chunk and finished by the rank merge (each chunk ranks
// its keys in the others: quadratic in the chunk count, cheapest up to ~6 chunks).  A longer list
// ("heavy", more than part_min instances) is first partitioned by depth into buckets of at most kChunk
// instances (minmax -> slab histogram -> plan -> scatter below), each bucket one final item; a heavy
// tile whose depths are too degenerate to partition falls back to chunks + merge as well.
// Item = {view*T + tile, offset inside the tile list, count, flags}.
constexpr uint32_t kItemSrcB = 1u;     // instances are read from the partitioned copy (inst_b)
constexpr uint32_t kItemFinal = 2u;    // the sorted item is in its final place: gather its records
constexpr int kSlabs = 256;            // depth slabs of a heavy tile (fine histogram bins)

// ranges[t] = [start, end) of tile t in the instance arrays (clamped to R_cap; (0,0) for empty tiles,
// as upstream leaves them), order[] = tile ids by descending size class (the blend launch order),
// items[] = one sort item per light tile (misc[0] = their number; the plan kernel appends the heavy
// tiles' buckets), hchunks[] = (tile, chunk index) pieces of the heavy tiles for the partition passes
// (misc[1]), heavy[] = the heavy tiles (misc[2]) with heavy_id[tile] their index, status = {R, overflow}.
// gridDim.x CTAs, each owning a span of consecutive tiles, with NO communication between them: every CTA
// first reads ALL V*T counts (a few coalesced, independent loads per thread) and reduces them twice -- over
// everything (R, the class histogram that lays out order[], the list totals) and over the tiles in front of
// its span (its base offsets) -- then scans its own span (a thread owns a run of consecutive tiles, one block
// scan of the run totals, a second walk over the run).  One CTA per ~1024 tiles: 8 views of 672 tiles are
// one tile per thread on 6 SMs instead of 6 tiles per thread on one (round 2: 17 -> ~7 us).
struct ScanTotals {
  unsigned long long sum;   // instances
  uint32_t l, h, hc, nbig;  // light items, heavy tiles, heavy chunks, lists longer than a chunk
  uint32_t merge;           // light items that are one of several chunks of their list (finished by merge_gather)
};

__global__ void __launch_bounds__(kScanThreads1)
tile_scan_schedule_kernel(int VT, uint64_t R_cap, uint32_t item_cap, uint32_t hchunk_cap, uint32_t heavy_cap,
                          uint32_t part_min, uint32_t big_max, const uint32_t *__restrict__ tile_count,
                          uint2 *__restrict__ ranges,
                          uint32_t *__restrict__ order, uint4 *__restrict__ items, uint2 *__restrict__ hchunks,
                          uint32_t *__restrict__ heavy, uint32_t *__restrict__ heavy_id,
                          uint32_t *__restrict__ misc, GhrStatus *__restrict__ status, uint64_t seq) {
  __shared__ uint64_t s_warp[kScanThreads1 / 32];
  __shared__ uint32_t s_wl[kScanThreads1 / 32], s_wh[kScanThreads1 / 32], s_whc[kScanThreads1 / 32];
  __shared__ uint32_t s_cls[kClasses + 1], s_before[kClasses + 1], s_cur[kClasses + 1];
  __shared__ ScanTotals s_all, s_pre;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid <= kClasses) {
    s_cls[tid] = 0;
    s_before[tid] = 0;
  }
  if (tid == 0) {
    s_all = ScanTotals{0ull, 0u, 0u, 0u, 0u, 0u};
    s_pre = ScanTotals{0ull, 0u, 0u, 0u, 0u, 0u};
  }
  __syncthreads();
  // spans are multiples of the block size, so a pass-1 round lies entirely in front of the span or not at all
  const int span = ((VT + (int)gridDim.x - 1) / (int)gridDim.x + kScanThreads1 - 1) / kScanThreads1 * kScanThreads1;
  const int lo = min(VT, (int)blockIdx.x * span), hi = min(VT, lo + span);

  // ---- pass 1: every count, reduced over all tiles and over the tiles in front of the span ----
  {
    unsigned long long a_sum = 0, b_sum = 0;
    uint32_t a_l = 0, a_h = 0, a_hc = 0, a_nbig = 0, a_merge = 0, b_l = 0, b_h = 0, b_hc = 0;
    // (the loads of up to kPre rounds are issued together: one memory round trip instead of one per round)
    constexpr int kPre = 8;
    for (int tb = 0; tb < VT; tb += kPre * kScanThreads1) {
      uint32_t cv[kPre];
#pragma unroll
      for (int u = 0; u < kPre; u++) {
        const int t = tb + u * kScanThreads1 + tid;
        cv[u] = t < VT ? tile_count[t] : 0u;
      }
#pragma unroll
      for (int u = 0; u < kPre; u++) {
        if (tb + u * kScanThreads1 >= VT) break;      // (block-uniform)
        const int t = tb + u * kScanThreads1 + tid;
        const bool valid = t < VT;
        const uint32_t c = cv[u];
        const bool pre = tb + u * kScanThreads1 < lo;   // block-uniform (lo is a multiple of the block size)
        uint32_t l = 0, h = 0, hc = 0;
        if (c > part_min) {
          h = 1;
          hc = (c + kChunk - 1) / kChunk;
        } else if (!(c > (uint32_t)kChunk && c <= big_max)) {   // (a big list is ONE item of sort_big_kernel)
          l = (c + kChunk - 1) / kChunk;           // one final item, or plain chunks finished by the rank merge
        }
        a_sum += c; a_l += l; a_h += h; a_hc += hc;
        a_nbig += c > (uint32_t)kChunk;
        a_merge += l > 1u ? l : 0u;
        if (pre) { b_sum += c; b_l += l; b_h += h; b_hc += hc; }
        // class counters are warp-aggregated (match.any): most tiles are empty, and thousands of shared-memory
        // atomics on one address would serialise
        const int cls = valid ? size_class(c) : kClasses;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cls);
        if (lane == __ffs(peers) - 1) {
          atomicAdd(&s_cls[cls], (uint32_t)__popc(peers));
          if (pre) atomicAdd(&s_before[cls], (uint32_t)__popc(peers));
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a_sum += __shfl_xor_sync(0xFFFFFFFFu, a_sum, o);
      b_sum += __shfl_xor_sync(0xFFFFFFFFu, b_sum, o);
    }
    a_l = __reduce_add_sync(0xFFFFFFFFu, a_l);
    a_h = __reduce_add_sync(0xFFFFFFFFu, a_h);
    a_hc = __reduce_add_sync(0xFFFFFFFFu, a_hc);
    a_nbig = __reduce_add_sync(0xFFFFFFFFu, a_nbig);
    a_merge = __reduce_add_sync(0xFFFFFFFFu, a_merge);
    b_l = __reduce_add_sync(0xFFFFFFFFu, b_l);
    b_h = __reduce_add_sync(0xFFFFFFFFu, b_h);
    b_hc = __reduce_add_sync(0xFFFFFFFFu, b_hc);
    if (lane == 0) {
      if (a_sum) atomicAdd(&s_all.sum, a_sum);
      if (a_l) atomicAdd(&s_all.l, a_l);
      if (a_h) atomicAdd(&s_all.h, a_h);
      if (a_hc) atomicAdd(&s_all.hc, a_hc);
      if (a_nbig) atomicAdd(&s_all.nbig, a_nbig);
      if (a_merge) atomicAdd(&s_all.merge, a_merge);
      if (b_sum) atomicAdd(&s_pre.sum, b_sum);
      if (b_l) atomicAdd(&s_pre.l, b_l);
      if (b_h) atomicAdd(&s_pre.h, b_h);
      if (b_hc) atomicAdd(&s_pre.hc, b_hc);
    }
  }

  // ---- pass 2: the span ----
  const int per = (hi - lo + kScanThreads1 - 1) / kScanThreads1;
  const int t0 = min(hi, lo + tid * per), t1 = min(hi, t0 + per);
  uint64_t sum = 0;
  uint32_t lsum = 0, hsum = 0, hcsum = 0;      // light tiles, heavy tiles, heavy chunks of this run
  for (int t = t0; t < t1; t++) {
    const uint32_t c = tile_count[t];
    sum += c;
    if (c > part_min) {
      hsum++;
      hcsum += (c + kChunk - 1) / kChunk;
    } else if (!(c > (uint32_t)kChunk && c <= big_max)) {
      lsum += (c + kChunk - 1) / kChunk;
    }
  }
  uint64_t incl = sum;
  uint32_t lincl = lsum, hincl = hsum, hcincl = hcsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    const uint32_t lup = __shfl_up_sync(0xFFFFFFFFu, lincl, o);
    const uint32_t hup = __shfl_up_sync(0xFFFFFFFFu, hincl, o);
    const uint32_t hcup = __shfl_up_sync(0xFFFFFFFFu, hcincl, o);
    if (lane >= o) {
      incl += up;
      lincl += lup;
      hincl += hup;
      hcincl += hcup;
    }
  }
  if (lane == 31) {
    s_warp[warp] = incl;
    s_wl[warp] = lincl;
    s_wh[warp] = hincl;
    s_whc[warp] = hcincl;
  }
  __syncthreads();                             // (and the pass-1 totals are complete)
  if (warp == 0) {
    // exclusive prefix of the 32 warp totals, in place (one warp scan instead of a serial walk in every thread)
    uint64_t a = s_warp[lane];
    uint32_t b = s_wl[lane], c = s_wh[lane], d = s_whc[lane];
    const uint64_t a0 = a;
    const uint32_t b0 = b, c0 = c, d0 = d;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t ua = __shfl_up_sync(0xFFFFFFFFu, a, o);
      const uint32_t ub = __shfl_up_sync(0xFFFFFFFFu, b, o);
      const uint32_t uc = __shfl_up_sync(0xFFFFFFFFu, c, o);
      const uint32_t ud = __shfl_up_sync(0xFFFFFFFFu, d, o);
      if (lane >= o) {
        a += ua;
        b += ub;
        c += uc;
        d += ud;
      }
    }
    s_warp[lane] = a - a0;
    s_wl[lane] = b - b0;
    s_wh[lane] = c - c0;
    s_whc[lane] = d - d0;
  }
  if (warp == 0) {
    // launch-order offsets: largest class first, empty tiles (class 0) last.  Lane l owns the classes at the
    // reversed positions 2l and 2l+1 (class = kClasses - 1 - position): one warp scan instead of a serial walk
    static_assert(kClasses <= 64, "two classes per lane");
    const int r0 = 2 * lane, r1 = r0 + 1;
    const int c0 = kClasses - 1 - r0, c1 = kClasses - 1 - r1;
    const uint32_t n0 = c0 >= 0 ? s_cls[c0] : 0u, n1 = c1 >= 0 ? s_cls[c1] : 0u;
    uint32_t incl2 = n0 + n1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl2, o);
      if (lane >= o) incl2 += up;
    }
    const uint32_t base2 = incl2 - (n0 + n1);
    if (c0 >= 0) s_cur[c0] = base2 + s_before[c0];
    if (c1 >= 0) s_cur[c1] = base2 + n0 + s_before[c1];
  }
  if (tid == 0) {
    if (blockIdx.x == 0) {
      const uint64_t total = s_all.sum;
      // the whole status: preprocess left its two words in misc (zeroed with the rest of the temp prefix), the
      // forward blend appends its backward units to reserved[1]
      GhrStatus st;
      st.R = total;
      st.overflow = (misc[kMiscPrefilter] ? GHR_STATUS_PREFILTER : 0u) | (total > R_cap ? GHR_STATUS_OVERFLOW : 0u);
      st.n_visible = misc[kMiscVisible];
      st.reserved[0] = seq;
      st.reserved[1] = 0;
      *status = st;
      misc[0] = s_all.l < item_cap ? s_all.l : item_cap;
      misc[1] = s_all.hc < hchunk_cap ? s_all.hc : hchunk_cap;
      misc[2] = s_all.h < heavy_cap ? s_all.h : heavy_cap;
      misc[3] = s_all.nbig;
      misc[4] = s_all.merge;       // (the plan kernel adds the fallback tiles' chunks)
    }
  }
  __syncthreads();
  // this thread's run starts at: tiles in front of the span + earlier warps + earlier lanes
  uint64_t start = s_pre.sum + s_warp[warp] + (incl - sum);
  uint32_t loff = s_pre.l + s_wl[warp] + (lincl - lsum), hoff = s_pre.h + s_wh[warp] + (hincl - hsum),
           hcoff = s_pre.hc + s_whc[warp] + (hcincl - hcsum);
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int j = 0; j < per; j++) {
    const int t = t0 + j;
    const bool valid = t < t1;
    const uint32_t c = valid ? tile_count[t] : 0u;
    const int cls = valid ? size_class(c) : kClasses;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cls);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) base = atomicAdd(&s_cur[cls], (uint32_t)__popc(peers));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (valid) {
      order[base + __popc(peers & lt_mask)] = (uint32_t)t;
      const uint64_t end = start + c;
      ranges[t] = c ? make_uint2((uint32_t)(start < R_cap ? start : R_cap), (uint32_t)(end < R_cap ? end : R_cap))
                    : make_uint2(0u, 0u);
      const uint32_t m = (c + kChunk - 1) / kChunk;
      if (c > part_min) {
        if (hoff < heavy_cap) {
          heavy[hoff] = (uint32_t)t;
          heavy_id[t] = hoff;
          for (uint32_t q = 0; q < m && hcoff + q < hchunk_cap; q++) hchunks[hcoff + q] = make_uint2((uint32_t)t, q);
        }
        hoff++;
        hcoff += m;
      } else if (!(c > (uint32_t)kChunk && c <= big_max)) {
        for (uint32_t q = 0; q < m && loff + q < item_cap; q++)
          items[loff + q] = make_uint4((uint32_t)t, q * kChunk, min((uint32_t)kChunk, c - q * kChunk),
                                       m == 1 ? kItemFinal : 0u);
        loff += m;
      }
      start = end;
    }
  }
}

// One thread per (view, Gaussian), a block = 256 consecutive Gaussians of one view.
__global__ void __launch_bounds__(256)
duplicate_kernel(int P, int gx, int gy, int T, int smem_tiles, uint64_t R_cap, const float4 *__restrict__ geom,
                 const uint2 *__restrict__ ranges, uint32_t *__restrict__ cursor, uint2 *__restrict__ inst) {
  extern __shared__ uint32_t s_mem[];
  uint32_t *s_cnt = s_mem, *s_base = s_mem + smem_tiles;
  const int v = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int k = threadIdx.x; k < smem_tiles; k += blockDim.x) s_cnt[k] = 0;
  __syncthreads();
  int minx = 0, miny = 0, maxx = 0, maxy = 0;
  uint32_t depth_bits = 0, gid = 0;
  if (i < P) {
    gid = (uint32_t)v * (uint32_t)P + (uint32_t)i;
    // (three independent loads: issued together, not one after the visibility test)
    const float4 q3 = geom[4 * (size_t)gid + 3];
    const float4 q0 = geom[4 * (size_t)gid];
    const float q2w = geom[4 * (size_t)gid + 2].w;
    if (__float_as_uint(q3.y)) {
      depth_bits = __float_as_uint(q2w);
      rect_of(q0.x, q0.y, __float_as_int(q3.x), gx, gy, minx, miny, maxx, maxy);
    }
  }
  const uint2 *vr = ranges + (size_t)v * T;
  uint32_t *vc = cursor + (size_t)v * T;
  if (smem_tiles) {
    for (int y = miny; y < maxy; y++)
      for (int x = minx; x < maxx; x++) atomicAdd(&s_cnt[y * gx + x], 1u);
    __syncthreads();
    for (int k = threadIdx.x; k < smem_tiles; k += blockDim.x) {
      const uint32_t c = s_cnt[k];
      if (c) {
        s_base[k] = vr[k].x + atomicAdd(&vc[k], c);
        s_cnt[k] = 0;
      }
    }
    __syncthreads();
    for (int y = miny; y < maxy; y++)
      for (int x = minx; x < maxx; x++) {
        const int t = y * gx + x;
        const uint64_t slot = (uint64_t)s_base[t] + atomicAdd(&s_cnt[t], 1u);
        if (slot < R_cap) inst[slot] = make_uint2(depth_bits, gid);
      }
  } else {
    // more tiles per view than fit in shared memory: one global atomic per instance
    for (int y = miny; y < maxy; y++)
      for (int x = minx; x < maxx; x++) {
        const int t = y * gx + x;
        const uint64_t slot = (uint64_t)vr[t].x + atomicAdd(&vc[t], 1u);
        if (slot < R_cap) inst[slot] = make_uint2(depth_bits, gid);
      }
  }
}

// ---- depth partition of the heavy tiles ----
// A heavy tile's instances sit unordered in its range.  They are cut into at most kSlabs slabs of equal
// width in depth bits between the tile's own minimum and maximum, consecutive slabs are grouped into
// buckets of at most kChunk instances, and the instances are copied bucket by bucket (unordered inside a
// bucket) into inst_b at the same range: buckets are depth-disjoint and in depth order, so sorting each
// one in shared memory finishes the tile -- linear in the list length, where ranking every chunk in
// every other chunk of its tile is quadratic (77k-instance lists at 1M Gaussians: 38 chunks).
constexpr int kPartThreads = 256;
constexpr int kPartItems = kChunk / kPartThreads;

// slab of a depth inside its tile: (depth bits - dmin) >> shift with shift such that the span fits kSlabs
__device__ __forceinline__ uint32_t slab_shift(uint32_t span) {
  const int hi = 32 - __clz(span);
  return hi > 8 ? (uint32_t)(hi - 8) : 0u;
}

// pass 1: per heavy tile, minimum (stored inverted: zero-initialised maximum of ~depth) and maximum depth
__global__ void __launch_bounds__(kPartThreads)
heavy_minmax_kernel(const uint2 *__restrict__ hchunks, const uint32_t *__restrict__ misc,
                    const uint2 *__restrict__ ranges, const uint32_t *__restrict__ heavy_id,
                    const uint2 *__restrict__ inst, uint2 *__restrict__ dmm) {
  if (blockIdx.x >= misc[1]) return;
  const uint2 hc = hchunks[blockIdx.x];
  const uint2 range = ranges[hc.x];
  const uint32_t cstart = range.x + hc.y * kChunk;
  if (cstart >= range.y) return;
  const uint32_t n = min((uint32_t)kChunk, range.y - cstart);
  uint32_t inv = 0u, mx = 0u;
#pragma unroll
  for (int i = 0; i < kPartItems; i++) {
    const uint32_t k = i * kPartThreads + threadIdx.x;
    if (k < n) {
      const uint32_t d = inst[cstart + k].x;
      inv = max(inv, ~d);
      mx = max(mx, d);
    }
  }
  inv = __reduce_max_sync(0xFFFFFFFFu, inv);
  mx = __reduce_max_sync(0xFFFFFFFFu, mx);
  __shared__ uint32_t s_inv, s_mx;
  if (threadIdx.x == 0) {
    s_inv = 0u;
    s_mx = 0u;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&s_inv, inv);
    atomicMax(&s_mx, mx);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint2 *d = dmm + heavy_id[hc.x];
    atomicMax(&d->x, s_inv);
    atomicMax(&d->y, s_mx);
  }
}

// pass 2 (kScatter = false): slab histogram of every heavy tile; pass 4 (kScatter = true): copy the
// instances to their bucket in inst_b -- slots inside a slab are reserved per block from a shared-memory
// histogram, one global atomic per (block, slab), like duplicate_kernel
template <bool kScatter>
__global__ void __launch_bounds__(kPartThreads)
heavy_slab_kernel(const uint2 *__restrict__ hchunks, const uint32_t *__restrict__ misc,
                  const uint2 *__restrict__ ranges, const uint32_t *__restrict__ heavy_id,
                  const uint2 *__restrict__ inst, const uint2 *__restrict__ dmm, uint32_t *__restrict__ slab_count,
                  const uint32_t *__restrict__ slab_off, const uint32_t *__restrict__ heavy_flag,
                  uint2 *__restrict__ inst_b) {
  __shared__ uint32_t s_cnt[kSlabs], s_base[kSlabs];
  if (blockIdx.x >= misc[1]) return;
  const uint2 hc = hchunks[blockIdx.x];
  const uint2 range = ranges[hc.x];
  const uint32_t cstart = range.x + hc.y * kChunk;
  if (cstart >= range.y) return;
  const uint32_t hid = heavy_id[hc.x];
  if (kScatter && heavy_flag[hid]) return;             // fallback tile: stays in place, plain chunks + merge
  const uint32_t n = min((uint32_t)kChunk, range.y - cstart);
  const uint2 mm = dmm[hid];
  const uint32_t dmin = ~mm.x, shift = slab_shift(mm.y - dmin);
  for (int k = threadIdx.x; k < kSlabs; k += kPartThreads) s_cnt[k] = 0;
  __syncthreads();
  uint2 e[kPartItems];
  uint32_t slot[kPartItems];
#pragma unroll
  for (int i = 0; i < kPartItems; i++) {
    const uint32_t k = i * kPartThreads + threadIdx.x;
    slot[i] = 0;
    if (k < n) {
      e[i] = inst[cstart + k];
      slot[i] = atomicAdd(&s_cnt[(e[i].x - dmin) >> shift], 1u);
    }
  }
  __syncthreads();
  uint32_t *cnt = slab_count + (size_t)hid * kSlabs;
  for (int k = threadIdx.x; k < kSlabs; k += kPartThreads) {
    const uint32_t c = s_cnt[k];
    if (c) {
      const uint32_t at = atomicAdd(&cnt[k], c);       // histogram; in the scatter pass: the slab's cursor
      if (kScatter) s_base[k] = range.x + slab_off[(size_t)hid * kSlabs + k] + at;
    }
  }
  if (!kScatter) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kPartItems; i++) {
    const uint32_t k = i * kPartThreads + threadIdx.x;
    if (k < n) {
      const uint32_t dst = s_base[(e[i].x - dmin) >> shift] + slot[i];
      if (dst < range.y) inst_b[dst] = e[i];
    }
  }
}

// pass 3: one CTA per heavy tile.  Exclusive scan of its slab counts (= slab offsets inside the tile
// range; the counters are zeroed to serve as the scatter's cursors), consecutive slabs grouped greedily
// into buckets of at most kChunk instances, one sort item per bucket appended to the item list.  A slab
// that alone exceeds kChunk (thousands of equal depths) cannot be cut: the tile then keeps the plain
// chunking and is finished by merge_gather.
__global__ void __launch_bounds__(kSlabs)
heavy_plan_kernel(const uint32_t *__restrict__ heavy, const uint2 *__restrict__ ranges, uint32_t *__restrict__ slab_count,
                  uint32_t *__restrict__ slab_off, uint32_t *__restrict__ heavy_flag, uint4 *__restrict__ items,
                  uint32_t *__restrict__ misc, uint32_t item_cap, int force_fallback) {
  __shared__ uint32_t s_c[kSlabs], s_o[kSlabs], s_warp[kSlabs / 32];
  __shared__ uint2 s_bucket[kSlabs];
  __shared__ uint32_t s_nb, s_at, s_over;
  if (blockIdx.x >= misc[2]) return;
  const uint32_t hid = blockIdx.x, vt = heavy[hid];
  const uint2 range = ranges[vt];
  const uint32_t nt = range.y - range.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t *cnt = slab_count + (size_t)hid * kSlabs;
  const uint32_t c = cnt[tid];
  cnt[tid] = 0u;
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) s_warp[warp] = incl;
  if (tid == 0) s_over = force_fallback ? 1u : 0u;
  __syncthreads();
  uint32_t off = incl - c;
  for (int w = 0; w < warp; w++) off += s_warp[w];
  s_c[tid] = c;
  s_o[tid] = off;
  slab_off[(size_t)hid * kSlabs + tid] = off;
  if (c > (uint32_t)kChunk) s_over = 1u;
  __syncthreads();
  const bool fallback = s_over != 0u;
  if (tid == 0) {
    uint32_t nb = 0;
    if (fallback) {
      nb = (nt + kChunk - 1) / kChunk;
    } else {
      uint32_t start = 0, cur = 0;
      for (int k = 0; k < kSlabs; k++) {
        const uint32_t ck = s_c[k];
        if (cur && cur + ck > (uint32_t)kChunk) {
          s_bucket[nb++] = make_uint2(start, cur);
          start = s_o[k];
          cur = 0;
        }
        cur += ck;
      }
      if (cur) s_bucket[nb++] = make_uint2(start, cur);
    }
    heavy_flag[hid] = fallback ? 1u : 0u;
    s_nb = nb;
    s_at = atomicAdd(&misc[0], nb);
    if (fallback) atomicAdd(&misc[4], nb);
  }
  __syncthreads();
  const uint32_t nb = s_nb, at = s_at;
  for (uint32_t b = tid; b < nb; b += kSlabs) {
    if (at + b >= item_cap) break;
    if (fallback) items[at + b] = make_uint4(vt, b * kChunk, min((uint32_t)kChunk, nt - b * kChunk), 0u);
    else items[at + b] = make_uint4(vt, s_bucket[b].x, s_bucket[b].y, kItemFinal | kItemSrcB);
  }
}

// ---- chunk sort: one CTA per chunk of <= kChunk instances of one tile ----
// Keys are 64-bit ((depth bits - min depth bits of the chunk) << 32 | Gaussian index within the view).
// One stable LSD pass on the 8-bit digit at `shift`, `a` -> `b` (shared memory): items in warp-blocked
// order (warp, iteration, lane), ranks by match.any + per-warp digit counters.
// A chunk is kSortItems instances per thread of its CTA: 2048 with 256 threads, 4096 with 512 (larger
// chunks = fewer multi-chunk tiles and fewer rank passes in merge_gather).
constexpr int kSortItems = 8;
template <int kSortThreads>
struct ChunkSort {
  uint16_t wcnt[kSortThreads / 32][256];   // per-warp digit counts (<= 32*kSortItems each)
  uint32_t dbase[256];                     // output offset of every digit
  uint32_t scan[8];
};

template <int kSortThreads>
__device__ __noinline__ void chunk_radix_pass(ChunkSort<kSortThreads> &S, const uint64_t *a, uint64_t *b, uint32_t n,
                                              int shift, int tid) {
  constexpr int kSortWarps = kSortThreads / 32;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  // items per thread this chunk needs (warp-blocked: warp w owns [w*32*it, (w+1)*32*it))
  const uint32_t it = (n + kSortThreads - 1) / kSortThreads;
  for (int k = tid; k < kSortWarps * 128; k += kSortThreads) reinterpret_cast<uint32_t *>(&S.wcnt[0][0])[k] = 0;
  __syncthreads();
  uint64_t key[kSortItems];
  uint32_t rank[kSortItems];
#pragma unroll
  for (int i = 0; i < kSortItems; i++) {
    if (i < (int)it) {
      const uint32_t idx = warp * 32 * it + i * 32 + lane;
      const bool ok = idx < n;
      key[i] = ok ? a[idx] : ~0ull;
      const uint32_t d = ok ? (uint32_t)(key[i] >> shift) & 255u : 256u;   // 256 = padding, never counted
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader && ok) {
        old = S.wcnt[warp][d];
        S.wcnt[warp][d] = (uint16_t)(old + __popc(peers));
      }
      old = __shfl_sync(0xFFFFFFFFu, old, leader);
      rank[i] = old + __popc(peers & lt_mask);
      __syncwarp();
    }
  }
  __syncthreads();
  // thread d (< 256) owns digit d: exclusive prefix over warps, then exclusive scan over digits
  uint32_t total = 0, incl = 0;
  if (tid < 256) {
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) {
      const uint32_t cw = S.wcnt[w][tid];
      S.wcnt[w][tid] = (uint16_t)total;
      total += cw;
    }
    incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) S.scan[warp] = incl;
  }
  __syncthreads();
  if (tid < 256) {
    uint32_t wbase = 0;
    for (int w = 0; w < warp; w++) wbase += S.scan[w];
    S.dbase[tid] = wbase + incl - total;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kSortItems; i++) {
    if (i < (int)it) {
      const uint32_t idx = warp * 32 * it + i * 32 + lane;
      if (idx < n) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
        b[S.dbase[d] + S.wcnt[warp][d] + rank[i]] = key[i];
      }
    }
  }
  __syncthreads();
}

// Writes one sorted instance: gathers its geometry record, computes the sub-block mask.
__device__ __forceinline__ void emit_instance(size_t r, uint32_t depth_bits, uint32_t id, uint32_t gbase, uint32_t tile,
                                              int tile_x0, int tile_y0, const float4 *__restrict__ geom,
                                              float4 *__restrict__ records, uint8_t *__restrict__ masks,
                                              uint64_t *__restrict__ dbg_keys, uint32_t *__restrict__ dbg_plist) {
  const size_t g = (size_t)gbase + id;
  const float4 q0 = geom[4 * g], q1 = geom[4 * g + 1];
  float4 q2 = geom[4 * g + 2];
  masks[r] = (uint8_t)subblock_mask(q0, q1, tile_x0, tile_y0);
  if (dbg_keys) dbg_keys[r] = ((uint64_t)tile << 32) | depth_bits;
  if (dbg_plist) dbg_plist[r] = id;
  q2.w = __uint_as_float(id);
  // record layout: {x, y, A, C} {B, opacity, thr, 0} {r, g, b, id} -- (x, y) and (A, C) are the operand
  // pairs of the blend kernels' packed (dx, dy) math
  records[3 * r] = make_float4(q0.x, q0.y, q0.z, q1.x);
  records[3 * r + 1] = make_float4(q0.w, q1.y, q1.z, 0.f);
  records[3 * r + 2] = q2;
}

// Sort of one chunk = ONE counting pass on the kBinBits leading significant bits of (depth bits - min
// depth bits of the chunk) -- as many bins as the chunk has slots, slot inside a bin taken with a
// shared-memory atomic (unordered) -- followed by a local fix: every bin with more than one key
// (equal depths land here too) is insertion-sorted on the full (depth bits, index) key by the thread
// that owns the bin's first slot.  A chunk with a bin longer than kMaxRun (degenerate depth
// distribution) is re-sorted by stable LSD radix passes over every key byte that varies.  Either way
// the result is the total order on (depth bits, index).
constexpr uint32_t kMaxRun = 32;

// shared memory of one sort CTA (dynamic): keys | bin counts / offsets (aliased by the fallback passes'
// counters) | per-warp record staging of the final gather
template <int kSortThreads>
struct SortSmem {
  static constexpr int kWarps = kSortThreads / 32;
  static constexpr int kItemMax = kSortThreads * kSortItems;
  static constexpr size_t kHistBytes =
      ((sizeof(ChunkSort<kSortThreads>) > (size_t)kItemMax * 4 ? sizeof(ChunkSort<kSortThreads>) : (size_t)kItemMax * 4) + 15) / 16 * 16;
  static constexpr size_t kOffHist = (size_t)kItemMax * 8;
  static constexpr size_t kOffStage = kOffHist + kHistBytes;
  static constexpr size_t kBytes = kOffStage + (size_t)kWarps * 96 * 16;
};

template <int kSortThreads>
__device__ __forceinline__ void sort_item(const uint4 item, const uint2 range, int P, FastDiv dT, FastDiv dgx, uint2 *inst,
                                          uint2 *inst_b, const float4 *__restrict__ geom, float4 *__restrict__ records,
                                          uint8_t *__restrict__ masks, uint64_t *__restrict__ dbg_keys,
                                          uint32_t *__restrict__ dbg_plist) {
  constexpr int kSortWarps = kSortThreads / 32;
  constexpr int kChunk = kSortThreads * kSortItems;
  constexpr int kBins = kChunk;                       // 8 bins per thread in the bin scan
  constexpr int kBinBits = kSortThreads == 256 ? 11 : (kSortThreads == 512 ? 12 : 13);
  static_assert(kBins == 1 << kBinBits, "chunk size");
  using SM = SortSmem<kSortThreads>;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  uint64_t *s_buf = reinterpret_cast<uint64_t *>(s_dyn);                       // [kChunk] keys
  struct HistU {
    union {
      ChunkSort<kSortThreads> cs;   // fallback passes only
      uint32_t hist[kBins];         // bin counts, then bin offsets
    };
  };
  HistU &U = *reinterpret_cast<HistU *>(s_dyn + SM::kOffHist);
  float4 (*s_stage)[96] = reinterpret_cast<float4 (*)[96]>(s_dyn + SM::kOffStage);   // record staging of the final gather
  __shared__ uint32_t s_scan[kSortWarps];
  __shared__ uint32_t s_dmin, s_dmax;
  __shared__ unsigned long long s_and, s_or;
  const uint32_t vt = item.x;
  const uint32_t cstart = range.x + item.y;
  if (cstart >= range.y) return;                       // item of a list clamped by an instance-capacity overflow
  const uint32_t n = min(min((uint32_t)kChunk, item.z), range.y - cstart);
  const bool single = item.w & kItemFinal;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t v = dT.div(vt), tile = vt - v * dT.d, gbase = v * (uint32_t)P;
  if (item.w & kItemSrcB) inst = inst_b;
  uint2 *src = inst + cstart;
  // keys per thread this item needs: every per-key loop below stops there (a 300-instance list must not pay
  // for eight rounds; uniform over the CTA)
  const int it = (int)((n + kSortThreads - 1) / kSortThreads);
  if (tid == 0) {
    s_dmin = 0xFFFFFFFFu;
    s_dmax = 0u;
    s_and = ~0ull;
    s_or = 0ull;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) U.hist[tid * 8 + i] = 0;
  __syncthreads();
  uint2 e[kSortItems];
  {
    uint32_t dmin = 0xFFFFFFFFu, dmax = 0u;
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
      const uint32_t k = i * kSortThreads + tid;
      e[i] = make_uint2(0u, 0u);
      if (i < it && k < n) {
        e[i] = src[k];
        dmin = min(dmin, e[i].x);
        dmax = max(dmax, e[i].x);
      }
    }
    dmin = __reduce_min_sync(0xFFFFFFFFu, dmin);
    dmax = __reduce_max_sync(0xFFFFFFFFu, dmax);
    if (lane == 0) {
      atomicMin(&s_dmin, dmin);
      atomicMax(&s_dmax, dmax);
    }
  }
  __syncthreads();
  const uint32_t dmin = s_dmin, span = s_dmax - dmin;
  const int hi = 32 - __clz(span);                     // significant bits of (depth - dmin); 0 if all equal
  const int shift0 = hi > kBinBits ? hi - kBinBits : 0;
  // bin + slot of every key; AND/OR of the packed keys for the fallback's byte skipping
  uint32_t slot[kSortItems];
  {
    uint64_t k_and = ~0ull, k_or = 0ull;
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
      const uint32_t k = i * kSortThreads + tid;
      slot[i] = 0;
      if (i < it && k < n) {
        const uint32_t rel = e[i].x - dmin;
        slot[i] = atomicAdd(&U.hist[rel >> shift0], 1u);
        const uint64_t key = ((uint64_t)rel << 32) | (e[i].y - gbase);
        k_and &= key;
        k_or |= key;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      k_and &= __shfl_xor_sync(0xFFFFFFFFu, k_and, o);
      k_or |= __shfl_xor_sync(0xFFFFFFFFu, k_or, o);
    }
    if (lane == 0) {
      atomicAnd(&s_and, (unsigned long long)k_and);
      atomicOr(&s_or, (unsigned long long)k_or);
    }
  }
  __syncthreads();
  // exclusive scan of the bin counts: 8 bins per thread, warp scan, warp totals
  bool too_long = false;
  {
    uint32_t c[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      c[i] = U.hist[tid * 8 + i];
      too_long |= c[i] > kMaxRun;
      sum += c[i];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    uint32_t base = incl - sum;
    for (int w = 0; w < warp; w++) base += s_scan[w];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      U.hist[tid * 8 + i] = base;
      base += c[i];
    }
  }
  __syncthreads();
  // the fallback's second buffer is the chunk's own (already consumed) slice of inst in global memory
  uint64_t *a = s_buf, *b = reinterpret_cast<uint64_t *>(src);
#pragma unroll
  for (int i = 0; i < kSortItems; i++) {
    const uint32_t k = i * kSortThreads + tid;
    if (i < it && k < n) {
      const uint32_t rel = e[i].x - dmin;
      a[U.hist[rel >> shift0] + slot[i]] = ((uint64_t)rel << 32) | (e[i].y - gbase);
    }
  }
  too_long = __syncthreads_or(too_long);
  if (!too_long) {
    // local fix: every key ranks itself inside its bin on the full (depth bits, index) key -- keys are
    // unique, so bin start + #smaller keys of the bin is its final slot.  One key per thread and
    // iteration, bins hold one or two keys on average: short uniform loops instead of per-bin owners.
    const int ps = 32 + shift0;
    uint64_t x[kSortItems];
    uint32_t pos[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
      const uint32_t k = i * kSortThreads + tid;
      x[i] = 0;
      pos[i] = 0;
      if (i < it && k < n) {
        x[i] = a[k];
        const uint32_t bin = (uint32_t)(x[i] >> ps);
        const uint32_t start = U.hist[bin], end = bin + 1 < (uint32_t)kBins ? U.hist[bin + 1] : n;
        uint32_t rank = 0;
        for (uint32_t j = start; j < end; j++) rank += a[j] < x[i];
        pos[i] = start + rank;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; i++)
      if (i < it && i * kSortThreads + tid < n) a[pos[i]] = x[i];
  } else {
    const uint64_t vary = s_and ^ s_or;
    for (int shift = 0; shift < 64; shift += 8) {
      if (((vary >> shift) & 255ull) == 0) continue;   // digit constant over the chunk: identity pass
      chunk_radix_pass(U.cs, a, b, n, shift, tid);
      uint64_t *t = a; a = b; b = t;
    }
  }
  __syncthreads();
  if (single) {
    // the whole tile list: gather in sorted order
    const uint32_t ty = dgx.div(tile);
    const int tile_x0 = (int)(tile - ty * dgx.d) * kTile, tile_y0 = (int)ty * kTile;
    // A warp's 32 records are one contiguous 1536-byte block of the slab: they are staged in shared memory
    // (48-byte stride: conflict-free 128-bit stores) and leave as three fully coalesced warp stores instead
    // of three stores that each touch all 48 sectors of the block.
    // Gather of a warp's 32 records.  A thread loading "its" record's three float4 touches 32 different cache
    // lines per load instruction (96 L1 wavefronts per 32 records: the LSU data pipe was the kernel's busiest
    // unit).  Instead 4 lanes share a record -- lane piece p < 3 loads float4 p, so one load instruction covers 8
    // records with 8 wavefronts -- and the raw pieces meet in the warp's staging block, where the record's owner
    // lane computes the cull mask and rewrites the first two float4 in the blend kernels' order.
    float4 *stg = &s_stage[warp][0];
    const uint32_t piece = lane & 3u, sub = lane >> 2;
    float4 raw[4];
    auto load_raw = [&](uint32_t kb_) {
      const uint32_t cnt_ = min(32u, n - kb_);
#pragma unroll
      for (int rd = 0; rd < 4; rd++) {
        const uint32_t rec = rd * 8 + sub;
        raw[rd] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rec < cnt_ && piece < 3u) raw[rd] = geom[4 * ((size_t)gbase + (uint32_t)a[kb_ + rec]) + piece];
      }
    };
    if (warp * 32u < n) load_raw(warp * 32u);
    for (uint32_t kb = warp * 32; kb < n; kb += kSortThreads) {
      const uint32_t cnt = min(32u, n - kb);
#pragma unroll
      for (int rd = 0; rd < 4; rd++) {
        const uint32_t rec = rd * 8 + sub;
        if (rec < cnt && piece < 3u) stg[3 * rec + piece] = raw[rd];
      }
      // the loads of the warp's next 32 records fly while these are finished
      if (kb + kSortThreads < n) load_raw(kb + kSortThreads);
      __syncwarp();
      if (lane < cnt) {
        const uint32_t k = kb + lane;
        const uint64_t key = a[k];
        const uint32_t id = (uint32_t)key;
        const size_t r = (size_t)cstart + k;
        const float4 q0 = stg[3 * lane], q1 = stg[3 * lane + 1];
        masks[r] = (uint8_t)subblock_mask(q0, q1, tile_x0, tile_y0);
        if (dbg_keys) dbg_keys[r] = ((uint64_t)tile << 32) | ((uint32_t)(key >> 32) + dmin);
        if (dbg_plist) dbg_plist[r] = id;
        stg[3 * lane] = make_float4(q0.x, q0.y, q0.z, q1.x);       // {x, y, A, C}
        stg[3 * lane + 1] = make_float4(q0.w, q1.y, q1.z, 0.f);   // {B, opacity, thr, 0}
        reinterpret_cast<uint32_t *>(&stg[3 * lane + 2])[3] = id; // {r, g, b, id}
      }
      __syncwarp();
      float4 *out = records + 3 * ((size_t)cstart + kb);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const uint32_t idx = i * 32 + lane;
        if (idx < 3u * cnt) out[idx] = stg[idx];
      }
      __syncwarp();
    }
  } else {
    // one of several chunks of its tile: leave the sorted absolute keys in place for merge_gather
    // (a may already be that slice after an odd number of fallback passes: same index, in place)
    uint64_t *dst = reinterpret_cast<uint64_t *>(src);
    for (uint32_t k = tid; k < n; k += kSortThreads) dst[k] = a[k] + ((uint64_t)dmin << 32);
  }
}

template <int kSortThreads>
__global__ void __launch_bounds__(kSortThreads, 1024 / kSortThreads)
sort_chunks_kernel(int P, FastDiv dT, FastDiv dgx, const uint4 *__restrict__ items, const uint32_t *__restrict__ misc,
                   const uint2 *__restrict__ ranges, uint2 *inst, uint2 *inst_b, const float4 *__restrict__ geom,
                   float4 *__restrict__ records, uint8_t *__restrict__ masks, uint64_t *__restrict__ dbg_keys,
                   uint32_t *__restrict__ dbg_plist) {
  // (one item per CTA: taking items from a work counter in a loop was slower -- 91 vs 86 us per 8 views; the
  // hardware starts the next CTA faster than a CTA can fetch its next item)
  if (blockIdx.x >= misc[0]) return;
  const uint4 item = items[blockIdx.x];
  sort_item<kSortThreads>(item, ranges[item.x], P, dT, dgx, inst, inst_b, geom, records, masks, dbg_keys, dbg_plist);
}

// Lists of kChunk < n <= kBigChunk instances (two to four chunks: 2/3 of the instances of the two-hand scene)
// are ONE item of a 1024-thread CTA and final after it -- no key write-back, no rank merge, coalesced record
// stores.  The tiles come from the head of order[] (descending size class: every list longer than a chunk
// is there, misc[3] of them, longest first).
__global__ void __launch_bounds__(1024, 1)
sort_big_kernel(int P, FastDiv dT, FastDiv dgx, uint32_t big_max, const uint32_t *__restrict__ order,
                const uint32_t *__restrict__ tile_count, const uint32_t *__restrict__ misc,
                const uint2 *__restrict__ ranges, uint2 *inst, const float4 *__restrict__ geom,
                float4 *__restrict__ records, uint8_t *__restrict__ masks, uint64_t *__restrict__ dbg_keys,
                uint32_t *__restrict__ dbg_plist) {
  if (blockIdx.x >= misc[3]) return;
  const uint32_t vt = order[blockIdx.x], c = tile_count[vt];
  if (c <= (uint32_t)kChunk || c > big_max) return;    // (longer lists: depth partition or chunks + merge)
  sort_item<1024>(make_uint4(vt, 0u, c, kItemFinal), ranges[vt], P, dT, dgx, inst, inst, geom, records, masks, dbg_keys,
                  dbg_plist);
}

// Multi-chunk tiles: one CTA per chunk.  Final position of a key = its index in its own (sorted) chunk
// + the number of smaller keys in every other chunk of the tile (keys are unique: (depth bits, index)).
template <int kSortThreads>
__global__ void __launch_bounds__(kSortThreads, 1024 / kSortThreads)
merge_gather_kernel(int P, FastDiv dT, FastDiv dgx, const uint4 *__restrict__ items, const uint32_t *__restrict__ misc,
                    const uint2 *__restrict__ ranges, const uint2 *__restrict__ inst, const float4 *__restrict__ geom,
                    float4 *__restrict__ records, uint8_t *__restrict__ masks, uint64_t *__restrict__ dbg_keys,
                    uint32_t *__restrict__ dbg_plist) {
  constexpr int kChunk = kSortThreads * kSortItems;
  __shared__ __align__(16) uint64_t s_other[kChunk];
  // grid-stride over the item list: with the grid at its upper bound every CTA has one item; when merge items
  // are the exception (lists up to kBigChunk go to sort_big_kernel) a small grid skims the list instead of
  // thousands of CTAs being launched to find out that their item is final
  if (misc[4] == 0u) return;                           // no list needs the rank merge (the usual case)
  const uint32_t n_items = misc[0];
  for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
  const uint4 item = items[it];
  if (item.w & kItemFinal) continue;                   // light tile or depth bucket: finished by sort_chunks
  const uint32_t vt = item.x;
  const uint2 range = ranges[vt];
  const uint32_t nt = range.y - range.x;
  const uint32_t m = (nt + kChunk - 1) / kChunk;       // the tile's plain chunks (plan_kernel's fallback)
  const uint32_t own = item.y / kChunk;
  const uint32_t cstart = range.x + item.y;
  if (cstart >= range.y) continue;
  const uint32_t n = min((uint32_t)kChunk, range.y - cstart);
  const int tid = threadIdx.x;
  const uint32_t v = dT.div(vt), tile = vt - v * dT.d, gbase = v * (uint32_t)P;
  const uint64_t *keys = reinterpret_cast<const uint64_t *>(inst);
  // a thread owns keys k = i*256 + tid (consecutive lanes = consecutive keys, so the final positions
  // of a warp are nearly consecutive and the record stores coalesce); its keys ascend with i, and so
  // do their lower bounds in another sorted chunk
  uint64_t key[kSortItems];
  uint32_t pos[kSortItems];
#pragma unroll
  for (int i = 0; i < kSortItems; i++) {
    const uint32_t k = i * kSortThreads + tid;
    key[i] = k < n ? keys[cstart + k] : ~0ull;
    pos[i] = k;
  }
  for (uint32_t c = 0; c < m; c++) {
    if (c == own) continue;
    const uint32_t ostart = range.x + c * kChunk, on = min((uint32_t)kChunk, range.y - ostart);
    __syncthreads();
    for (uint32_t k = tid; k < on; k += kSortThreads) s_other[k] = keys[ostart + k];
    __syncthreads();
    // lower bound of every key in s_other[0, on): branch-free halving with the thread's kSortItems
    // searches advancing in lockstep, so their dependent shared-memory probes overlap
    uint32_t lb[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; i++) lb[i] = 0;
#pragma unroll 1
    for (uint32_t step = kChunk / 2; step > 0; step >>= 1) {
#pragma unroll
      for (int i = 0; i < kSortItems; i++) {
        const uint32_t probe = lb[i] + step;
        if (probe <= on && s_other[probe - 1] < key[i]) lb[i] = probe;
      }
    }
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
      // (the halving covers [0, kChunk): one last probe settles a full other chunk's final element)
      if (lb[i] < on && s_other[lb[i]] < key[i]) lb[i]++;
      pos[i] += lb[i];
    }
  }
  const uint32_t ty = dgx.div(tile);
  const int tile_x0 = (int)(tile - ty * dgx.d) * kTile, tile_y0 = (int)ty * kTile;
#pragma unroll
  for (int i = 0; i < kSortItems; i++) {
    const uint32_t k = i * kSortThreads + tid;
    if (k < n)
      emit_instance((size_t)range.x + pos[i], (uint32_t)(key[i] >> 32), (uint32_t)key[i], gbase, tile, tile_x0, tile_y0,
                    geom, records, masks, dbg_keys, dbg_plist);
  }
  }
}

// Geometry reuse: the earlier call's status with this call's sequence number and an empty backward unit list
__global__ void reuse_status_kernel(const GhrStatus *old, GhrStatus *st, uint64_t seq) {
  GhrStatus v = *old;
  v.reserved[0] = seq;
  v.reserved[1] = 0;
  *st = v;
}

// Geometry reuse: sorted records and cull masks are copied (they depend on geometry only); the colour of
// every record comes from this call's geometry block.  One thread per sorted instance, R from the status.
__global__ void __launch_bounds__(256)
recolor_records_kernel(uint64_t R_cap, const GhrStatus *__restrict__ old_status, const float4 *__restrict__ old_rec,
                       const uint8_t *__restrict__ old_masks, const float4 *__restrict__ geom,
                       float4 *__restrict__ rec, uint8_t *__restrict__ masks) {
  const uint64_t R = old_status->R < R_cap ? old_status->R : R_cap;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (uint64_t)gridDim.x * blockDim.x) {
    const float4 q0 = old_rec[3 * r], q1 = old_rec[3 * r + 1];
    float4 q2 = old_rec[3 * r + 2];
    const float4 c = geom[4 * (size_t)__float_as_uint(q2.w) + 2];
    q2.x = c.x;
    q2.y = c.y;
    q2.z = c.z;
    rec[3 * r] = q0;
    rec[3 * r + 1] = q1;
    rec[3 * r + 2] = q2;
    masks[r] = old_masks[r];
  }
}

}  // namespace

// Lists longer than kPartMin are depth-partitioned instead of rank-merged.  The partition passes are
// launched only when the instance capacity says long lists are likely (average list >= kChunk; the
// Python layer sizes R_cap from the instance counts it has seen): at the two-hand sizes no list reaches
// kPartMin and four empty launches per forward would be pure overhead.  Either way every list is
// handled exactly -- the merge has no length limit.  GHR_PARTITION=0/1 forces the choice (A/B).
constexpr int kBigChunk = 8192;       // longest list one 1024-thread CTA sorts (sort_big_kernel)
// A big item occupies a whole SM for ~20 us; with few of them (a single view: ~45) most SMs idle and the launch
// lasts as long as its longest list, where 2048-instance chunks + the rank merge spread the same lists over five
// times as many CTAs (single view: 37 vs 46 us).  The big items pay off once they fill the machine twice.
static uint32_t big_max(const GhrDims &d) { return (uint64_t)d.R_cap / kChunk >= 2u * 148u ? (uint32_t)kBigChunk : 0u; }
static uint32_t part_min(const GhrDims &d) { return big_max(d) ? big_max(d) : (uint32_t)kChunk; }
static bool use_partition(const GhrDims &d, const Layout &L) {
  static const int forced = [] { const char *e = getenv("GHR_PARTITION"); return e ? atoi(e) : -1; }();
  if (forced >= 0) return forced != 0;
  return (uint64_t)d.R_cap >= (uint64_t)d.V * L.T * kChunk;
}

cudaError_t launch_reuse_binning(const GhrDims &d, const Layout &L, const Layout &Lold, const char *old_state,
                                 char *state, uint64_t seq, cudaStream_t s) {
  const size_t VT = (size_t)d.V * L.T;
  reuse_status_kernel<<<1, 1, 0, s>>>((const GhrStatus *)(old_state + Lold.pub.off_status),
                                      (GhrStatus *)(state + L.pub.off_status), seq);
  cudaError_t e = cudaMemcpyAsync(state + L.pub.off_ranges, old_state + Lold.pub.off_ranges, VT * 8, cudaMemcpyDeviceToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(state + L.pub.off_order, old_state + Lold.pub.off_order, VT * 4, cudaMemcpyDeviceToDevice, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(state + L.pub.off_tilemax, 0, VT * 8, s);
  if (e != cudaSuccess) return e;
  if (d.R_cap > 0) {
    int grid = (int)((d.R_cap + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    recolor_records_kernel<<<grid, 256, 0, s>>>((uint64_t)d.R_cap, (const GhrStatus *)(old_state + Lold.pub.off_status),
                                                (const float4 *)(old_state + Lold.pub.off_records),
                                                (const uint8_t *)(old_state + Lold.pub.off_masks),
                                                (const float4 *)(state + L.pub.off_geom),
                                                (float4 *)(state + L.pub.off_records), (uint8_t *)(state + L.pub.off_masks));
  }
  return cudaGetLastError();
}

cudaError_t launch_tile_scan_schedule(const GhrDims &d, const Layout &L, char *state, char *temp, uint64_t seq,
                                      cudaStream_t s) {
  const int VT = d.V * L.T;
  if (VT == 0) return cudaSuccess;
  // one CTA per ~1024 tiles (every CTA reads all counts first: bounded so that pass stays short)
  const int scan_ctas = std::min(32, std::max(1, (VT + kScanThreads1 - 1) / kScanThreads1));
  tile_scan_schedule_kernel<<<scan_ctas, kScanThreads1, 0, s>>>(
      VT, (uint64_t)d.R_cap, (uint32_t)L.n_chunks, (uint32_t)L.n_hchunks, (uint32_t)L.n_heavy,
      use_partition(d, L) ? part_min(d) : 0xFFFFFFFFu, big_max(d),
      (const uint32_t *)(temp + L.t_tile_count), (uint2 *)(state + L.pub.off_ranges),
      (uint32_t *)(state + L.pub.off_order), (uint4 *)(temp + L.t_chunks), (uint2 *)(temp + L.t_hchunks),
      (uint32_t *)(temp + L.t_heavy), (uint32_t *)(temp + L.t_heavy_id), (uint32_t *)(temp + L.t_misc),
      (GhrStatus *)(state + L.pub.off_status), seq);
  return cudaGetLastError();
}

cudaError_t launch_duplicate(const GhrDims &d, const Layout &L, char *state, char *temp, cudaStream_t s) {
  if (d.P == 0 || d.R_cap <= 0) return cudaSuccess;
  dim3 grid((d.P + 255) / 256, d.V), block(256);
  const int smem_tiles = L.T <= kMaxSmemTiles ? L.T : 0;
  const size_t smem = (size_t)smem_tiles * 8;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(duplicate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  duplicate_kernel<<<grid, block, smem, s>>>(d.P, L.gx, L.gy, L.T, smem_tiles, (uint64_t)d.R_cap,
                                             (const float4 *)(state + L.pub.off_geom),
                                             (const uint2 *)(state + L.pub.off_ranges),
                                             (uint32_t *)(temp + L.t_cursor), (uint2 *)(temp + L.t_inst));
  return cudaGetLastError();
}

cudaError_t launch_sort_gather(const GhrDims &d, const Layout &L, char *state, char *temp, uint64_t *dbg_keys,
                               uint32_t *dbg_plist, cudaStream_t s) {
  const int VT = d.V * L.T;
  if (VT == 0 || d.R_cap <= 0) return cudaSuccess;
  const FastDiv dT = make_fastdiv((uint32_t)L.T), dgx = make_fastdiv((uint32_t)L.gx);
  const bool partition = use_partition(d, L);
  const uint32_t *misc = (const uint32_t *)(temp + L.t_misc);
  const uint2 *ranges = (const uint2 *)(state + L.pub.off_ranges);
  uint2 *inst = (uint2 *)(temp + L.t_inst), *inst_b = (uint2 *)(temp + L.t_inst_b);
  uint4 *items = (uint4 *)(temp + L.t_chunks);
  if (partition) {
    const uint2 *hchunks = (const uint2 *)(temp + L.t_hchunks);
    const uint32_t *heavy_id = (const uint32_t *)(temp + L.t_heavy_id);
    uint2 *dmm = (uint2 *)(temp + L.t_dmm);
    uint32_t *slab_count = (uint32_t *)(temp + L.t_slab_count), *slab_off = (uint32_t *)(temp + L.t_slab_off);
    uint32_t *heavy_flag = (uint32_t *)(temp + L.t_heavy_flag);
    // grids are upper bounds (the scan wrote the exact counts to misc[]), surplus CTAs exit on their first
    // instruction
    const int hgrid = (int)L.n_hchunks;
    heavy_minmax_kernel<<<hgrid, kPartThreads, 0, s>>>(hchunks, misc, ranges, heavy_id, inst, dmm);
    heavy_slab_kernel<false><<<hgrid, kPartThreads, 0, s>>>(hchunks, misc, ranges, heavy_id, inst, dmm, slab_count,
                                                            slab_off, heavy_flag, inst_b);
    heavy_plan_kernel<<<(int)L.n_heavy, kSlabs, 0, s>>>((const uint32_t *)(temp + L.t_heavy), ranges, slab_count,
                                                        slab_off, heavy_flag, items, (uint32_t *)(temp + L.t_misc),
                                                        (uint32_t)L.n_chunks, 0);
    heavy_slab_kernel<true><<<hgrid, kPartThreads, 0, s>>>(hchunks, misc, ranges, heavy_id, inst, dmm, slab_count,
                                                           slab_off, heavy_flag, inst_b);
  }
  // chunk sort of every item (light tiles and depth buckets are final), rank merge for the fallback tiles
  const int grid = partition ? (int)L.n_chunks : (int)((size_t)d.R_cap / kChunk + VT + 1);
  constexpr int kThreads = 256;
  const size_t sort_smem = SortSmem<kThreads>::kBytes, big_smem = SortSmem<1024>::kBytes;
  cudaError_t e = cudaFuncSetAttribute(sort_chunks_kernel<kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sort_smem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(sort_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_smem);
  if (e != cudaSuccess) return e;
  // lists of two to four chunks first (the longest work items), one CTA each
  size_t big_grid = (size_t)d.R_cap / kChunk + 1;
  if (big_grid > (size_t)VT) big_grid = (size_t)VT;
  if (big_max(d))
    sort_big_kernel<<<(int)big_grid, 1024, big_smem, s>>>(
      d.P, dT, dgx, big_max(d), (const uint32_t *)(state + L.pub.off_order),
      (const uint32_t *)(temp + L.t_tile_count), misc, ranges, inst, (const float4 *)(state + L.pub.off_geom),
      (float4 *)(state + L.pub.off_records), (uint8_t *)(state + L.pub.off_masks), dbg_keys, dbg_plist);
  sort_chunks_kernel<kThreads><<<grid, kThreads, sort_smem, s>>>(
      d.P, dT, dgx, items, misc, ranges, inst, inst_b, (const float4 *)(state + L.pub.off_geom),
      (float4 *)(state + L.pub.off_records), (uint8_t *)(state + L.pub.off_masks), dbg_keys, dbg_plist);
  const int merge_grid = big_max(d) && grid > 148 * 4 ? 148 * 4 : grid;
  merge_gather_kernel<kThreads><<<merge_grid, kThreads, 0, s>>>(
      d.P, dT, dgx, items, misc, ranges, inst, (const float4 *)(state + L.pub.off_geom),
      (float4 *)(state + L.pub.off_records), (uint8_t *)(state + L.pub.off_masks), dbg_keys, dbg_plist);
  return cudaGetLastError();
}

}  // namespace ghr
