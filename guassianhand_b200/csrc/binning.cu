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
//   sort_chunks        : every tile list is cut into chunks of kChunk instances; one CTA sorts one chunk by
//                        (depth bits, id) in shared memory (two radix passes on the 16 leading significant
//                        bits + a local fix).  A single-chunk tile is finished here: its 48-byte geometry
//                        records are gathered in sorted order with their sub-block cull masks.
//   merge_gather       : for a tile of several chunks, one CTA per chunk ranks its keys in the other
//                        (sorted) chunks by binary search -- final position = own index + ranks -- and
//                        gathers its records straight to their final place.
// No global sort, no cross-block prefix: all work items are uniform chunks, independent after the scan.  Result: bit for
// bit the upstream order (checked against the oracle's sorted keys / point list / ranges).
#include <cstdlib>

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

// ranges[t] = [start, end) of tile t in the instance arrays (clamped to R_cap; (0,0) for empty tiles,
// as upstream leaves them), order[] = tile ids by descending size class (the blend launch order),
// chunks[] = (tile, chunk index) work items of the sort, misc[0] = their number, status = {R, overflow}.
// One CTA: every thread owns a run of consecutive tiles, sums it, one block scan of the 1024 run
// totals, then walks its run again.
__global__ void __launch_bounds__(kScanThreads1)
tile_scan_schedule_kernel(int VT, uint64_t R_cap, uint32_t chunk_cap, int chunk_log2,
                          const uint32_t *__restrict__ tile_count,
                          uint2 *__restrict__ ranges, uint32_t *__restrict__ order, uint2 *__restrict__ chunks,
                          uint32_t *__restrict__ misc, GhrStatus *__restrict__ status) {
  __shared__ uint64_t s_warp[kScanThreads1 / 32];
  __shared__ uint32_t s_wchunk[kScanThreads1 / 32];
  __shared__ uint32_t s_cls[kClasses + 1], s_cur[kClasses + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < kClasses) s_cls[tid] = 0;
  const int per = (VT + kScanThreads1 - 1) / kScanThreads1;
  const int t0 = min(VT, tid * per), t1 = min(VT, t0 + per);
  const uint32_t chunk_mask = (1u << chunk_log2) - 1u;
  uint64_t sum = 0;
  uint32_t csum = 0;
  for (int t = t0; t < t1; t++) {
    const uint32_t c = tile_count[t];
    sum += c;
    csum += (c + chunk_mask) >> chunk_log2;
  }
  uint64_t incl = sum;
  uint32_t cincl = csum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    const uint32_t cup = __shfl_up_sync(0xFFFFFFFFu, cincl, o);
    if (lane >= o) {
      incl += up;
      cincl += cup;
    }
  }
  if (lane == 31) {
    s_warp[warp] = incl;
    s_wchunk[warp] = cincl;
  }
  __syncthreads();
  uint64_t start = incl - sum, total = 0;
  uint32_t coff = cincl - csum, ctotal = 0;
  for (int w = 0; w < kScanThreads1 / 32; w++) {
    if (w < warp) {
      start += s_warp[w];
      coff += s_wchunk[w];
    }
    total += s_warp[w];
    ctotal += s_wchunk[w];
  }
  // class counters are warp-aggregated (match.any): most tiles are empty, and 5000 shared-memory atomics
  // on one address would serialise
  for (int j = 0; j < per; j++) {
    const int t = t0 + j;
    const bool valid = t < t1;
    const uint32_t c = valid ? tile_count[t] : 0u;
    const int cls = valid ? size_class(c) : kClasses;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cls);
    if (valid) {
      const uint64_t end = start + c;
      ranges[t] = c ? make_uint2((uint32_t)(start < R_cap ? start : R_cap), (uint32_t)(end < R_cap ? end : R_cap))
                    : make_uint2(0u, 0u);
      if (lane == __ffs(peers) - 1) atomicAdd(&s_cls[cls], (uint32_t)__popc(peers));
      const uint32_t m = (c + chunk_mask) >> chunk_log2;
      for (uint32_t q = 0; q < m && coff + q < chunk_cap; q++) chunks[coff + q] = make_uint2((uint32_t)t, q);
      coff += m;
      start = end;
    }
  }
  __syncthreads();
  if (tid == 0) {
    status->R = total;
    status->overflow = total > R_cap ? 1u : 0u;
    misc[0] = ctotal < chunk_cap ? ctotal : chunk_cap;
    uint32_t acc = 0;
    for (int c = kClasses - 1; c >= 0; c--) {     // largest class first; empty tiles (class 0) last
      s_cur[c] = acc;
      acc += s_cls[c];
    }
  }
  __syncthreads();
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int j = 0; j < per; j++) {
    const int t = t0 + j;
    const bool valid = t < t1;
    const int cls = valid ? size_class(tile_count[t]) : kClasses;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cls);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) base = atomicAdd(&s_cur[cls], (uint32_t)__popc(peers));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (valid) order[base + __popc(peers & lt_mask)] = (uint32_t)t;
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
    const float4 q3 = geom[4 * (size_t)gid + 3];
    if (__float_as_uint(q3.y)) {
      const float4 q0 = geom[4 * (size_t)gid];
      depth_bits = __float_as_uint(geom[4 * (size_t)gid + 2].w);
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
  records[3 * r] = q0;
  records[3 * r + 1] = q1;
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
template <int kSortThreads>
__global__ void __launch_bounds__(kSortThreads, 1024 / kSortThreads)
sort_chunks_kernel(int P, FastDiv dT, FastDiv dgx, const uint2 *__restrict__ chunks, const uint32_t *__restrict__ misc,
                   const uint2 *__restrict__ ranges, uint2 *inst, const float4 *__restrict__ geom,
                   float4 *__restrict__ records, uint8_t *__restrict__ masks, uint64_t *__restrict__ dbg_keys,
                   uint32_t *__restrict__ dbg_plist) {
  constexpr int kSortWarps = kSortThreads / 32;
  constexpr int kChunk = kSortThreads * kSortItems;
  constexpr int kBins = kChunk;                       // 8 bins per thread in the bin scan
  constexpr int kBinBits = kSortThreads == 256 ? 11 : 12;
  static_assert(kBins == 1 << kBinBits, "chunk size");
  extern __shared__ __align__(16) uint64_t s_buf[];   // [kChunk] keys
  __shared__ union {
    ChunkSort<kSortThreads> cs;   // fallback passes only
    uint32_t hist[kBins];         // bin counts, then bin offsets
  } U;
  __shared__ uint32_t s_scan[kSortWarps];
  __shared__ uint32_t s_dmin, s_dmax;
  __shared__ unsigned long long s_and, s_or;
  if (blockIdx.x >= misc[0]) return;
  const uint2 chunk = chunks[blockIdx.x];
  const uint32_t vt = chunk.x;
  const uint2 range = ranges[vt];
  const uint32_t cstart = range.x + chunk.y * kChunk;
  if (cstart >= range.y) return;                       // chunk of a list clamped by an instance-capacity overflow
  const uint32_t n = min((uint32_t)kChunk, range.y - cstart);
  const bool single = range.y - range.x <= (uint32_t)kChunk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t v = dT.div(vt), tile = vt - v * dT.d, gbase = v * (uint32_t)P;
  uint2 *src = inst + cstart;
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
      e[i] = k < n ? src[k] : make_uint2(0u, 0u);
      if (k < n) {
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
      if (k < n) {
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
    if (k < n) {
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
      if (k < n) {
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
      if (i * kSortThreads + tid < n) a[pos[i]] = x[i];
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
#pragma unroll 2
    for (uint32_t k = tid; k < n; k += kSortThreads) {
      const uint64_t key = a[k];
      emit_instance((size_t)cstart + k, (uint32_t)(key >> 32) + dmin, (uint32_t)key, gbase, tile, tile_x0, tile_y0,
                    geom, records, masks, dbg_keys, dbg_plist);
    }
  } else {
    // one of several chunks of its tile: leave the sorted absolute keys in place for merge_gather
    // (a may already be that slice after an odd number of fallback passes: same index, in place)
    uint64_t *dst = reinterpret_cast<uint64_t *>(src);
    for (uint32_t k = tid; k < n; k += kSortThreads) dst[k] = a[k] + ((uint64_t)dmin << 32);
  }
}

// Multi-chunk tiles: one CTA per chunk.  Final position of a key = its index in its own (sorted) chunk
// + the number of smaller keys in every other chunk of the tile (keys are unique: (depth bits, index)).
template <int kSortThreads>
__global__ void __launch_bounds__(kSortThreads, 1024 / kSortThreads)
merge_gather_kernel(int P, FastDiv dT, FastDiv dgx, const uint2 *__restrict__ chunks, const uint32_t *__restrict__ misc,
                    const uint2 *__restrict__ ranges, const uint2 *__restrict__ inst, const float4 *__restrict__ geom,
                    float4 *__restrict__ records, uint8_t *__restrict__ masks, uint64_t *__restrict__ dbg_keys,
                    uint32_t *__restrict__ dbg_plist) {
  constexpr int kChunk = kSortThreads * kSortItems;
  __shared__ __align__(16) uint64_t s_other[kChunk];
  if (blockIdx.x >= misc[0]) return;
  const uint2 chunk = chunks[blockIdx.x];
  const uint32_t vt = chunk.x;
  const uint2 range = ranges[vt];
  const uint32_t nt = range.y - range.x;
  if (nt <= (uint32_t)kChunk) return;                  // single-chunk tile: finished by sort_chunks
  const uint32_t m = (nt + kChunk - 1) / kChunk;
  const uint32_t cstart = range.x + chunk.y * kChunk;
  if (cstart >= range.y) return;
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
    if (c == chunk.y) continue;
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

__global__ void init_status_kernel(GhrStatus *st, GhrStatus v) { *st = v; }

}  // namespace

// log2 of the sort chunk size: 11 (2048, 256-thread CTAs) or 12 (4096, 512-thread CTAs); the layout's
// chunk-list capacity is sized for the smaller one.  GHR_CHUNK_LOG2 overrides (A/B only).
static int chunk_log2() {
  static const int v = [] {
    const char *e = getenv("GHR_CHUNK_LOG2");
    const int x = e ? atoi(e) : 11;
    return x == 11 ? 11 : 12;
  }();
  return v;
}

cudaError_t launch_init_status(char *status, GhrStatus st0, cudaStream_t s) {
  init_status_kernel<<<1, 1, 0, s>>>((GhrStatus *)status, st0);
  return cudaGetLastError();
}

cudaError_t launch_tile_scan_schedule(const GhrDims &d, const Layout &L, char *state, char *temp, cudaStream_t s) {
  const int VT = d.V * L.T;
  if (VT == 0) return cudaSuccess;
  tile_scan_schedule_kernel<<<1, kScanThreads1, 0, s>>>(
      VT, (uint64_t)d.R_cap, (uint32_t)L.n_chunks, chunk_log2(), (const uint32_t *)(temp + L.t_tile_count),
      (uint2 *)(state + L.pub.off_ranges), (uint32_t *)(state + L.pub.off_order), (uint2 *)(temp + L.t_chunks),
      (uint32_t *)(temp + L.t_misc), (GhrStatus *)(state + L.pub.off_status));
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
  // grid = upper bound of the chunk count (the scan wrote the exact one to misc[0]); surplus CTAs exit
  const int grid = (int)L.n_chunks;
  auto launch = [&](auto sort_k, auto merge_k, int threads) -> cudaError_t {
    const size_t sort_smem = (size_t)threads * kSortItems * 8;
    cudaError_t e = cudaFuncSetAttribute(sort_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem);
    if (e != cudaSuccess) return e;
    sort_k<<<grid, threads, sort_smem, s>>>(
        d.P, dT, dgx, (const uint2 *)(temp + L.t_chunks), (const uint32_t *)(temp + L.t_misc),
        (const uint2 *)(state + L.pub.off_ranges), (uint2 *)(temp + L.t_inst),
        (const float4 *)(state + L.pub.off_geom), (float4 *)(state + L.pub.off_records),
        (uint8_t *)(state + L.pub.off_masks), dbg_keys, dbg_plist);
    merge_k<<<grid, threads, 0, s>>>(
        d.P, dT, dgx, (const uint2 *)(temp + L.t_chunks), (const uint32_t *)(temp + L.t_misc),
        (const uint2 *)(state + L.pub.off_ranges), (const uint2 *)(temp + L.t_inst),
        (const float4 *)(state + L.pub.off_geom), (float4 *)(state + L.pub.off_records),
        (uint8_t *)(state + L.pub.off_masks), dbg_keys, dbg_plist);
    return cudaSuccess;
  };
  cudaError_t e = chunk_log2() == 12 ? launch(sort_chunks_kernel<512>, merge_gather_kernel<512>, 512)
                                     : launch(sort_chunks_kernel<256>, merge_gather_kernel<256>, 256);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace ghr
