// Tile binning around the sorts:
//   scan_duplicate : single-pass prefix sum (decoupled look-back) of tiles_touched taken in
//                    depth-sorted order, fused with the emission of one ((view,tile), id) pair per
//                    touched tile and with the digit histograms of the following tile sort.
//                    Replaces upstream InclusiveSum + blocking D2H of num_rendered +
//                    duplicateWithKeys (SURVEY.md §2a K2,K3; A.4).  R stays on the device.
//   gather_ranges  : after the tile sort, copies each instance's 48-byte geometry record into
//                    sorted order (so the blend kernels stream contiguous slabs with bulk async
//                    copies) and marks tile range boundaries.  Replaces identifyTileRanges (K5).
#include "ghr_internal.cuh"

namespace ghr {

namespace {

constexpr uint64_t kScanLocal = 1ull << 62;
constexpr uint64_t kScanIncl = 2ull << 62;
constexpr uint64_t kScanMask = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_volatile64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile64(uint64_t *p, uint64_t v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void rect_of(float px, float py, int radius, int gx, int gy, int &minx, int &miny,
                                        int &maxx, int &maxy) {
  float rf = (float)radius;
  // x / 16 == x * 0.0625 bit for bit (power of two)
  minx = min(gx, max(0, (int)fmul(fsub(px, rf), 0.0625f)));
  miny = min(gy, max(0, (int)fmul(fsub(py, rf), 0.0625f)));
  maxx = min(gx, max(0, (int)fmul(fsub(fadd(fadd(px, rf), 16.0f), 1.0f), 0.0625f)));
  maxy = min(gy, max(0, (int)fmul(fsub(fadd(fadd(py, rf), 16.0f), 1.0f), 0.0625f)));
}

__global__ void __launch_bounds__(kScanThreads)
scan_duplicate_kernel(int P, FastDiv dP, int V, int gx, int gy, int T, int npt, uint64_t R_cap,
                      const float4 *__restrict__ geom, const uint32_t *__restrict__ order /*[V,P] depth-sorted ids*/,
                      uint64_t *__restrict__ scan_status, uint32_t *__restrict__ ticket,
                      uint32_t *__restrict__ tkeys, uint32_t *__restrict__ tvals, uint32_t *__restrict__ thist,
                      GhrStatus *__restrict__ status, uint32_t nblk) {
  constexpr int kItems = kScanItems;
  constexpr int kBlockItems = kScanThreads * kItems;
  __shared__ uint32_t s_hist[4][256];
  __shared__ uint64_t s_warp[kScanThreads / 32];
  __shared__ uint64_t s_prefix;
  __shared__ uint32_t s_blk;
  __shared__ uint32_t s_vis[kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_blk = atomicAdd(ticket, 1u);
  for (int k = tid; k < 4 * 256; k += kScanThreads) (&s_hist[0][0])[k] = 0;
  __syncthreads();
  const uint32_t blk = s_blk;
  const uint64_t total_elems = (uint64_t)V * P;

  // each thread owns kItems consecutive elements of the (view-major, depth-sorted) sequence
  uint32_t tt[kItems];
  uint32_t gid[kItems];
  float2 xy[kItems];
  int rad[kItems];
  uint32_t sum = 0, nvis = 0;
#pragma unroll
  for (int k = 0; k < kItems; k++) {
    uint64_t e = (uint64_t)blk * kBlockItems + (uint64_t)tid * kItems + k;
    tt[k] = 0;
    gid[k] = 0;
    rad[k] = 0;
    xy[k] = make_float2(0.f, 0.f);
    if (e < total_elems) {
      uint32_t v = dP.div((uint32_t)e);
      uint32_t id = order[e];
      uint32_t g = v * (uint32_t)P + id;
      float4 q3 = geom[4 * (size_t)g + 3];
      tt[k] = __float_as_uint(q3.y);
      if (tt[k]) {
        float4 q0 = geom[4 * (size_t)g + 0];
        xy[k] = make_float2(q0.x, q0.y);
        rad[k] = __float_as_int(q3.x);
        gid[k] = g;
        nvis++;
      }
    }
    sum += tt[k];
  }
  // block exclusive scan of per-thread sums
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  uint32_t wv = nvis;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wv += __shfl_xor_sync(0xFFFFFFFFu, wv, o);
  if (lane == 31) s_warp[warp] = incl;
  if (lane == 0) s_vis[warp] = wv;
  __syncthreads();
  uint64_t wbase = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; w++) {
    if (w < warp) wbase += s_warp[w];
    block_total += s_warp[w];
  }
  // decoupled look-back on block totals (warp 0)
  if (warp == 0) {
    uint64_t excl = 0;
    if (blk == 0) {
      if (lane == 0) st_volatile64(&scan_status[0], kScanIncl | block_total);
    } else {
      if (lane == 0) st_volatile64(&scan_status[blk], kScanLocal | block_total);
      if (nblk <= 2048u) {
        // every block is resident at once: a look-back would be a serial chain over the blocks, so the
        // warp sums the local totals of ALL predecessors directly (blk/32 independent loads per lane;
        // they only wait for the predecessors' block sums, never for their look-back)
        uint64_t part = 0;
        for (int64_t j = (int64_t)blk - 1 - lane; j >= 0; j -= 32) {
          uint64_t sv;
          do { sv = ld_volatile64(&scan_status[j]); } while ((sv >> 62) == 0);
          part += sv & kScanMask;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
        excl = part;
      } else {
      int64_t b = (int64_t)blk - 1;
      while (true) {
        int64_t mine = b - lane;
        uint64_t sv = 0;
        if (mine >= 0) {
          do { sv = ld_volatile64(&scan_status[mine]); } while ((sv >> 62) == 0);
        } else {
          sv = kScanIncl;   // virtual predecessor of block 0 with value 0
        }
        uint32_t incl_mask = __ballot_sync(0xFFFFFFFFu, (sv >> 62) == 2);
        int first = incl_mask ? (__ffs(incl_mask) - 1) : 32;
        uint64_t contrib = (lane <= first) ? (sv & kScanMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, o);
        excl += contrib;
        if (incl_mask) break;
        b -= 32;
      }
      if (lane == 0) st_volatile64(&scan_status[blk], kScanIncl | (excl + block_total));
      }
    }
    if (lane == 0) {
      s_prefix = excl;
      uint32_t bv = 0;
      for (int w = 0; w < kScanThreads / 32; w++) bv += s_vis[w];
      if (bv) atomicAdd(&status->n_visible, bv);
      if (blk == nblk - 1) {
        uint64_t Rtot = excl + block_total;
        status->R = Rtot;
        status->overflow = Rtot > R_cap ? 1u : 0u;
      }
    }
  }
  __syncthreads();
  uint64_t off = s_prefix + wbase + (incl - sum);

  // emission: row-major over the tile rectangle (A.4), in depth-sorted Gaussian order.  Histogram
  // of the tile key's digits for the tile sort: the low digit per instance; the upper digits change
  // rarely along a thread's emissions (same view, neighbouring tiles), so they are counted as runs
  // and flushed once per change instead of one (warp-wide conflicting) shared atomic per instance.
  uint32_t run_hi = 0xFFFFFFFFu, run_cnt = 0;
  auto flush = [&]() {
    if (run_cnt)
      for (int p = 1; p < npt; p++) atomicAdd(&s_hist[p][(run_hi >> (8 * (p - 1))) & 255u], run_cnt);
    run_cnt = 0;
  };
#pragma unroll
  for (int k = 0; k < kItems; k++) {
    if (tt[k]) {
      int minx, miny, maxx, maxy;
      rect_of(xy[k].x, xy[k].y, rad[k], gx, gy, minx, miny, maxx, maxy);
      uint32_t v = dP.div(gid[k]);
      uint32_t tbase = v * (uint32_t)T;
      for (int y = miny; y < maxy; y++)
        for (int x = minx; x < maxx; x++) {
          uint32_t tk = tbase + (uint32_t)(y * gx + x);
          if (off < R_cap) {
            tkeys[off] = tk;
            tvals[off] = gid[k];
            atomicAdd(&s_hist[0][tk & 255u], 1u);
            if ((tk >> 8) != run_hi) {
              flush();
              run_hi = tk >> 8;
            }
            run_cnt++;
          }
          off++;
        }
    }
  }
  flush();
  __syncthreads();
  for (int k = tid; k < npt * 256; k += kScanThreads) {
    uint32_t c = (&s_hist[0][0])[k];
    if (c) atomicAdd(&thist[k], c);
  }
}

// One thread per sorted instance: loads the instance's 48-byte geometry record (3 independent
// 128-bit loads), computes the 8-bit mask of 8x4 sub-blocks of its tile the alpha >= 1/255 ellipse
// reaches (subblock_mask; compact byte array staged by the blend kernels next to the records),
// marks tile range boundaries, and stages the record in shared memory so that the block writes the
// sorted slab with fully coalesced 128-bit stores.
constexpr int kGatherThreads = 256;
__global__ void __launch_bounds__(kGatherThreads)
gather_ranges_kernel(FastDiv dP, FastDiv dT, FastDiv dgx, uint64_t R_cap,
                     const GhrStatus *__restrict__ status, const uint32_t *__restrict__ tkeys,
                     const uint32_t *__restrict__ tvals, const float4 *__restrict__ geom,
                     float4 *__restrict__ records, uint8_t *__restrict__ masks, uint2 *__restrict__ ranges,
                     uint64_t *__restrict__ dbg_keys, uint32_t *__restrict__ dbg_plist) {
  __shared__ float4 s_rec[kGatherThreads * 3];
  uint64_t R = status->R;
  if (R > R_cap) R = R_cap;
  const int tid = threadIdx.x;
  for (uint64_t base = (uint64_t)blockIdx.x * kGatherThreads; base < R; base += (uint64_t)gridDim.x * kGatherThreads) {
    const uint64_t r = base + tid;
    if (r < R) {
      const uint32_t g = tvals[r], tk = tkeys[r];
      const uint32_t tk_prev = r ? tkeys[r - 1] : 0xFFFFFFFFu, tk_next = r + 1 < R ? tkeys[r + 1] : 0xFFFFFFFFu;
      const float4 q0 = geom[4 * (size_t)g], q1 = geom[4 * (size_t)g + 1];
      float4 q2 = geom[4 * (size_t)g + 2];
      const uint32_t tile = dT.mod(tk), ty = dgx.div(tile), id = dP.mod(g);
      masks[r] = (uint8_t)subblock_mask(q0, q1, (int)(tile - ty * dgx.d) * kTile, (int)ty * kTile);
      if (dbg_keys) dbg_keys[r] = ((uint64_t)tile << 32) | __float_as_uint(q2.w);
      if (dbg_plist) dbg_plist[r] = id;
      q2.w = __uint_as_float(id);
      s_rec[3 * tid] = q0;
      s_rec[3 * tid + 1] = q1;
      s_rec[3 * tid + 2] = q2;
      if (tk_prev != tk) ranges[tk].x = (uint32_t)r;
      if (tk_next != tk) ranges[tk].y = (uint32_t)(r + 1);
    }
    __syncthreads();
    const uint64_t left = R - base;
    const int nrec = left < (uint64_t)kGatherThreads ? (int)left : kGatherThreads;
    float4 *dst = records + 3 * base;
    for (int k = tid; k < 3 * nrec; k += kGatherThreads) dst[k] = s_rec[k];
    __syncthreads();
  }
}

__global__ void init_status_kernel(GhrStatus *st, GhrStatus v) { *st = v; }

}  // namespace

cudaError_t launch_init_status(char *status, GhrStatus st0, cudaStream_t s) {
  init_status_kernel<<<1, 1, 0, s>>>((GhrStatus *)status, st0);
  return cudaGetLastError();
}

cudaError_t launch_scan_duplicate(const GhrDims &d, const Layout &L, char *state, char *temp, uint64_t seq,
                                  cudaStream_t s) {
  (void)seq;
  if (L.nblk_scan == 0) return cudaSuccess;
  scan_duplicate_kernel<<<L.nblk_scan, kScanThreads, 0, s>>>(
      d.P, make_fastdiv((uint32_t)d.P), d.V, L.gx, L.gy, L.T, L.npt, (uint64_t)d.R_cap, (const float4 *)(state + L.pub.off_geom),
      (const uint32_t *)(temp + L.t_dvals[depth_sorted_buf()]), (uint64_t *)(temp + L.t_scan_status),
      (uint32_t *)(temp + L.t_tickets) + 4 * (size_t)d.V + 4, (uint32_t *)(temp + L.t_tkeys[0]),
      (uint32_t *)(temp + L.t_tvals[0]), (uint32_t *)(temp + L.t_thist), (GhrStatus *)(state + L.pub.off_status),
      (uint32_t)L.nblk_scan);
  return cudaGetLastError();
}

cudaError_t launch_gather_ranges(const GhrDims &d, const Layout &L, char *state, char *temp,
                                 uint64_t *dbg_keys, uint32_t *dbg_plist, cudaStream_t s) {
  if (d.R_cap <= 0) return cudaSuccess;
  int buf = tile_sorted_buf(L);
  const uint64_t kMax = 148 * 8;
  uint64_t want = ((uint64_t)d.R_cap + kGatherThreads - 1) / kGatherThreads;
  int nb = (int)(want < kMax ? want : kMax);
  gather_ranges_kernel<<<nb, kGatherThreads, 0, s>>>(
      make_fastdiv((uint32_t)d.P), make_fastdiv((uint32_t)L.T), make_fastdiv((uint32_t)L.gx), (uint64_t)d.R_cap,
      (const GhrStatus *)(state + L.pub.off_status), (const uint32_t *)(temp + L.t_tkeys[buf]),
      (const uint32_t *)(temp + L.t_tvals[buf]), (const float4 *)(state + L.pub.off_geom),
      (float4 *)(state + L.pub.off_records), (uint8_t *)(state + L.pub.off_masks),
      (uint2 *)(state + L.pub.off_ranges), dbg_keys, dbg_plist);
  return cudaGetLastError();
}

}  // namespace ghr
