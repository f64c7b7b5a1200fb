// Gradient all-reduce of the camera-sharded step over NVLink peer memory (SURVEY.md §8(e) stage 2).
//
// The reference's only multi-GPU exchange is Lightning DDP's NCCL all-reduce of ~444 MB of parameter
// gradients (/root/reference/infer_one_shot.py:631,638).  Here the exchange is the packed
// Gaussian-attribute gradient buffer (56 B x P, 3.4 MB at 60k Gaussians): latency-bound, so it is done by
// ONE kernel per rank over peer-mapped memory instead of a library collective:
//   - every rank's buffer lives in a cudaMalloc allocation of this library whose IPC handle the host side
//     exchanges once (ghr_comm_handle / ghr_comm_connect); preprocess_backward writes its result straight
//     into it (the buffer IS PackedGrads.flat), so there is no staging copy;
//   - two-shot, in place, deterministic: after a flag barrier ("my gradients are written"), rank r sums
//     slice r of all ranks in rank order with 128-bit peer loads and pushes the sum into slice r of every
//     rank with 128-bit peer stores; a second flag barrier ("my pushes have landed") ends the kernel.
//     Slice j of a rank's buffer is read only by rank j and overwritten only by rank j afterwards, so no
//     intermediate barrier or second buffer is needed.  Per rank (N-1)/N of the buffer crosses NVLink in
//     each direction: 2.9 MB each way at N = 8, ~4 us at the measured 770 GB/s; the rest is two flag round
//     trips.
//   - flags carry a device-resident epoch (incremented by the kernel itself), so the launch has no
//     per-call host argument and can be captured in a CUDA graph and replayed.
//   - every spin has a clock bound (a peer that never arrives sets the error word instead of hanging
//     the GPU); ghr_comm_status reports it.
#include <cstring>

#include "ghr_internal.cuh"

namespace ghr {

namespace {

constexpr int kMaxRanks = GHR_COMM_MAX_RANKS;
constexpr int kCommThreads = 512;
constexpr long long kSpinLimit = 4000000000ll;   // ~2 s of SM clocks

struct CommDev {
  float *data[kMaxRanks];        // every rank's buffer (index = rank; own entry = local pointer)
  uint32_t *flags[kMaxRanks];    // every rank's flag words: [phase * kMaxRanks + source rank]
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Flag polls are RELAXED system-scope loads; the acquire is ONE fence after the flag has been seen (a spinning
// kernel should not issue a system-scope acquire per poll).
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ float4 ld_peer(const float4 *p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float4 *p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until every rank's word of `phase` in MY flag array equals `want` (threads 0..world-1 poll local memory)
__device__ __forceinline__ void wait_flags(const CommDev &c, int phase, uint32_t want, uint32_t *err, bool acquire) {
  if ((int)threadIdx.x < c.world) {
    const uint32_t *f = c.flags[c.rank] + phase * kMaxRanks + threadIdx.x;
    const long long t0 = clock64();
    while (ld_relaxed_sys(f) != want) {
      __nanosleep(40);
      if (clock64() - t0 > kSpinLimit) {
        atomicExch(err, 1u + (uint32_t)phase);
        break;
      }
    }
    if (acquire) fence_acq_rel_sys();
  }
  __syncthreads();
}

// The kernel's critical path is a chain of system-scope fences and NVLink flag trips (a 256-byte all-reduce takes
// as long as 60 % of the 3.4 MB one), so it has as few of them as the protocol allows:
//   A  "my gradients are written": CTA 0 releases the epoch into every rank's phase-0 word (the gradients come
//      from earlier kernels of the stream); EVERY CTA polls its own rank's words and acquires once;
//   B  "my pushes have landed": every CTA, after its stores and one CTA barrier, adds 1 with release semantics to
//      its phase-1 word on every rank; the words count CTA arrivals since the communicator was created, and every
//      CTA waits until each rank's count has reached the running target -- no hand-over through a last CTA.  No
//      acquire after B: the kernel ends there and the next kernel of the stream reads the buffer.
// state words (local): [0] all-reduces completed (= epoch), [1] CTAs finished, [2] error, [3] CTA arrivals expected
// from every rank so far; [0] and [3] are advanced by the grid's last CTA after every CTA has read them.
__global__ void __launch_bounds__(kCommThreads)
peer_allreduce_kernel(CommDev c, size_t n4, uint32_t *__restrict__ st) {
  const uint32_t epoch = *(volatile uint32_t *)&st[0] + 1u;
  const uint32_t target = *(volatile uint32_t *)&st[3] + gridDim.x;
  if (blockIdx.x == 0 && (int)threadIdx.x < c.world)
    st_release_sys(c.flags[threadIdx.x] + 0 * kMaxRanks + c.rank, epoch);
  wait_flags(c, 0, epoch, &st[2], true);
  // reduce-scatter + all-gather of my slice: sum in rank order, push to every rank
  const size_t slice = (n4 + c.world - 1) / c.world;
  const size_t lo = (size_t)c.rank * slice, hi = lo + slice < n4 ? lo + slice : n4;
  for (size_t i = lo + (size_t)blockIdx.x * kCommThreads + threadIdx.x; i < hi; i += (size_t)gridDim.x * kCommThreads) {
    float4 v[kMaxRanks];
#pragma unroll
    for (int p = 0; p < kMaxRanks; p++)
      if (p < c.world) v[p] = ld_peer(reinterpret_cast<const float4 *>(c.data[p]) + i);
    float4 acc = v[0];
#pragma unroll
    for (int p = 1; p < kMaxRanks; p++)
      if (p < c.world) {
        acc.x += v[p].x; acc.y += v[p].y; acc.z += v[p].z; acc.w += v[p].w;
      }
#pragma unroll
    for (int p = 0; p < kMaxRanks; p++)
      if (p < c.world) st_peer(reinterpret_cast<float4 *>(c.data[p]) + i, acc);
  }
  __syncthreads();                   // the CTA's stores are ordered before the releases below (cumulativity)
  if ((int)threadIdx.x < c.world) red_add_release_sys(c.flags[threadIdx.x] + 1 * kMaxRanks + c.rank, 1u);
  wait_flags(c, 1, target, &st[2], false);
  if (threadIdx.x == 0 && atomicAdd(&st[1], 1u) == gridDim.x - 1u) {
    st[1] = 0u;
    *(volatile uint32_t *)&st[3] = target;
    *(volatile uint32_t *)&st[0] = epoch;
  }
}

}  // namespace

}  // namespace ghr

using namespace ghr;

struct GhrComm {
  int rank, world, device;
  size_t bytes;
  char *base;                 // one allocation: [data | flags (2 * kMaxRanks words) | state (4 words)]
  size_t off_flags, off_state;
  void *peer_base[kMaxRanks];
  CommDev dev;
  bool connected;
};

struct CommHandle {          // what ranks exchange (GHR_COMM_HANDLE_BYTES)
  cudaIpcMemHandle_t mem;
  int32_t rank, world;
  uint64_t bytes;
  char pad[GHR_COMM_HANDLE_BYTES - sizeof(cudaIpcMemHandle_t) - 16];
};
static_assert(sizeof(CommHandle) == GHR_COMM_HANDLE_BYTES, "handle size");

extern "C" {

int ghr_comm_create(int32_t rank, int32_t world, size_t bytes, GhrComm **out) {
  if (!out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || bytes == 0) {
    set_error("ghr_comm_create: bad arguments (rank %d, world %d, bytes %zu; at most %d ranks)", rank, world, bytes,
              kMaxRanks);
    return GHR_EINVAL;
  }
  GhrComm *c = new GhrComm();
  memset(c, 0, sizeof(*c));
  c->rank = rank;
  c->world = world;
  c->bytes = (bytes + 255) / 256 * 256;
  c->off_flags = c->bytes;
  c->off_state = c->off_flags + 2 * kMaxRanks * sizeof(uint32_t);
  const size_t total = c->off_state + 256;
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaMalloc((void **)&c->base, total);
  if (e == cudaSuccess) e = cudaMemset(c->base, 0, total);
  if (e != cudaSuccess) {
    set_error("ghr_comm_create: %s", cudaGetErrorString(e));
    delete c;
    return GHR_ECUDA;
  }
  c->peer_base[rank] = c->base;
  *out = c;
  return GHR_OK;
}

int ghr_comm_handle(GhrComm *c, void *handle_out) {
  if (!c || !handle_out) { set_error("ghr_comm_handle: NULL argument"); return GHR_EINVAL; }
  CommHandle h;
  memset(&h, 0, sizeof(h));
  cudaError_t e = cudaIpcGetMemHandle(&h.mem, c->base);
  if (e != cudaSuccess) { set_error("ghr_comm_handle: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return GHR_ECUDA; }
  h.rank = c->rank;
  h.world = c->world;
  h.bytes = c->bytes;
  memcpy(handle_out, &h, sizeof(h));
  return GHR_OK;
}

int ghr_comm_connect(GhrComm *c, const void *all_handles) {
  if (!c || !all_handles) { set_error("ghr_comm_connect: NULL argument"); return GHR_EINVAL; }
  const CommHandle *hs = (const CommHandle *)all_handles;
  for (int p = 0; p < c->world; p++) {
    if (hs[p].rank != p || hs[p].world != c->world || hs[p].bytes != c->bytes) {
      set_error("ghr_comm_connect: handle %d does not match (rank %d, world %d, bytes %llu vs %zu)", p, hs[p].rank,
                hs[p].world, (unsigned long long)hs[p].bytes, c->bytes);
      return GHR_EINVAL;
    }
    if (p == c->rank) continue;
    cudaError_t e = cudaIpcOpenMemHandle(&c->peer_base[p], hs[p].mem, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("ghr_comm_connect: cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
      return GHR_ECUDA;
    }
  }
  memset(&c->dev, 0, sizeof(c->dev));
  for (int p = 0; p < c->world; p++) {
    c->dev.data[p] = (float *)c->peer_base[p];
    c->dev.flags[p] = (uint32_t *)((char *)c->peer_base[p] + c->off_flags);
  }
  c->dev.rank = c->rank;
  c->dev.world = c->world;
  c->connected = true;
  return GHR_OK;
}

void *ghr_comm_buffer(GhrComm *c) { return c ? (void *)c->base : nullptr; }

int ghr_comm_allreduce(GhrComm *c, size_t nfloats, void *cuda_stream) {
  if (!c || !c->connected) { set_error("ghr_comm_allreduce: communicator not connected"); return GHR_EINVAL; }
  if (nfloats % 4 || nfloats * sizeof(float) > c->bytes) {
    set_error("ghr_comm_allreduce: %zu floats (must be a multiple of 4 and fit the %zu-byte buffer)", nfloats, c->bytes);
    return GHR_EINVAL;
  }
  if (nfloats == 0) return GHR_OK;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const size_t n4 = nfloats / 4, slice = (n4 + c->world - 1) / c->world;
  int grid = (int)((slice + kCommThreads - 1) / kCommThreads);
  if (grid > 148) grid = 148;
  if (grid < 1) grid = 1;
  peer_allreduce_kernel<<<grid, kCommThreads, 0, s>>>(c->dev, n4, (uint32_t *)(c->base + c->off_state));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("ghr_comm_allreduce: %s", cudaGetErrorString(e)); return GHR_ECUDA; }
  return GHR_OK;
}

int ghr_comm_status(GhrComm *c, uint32_t *epoch_out, uint32_t *error_out) {
  if (!c) { set_error("ghr_comm_status: NULL"); return GHR_EINVAL; }
  uint32_t st[4];
  cudaError_t e = cudaMemcpy(st, c->base + c->off_state, sizeof(st), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { set_error("ghr_comm_status: %s", cudaGetErrorString(e)); return GHR_ECUDA; }
  if (epoch_out) *epoch_out = st[0];
  if (error_out) *error_out = st[2];
  return GHR_OK;
}

int ghr_comm_destroy(GhrComm *c) {
  if (!c) return GHR_OK;
  for (int p = 0; p < c->world; p++)
    if (p != c->rank && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
  if (c->base) cudaFree(c->base);
  delete c;
  return GHR_OK;
}

}  // extern "C"
