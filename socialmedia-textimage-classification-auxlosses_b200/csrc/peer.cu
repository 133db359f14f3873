// Peer (symmetric) memory over NVLink 5 / NVSwitch for the row-sharded ITC exchange: every rank owns one cudaMalloc'd
// block, exported as a CUDA IPC handle and mapped by its peers, so a rank's kernels load the other ranks' embeddings and
// softmax statistics directly through NVLink — the data path has no NCCL call, no host synchronisation, and is CUDA-graph
// capturable.  The reference is single-device (models/mm_late.py:30); this is the B200 exchange step of SURVEY.md §8(e).
//
//   tic_peer_alloc / export / open / close / free : plumbing (host)
//   tic_peer_exchange                             : ONE kernel = cross-rank barrier (release/acquire flags at system scope)
//                                                   + pull of up to kMaxSeg byte ranges from every peer into local buffers
#include <cstdlib>
#include "common.cuh"

namespace tic {

constexpr int kMaxPeers = 8;
constexpr int kMaxSeg = 8;

struct PeerPtrs {
  uint8_t* base[kMaxPeers];
};
struct ExchangeSeg {
  int64_t src_off;      // byte offset inside every rank's symmetric block
  int64_t bytes;        // multiple of 16
  uint8_t* dst;         // local destination of peer 0's range
  int64_t dst_stride;   // peer p's range lands at dst + p * dst_stride
};
struct ExchangeArgs {
  ExchangeSeg seg[kMaxSeg];
  int nseg;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Streaming 16-byte load of a peer-published payload.  NOT `.nc`: the remote ranks write these blocks while this kernel is
// already resident (it spins on their flags), and PTX only allows the non-coherent path for data that is read-only for the
// kernel's whole lifetime.  A relaxed system-scope load after the ld.acquire.sys of the flag is ordered by that acquire and
// is served from the point of coherence; L1::no_allocate keeps the never-re-read payload out of L1.
__device__ __forceinline__ uint4 ld_nc_na(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// ctr[0] = epoch of the last completed exchange on this rank, ctr[1] = ticket of blocks done (both local, zero-initialised).
// flags live at base[r] + flag_off: uint32 flags[kMaxPeers]; slot q of rank r's flags is written by rank q only.
__global__ void __launch_bounds__(256)
peer_exchange_kernel(PeerPtrs sym, int world, int rank, int64_t flag_off, uint32_t* __restrict__ ctr, ExchangeArgs xa,
                     unsigned long long timeout_ns) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(ctr) + 1u;
  // -- signal: everything this rank published (earlier kernels on this stream) is visible before the flag is
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(sym.base[threadIdx.x] + flag_off) + rank, epoch);
  }
  // -- wait: every block observes every peer's flag itself (no grid-wide dependency inside the kernel)
  if (threadIdx.x < world) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(sym.base[rank] + flag_off) + threadIdx.x;
    const uint64_t t0 = globaltimer_ns();
    while (static_cast<int32_t>(ld_acquire_sys(f) - epoch) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        printf("tic: peer barrier timeout (rank %d waiting for rank %d, epoch %u)\n", rank, threadIdx.x, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  // -- pull: 16-byte words, all (segment, peer) ranges flattened over the grid
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int s = 0; s < xa.nseg; ++s) {
    const ExchangeSeg sg = xa.seg[s];
    const int64_t words = sg.bytes >> 4;
    const int64_t total = words * world;
    int64_t i = tid;
    // 4 independent loads in flight per thread.  (8 were tried for the 59 MB "dv" reduction at 8 GPUs, which runs at 270 GB/s:
    // no gain there — the transfer is not bound by the loads in flight — and the short exchanges fell into the scalar tail loop.)
    for (; i + 3 * nthr < total; i += 4 * nthr) {
      uint4 v[4];
      int64_t pw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t j = i + u * nthr;
        const int p = static_cast<int>(j / words);
        pw[u] = j;
        v[u] = ld_nc_na(reinterpret_cast<const uint4*>(sym.base[p] + sg.src_off) + (j - p * words));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = static_cast<int>(pw[u] / words);
        reinterpret_cast<uint4*>(sg.dst + p * sg.dst_stride)[pw[u] - p * words] = v[u];
      }
    }
    for (; i < total; i += nthr) {
      const int p = static_cast<int>(i / words);
      reinterpret_cast<uint4*>(sg.dst + p * sg.dst_stride)[i - p * words] =
          ld_nc_na(reinterpret_cast<const uint4*>(sym.base[p] + sg.src_off) + (i - p * words));
    }
  }
  // -- the last block to finish publishes the epoch for the next exchange on this rank
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) {
      ctr[1] = 0u;
      ctr[0] = epoch;
      __threadfence();
    }
  }
}

// Pull WITHOUT a barrier (the barrier of the preceding tic_peer_exchange already ordered the ranks), peer by peer starting
// with the rank itself: when the last block has copied peer p's ranges it publishes ready[p] = epoch (release, gpu scope).
// A GEMM running beside this kernel consumes the segments as they land (SegOrder in tic_umma.cuh).
// Register budget: this kernel must be able to co-reside with the persistent ITC tile kernel (320 threads x 168 registers
// per SM leave 11.7 K registers) — otherwise the tile kernel, which waits for the segments, would starve it: 128 x 64.
__global__ void __maxnreg__(64)
peer_pull_kernel(PeerPtrs sym, int world, int rank, const uint32_t* __restrict__ ctr, uint32_t* __restrict__ ready,
                 uint32_t* __restrict__ tickets, ExchangeArgs xa) {
  pdl_trigger();
  pdl_wait();
  const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(ctr);
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int k = 0; k < world; ++k) {
    int p = rank + k;
    if (p >= world) p -= world;
    for (int s = 0; s < xa.nseg; ++s) {
      const ExchangeSeg sg = xa.seg[s];
      const int64_t words = sg.bytes >> 4;
      const uint4* src = reinterpret_cast<const uint4*>(sym.base[p] + sg.src_off);
      uint4* dst = reinterpret_cast<uint4*>(sg.dst + p * sg.dst_stride);
      int64_t i = tid;
      for (; i + 3 * nthr < words; i += 4 * nthr) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_nc_na(src + i + u * nthr);
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[i + u * nthr] = v[u];
      }
      for (; i < words; i += nthr) dst[i] = ld_nc_na(src + i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(tickets + p, 1u) == gridDim.x - 1) {
        tickets[p] = 0u;
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ready + p), "r"(epoch) : "memory");
      }
    }
  }
}

// TIC_PEER_SYSFENCE=1: explicit system-scope fences in the push / signal kernels (A/B measurement switch)
inline int peer_sys_fence() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TIC_PEER_SYSFENCE"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}

// ------------------------------------------------------------------ push form (small global batches)
// The pull form above costs a rank two NVLink round trips per exchange (flag out -> flags in, then loads from the peers) inside
// a kernel of its own that the next kernel has to wait for: 11-16 us at 2 GPUs for 1 MB (CUPTI timeline,
// profiles/r02_timeline_c2_g2_before.txt), two of them on the critical chain of the small-batch step.  The push form turns the
// exchange around: the PRODUCER writes its ranges straight into every peer's gathered buffers (posted stores over NVLink: no
// round trip), then releases one flag word per peer; nobody waits in this kernel.  The CONSUMER is the tile kernel itself:
// its TMA producer thread polls the (local) flag of a column segment right before its first load from it (SegOrder,
// tic_umma.cuh) and visits the local segment first, so the remote data has most of the kernel to arrive.
//   ctr[0] = epoch of this phase (incremented here once per step; the consumers compare the flags against it), ctr[1] = ticket.
//   flags of this phase: uint32[world] at flag_off of EVERY block; slot r is written by rank r only.
//   wait_off >= 0: before touching a peer's buffers, wait until that peer's slot in MY uint32[world] at wait_off reaches
//   epoch - 1: the peer has finished reading what the previous step pushed (tic_peer_signal at the tail of its step).
struct PushSeg {
  const uint8_t* src;     // local source = this rank's slot of its own gathered buffer
  int64_t bytes;          // multiple of 16
  int64_t dst_off;        // byte offset of rank 0's slot inside every rank's block
  int64_t dst_stride;     // this rank's slot: dst_off + rank * dst_stride
};
struct PushArgs {
  PushSeg seg[kMaxSeg];
  int nseg;
};
// step[0] = number of completed steps (written by tic_peer_signal at the tail of a step), step[1] = step[0] + 1 = the epoch
// of the step in flight: every push of a step publishes flags = step[1], every consumer of that step waits for flags >=
// step[1] (SegOrder.epoch = step + 1) — the expected value does not depend on when the push kernel runs, so the consumers
// need NO stream dependency on it: they are launched beside it and start on their local segment at once.
__global__ void __launch_bounds__(256)
peer_push_kernel(PeerPtrs sym, int world, int rank, PushArgs xa, int64_t flag_off, int64_t wait_off, const uint32_t* __restrict__ step,
                 uint32_t* __restrict__ ticket, unsigned long long timeout_ns, int sys_fence) {
  pdl_trigger();
  pdl_wait();
  const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(step + 1);
  if (wait_off >= 0) {
    if (threadIdx.x < world && threadIdx.x != rank) {
      const uint32_t* f = reinterpret_cast<const uint32_t*>(sym.base[rank] + wait_off) + threadIdx.x;
      const uint64_t t0 = globaltimer_ns();
      while (static_cast<int32_t>(ld_acquire_sys(f) - (epoch - 1u)) < 0) {
        if (globaltimer_ns() - t0 > timeout_ns) {
          printf("tic: peer push timeout (rank %d waiting for rank %d to release its buffers, epoch %u)\n", rank, threadIdx.x, epoch);
          __trap();
        }
      }
    }
    __syncthreads();
  }
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int s = 0; s < xa.nseg; ++s) {
    const PushSeg sg = xa.seg[s];
    const int64_t words = sg.bytes >> 4;
    const uint4* src = reinterpret_cast<const uint4*>(sg.src);
    for (int64_t i = tid; i < words; i += nthr) {
      const uint4 v = src[i];
      for (int k = 1; k < world; ++k) {          // the REMOTE peers only: the local slot is where the data was produced
        int q = rank + k;
        if (q >= world) q -= world;
        reinterpret_cast<uint4*>(sym.base[q] + sg.dst_off + rank * sg.dst_stride)[i] = v;
      }
    }
  }
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    // Block-local stores -> (bar.sync) -> gpu-scope fence + ticket -> the LAST block's release stores at system scope: the
    // release is cumulative over everything that happens-before it, including the other blocks' stores it observed through
    // the ticket.  No fence.sc.sys: explicit system-scope fences cost this kernel ~8 us (19.2 -> 10.4 us at 2 GPUs,
    // profiles/r02_timeline_c2_g2_push_v1.txt vs _v2); sys_fence != 0 restores them (A/B switch).
    if (sys_fence) __threadfence_system(); else __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (s_last) {
      *ticket = 0u;
      if (sys_fence) __threadfence_system(); else __threadfence();
    }
  }
  __syncthreads();
  // one releasing THREAD per peer (ordered behind thread 0's fence by the bar.sync): each release store waits a fabric round
  // trip, and a single thread looping over 7 peers made that 7 round trips (push 38 us, signal 17 us at 8 GPUs;
  // profiles/r02_timeline_c2_g8_push_serial_flags.txt)
  if (s_last && threadIdx.x >= 1 && threadIdx.x < world) {
    int q = rank + threadIdx.x;
    if (q >= world) q -= world;
    st_release_sys(reinterpret_cast<uint32_t*>(sym.base[q] + flag_off) + rank, epoch);
  }
}

// One small block, tail of a step: step[0] = the finished step, step[1] = the next epoch, and flags[rank] = step[0] in every peer's
// block (release at system scope): "this rank has finished reading what step[0]'s pushes delivered".
__global__ void peer_signal_kernel(PeerPtrs sym, int world, int rank, int64_t flag_off, uint32_t* __restrict__ step, int sys_fence) {
  pdl_trigger();
  pdl_wait();
  if (blockIdx.x != 0) return;
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(step + 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    step[0] = epoch;
    step[1] = epoch + 1u;
  }
  if (threadIdx.x >= 1 && threadIdx.x < world) {      // one releasing thread per peer (see peer_push_kernel)
    if (sys_fence) __threadfence_system();
    int q = rank + threadIdx.x;
    if (q >= world) q -= world;
    st_release_sys(reinterpret_cast<uint32_t*>(sym.base[q] + flag_off) + rank, epoch);
  }
}

}  // namespace tic

using namespace tic;

#define TIC_CUDA_OK(expr, name)                                                   \
  do {                                                                            \
    cudaError_t e__ = (expr);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      ::tic::set_error("%s: %s", name, cudaGetErrorString(e__));                  \
      return TIC_E_CUDA;                                                          \
    }                                                                             \
  } while (0)

extern "C" {

int tic_peer_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

int tic_peer_alloc(int64_t bytes, void** out_ptr) {
  TIC_CHECK_ARG(bytes > 0 && out_ptr, "tic_peer_alloc: bad arguments");
  void* p = nullptr;
  TIC_CUDA_OK(cudaMalloc(&p, static_cast<size_t>(bytes)), "tic_peer_alloc(cudaMalloc)");
  TIC_CUDA_OK(cudaMemset(p, 0, static_cast<size_t>(bytes)), "tic_peer_alloc(cudaMemset)");
  TIC_CUDA_OK(cudaDeviceSynchronize(), "tic_peer_alloc(sync)");
  *out_ptr = p;
  return TIC_OK;
}

int tic_peer_free(void* ptr) {
  if (ptr) TIC_CUDA_OK(cudaFree(ptr), "tic_peer_free");
  return TIC_OK;
}

int tic_peer_export(const void* ptr, void* handle_out_host) {
  TIC_CHECK_ARG(ptr && handle_out_host, "tic_peer_export: null pointer");
  cudaIpcMemHandle_t h;
  TIC_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)), "tic_peer_export(cudaIpcGetMemHandle)");
  memcpy(handle_out_host, &h, sizeof(h));
  return TIC_OK;
}

int tic_peer_open(const void* handle_host, void** out_ptr) {
  TIC_CHECK_ARG(handle_host && out_ptr, "tic_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  void* p = nullptr;
  TIC_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "tic_peer_open(cudaIpcOpenMemHandle)");
  *out_ptr = p;
  return TIC_OK;
}

int tic_peer_close(void* ptr) {
  if (ptr) TIC_CUDA_OK(cudaIpcCloseMemHandle(ptr), "tic_peer_close");
  return TIC_OK;
}

int tic_peer_exchange(void* const* bases_host, int world, int rank, int64_t flag_off, uint32_t* ctr, int nseg,
                      const int64_t* src_off_host, const int64_t* bytes_host, void* const* dst_host,
                      const int64_t* dst_stride_host, void* stream) {
  TIC_CHECK_ARG(bases_host && ctr && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world,
                "tic_peer_exchange: bad group (world=%d rank=%d, at most %d peers)", world, rank, kMaxPeers);
  TIC_CHECK_ARG(nseg >= 0 && nseg <= kMaxSeg && (flag_off & 15) == 0, "tic_peer_exchange: bad segment list");
  PeerPtrs sym{};
  for (int p = 0; p < world; ++p) {
    TIC_CHECK_ARG(bases_host[p] != nullptr, "tic_peer_exchange: rank %d has no mapped block", p);
    sym.base[p] = static_cast<uint8_t*>(bases_host[p]);
  }
  ExchangeArgs xa{};
  xa.nseg = nseg;
  int64_t total = 0;
  for (int s = 0; s < nseg; ++s) {
    TIC_CHECK_ARG((src_off_host[s] & 15) == 0 && (bytes_host[s] & 15) == 0 && (dst_stride_host[s] & 15) == 0 &&
                      aligned16(dst_host[s]) && bytes_host[s] >= 0,
                  "tic_peer_exchange: segment %d is not 16-byte aligned", s);
    xa.seg[s] = ExchangeSeg{src_off_host[s], bytes_host[s], static_cast<uint8_t*>(dst_host[s]), dst_stride_host[s]};
    total += bytes_host[s] * world;
  }
  // latency-bound below ~1 MB (few blocks: every block spins on the flags); NVLink-bound above (many loads in flight)
  int grid = static_cast<int>((total / 16 + 256 * 4 - 1) / (256 * 4));
  if (grid < 1) grid = 1;
  if (grid > 4 * 148) grid = 4 * 148;
  launch_k(peer_exchange_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), sym, world, rank, flag_off, ctr, xa,
                                                                           20ull * 1000ull * 1000ull * 1000ull);
  TIC_CHECK_LAUNCH("tic_peer_exchange");
  return TIC_OK;
}

int tic_peer_push(void* const* bases_host, int world, int rank, int64_t flag_off, int64_t wait_off, const uint32_t* step,
                  uint32_t* ticket, int nseg, void* const* src_host, const int64_t* bytes_host, const int64_t* dst_off_host,
                  const int64_t* dst_stride_host, void* stream) {
  TIC_CHECK_ARG(bases_host && step && ticket && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world,
                "tic_peer_push: bad group (world=%d rank=%d, at most %d peers)", world, rank, kMaxPeers);
  TIC_CHECK_ARG(nseg >= 1 && nseg <= kMaxSeg && (flag_off & 3) == 0 && (wait_off < 0 || (wait_off & 3) == 0), "tic_peer_push: bad arguments");
  PeerPtrs sym{};
  for (int p = 0; p < world; ++p) {
    TIC_CHECK_ARG(bases_host[p] != nullptr, "tic_peer_push: rank %d has no mapped block", p);
    sym.base[p] = static_cast<uint8_t*>(bases_host[p]);
  }
  PushArgs xa{};
  xa.nseg = nseg;
  int64_t total = 0;
  for (int s = 0; s < nseg; ++s) {
    TIC_CHECK_ARG(aligned16(src_host[s]) && (bytes_host[s] & 15) == 0 && (dst_off_host[s] & 15) == 0 && (dst_stride_host[s] & 15) == 0 &&
                      bytes_host[s] >= 0,
                  "tic_peer_push: segment %d is not 16-byte aligned", s);
    xa.seg[s] = PushSeg{static_cast<const uint8_t*>(src_host[s]), bytes_host[s], dst_off_host[s], dst_stride_host[s]};
    total += bytes_host[s];
  }
  int grid = static_cast<int>((total / 16 + 256 * 2 - 1) / (256 * 2));
  if (grid < 1) grid = 1;
  if (grid > 2 * 148) grid = 2 * 148;
  launch_k(peer_push_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), sym, world, rank, xa, flag_off, wait_off, step,
           ticket, 20ull * 1000ull * 1000ull * 1000ull, peer_sys_fence());
  TIC_CHECK_LAUNCH("tic_peer_push");
  return TIC_OK;
}

int tic_peer_signal(void* const* bases_host, int world, int rank, int64_t flag_off, uint32_t* step, void* stream) {
  TIC_CHECK_ARG(bases_host && step && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world && (flag_off & 3) == 0,
                "tic_peer_signal: bad arguments");
  PeerPtrs sym{};
  for (int p = 0; p < world; ++p) {
    TIC_CHECK_ARG(bases_host[p] != nullptr, "tic_peer_signal: rank %d has no mapped block", p);
    sym.base[p] = static_cast<uint8_t*>(bases_host[p]);
  }
  launch_k(peer_signal_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream), sym, world, rank, flag_off, step, peer_sys_fence());
  TIC_CHECK_LAUNCH("tic_peer_signal");
  return TIC_OK;
}

int tic_peer_pull(void* const* bases_host, int world, int rank, const uint32_t* ctr, uint32_t* ready, uint32_t* tickets, int nseg,
                  const int64_t* src_off_host, const int64_t* bytes_host, void* const* dst_host, const int64_t* dst_stride_host,
                  int max_blocks, void* stream) {
  TIC_CHECK_ARG(bases_host && ctr && ready && tickets && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world,
                "tic_peer_pull: bad group (world=%d rank=%d)", world, rank);
  TIC_CHECK_ARG(nseg >= 1 && nseg <= kMaxSeg, "tic_peer_pull: bad segment list");
  PeerPtrs sym{};
  for (int p = 0; p < world; ++p) {
    TIC_CHECK_ARG(bases_host[p] != nullptr, "tic_peer_pull: rank %d has no mapped block", p);
    sym.base[p] = static_cast<uint8_t*>(bases_host[p]);
  }
  ExchangeArgs xa{};
  xa.nseg = nseg;
  int64_t per_peer = 0;
  for (int s = 0; s < nseg; ++s) {
    TIC_CHECK_ARG((src_off_host[s] & 15) == 0 && (bytes_host[s] & 15) == 0 && (dst_stride_host[s] & 15) == 0 &&
                      aligned16(dst_host[s]) && bytes_host[s] >= 0,
                  "tic_peer_pull: segment %d is not 16-byte aligned", s);
    xa.seg[s] = ExchangeSeg{src_off_host[s], bytes_host[s], static_cast<uint8_t*>(dst_host[s]), dst_stride_host[s]};
    per_peer += bytes_host[s];
  }
  int grid = static_cast<int>((per_peer / 16 + 128 * 4 - 1) / (128 * 4));
  if (max_blocks <= 0) max_blocks = 148;   // one small block per SM: enough loads in flight for NVLink beside a persistent GEMM
  if (grid > max_blocks) grid = max_blocks;
  if (grid < 1) grid = 1;
  launch_k(peer_pull_kernel, dim3(grid), dim3(128), 0, static_cast<cudaStream_t>(stream), sym, world, rank, ctr, ready, tickets, xa);
  TIC_CHECK_LAUNCH("tic_peer_pull");
  return TIC_OK;
}

}  // extern "C"
