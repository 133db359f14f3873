// Warp-specialised persistent tcgen05 GEMM mainloop for sm_100a, shared by every dense contraction on the
// hot path (ITC similarity tiles, ITC gradient GEMMs, projection / fusion linears and their gradients).
//
//   D[128 x BN] (fp32, TMEM) = sum_k A[128 x 64] * B[BN x 64]     (bf16 operands, 128B-swizzled smem tiles)
//
//   warp 0      : TMA producer   (one lane)  global -> smem ring, mbarrier complete_tx
//   warp 1      : MMA issuer     (one lane)  tcgen05.mma into one of two TMEM accumulator stages; owns TMEM alloc
//   warps 2..   : epilogue       (4 or 8 warps) tcgen05.ld -> registers -> Epi::tile(...)
//
// Operands may be K-major (row-major [rows, K], the "TN" form) or MN-major (row-major [K, rows]); the choice
// only changes the TMA box shape and the UMMA descriptors, never the data in HBM (no transposed copies).
#pragma once
#include <cuda.h>
#include "tic_ptx.cuh"
#include "common.cuh"

namespace tic {

constexpr int kBM = 128;       // UMMA M (cta_group::1)
constexpr int kBK = 64;        // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;     // bf16 UMMA K
constexpr int kABytes = kBM * kBK * 2;
// Epilogue scratch: [0, 6912) per-tile vectors of the ITC epilogues (6 * BN floats), [6912, 6912 + 8 * 2048) one
// 1024-byte-aligned 32x32 bf16 staging tile per epilogue warp for TMA stores (scratch starts at 256 mod 1024).
constexpr int kEpiStageOff = 6912;
constexpr int kEpiScratchBytes = kEpiStageOff + 8 * 2048 + 256;

template <int BN>
struct UmmaCfg {
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesFit = (227 * 1024 - 1024 - 256 - kEpiScratchBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiScratchBytes;
};

struct EpiCtx {
  int m0, n0;          // tile origin
  int m_blk, n_blk;    // tile indices
  int M, N;            // problem extents
  uint32_t tmem_acc;   // TMEM address (lane 0, first column) of this accumulator stage
  int quad;            // TMEM lane quadrant of this warp: rows m0 + 32*quad + lane
  int part, nparts;    // column split across epilogue warp groups: columns [part*BN/nparts, (part+1)*BN/nparts)
  int epi_tid;         // 0 .. 32*EPI_WARPS-1
  int epi_threads;
  int iter;             // tiles already processed by this CTA (for double-buffering the scratch)
  int ks, ksplit;       // split-K: this work item covers K-slice ks of ksplit (epilogue must accumulate atomically)
  uint8_t* scratch;    // kEpiScratchBytes of smem shared by the epilogue warps
  bool coherent = false;      // the epilogue's inputs may be written (by peers) while this kernel runs: no ld.global.nc on them
  uint8_t* stage = nullptr;   // optional 2 KB per epilogue warp of 1024-byte-aligned staging space (e.g. the idle operand ring)
  float pre[8];        // per-thread values loaded by Epi::prefetch for Epi::tile (row norms, row lse, logit scale)
};

// Optional tile order for a B operand that ARRIVES segment by segment (multi-GPU: the gathered embeddings are pulled from the
// peers by a concurrent kernel): tiles are visited segment-major starting with the local segment, and the TMA producer waits
// for a segment's ready word (>= *epoch, acquire) before its first load from it — the GEMM consumes the all-gather as it lands.
constexpr int kSegFreeSms = 8;
struct SegOrder {
  const uint32_t* ready;   // [nseg] device words, nullptr = ordinary order / operand complete
  const uint32_t* epoch;
  int tiles_per_seg;       // N tiles per segment; 0 = layout not tile-aligned: wait for every segment up front
  int my_seg;              // visited first (the pull copies it first: a local copy, no NVLink hop)
  int nseg;
  int remote;              // 1: the ready words are written by the PEERS themselves (push form, tic_peer_push): no pull kernel runs
                           // beside this one (the launch keeps the whole machine) and the data may land while this kernel runs
};
// system scope: in the push form the flag is released by another GPU
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void seg_wait(const SegOrder& so, int seg, uint32_t epoch) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (static_cast<int32_t>(ld_acquire_sys_u32(so.ready + seg) - epoch) < 0) {
    if ((++spins & 0xFFFu) == 0u) {        // wall-time limit, as mbar_wait (tic_ptx.cuh)
      const uint64_t t = watchdog_now_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > kWatchdogNs) { printf("tic: segment %d never became ready\n", seg); __trap(); }
    }
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");   // the pulled data is read by the TMA (async proxy) next
}

// Named barrier among the epilogue warps only (id 1); the producer / MMA warps never touch it.
__device__ __forceinline__ void epi_bar_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// CLUSTER = 2: two CTAs of a cluster work on M-adjacent tiles of the same N block and SHARE the B tile — each loads half
// of it and TMA-multicasts it into both CTAs' shared memory (operand traffic from L2 per tile drops by 1/3 for 128x256
// tiles; the K=768 similarity tiles are L2-bandwidth-bound otherwise).  Stage release is a multicast tcgen05.commit to
// the `empty` barrier of both CTAs.  In that mode tmap_b_lo holds the half-height box map of B (no split operands).
//
// KC > 1: CLUSTER SPLIT-K.  The KC CTAs of a cluster work on the SAME output tile, CTA r on K-slice r, each into its own
// TMEM accumulator; the non-leaders then push their fp32 partial tiles into the leader's shared memory (DSMEM stores into
// the operand ring, which is idle once the leader's MMAs have completed), the leader folds them into its TMEM accumulator
// and runs the ordinary epilogue.  Why: a CTA's k-loop runs at ~0.28 us per 64-deep k-block whatever the number of CTAs
// (measured, scripts/gemm_rate.py: 16 and 144 CTAs take the same time), so the small-batch GEMMs of the c2 step — 8..48
// tiles of 12..24 k-blocks on a 148-SM machine — are shortened by spreading K over idle SMs, and unlike the fp32-atomic
// split-K this works for every epilogue (bias / ReLU / bf16 output, the ITC softmax epilogues).  One tile per cluster
// (grid = tiles * KC, not persistent): the reduce buffer is used once, so there is no reuse handshake.
template <int BN, bool A_MN, bool B_MN, int EPI_WARPS, class Epi, int CLUSTER = 1, int KC = 1>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a_lo,
                 const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b_lo, int split, int ksplit,
                 int M, int N, int K, const __grid_constant__ typename Epi::Params ep, SegOrder so, int stages_rt = 0) {
  static_assert(CLUSTER == 1 || (CLUSTER == 2 && BN >= 128), "cluster multicast needs BN >= 128");
  using Cfg = UmmaCfg<BN>;
  constexpr int kMaxStages = Cfg::kStages;
  const int STAGES = (stages_rt > 0 && stages_rt < kMaxStages) ? stages_rt : kMaxStages;   // probe: shallower ring / less smem
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "epilogue warps must cover the 4 TMEM lane quadrants");
  static_assert(BN % 64 == 0 && BN <= 256, "BN in {64,128,192,256}");
  static_assert(KC == 1 || CLUSTER == 1, "cluster split-K and B-multicast pairs are exclusive");
  static_assert(KC == 1 || KC == 2 || KC == 4, "cluster split-K over 2 or 4 CTAs");
  static_assert((KC - 1) * kBM * BN * 4 <= kMaxStages * Cfg::kStageBytes, "split-K partial tiles must fit in the operand ring");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  const uint32_t red_full = bar_base + 8u * (2 * STAGES + 5);   // leader: every non-leader epilogue warp has delivered
  const uint32_t red_go = bar_base + 8u * (2 * STAGES + 6);     // non-leader: the leader's operand ring is free
  uint8_t* scratch = smem_gen + STAGES * Cfg::kStageBytes + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + kBM - 1) / kBM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + kBK - 1) / kBK;
  // Split-precision operands: an operand given as a bf16 (hi, lo) pair contributes an extra K-segment, so
  // D = A_hi*B_hi (+ A_lo*B_hi) (+ A_hi*B_lo) accumulates in TMEM with ~16 mantissa bits per split operand.
  const int nseg = 1 + (split & 1) + ((split >> 1) & 1);
  const int total_kb = num_kb * nseg;
  // work item = (tile, K-slice); in cluster mode an item is a PAIR of M-adjacent tiles, one per CTA of the cluster
  const int m_groups = (m_tiles + CLUSTER - 1) / CLUSTER;
  const int num_items = KC > 1 ? num_tiles : (CLUSTER == 1 ? num_tiles : m_groups * n_tiles) * ksplit;
  const uint32_t crank = CLUSTER == 1 ? 0u : cluster_ctarank();
  const uint32_t krank = KC == 1 ? 0u : cluster_ctarank();        // K-slice of this CTA in cluster split-K mode
  const int kslices = KC > 1 ? KC : ksplit;
  const int worker = (CLUSTER == 1 && KC == 1) ? blockIdx.x : blockIdx.x / (CLUSTER * KC);
  const int nworkers = (CLUSTER == 1 && KC == 1) ? gridDim.x : gridDim.x / (CLUSTER * KC);
  // work item -> (tile, K-slice): split-K items enumerate the slices, cluster split-K takes the slice from the CTA rank
  auto item_of = [&](int w, int& t, int& ks) {
    if constexpr (KC > 1) { t = w; ks = static_cast<int>(krank); }
    else { t = w / ksplit; ks = w - t * ksplit; }
  };
  const bool seg_major = so.ready != nullptr && so.tiles_per_seg > 0;
  auto tile_of = [&](int item, int& m_blk, int& n_blk) {
    if (seg_major) {   // segment-major, local segment first (ksplit == 1 in this mode)
      const int per_seg = m_groups * so.tiles_per_seg;
      const int sidx = item / per_seg, rem = item - sidx * per_seg;
      int seg = so.my_seg + sidx;
      if (seg >= so.nseg) seg -= so.nseg;
      m_blk = (rem / so.tiles_per_seg) * CLUSTER + static_cast<int>(crank);
      n_blk = seg * so.tiles_per_seg + rem % so.tiles_per_seg;
      return;
    }
    m_blk = (item / n_tiles) * CLUSTER + static_cast<int>(crank);
    n_blk = item % n_tiles;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (split & 1) tma_prefetch_desc(&tmap_a_lo);
    if (split & 2) tma_prefetch_desc(&tmap_b_lo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CLUSTER);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EPI_WARPS);
    }
    if constexpr (KC > 1) {
      mbar_init(red_full, (KC - 1) * EPI_WARPS);
      mbar_init(red_go, 1);
    }
    fence_mbar_init();
  }
  // one k-block of operands -> ring stage s (issued by the producer lane only)
  auto load_stage = [&](int kt, int m0, int n0, int s) {
    const int seg = kt / num_kb, kb = kt - seg * num_kb;
    const bool a_lo = (seg == 1) && (split & 1);
    const bool b_lo = (seg >= 1) && !a_lo;
    const CUtensorMap* ta = a_lo ? &tmap_a_lo : &tmap_a;
    const CUtensorMap* tb = b_lo ? &tmap_b_lo : &tmap_b;
    const uint32_t sa = smem_base + s * Cfg::kStageBytes;
    const uint32_t sb = sa + kABytes;
    mbar_arrive_expect_tx(full_bar(s), Cfg::kStageBytes);
    const int k0 = kb * kBK;
    if constexpr (!A_MN) {
      tma_load_2d(sa, ta, full_bar(s), k0, m0);
    } else {
      tma_load_2d(sa, ta, full_bar(s), m0, k0);
      tma_load_2d(sa + 8192, ta, full_bar(s), m0 + 64, k0);
    }
    if constexpr (CLUSTER == 2) {
      // my half of the shared B tile, multicast into both CTAs of the pair
      if constexpr (!B_MN) {
        tma_load_2d_mc(sb + crank * (BN / 2) * 128, &tmap_b_lo, full_bar(s), k0, n0 + static_cast<int>(crank) * (BN / 2),
                       uint16_t(3));
      } else {
#pragma unroll
        for (int i = 0; i < BN / 128; ++i) {
          const int bi = static_cast<int>(crank) * (BN / 128) + i;
          tma_load_2d_mc(sb + bi * 8192, tb, full_bar(s), n0 + bi * 64, k0, uint16_t(3));
        }
      }
    } else if constexpr (!B_MN) {
      tma_load_2d(sb, tb, full_bar(s), k0, n0);
    } else {
#pragma unroll
      for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * 8192, tb, full_bar(s), n0 + i * 64, k0);
    }
  };
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if constexpr (CLUSTER > 1) cluster_sync_all();   // peer barriers are initialised before any multicast / remote commit
  if constexpr (KC > 1) { if (!(split & 32)) cluster_sync_all(); }   // (bit 5 of `split`: timing probe without the cluster syncs)
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  pdl_wait();      // prologue done (barriers, TMEM, descriptors); from here on global memory of the predecessors is read

  if (split & 64) {
    // timing probe: prologue + teardown only
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int seg_ok = -1;
      uint32_t seg_epoch = 0;
      if (so.ready != nullptr) {
        seg_epoch = *reinterpret_cast<const volatile uint32_t*>(so.epoch);
        if (!seg_major)   // segments are not tile-aligned: the whole operand has to be there before the first load
          for (int g = 0; g < so.nseg; ++g)
            if (!(so.remote && g == so.my_seg)) seg_wait(so, g, seg_epoch);
      }
      for (int w = worker; w < num_items; w += nworkers) {
        int t, ks;
        item_of(w, t, ks);
        int m_blk, n_blk;
        tile_of(t, m_blk, n_blk);
        const int m0 = m_blk * kBM, n0 = n_blk * BN;
        if (seg_major) {
          const int seg = n_blk / so.tiles_per_seg;
          // pull form: the local segment is copied first, but copied too; push form: it was produced in place (stream order)
          if (seg != seg_ok) { if (!(so.remote && seg == so.my_seg)) seg_wait(so, seg, seg_epoch); seg_ok = seg; }
        }
        const int kt_begin = static_cast<int>(static_cast<int64_t>(total_kb) * ks / kslices);
        const int kt_end = static_cast<int>(static_cast<int64_t>(total_kb) * (ks + 1) / kslices);
        for (int kt = kt_begin; kt < kt_end; ++kt) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          load_stage(kt, m0, n0, s);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
      // all operand loads of this CTA are in flight: the next kernel of the stream may be launched now and run its prologue
      // while the last MMAs and the last epilogue drain (triggering earlier lets waiting CTAs of later kernels pile up on
      // the SMs and starve the parallel branches of the step — measured slower).
      pdl_trigger();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, A_MN, B_MN);
      // K-major: 8-row groups 1024 B apart, K advance = 32 B inside the swizzle row.
      // MN-major: 64-element MN groups 8192 B apart (one TMA box each), 8-k groups 1024 B apart, K advance = 2048 B.
      constexpr uint32_t a_lbo = A_MN ? 8192u : 16u, b_lbo = B_MN ? 8192u : 16u;
      constexpr uint32_t a_kadv = (A_MN ? 2048u : 32u) >> 4, b_kadv = (B_MN ? 2048u : 32u) >> 4;
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int w = worker; w < num_items; w += nworkers) {
        int t_unused, ks;
        item_of(w, t_unused, ks);
        const int kt_begin = static_cast<int>(static_cast<int64_t>(total_kb) * ks / kslices);
        const int kt_end = static_cast<int>(static_cast<int64_t>(total_kb) * (ks + 1) / kslices);
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < kt_end - kt_begin; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint64_t da = umma_smem_desc_sw128(sa, a_lbo, 1024u);
          const uint64_t db = umma_smem_desc_sw128(sa + kABytes, b_lbo, 1024u);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_bf16(d_tmem, da + static_cast<uint64_t>(k * a_kadv), db + static_cast<uint64_t>(k * b_kadv), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          if constexpr (CLUSTER == 2) umma_commit_mc(empty_bar(s), uint16_t(3));   // frees the stage in both CTAs
          else umma_commit(empty_bar(s));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        umma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue =====================
    EpiCtx cx;
    cx.M = M; cx.N = N;
    cx.quad = warp & 3;
    cx.part = (warp - 2) >> 2;
    cx.nparts = EPI_WARPS / 4;
    cx.epi_tid = threadIdx.x - 64;
    cx.epi_threads = 32 * EPI_WARPS;
    cx.scratch = scratch;
    cx.coherent = so.ready != nullptr;
    int as = 0;
    uint32_t aph = 0;
    cx.iter = 0;
    for (int w = worker; w < num_items; w += nworkers) {
      int t;
      item_of(w, t, cx.ks);
      cx.ksplit = ksplit;
      if constexpr (KC > 1) { cx.ks = 0; cx.ksplit = 1; }   // the leader's epilogue sees the complete K sum
      tile_of(t, cx.m_blk, cx.n_blk);
      cx.m0 = cx.m_blk * kBM; cx.n0 = cx.n_blk * BN;
      // Everything the epilogue reads from global memory that does not depend on the accumulator (bias, norms, softmax
      // statistics) is fetched BEFORE waiting for the MMAs, so its latency hides behind the k-loop: measured on a one-tile
      // GEMM (scripts/gemm_rate.py probes) the epilogue was 3.9 us of an 8.2 us launch, most of it exposed load latency.
      // Only for the FIRST tile of a CTA, though: there the epilogue warps would otherwise idle through the whole k-loop.  On
      // the later tiles of a persistent CTA the epilogue is the busy side (the MMAs run a tile ahead) and the early placement
      // measured slower (ITC forward at 16384x768: 0.306 -> 0.391 ms), so those keep the loads after the wait.
      const bool have_rows = cx.m0 < M && (KC == 1 || krank == 0);
      // (not in segment-consuming mode: what the prefetch reads — gathered norms / lse — belongs to the segment the producer
      // waits for, and only the accumulator barrier orders the epilogue behind that wait)
      const bool early = cx.iter == 0 && so.ready == nullptr;
      if (have_rows && early) Epi::template prefetch<BN>(ep, cx);
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      if (have_rows && !early) Epi::template prefetch<BN>(ep, cx);
      cx.tmem_acc = tmem_base + as * BN;
      if constexpr (KC > 1) {
        // ---- cluster split-K reduction through distributed shared memory (one tile per cluster)
        // partial tile of non-leader r: fp32 [BN columns][128 rows] at ring offset (r-1) * 128*BN*4 of the LEADER
        constexpr uint32_t kPartBytes = kBM * BN * 4;
        constexpr int kColsPerPart = BN / (EPI_WARPS / 4);
        const int quad = warp & 3, part = (warp - 2) >> 2;
        const uint32_t trow = cx.tmem_acc + (static_cast<uint32_t>(quad * 32) << 16);
        const uint32_t row_off = static_cast<uint32_t>(quad * 32 + lane) * 4u;
        if (split & 16) {                        // timing probe: no reduction at all (results are wrong by design)
          if (krank == 0 && cx.m0 < M) Epi::template tile<BN>(ep, cx);
        } else if (krank != 0) {
          mbar_wait_cluster(red_go, 0);          // the leader's MMAs are done reading its ring
          const uint32_t dst = mapa_cluster(smem_base, 0) + (krank - 1) * kPartBytes + row_off;
#pragma unroll 1
          for (int c = 0; c < kColsPerPart / 32; ++c) {
            const int cl = part * kColsPerPart + c * 32;
            uint32_t v[32];
            tmem_ld_32x32(trow + cl, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) st_cluster_f32(dst + static_cast<uint32_t>(cl + j) * (kBM * 4u), v[j]);
          }
          fence_acq_rel_cluster();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa_cluster(red_full, 0));
        } else {
          if (cx.epi_tid == 0) {                 // tfull observed: this CTA's MMAs have completed, its ring is free
#pragma unroll
            for (uint32_t r = 1; r < KC; ++r) mbar_arrive_remote(mapa_cluster(red_go, r));
          }
          mbar_wait_cluster(red_full, 0);
#pragma unroll 1
          for (int c = 0; c < kColsPerPart / 32; ++c) {
            const int cl = part * kColsPerPart + c * 32;
            uint32_t v[32];
            tmem_ld_32x32(trow + cl, v);
            tmem_ld_wait();
#pragma unroll
            for (uint32_t r = 0; r + 1 < KC; ++r) {
              const float* src = reinterpret_cast<const float*>(smem_gen + r * kPartBytes + row_off);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + src[(cl + j) * kBM]);
            }
            tmem_st_32x32(trow + cl, v);
          }
          tmem_st_wait();
          tc_fence_before();
          epi_bar_sync(32 * EPI_WARPS);          // every epilogue warp may read any column of the folded accumulator
          tc_fence_after();
          if (cx.m0 < M) Epi::template tile<BN>(ep, cx);
        }
      } else {
        if (cx.m0 < M && !(split & 128)) Epi::template tile<BN>(ep, cx);   // the odd tail tile of a cluster pair has no rows
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      ++cx.iter;
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
    tma_store_wait<0>();   // bulk stores issued by the epilogue (if any) have completed before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CLUSTER > 1) cluster_sync_all();   // no CTA exits while its peer may still multicast into it
  if constexpr (KC > 1) { if (!(split & 32)) cluster_sync_all(); }   // no CTA exits while a peer may still write into / signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
struct TmapEncoder {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn get();
};

// 2D bf16 tensor map over a row-major [outer, inner] matrix with pitch `pitch_elems`, 128B swizzle, zero OOB fill.
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer);

// 2D bf16 tensor map for TMA STORES of 32x32 tiles (64-byte rows, 64-byte swizzle) into a row-major [outer, inner] matrix.
int make_tmap_bf16_store32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems);

int device_sm_count();

template <int BN, bool A_MN, bool B_MN, int EPI_WARPS, class Epi>
int launch_umma_gemm(const void* A, const void* A_lo, int64_t lda, const void* B, const void* B_lo, int64_t ldb, int M, int N,
                     int K, const typename Epi::Params& ep, cudaStream_t stream, int ksplit = 1, int max_ctas = 0,
                     const SegOrder* seg = nullptr) {
  using Cfg = UmmaCfg<BN>;
  if (M <= 0 || N <= 0 || K <= 0) return -1;
  CUtensorMap ta, tb, ta_lo, tb_lo;
  auto map_a = [&](CUtensorMap* t, const void* p) {
    return !A_MN ? make_tmap_bf16_2d(t, p, (uint64_t)K, (uint64_t)M, (uint64_t)lda, kBK, kBM)
                 : make_tmap_bf16_2d(t, p, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, kBK);
  };
  auto map_b = [&](CUtensorMap* t, const void* p) {
    return !B_MN ? make_tmap_bf16_2d(t, p, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, kBK, BN)
                 : make_tmap_bf16_2d(t, p, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, kBK);
  };
  int rc;
  if ((rc = map_a(&ta, A))) return rc;
  if ((rc = map_b(&tb, B))) return rc;
  if ((rc = map_a(&ta_lo, A_lo ? A_lo : A))) return rc;
  if ((rc = map_b(&tb_lo, B_lo ? B_lo : B))) return rc;
  int split = (A_lo ? 1 : 0) | (B_lo ? 2 : 0);
  int stages_rt = 0;
  { const char* e = getenv("TIC_KC_PROBE"); if (e) split |= (atoi(e) & 15) << 4; }
  { const char* e = getenv("TIC_GEMM_STAGES"); if (e) stages_rt = atoi(e); }
  auto kern = umma_gemm_kernel<BN, A_MN, B_MN, EPI_WARPS, Epi>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return -3;
    attr_set = true;
  }
  const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + BN - 1) / BN;
  const int total_kb = ((K + kBK - 1) / kBK) * (1 + (A_lo ? 1 : 0) + (B_lo ? 1 : 0));
  if (ksplit < 1) ksplit = 1;
  if (ksplit > total_kb) ksplit = total_kb;
  int grid = m_tiles * n_tiles * ksplit;
  // Small launches (the latency-bound small-batch step runs several of them side by side on parallel graph branches): a
  // 4-deep ring (3 for 128-wide tiles) and only the scratch this epilogue needs, so that TWO CTAs fit on an SM.  With the
  // full 200 KB ring every CTA owned an SM, and a 16-CTA kernel of the critical chain queued behind the 192 CTAs of two
  // weight-gradient GEMMs of the other chain (CUPTI timeline, profiles/r02_timeline_c2_*.txt: a 6.6 us hole in the chain).
  // The k-loop rate does not depend on the ring depth (4 vs 8: scripts/gemm_rate.py).
  if (stages_rt == 0 && BN <= 128 && !(seg && seg->ready && !seg->remote) && grid <= 2 * device_sm_count()) stages_rt = BN == 64 ? 4 : 3;
  int cap = max_ctas > 0 ? max_ctas : device_sm_count();
  if (seg && seg->ready && !seg->remote && cap > device_sm_count() - kSegFreeSms) cap = device_sm_count() - kSegFreeSms;
  if (grid > cap) grid = cap;
  SegOrder so = seg ? *seg : SegOrder{nullptr, nullptr, 0, 0, 0, 0};
  if (so.ready && (ksplit != 1 || so.tiles_per_seg <= 0 || so.tiles_per_seg * so.nseg != n_tiles)) so.tiles_per_seg = 0;
  size_t smem = Cfg::kSmemBytes;
  if (stages_rt > 0 && stages_rt < Cfg::kStages)
    smem = static_cast<size_t>(stages_rt) * Cfg::kStageBytes + 1024 + 256 + Epi::scratch_bytes(BN);
  launch_k(kern, dim3(grid), dim3(64 + 32 * EPI_WARPS), smem, stream, ta, ta_lo, tb, tb_lo, split, ksplit, M, N, K, ep, so, stages_rt);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// Cluster split-K launch (KC CTAs per output tile, reduction through DSMEM; see the kernel comment).  One cluster per tile.
template <int BN, bool A_MN, bool B_MN, int EPI_WARPS, class Epi, int KC>
int launch_umma_gemm_kc(const void* A, const void* A_lo, int64_t lda, const void* B, const void* B_lo, int64_t ldb, int M, int N,
                        int K, const typename Epi::Params& ep, cudaStream_t stream) {
  using Cfg = UmmaCfg<BN>;
  if (M <= 0 || N <= 0 || K <= 0) return -1;
  CUtensorMap ta, tb, ta_lo, tb_lo;
  auto map_a = [&](CUtensorMap* t, const void* p) {
    return !A_MN ? make_tmap_bf16_2d(t, p, (uint64_t)K, (uint64_t)M, (uint64_t)lda, kBK, kBM)
                 : make_tmap_bf16_2d(t, p, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, kBK);
  };
  auto map_b = [&](CUtensorMap* t, const void* p) {
    return !B_MN ? make_tmap_bf16_2d(t, p, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, kBK, BN)
                 : make_tmap_bf16_2d(t, p, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, kBK);
  };
  int rc;
  if ((rc = map_a(&ta, A))) return rc;
  if ((rc = map_b(&tb, B))) return rc;
  if ((rc = map_a(&ta_lo, A_lo ? A_lo : A))) return rc;
  if ((rc = map_b(&tb_lo, B_lo ? B_lo : B))) return rc;
  int split = (A_lo ? 1 : 0) | (B_lo ? 2 : 0);
  { static int probe = -1; if (probe < 0) { const char* e = getenv("TIC_KC_PROBE"); probe = e ? atoi(e) : 0; } split |= (probe & 3) << 4; }
  const int total_kb = ((K + kBK - 1) / kBK) * (1 + (A_lo ? 1 : 0) + (B_lo ? 1 : 0));
  if (total_kb < KC) return -1;   // every CTA of the cluster needs at least one k-block (its accumulator must be defined)
  auto kern = umma_gemm_kernel<BN, A_MN, B_MN, EPI_WARPS, Epi, 1, KC>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return -3;
    attr_set = true;
  }
  const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + BN - 1) / BN;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(m_tiles * n_tiles * KC);
  cfg.blockDim = dim3(64 + 32 * EPI_WARPS);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = KC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  SegOrder so{nullptr, nullptr, 0, 0, 0, 0};
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, ta_lo, tb, tb_lo, split, 1, M, N, K, ep, so, 0);
  return e == cudaSuccess ? 0 : -4;
}

// Largest cluster split-K factor (4, 2 or 1) for a problem of `tiles` output tiles and `total_kb` k-blocks: every CTA keeps
// at least `min_kb` k-blocks and the launch stays within a CTA budget.  OFF unless TIC_CLUSTER_K=<CTA budget> is set (read
// at every call): measured on B200 (scripts/gemm_rate.py, profiles/r01_cluster_splitk.md) the k-loop does shrink by the
// split factor (0.28 us per k-block per CTA), but a cluster launch costs +1.4 us, its two cluster barriers +0.5 us and the
// DSMEM fold +3.5 us, so it only wins from ~40 k-blocks per tile on (K >= 2560) — none of the GEMMs of the c2 step (12-24).
inline int pick_cluster_k(int tiles, int total_kb, int min_kb = 3) {
  const char* e = getenv("TIC_CLUSTER_K");
  const int budget = e ? atoi(e) : 0;
  if (budget <= 0) return 1;
  for (int kc = 4; kc >= 2; kc /= 2)
    if (tiles * kc <= budget && total_kb >= kc * min_kb) return kc;
  return 1;
}

// Cluster (TMA multicast) launch: pairs of M-adjacent tiles share B. No split operands, no split-K.
template <int BN, bool A_MN, bool B_MN, int EPI_WARPS, class Epi>
int launch_umma_gemm_cluster2(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                              const typename Epi::Params& ep, cudaStream_t stream, const SegOrder* seg = nullptr) {
  using Cfg = UmmaCfg<BN>;
  if (M <= 0 || N <= 0 || K <= 0) return -1;
  CUtensorMap ta, tb, tb_half;
  int rc;
  rc = !A_MN ? make_tmap_bf16_2d(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, kBK, kBM)
             : make_tmap_bf16_2d(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, kBK);
  if (rc) return rc;
  rc = !B_MN ? make_tmap_bf16_2d(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, kBK, BN)
             : make_tmap_bf16_2d(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, kBK);
  if (rc) return rc;
  rc = !B_MN ? make_tmap_bf16_2d(&tb_half, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, kBK, BN / 2)
             : make_tmap_bf16_2d(&tb_half, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, kBK);
  if (rc) return rc;
  auto kern = umma_gemm_kernel<BN, A_MN, B_MN, EPI_WARPS, Epi, 2>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return -3;
    attr_set = true;
  }
  const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + BN - 1) / BN;
  const int items = ((m_tiles + 1) / 2) * n_tiles;
  // Segment-consuming mode: leave kSegFreeSms SMs to the concurrent pull kernel.  Co-residency of its blocks with this
  // kernel's CTAs (one per SM, ~all shared memory) is not something the hardware scheduler promises, and a pull that cannot
  // be scheduled while the producers wait for its segments is a deadlock.
  int clusters = (device_sm_count() - (seg && seg->ready && !seg->remote ? kSegFreeSms : 0)) / 2;
  if (clusters > items) clusters = items;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(64 + 32 * EPI_WARPS);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int split = 0, ksplit = 1;
  SegOrder so = seg ? *seg : SegOrder{nullptr, nullptr, 0, 0, 0, 0};
  if (so.ready && (so.tiles_per_seg <= 0 || so.tiles_per_seg * so.nseg != n_tiles)) so.tiles_per_seg = 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, ta, tb, tb_half, split, ksplit, M, N, K, ep, so, 0);
  return e == cudaSuccess ? 0 : -4;
}

// hi/lo split of an fp32 value into two bf16: hi = rn(x), lo = rn(x - hi)  (hi + lo carries ~16 mantissa bits)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

}  // namespace tic
