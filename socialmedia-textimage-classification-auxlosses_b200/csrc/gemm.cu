// tic_gemm_bf16: generic bf16 x bf16 -> fp32-accumulate GEMM on tcgen05/TMEM fed by TMA (see tic_umma.cuh),
// plus the library-wide host helpers (error string, tensor-map encoder) and a SIMT reference GEMM for self-tests.
#include <mutex>
#include "common.cuh"
#include "tic_umma.cuh"

namespace tic {

static thread_local char g_err[512] = "ok";
constexpr int kRowSsBN = 64;   // tile width of tic_gemm_bf16_rowss (fixes the layout of the row statistics)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local int g_pdl_override = -1;   // tic_set_pdl: -1 = follow TIC_PDL, 0 / 1 = force for this thread's next launches
bool pdl_enabled() {
  if (g_pdl_override >= 0) return g_pdl_override == 1;
  static int v = -1;
  if (v < 0) { const char* e = getenv("TIC_PDL"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
int set_pdl_override(int mode) {
  const int prev = g_pdl_override;
  g_pdl_override = mode < 0 ? -1 : (mode ? 1 : 0);
  return prev;
}

TmapEncoder::EncodeFn TmapEncoder::get() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer) {
  auto fn = TmapEncoder::get();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return TIC_E_TMAP;
  }
  if (!aligned16(ptr) || (pitch_elems % 8) != 0) {
    set_error("TMA operand needs 16-byte aligned base and leading dimension multiple of 8 (ptr=%p ld=%llu)", ptr,
              (unsigned long long)pitch_elems);
    return TIC_E_ARG;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu pitch=%llu box=%ux%u", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems, box_inner, box_outer);
    return TIC_E_TMAP;
  }
  return 0;
}

int make_tmap_bf16_store32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems) {
  auto fn = TmapEncoder::get();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return TIC_E_TMAP;
  }
  if (!aligned16(ptr) || (pitch_elems % 8) != 0) {
    set_error("TMA store target needs a 16-byte aligned base and a leading dimension multiple of 8");
    return TIC_E_ARG;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (store map) failed (%d)", (int)r);
    return TIC_E_TMAP;
  }
  return 0;
}

int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ------------------------------------------------------------------ generic store epilogue
struct StoreEpi {
  struct Params {
    void* D;
    void* D_lo;   // optional bf16 residual (x - bf16(x)) so the result can feed another GEMM as a split operand
    int64_t ldd;
    int d_bf16;
    float alpha;
    const float* bias;
    int relu;
    int accumulate;   // D += result with fp32 atomics (D holds the initial value); required for split-K
    float* row_ss_part;   // optional [n_tiles][M]: per-tile sum of squares of each output row (fused L2-norm statistics)
  };
  static constexpr int scratch_bytes(int bn) { return 2 * bn * 4 + 256; }   // double-buffered bias vector
  // bias of this tile's columns -> shared memory (double-buffered by tile parity), before the accumulator is waited for
  template <int BN>
  __device__ static void prefetch(const Params& p, EpiCtx& cx) {
    if (p.bias == nullptr || cx.ks != 0) return;   // CTA-uniform
    float* sbias = reinterpret_cast<float*>(cx.scratch) + (cx.iter & 1) * BN;
    for (int j = cx.epi_tid; j < BN; j += cx.epi_threads) sbias[j] = (cx.n0 + j < cx.N) ? __ldg(p.bias + cx.n0 + j) : 0.f;
    epi_bar_sync(cx.epi_threads);
  }
  template <int BN>
  __device__ static void tile(const Params& p, const EpiCtx& cx) {
    const int lane = threadIdx.x & 31;
    const int row = cx.m0 + cx.quad * 32 + lane;
    const int cols_per_part = BN / cx.nparts;
    const float* sbias = reinterpret_cast<const float*>(cx.scratch) + (cx.iter & 1) * BN;
    const bool add_bias = p.bias != nullptr && cx.ks == 0;
    const bool vec_ok = p.d_bf16 ? ((p.ldd & 7) == 0 && (reinterpret_cast<uintptr_t>(p.D) & 15) == 0)
                                 : ((p.ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(p.D) & 15) == 0);
    float ss = 0.f;
#pragma unroll 1
    for (int c = 0; c < cols_per_part / 32; ++c) {
      const int cl = cx.part * cols_per_part + c * 32;
      const int col0 = cx.n0 + cl;
      if (col0 >= cx.N) break;  // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32(cx.tmem_acc + (static_cast<uint32_t>(cx.quad * 32) << 16) + cl, v);
      tmem_ld_wait();
      if (row >= cx.M) continue;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = p.alpha * __uint_as_float(v[j]);
        if (add_bias) x += sbias[cl + j];
        if (p.relu) x = fmaxf(x, 0.f);
        f[j] = x;
        if (col0 + j < cx.N) ss = fmaf(x, x, ss);
      }
      const bool full = (col0 + 32 <= cx.N) && vec_ok;
      if (p.d_bf16) {
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + static_cast<int64_t>(row) * p.ldd + col0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = pack_bf16x2(f[j], f[j + 1]);
            u.y = pack_bf16x2(f[j + 2], f[j + 3]);
            u.z = pack_bf16x2(f[j + 4], f[j + 5]);
            u.w = pack_bf16x2(f[j + 6], f[j + 7]);
            *reinterpret_cast<uint4*>(d + j) = u;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < cx.N) d[j] = __float2bfloat16_rn(f[j]);
        }
        if (p.D_lo != nullptr) {
          __nv_bfloat16* dl = reinterpret_cast<__nv_bfloat16*>(p.D_lo) + static_cast<int64_t>(row) * p.ldd + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] -= __bfloat162float(__float2bfloat16_rn(f[j]));
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_bf16x2(f[j], f[j + 1]);
              u.y = pack_bf16x2(f[j + 2], f[j + 3]);
              u.z = pack_bf16x2(f[j + 4], f[j + 5]);
              u.w = pack_bf16x2(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(dl + j) = u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < cx.N) dl[j] = __float2bfloat16_rn(f[j]);
          }
        }
      } else if (p.accumulate) {
        float* d = reinterpret_cast<float*>(p.D) + static_cast<int64_t>(row) * p.ldd + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < cx.N) atomicAdd(d + j, f[j]);
      } else {
        float* d = reinterpret_cast<float*>(p.D) + static_cast<int64_t>(row) * p.ldd + col0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(d + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < cx.N) d[j] = f[j];
        }
      }
    }
    if (p.row_ss_part != nullptr && row < cx.M) {
      if (cx.nparts == 1) p.row_ss_part[static_cast<int64_t>(cx.n_blk) * cx.M + row] = ss;
      else atomicAdd(p.row_ss_part + static_cast<int64_t>(cx.n_blk) * cx.M + row, ss);
    }
  }
};

// cluster split-K variants (kc = 2 | 4): see the KC comment in tic_umma.cuh
template <int BN, int KC>
static int dispatch_major_kc(const void* A, const void* A_lo, int64_t lda, int a_mn, const void* B, const void* B_lo, int64_t ldb,
                             int b_mn, int M, int N, int K, const StoreEpi::Params& ep, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_umma_gemm_kc<BN, false, false, 4, StoreEpi, KC>(A, A_lo, lda, B, B_lo, ldb, M, N, K, ep, st);
  if (!a_mn && b_mn) return launch_umma_gemm_kc<BN, false, true, 4, StoreEpi, KC>(A, A_lo, lda, B, B_lo, ldb, M, N, K, ep, st);
  if (a_mn && !b_mn) return launch_umma_gemm_kc<BN, true, false, 4, StoreEpi, KC>(A, A_lo, lda, B, B_lo, ldb, M, N, K, ep, st);
  return launch_umma_gemm_kc<BN, true, true, 4, StoreEpi, KC>(A, A_lo, lda, B, B_lo, ldb, M, N, K, ep, st);
}

template <int BN>
static int dispatch_major(const void* A, const void* A_lo, int64_t lda, int a_mn, const void* B, const void* B_lo, int64_t ldb,
                          int b_mn, int M, int N, int K, const StoreEpi::Params& ep, cudaStream_t st, int ksplit, int kc = 1) {
  if constexpr (BN <= 128) {
    if (kc == 4) return dispatch_major_kc<BN, 4>(A, A_lo, lda, a_mn, B, B_lo, ldb, b_mn, M, N, K, ep, st);
    if (kc == 2) return dispatch_major_kc<BN, 2>(A, A_lo, lda, a_mn, B, B_lo, ldb, b_mn, M, N, K, ep, st);
  }
  if (!a_mn && !b_mn) return launch_umma_gemm<BN, false, false, 4, StoreEpi>(A, A_lo, lda, B, B_lo, ldb, M, N, K, ep, st, ksplit);
  if (!a_mn && b_mn) return launch_umma_gemm<BN, false, true, 4, StoreEpi>(A, A_lo, lda, B, B_lo, ldb, M, N, K, ep, st, ksplit);
  if (a_mn && !b_mn) return launch_umma_gemm<BN, true, false, 4, StoreEpi>(A, A_lo, lda, B, B_lo, ldb, M, N, K, ep, st, ksplit);
  return launch_umma_gemm<BN, true, true, 4, StoreEpi>(A, A_lo, lda, B, B_lo, ldb, M, N, K, ep, st, ksplit);
}

// 2-CTA clusters: M-adjacent tile pairs share the B tile through TMA multicast (one L2 read feeds two SMs).  Large GEMMs
// with 128x256 tiles need ~17 TB/s of L2->SM operand traffic at tensor peak, which the L2 cannot deliver; sharing B
// cuts the traffic per tile by a third.
template <int BN>
static int dispatch_major_cluster(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int M, int N, int K,
                                  const StoreEpi::Params& ep, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_umma_gemm_cluster2<BN, false, false, 4, StoreEpi>(A, lda, B, ldb, M, N, K, ep, st);
  if (!a_mn && b_mn) return launch_umma_gemm_cluster2<BN, false, true, 4, StoreEpi>(A, lda, B, ldb, M, N, K, ep, st);
  if (a_mn && !b_mn) return launch_umma_gemm_cluster2<BN, true, false, 4, StoreEpi>(A, lda, B, ldb, M, N, K, ep, st);
  return launch_umma_gemm_cluster2<BN, true, true, 4, StoreEpi>(A, lda, B, ldb, M, N, K, ep, st);
}
// Split-K GEMMs (weight gradients) run on side branches of the step beside the latency-critical chain; every CTA of this
// kernel owns a whole SM (shared memory), so a split that fills the machine makes the critical kernels queue for SMs.
// Measured on the c2 step (one multi-branch CUDA graph): 0.110 ms with split-K up to 148 CTAs, 0.098 ms capped at 64.
// Small GEMMs (< 4 GFLOP) are therefore capped at 64 CTAs; large ones may fill the machine.  TIC_SPLITK_MAX_CTAS overrides.
static int splitk_max_ctas(double flops) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TIC_SPLITK_MAX_CTAS"); v = e ? atoi(e) : 0; if (v < 0) v = 0; }
  if (v > 0) return v;
  return flops < 4e9 ? 64 : 1 << 30;
}
// TIC_GEMM_MULTICAST=0 disables the cluster variant (A/B measurement switch).
static bool gemm_multicast() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TIC_GEMM_MULTICAST"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// Smallest tile count for which the 2-CTA multicast form is used (TIC_GEMM_MULTICAST_MIN_TILES: measurement switch).
static int gemm_multicast_min_tiles(int sms) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TIC_GEMM_MULTICAST_MIN_TILES"); v = e ? atoi(e) : 0; }
  return v > 0 ? v : 2 * sms;
}

// ------------------------------------------------------------------ SIMT reference GEMM (self-test only)
__global__ void simt_gemm_kernel(const __nv_bfloat16* A, int64_t lda, int a_mn, const __nv_bfloat16* B, int64_t ldb,
                                 int b_mn, void* D, int64_t ldd, int d_bf16, int M, int N, int K, float alpha,
                                 const float* bias, int relu) {
  __shared__ float sa[16][17], sb[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx;
    const int ma = blockIdx.y * 16 + ty;
    sa[ty][tx] = (ma < M && ka < K) ? __bfloat162float(a_mn ? A[ka * lda + ma] : A[ma * lda + ka]) : 0.f;
    const int nb = blockIdx.x * 16 + ty;
    sb[ty][tx] = (nb < N && ka < K) ? __bfloat162float(b_mn ? B[ka * ldb + nb] : B[nb * ldb + ka]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[ty][k], sb[tx][k], acc);
    __syncthreads();
  }
  if (m < M && n < N) {
    float x = alpha * acc + (bias ? bias[n] : 0.f);
    if (relu) x = fmaxf(x, 0.f);
    if (d_bf16) reinterpret_cast<__nv_bfloat16*>(D)[m * ldd + n] = __float2bfloat16_rn(x);
    else reinterpret_cast<float*>(D)[m * ldd + n] = x;
  }
}


// Tile width / split-K / cluster split-K of one GEMM call (also exported as tic_gemm_plan so tests can see the choice).
static void gemm_config(int M, int N, int K, int n_split_operands, int accumulate, bool rowss, int& best_bn, int& best_ks, int& kc) {
  const int sms = device_sm_count();
  const int m_tiles = ceil_div(M, kBM);
  const int total_kb = ceil_div(K, kBK) * (1 + n_split_operands);
  // Tile width / split-K by a small cost model: waves x (k-blocks x cycles-per-k-block(BN) + fixed + epilogue(BN)).
  // Narrow tiles are shared-memory-bandwidth bound (A is re-read per N tile), wide ones leave SMs idle on small problems;
  // split-K (only when the caller lets us accumulate atomically) fills the machine for long-K weight-gradient GEMMs.
  static const int bns[3] = {256, 128, 64};
  static const double cyc_kb[3] = {520.0, 300.0, 200.0};
  best_bn = 128; best_ks = 1;
  double best = 1e30;
  for (int i = 0; i < 3; ++i) {
    const int tiles = m_tiles * ceil_div(N, bns[i]);
    for (int ks = 1; ks <= (accumulate ? 16 : 1); ks *= 2) {
      if (ks > total_kb) break;
      const int items = tiles * ks;
      if (ks > 1 && items > splitk_max_ctas(2.0 * M * N * K)) break;
      const int waves = ceil_div(items, sms);
      const double t = waves * (ceil_div(total_kb, ks) * cyc_kb[i] + 1500.0 + 8.0 * bns[i] * (accumulate ? 2.0 : 1.0));
      if (t < best) { best = t; best_bn = bns[i]; best_ks = ks; }
    }
  }
  if (rowss) { best_bn = kRowSsBN; best_ks = 1; }   // the partial layout [ceil(N/64)][M] is part of the ABI
  // Cluster split-K for the small (single-wave) problems whose result needs a real epilogue (no atomics): K over 2 or 4 SMs.
  kc = 1;
  if (!accumulate && best_ks == 1 && best_bn <= 128) kc = pick_cluster_k(m_tiles * ceil_div(N, best_bn), total_kb);
}

}  // namespace tic

using namespace tic;

extern "C" {

const char* tic_last_error_string(void) { return g_err; }
int tic_version(void) { return 100; }
int tic_sm_count(void) { return device_sm_count(); }
int tic_set_pdl(int mode) { return set_pdl_override(mode); }

static int gemm_impl(const void* A, const void* A_lo, int64_t lda, int a_mn, const void* B, const void* B_lo, int64_t ldb,
                     int b_mn, void* D, void* D_lo, int64_t ldd, int d_dtype, int M, int N, int K, float alpha, const float* bias,
                     int relu, int accumulate, float* row_ss_part, void* stream) {
  TIC_CHECK_ARG(A && B && D, "tic_gemm_bf16: null pointer");
  TIC_CHECK_ARG(M > 0 && N > 0 && K > 0, "tic_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  TIC_CHECK_ARG(d_dtype == 0 || d_dtype == 1, "tic_gemm_bf16: d_dtype must be 0 (fp32) or 1 (bf16)");
  TIC_CHECK_ARG(D_lo == nullptr || d_dtype == 1, "tic_gemm_bf16: D_lo needs a bf16 output");
  TIC_CHECK_ARG(!accumulate || (d_dtype == 0 && !relu), "tic_gemm_bf16: accumulate needs an fp32 output and no ReLU");
  TIC_CHECK_ARG(!(accumulate && row_ss_part), "tic_gemm_bf16_rowss: row statistics need the complete K sum (no accumulate)");
  StoreEpi::Params ep{D, D_lo, ldd, d_dtype, alpha, bias, relu, accumulate, row_ss_part};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = device_sm_count();
  const int m_tiles = ceil_div(M, kBM);
  int best_bn, best_ks, kc;
  gemm_config(M, N, K, (A_lo ? 1 : 0) + (B_lo ? 1 : 0), accumulate, row_ss_part != nullptr, best_bn, best_ks, kc);
  int rc;
  const bool cluster = gemm_multicast() && !A_lo && !B_lo && best_ks == 1 && best_bn == 256 && m_tiles >= 16 &&
                       static_cast<int64_t>(m_tiles) * ceil_div(N, best_bn) >= gemm_multicast_min_tiles(sms);
  if (cluster) rc = dispatch_major_cluster<256>(A, lda, a_mn, B, ldb, b_mn, M, N, K, ep, st);
  else if (best_bn == 256) rc = dispatch_major<256>(A, A_lo, lda, a_mn, B, B_lo, ldb, b_mn, M, N, K, ep, st, best_ks);
  else if (best_bn == 128) rc = dispatch_major<128>(A, A_lo, lda, a_mn, B, B_lo, ldb, b_mn, M, N, K, ep, st, best_ks, kc);
  else rc = dispatch_major<64>(A, A_lo, lda, a_mn, B, B_lo, ldb, b_mn, M, N, K, ep, st, best_ks, kc);
  if (rc == -3) { set_error("tic_gemm_bf16: cudaFuncSetAttribute(max dynamic smem) failed"); return TIC_E_ATTR; }
  if (rc == -4) { set_error("tic_gemm_bf16: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return TIC_E_LAUNCH; }
  return rc;
}

int tic_gemm_bf16(const void* A, const void* A_lo, int64_t lda, int a_mn, const void* B, const void* B_lo, int64_t ldb,
                  int b_mn, void* D, void* D_lo, int64_t ldd, int d_dtype, int M, int N, int K, float alpha, const float* bias,
                  int relu, int accumulate, void* stream) {
  return gemm_impl(A, A_lo, lda, a_mn, B, B_lo, ldb, b_mn, D, D_lo, ldd, d_dtype, M, N, K, alpha, bias, relu, accumulate, nullptr,
                   stream);
}

int tic_gemm_rowss_parts(int N) { return ceil_div(N, kRowSsBN); }

int tic_gemm_plan(int M, int N, int K, int n_split_operands, int accumulate, int* tile_n, int* ksplit, int* cluster_k) {
  TIC_CHECK_ARG(M > 0 && N > 0 && K > 0 && n_split_operands >= 0 && n_split_operands <= 2 && tile_n && ksplit && cluster_k,
                "tic_gemm_plan: bad arguments");
  gemm_config(M, N, K, n_split_operands, accumulate, false, *tile_n, *ksplit, *cluster_k);
  return TIC_OK;
}

int tic_gemm_bf16_rowss(const void* A, const void* A_lo, int64_t lda, int a_mn, const void* B, const void* B_lo, int64_t ldb,
                        int b_mn, void* D, void* D_lo, int64_t ldd, int d_dtype, int M, int N, int K, float alpha,
                        const float* bias, int relu, float* row_ss_part, void* stream) {
  TIC_CHECK_ARG(row_ss_part != nullptr, "tic_gemm_bf16_rowss: row_ss_part is NULL");
  return gemm_impl(A, A_lo, lda, a_mn, B, B_lo, ldb, b_mn, D, D_lo, ldd, d_dtype, M, N, K, alpha, bias, relu, 0, row_ss_part,
                   stream);
}

int tic_gemm_bf16_simt(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* D, int64_t ldd,
                       int d_dtype, int M, int N, int K, float alpha, const float* bias, int relu, void* stream) {
  TIC_CHECK_ARG(A && B && D && M > 0 && N > 0 && K > 0, "tic_gemm_bf16_simt: bad arguments");
  dim3 grid(ceil_div(N, 16), ceil_div(M, 16)), block(16, 16);
  simt_gemm_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(A), lda, a_mn, static_cast<const __nv_bfloat16*>(B), ldb, b_mn, D, ldd, d_dtype, M,
      N, K, alpha, bias, relu);
  TIC_CHECK_LAUNCH("tic_gemm_bf16_simt");
  return TIC_OK;
}

}  // extern "C"
