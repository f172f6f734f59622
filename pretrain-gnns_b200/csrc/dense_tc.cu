// Dense node transforms on the 5th-generation tensor cores (precision 1): error-compensated 3xTF32.
//
// fp32 parity (1e-4, north_star) rules out plain TF32 (10-bit mantissa over K = 300/600).  Every fp32
// operand x is split in registers into hi = tf32(x) and lo = tf32(x - hi) (cvt.rna), both halves are
// staged in shared memory, and each k-step issues three tcgen05.mma (lo*hi, hi*lo, hi*hi) into one fp32
// TMEM accumulator: the dropped lo*lo term and the rounding of lo are ~2^-22 relative, i.e. fp32-class.
//
// One kernel template serves the three operand layouts of dense.cu (same roles, same epilogues):
//   fwd    y[M,N]  = x[M,K]  . w[N,K]^T      A K-major,  B K-major
//   dgrad  gx[M,K] = gy[M,N] . w[N,K]        A K-major,  B MN-major (w rows are the reduction)
//   wgrad  gw[N,K] = gy[M,N]^T . x[M,K]      A MN-major, B MN-major (node rows are the reduction; split-K)
//
// Structure (per CTA: one 128 x BN output tile, 256 threads):
//   * operands are converted + written to smem by all threads in the UMMA canonical NO-SWIZZLE layouts
//     (8x16B core matrices; K-major: rows 16 B apart, MN-major: k 16 B apart), two stages;
//   * fence.proxy.async + __syncthreads hands a stage to the tensor core; one elected thread issues the
//     4 k-steps x 3 products and tcgen05.commit's onto that stage's mbarrier, which the writers of the
//     stage after next wait on -> loads/splits of block k+1 overlap the MMAs of block k;
//   * epilogue: tcgen05.ld (32 lanes x 16 columns per warp and step) -> bias / ReLU / mask -> global.
// TMA is not used for the operands because the hi/lo split has to happen between global and shared
// memory; the fused layer kernel reuses this staging code with the gather as its A producer.
#include "common.cuh"

namespace {

constexpr int BM = 128;       // UMMA M (TMEM lanes)
constexpr int BK = 32;        // fp32 elements of the reduction per stage = 4 UMMA k-steps of 8
constexpr int NTHREADS = 256;
constexpr int NSTAGE = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte sizes of one operand buffer (hi or lo) for a tile of R rows (MN extent) x BK
__host__ __device__ constexpr int kmajor_lbo(int R) { return R * 16 + 16; }          // stride between 16-byte k-chunks (+16: the 8 chunks of a row hit 8 distinct bank groups)
__host__ __device__ constexpr int kmajor_bytes(int R) { return kmajor_lbo(R) * (BK / 4); }
// MN-major tf32 operands must use the 128B-swizzle-with-32B-base layout (UMMA layout type 1): atoms of
// [4 k][32 consecutive row indices] = 4 rows of 128 B, the 32-byte chunk index XOR-ed with (k mod 4).
// A tile keeps the BK/4 atoms of one 32-row block contiguous: k-group stride (SBO) 512 B, block stride (LBO) 4 KiB.
constexpr int MN_SBO = 512;
constexpr int MN_LBO = (BK / 4) * MN_SBO;
__host__ __device__ constexpr int mnmajor_bytes(int R) { return (R / 32) * MN_LBO; }

// UMMA shared-memory matrix descriptor, version 1 (sm_100).
// bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version | [61,64) layout type
// (0 = no swizzle, 1 = 128B swizzle with 32B base)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type = 0) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}

// instruction descriptor for kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Time-bounded wait: a lost arrival traps after ~2 s (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const uint64_t t0 = globaltimer_ns();
#pragma unroll 1
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 1023u) == 1023u && globaltimer_ns() - t0 > 2000000000ull) break;
  }
  asm volatile("trap;");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(float4 v, float4& hi, float4& lo) {
  hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
  lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
}

__device__ int g_tc_mn_variant = 0;  // development switch for the MN-major descriptor convention

struct TcEpilogue {
  const float* bias;      // [N] or null
  int relu;
  const float* mask_src;  // [M,N] (ld = ldm): zero where mask_src <= 0
  int64_t ldm;
  int atomic;             // split-K: accumulate with atomics into a zeroed output
};

// Operand tile loader.  KC: source is contiguous along the reduction (element (r,k) at src[r*ld + k]);
// otherwise contiguous along the row index (element (r,k) at src[k*ld + r]).  R = rows in the tile.
// Each thread owns NV float4 vectors; `fetch` pulls them from global (zero-filled out of bounds),
// `stash` splits them and writes hi/lo into the canonical smem layout.
template <bool KC, int R>
struct Operand {
  static constexpr int VEC = R * BK / 4;
  static constexpr int NV = (VEC + NTHREADS - 1) / NTHREADS;
  // rounded to 1 KiB so that every buffer (the swizzled MN-major ones need 512 B atoms) starts aligned
  static constexpr int BYTES = ((KC ? kmajor_bytes(R) : mnmajor_bytes(R)) + 1023) / 1024 * 1024;
  float4 v[NV];

  __device__ __forceinline__ void fetch(const float* __restrict__ src, int64_t ld, int r0, int rows, int k0, int kend) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int f = threadIdx.x + i * NTHREADS;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < VEC) {
        if (KC) {
          const int r = f / (BK / 4), kc = f % (BK / 4);
          const int gr = r0 + r, gk = k0 + kc * 4;
          if (gr < rows && gk < kend) t = *reinterpret_cast<const float4*>(src + (int64_t)gr * ld + gk);  // kend % 4 == 0
        } else {
          const int k = f / (R / 4), rc = f % (R / 4);
          const int gk = k0 + k, gr = r0 + rc * 4;
          if (gk < kend && gr < rows) t = *reinterpret_cast<const float4*>(src + (int64_t)gk * ld + gr);  // rows % 4 == 0
        }
      }
      v[i] = t;
    }
  }
  __device__ __forceinline__ void stash(uint8_t* hi_buf, uint8_t* lo_buf) const {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int f = threadIdx.x + i * NTHREADS;
      if (f < VEC) {
        int off;
        if (KC) {
          const int r = f / (BK / 4), kc = f % (BK / 4);
          off = kc * kmajor_lbo(R) + (r >> 3) * 128 + (r & 7) * 16;
        } else {
          const int k = f / (R / 4), rc = f % (R / 4);  // rc: group of 4 consecutive row indices
          off = (rc >> 3) * MN_LBO + (k >> 2) * MN_SBO + (k & 3) * 128 + ((((rc >> 1) & 3) ^ (k & 3)) << 5) + (rc & 1) * 16;
        }
        float4 hi, lo;
        split4(v[i], hi, lo);
        *reinterpret_cast<float4*>(hi_buf + off) = hi;
        *reinterpret_cast<float4*>(lo_buf + off) = lo;
      }
    }
  }
  // descriptor of k-step j (8 reduction elements) inside a staged buffer
  __device__ __forceinline__ static uint64_t desc(uint32_t base, int j) {
    if (KC) return umma_desc(base + 2 * j * kmajor_lbo(R), kmajor_lbo(R), 128);
    if (g_tc_mn_variant == 1) return umma_desc(base + j * 2 * MN_SBO, MN_SBO, MN_LBO, 1);
    return umma_desc(base + j * 2 * MN_SBO, MN_LBO, MN_SBO, 1);  // k-step = 8 k = two 4-deep atoms
  }
};

// Two fp32 accumulators per tile: hi*hi in columns [0,BN), the two cross terms in [BN,2BN).  The tensor
// core adds each k-step into the accumulator with truncation, so the error of a chain grows with its
// length; keeping the (2^-11 times smaller) cross terms out of the main chain cuts its length by 3x and
// brings the GEMM to plain-fp32 accuracy (measured, tools/check_tc.py).
template <int BN>
__host__ __device__ constexpr int tmem_cols() { return 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512; }

template <bool A_KC, bool B_KC, int BN>
__host__ __device__ constexpr int smem_bytes() { return NSTAGE * 2 * (Operand<A_KC, BM>::BYTES + Operand<B_KC, BN>::BYTES) + 1024; }

// C[m, n] = sum_r A(m, r) * B(n, r) over r in [kbeg, kend) of this split.
template <bool A_KC, bool B_KC, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_3xtf32(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, float* __restrict__ C, int64_t ldc,
              int M, int N, int K, int k_per_split, TcEpilogue ep) {
  using OpA = Operand<A_KC, BM>;
  using OpB = Operand<B_KC, BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t mma_done[NSTAGE];
  __shared__ uint32_t tmem_base_s;

  uint8_t* bufs = smem;
  auto a_hi = [&](int s) { return bufs + s * 2 * (OpA::BYTES + OpB::BYTES); };
  auto a_lo = [&](int s) { return a_hi(s) + OpA::BYTES; };
  auto b_hi = [&](int s) { return a_hi(s) + 2 * OpA::BYTES; };
  auto b_lo = [&](int s) { return b_hi(s) + OpB::BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int nkb = (kend - kbeg + BK - 1) / BK;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)tmem_cols<BN>())
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(smem_u32(&mma_done[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_s;
  constexpr uint32_t idesc = umma_idesc(BM, BN, !A_KC, !B_KC);

  // Register double-buffering: the global loads of block kb+2 are issued before block kb+1 is split, so a
  // load has two block-times (2 x 12 MMAs) to arrive and HBM/L2 latency stays off the critical path.
  OpA ra[2];
  OpB rb[2];
  if (nkb > 0) {
    ra[0].fetch(A, lda, m0, M, kbeg, kend);
    rb[0].fetch(B, ldb, n0, N, kbeg, kend);
  }
  if (nkb > 1) {
    ra[1].fetch(A, lda, m0, M, kbeg + BK, kend);
    rb[1].fetch(B, ldb, n0, N, kbeg + BK, kend);
  }
  auto block = [&](int kb, OpA& qa, OpB& qb) {
    const int s = kb & 1;
    if (kb >= NSTAGE) mbar_wait(smem_u32(&mma_done[s]), ((kb / NSTAGE) - 1) & 1);  // MMAs of block kb-2 released stage s
    qa.stash(a_hi(s), a_lo(s));
    qb.stash(b_hi(s), b_lo(s));
    if (kb + 2 < nkb) {
      qa.fetch(A, lda, m0, M, kbeg + (kb + 2) * BK, kend);
      qb.fetch(B, ldb, n0, N, kbeg + (kb + 2) * BK, kend);
    }
    fence_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (lane == 0) {
        const uint32_t ah = smem_u32(a_hi(s)), al = smem_u32(a_lo(s)), bh = smem_u32(b_hi(s)), bl = smem_u32(b_lo(s));
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          const uint32_t first = (kb == 0 && j == 0) ? 0u : 1u;
          umma_tf32(tmem_acc + BN, OpA::desc(al, j), OpB::desc(bh, j), idesc, first);  // cross terms -> second accumulator
          umma_tf32(tmem_acc + BN, OpA::desc(ah, j), OpB::desc(bl, j), idesc, 1u);
          umma_tf32(tmem_acc, OpA::desc(ah, j), OpB::desc(bh, j), idesc, first);
        }
        umma_commit(smem_u32(&mma_done[s]));  // implies tcgen05.fence::before_thread_sync
      }
      __syncwarp();
    }
  };
  for (int kb = 0; kb < nkb; kb += 2) {  // unrolled by two so the register sets are addressed statically
    block(kb, ra[0], rb[0]);
    if (kb + 1 < nkb) block(kb + 1, ra[1], rb[1]);
  }
  if (nkb > 0) {
    const int last = nkb - 1;
    mbar_wait(smem_u32(&mma_done[last & 1]), (last / NSTAGE) & 1);  // the last commit covers every earlier MMA
  }
  tc_fence_after();

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4).., column half w/4 ----
  const int row = (warp & 3) * 32 + lane;
  const int gm = m0 + row;
  const int cbeg = (warp >> 2) * (BN / 2);
#pragma unroll 1
  for (int c = 0; c < BN / 2; c += 16) {
    float v[16];
    if (nkb > 0) {
      float x[16];
      tmem_ld16(tmem_acc + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cbeg + c), v);
      tmem_ld16(tmem_acc + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(BN + cbeg + c), x);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += x[i];
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    const int gn0 = n0 + cbeg + c;
    if (gm < M && gn0 < N) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const int gn = gn0 + i;
        if (gn >= N) break;
        float o[4] = {v[i], v[i + 1], v[i + 2], v[i + 3]};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (gn + q < N) {
            if (ep.bias && blockIdx.z == 0) o[q] += ep.bias[gn + q];
            if (ep.relu) o[q] = fmaxf(o[q], 0.f);
            if (ep.mask_src) o[q] = ep.mask_src[(int64_t)gm * ep.ldm + gn + q] > 0.f ? o[q] : 0.f;
          }
        }
        float* dst = C + (int64_t)gm * ldc + gn;
        if (gn + 3 < N && ((ldc & 3) == 0)) {
          if (ep.atomic) atomicAdd(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
          else *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (gn + q < N) {
              if (ep.atomic) atomicAdd(dst + q, o[q]); else dst[q] = o[q];
            }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)tmem_cols<BN>()) : "memory");
  }
}

__global__ void __launch_bounds__(128)
k_colsum_tc(const float* __restrict__ gy, int64_t ld, int M, int N, int rows_per_split, float* __restrict__ gb) {
  const int n = blockIdx.x * 128 + threadIdx.x;
  if (n >= N) return;
  const int r0 = blockIdx.y * rows_per_split, r1 = min(M, r0 + rows_per_split);
  float a = 0.f;
  for (int r = r0; r < r1; ++r) a += gy[(int64_t)r * ld + n];
  atomicAdd(&gb[n], a);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <bool A_KC, bool B_KC, int BN>
int launch(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N, int K, int splits,
           int k_per_split, const TcEpilogue& ep, cudaStream_t st) {
  constexpr int smem = smem_bytes<A_KC, B_KC, BN>();
  static bool configured = false;
  if (!configured) {
    PGNN_CUDA(cudaFuncSetAttribute(k_gemm_3xtf32<A_KC, B_KC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM), (unsigned)splits);
  k_gemm_3xtf32<A_KC, B_KC, BN><<<grid, NTHREADS, smem, st>>>(A, lda, B, ldb, C, ldc, M, N, K, k_per_split, ep);
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

// Tile width: the candidate that minimises (waves over 148 SMs) x (tile width), i.e. the tensor time of the slowest SM.
inline int pick_bn(int M, int N, int splits) {
  const int cand[4] = {64, 128, 160, 224};
  int best = 64;
  double best_cost = 1e30;
  for (int bn : cand) {
    const int64_t tiles = ceil_div(M, BM) * ceil_div(N, bn) * splits;
    const double cost = (double)ceil_div(tiles, kNumSMs) * (bn + 48);  // +48: per-tile fixed cost (prologue/epilogue)
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

template <bool A_KC, bool B_KC>
int dispatch(int bn, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N, int K, int splits,
             int k_per_split, const TcEpilogue& ep, cudaStream_t st) {
  switch (bn) {
    case 64: return launch<A_KC, B_KC, 64>(A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
    case 128: return launch<A_KC, B_KC, 128>(A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
    case 160: return launch<A_KC, B_KC, 160>(A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
    default: return launch<A_KC, B_KC, 224>(A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
  }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int pgnn_debug_set_tc_variant(int v) {
  return cudaMemcpyToSymbol(g_tc_mn_variant, &v, sizeof(int)) == cudaSuccess ? 0 : -2;
}

// development entry: C[M,N] = sum_r A(m,r) B(n,r) with explicit operand majors (1 = reduction-contiguous)
extern "C" __attribute__((visibility("default"))) int pgnn_debug_tc_gemm(int a_kc, int b_kc, int bn, const float* A, int64_t lda,
                                                                        const float* B, int64_t ldb, float* C, int64_t ldc, int M,
                                                                        int N, int K, void* stream) {
  TcEpilogue ep{nullptr, 0, nullptr, 0, 0};
  cudaStream_t st = as_stream(stream);
  if (a_kc && b_kc) return dispatch<true, true>(bn, A, lda, B, ldb, C, ldc, M, N, K, 1, K, ep, st);
  if (a_kc && !b_kc) return dispatch<true, false>(bn, A, lda, B, ldb, C, ldc, M, N, K, 1, K, ep, st);
  if (!a_kc && b_kc) return dispatch<false, true>(bn, A, lda, B, ldb, C, ldc, M, N, K, 1, K, ep, st);
  return dispatch<false, false>(bn, A, lda, B, ldb, C, ldc, M, N, K, 1, K, ep, st);
}

// y[M,N] = act(x[M,K] . w[N,K]^T + bias)
int pgnn_tc_linear_fwd(const float* x, int64_t ldx, const float* w, const float* bias, int64_t M, int64_t N, int64_t K, int relu,
                       float* y, int64_t ldy, cudaStream_t st) {
  if (K % 4 || ldx % 4 || !aligned16(x) || !aligned16(w) || !aligned16(y) || M < 1) return PGNN_EUNSUPPORTED;
  TcEpilogue ep{bias, relu, nullptr, 0, 0};
  return dispatch<true, true>(pick_bn((int)M, (int)N, 1), x, ldx, w, K, y, ldy, (int)M, (int)N, (int)K, 1, (int)K, ep, st);
}

// gx[M,K] = (gy[M,N] . w[N,K]) masked by relu_src > 0
int pgnn_tc_linear_bwd_x(const float* gy, int64_t ldgy, const float* w, int64_t M, int64_t N, int64_t K, const float* relu_src,
                         int64_t ldr, float* gx, int64_t ldgx, cudaStream_t st) {
  if (N % 4 || K % 4 || ldgy % 4 || !aligned16(gy) || !aligned16(w) || !aligned16(gx) || M < 1) return PGNN_EUNSUPPORTED;
  TcEpilogue ep{nullptr, 0, relu_src, ldr, 0};
  // output columns are K; the reduction runs over N; B(n_out = k, r = n) = w[r*K + k] is row-index contiguous
  return dispatch<true, false>(pick_bn((int)M, (int)K, 1), gy, ldgy, w, K, gx, ldgx, (int)M, (int)K, (int)N, 1, (int)N, ep, st);
}

// gw[N,K] = gy[M,N]^T . x[M,K]; gb[N] = column sums of gy
int pgnn_tc_linear_bwd_w(const float* gy, int64_t ldgy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K, float* gw,
                         float* gb, cudaStream_t st) {
  if (N % 4 || K % 4 || ldgy % 4 || ldx % 4 || !aligned16(gy) || !aligned16(x) || !aligned16(gw) || M < 1) return PGNN_EUNSUPPORTED;
  // output [N, K] (rows N = "M" of the MMA), reduction over the M node rows, split so the grid fills the chip
  const int bn = pick_bn((int)N, (int)K, 1);
  const int tiles = (int)(ceil_div(N, BM) * ceil_div(K, bn));
  int splits = (int)ceil_div(kNumSMs, tiles);
  // the tensor core accumulates with truncation: keep each TMEM chain <= 1024 rows, fold the rest in fp32 atomics
  if (splits < (int)ceil_div(M, 1024)) splits = (int)ceil_div(M, 1024);
  const int max_splits = (int)ceil_div(M, 2 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int per = (int)align_up(ceil_div(M, splits), BK);
  splits = (int)ceil_div(M, per);
  if (splits > 1) PGNN_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * N * K, st));
  TcEpilogue ep{nullptr, 0, nullptr, 0, splits > 1};
  int rc = dispatch<false, false>(bn, gy, ldgy, x, ldx, gw, K, (int)N, (int)K, (int)M, splits, per, ep, st);
  if (rc != PGNN_OK) return rc;
  if (gb) {
    PGNN_CUDA(cudaMemsetAsync(gb, 0, sizeof(float) * N, st));
    int rsplit = (int)ceil_div(M, 256);
    if (rsplit > 64) rsplit = 64;
    const int rows_per = (int)ceil_div(M, rsplit);
    dim3 g2((unsigned)ceil_div(N, 128), (unsigned)ceil_div(M, rows_per));
    k_colsum_tc<<<g2, 128, 0, st>>>(gy, ldgy, (int)M, (int)N, rows_per, gb);
    PGNN_LAUNCH_CHECK();
  }
  return PGNN_OK;
}
