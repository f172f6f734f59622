// Dense node transforms on the 5th-generation tensor cores (precision 1): error-compensated 3xTF32.
//
// fp32 parity (1e-4, north_star) rules out plain TF32 (10-bit mantissa over K = 300/600).  Every fp32
// operand x is used as hi = the raw fp32 word (the tensor core reads its top 19 bits, i.e. truncates to
// tf32) and lo = x - trunc_tf32(x) (exact in fp32); each k-step issues three tcgen05.mma (lo*hi, hi*lo,
// hi*hi).  The dropped lo*lo term and the truncation of lo are ~2^-20 relative; measured GEMM error is
// 1-3e-6 of the output scale, the same class as an fp32 FFMA GEMM (tools/check_tc.py).
//
// One kernel template serves the three operand layouts of dense.cu (same roles, same epilogues):
//   fwd    y[M,N]  = x[M,K]  . w[N,K]^T      A K-major,  B K-major
//   dgrad  gx[M,K] = gy[M,N] . w[N,K]        A K-major,  B MN-major (w rows are the reduction)
//   wgrad  gw[N,K] = gy[M,N]^T . x[M,K]      A MN-major, B MN-major (node rows are the reduction; split-K)
//
// Structure (per CTA: one 128 x BN output tile; 8 producer warps + 1 MMA warp; 4 smem stages of BK = 16):
//   * producers cp.async (16 B, L2-only, zero-filled out of bounds) the raw fp32 operand pieces straight into
//     the UMMA canonical smem layouts (K-major: no-swizzle 8x16B core matrices; MN-major: 128B swizzle with
//     32B base, the only layout tf32 accepts transposed), two blocks ahead; when a thread's own pieces of a
//     block have landed (cp.async.wait_group) it re-reads them, writes lo into the stage's second buffer,
//     fence.proxy.async's, and its warp arrives on full[stage];
//   * the MMA warp waits full[stage], one lane issues 2 k-steps x 3 products and tcgen05.commit's onto
//     empty[stage]; the hi*hi chain and the cross terms use separate TMEM accumulators (truncating adds);
//   * epilogue (producer warps): tcgen05.ld both accumulators -> add -> bias / ReLU / mask -> global.
// No block-wide barrier and no register-staged global load sits in the main loop: the only waits are
// the two mbarrier rings and the thread's own cp.async group.
#include "common.cuh"

namespace {

constexpr int BM = 128;       // UMMA M (TMEM lanes)
constexpr int BK = 16;        // fp32 elements of the reduction per stage = 2 UMMA k-steps of 8
constexpr int NPRODUCER = 256;  // 8 producer warps (they also run the epilogue)
constexpr int NTHREADS = NPRODUCER + 32;  // + one MMA-issuing warp
constexpr int NRAW = 6;   // ring of raw (= hi) operand stages filled by cp.async: deep, because L2 round trips are ~1-2 us under load
constexpr int NLO = 2;    // ring of lo stages written by the producers just before a block is published
constexpr int AHEAD = 4;  // cp.async groups in flight per thread

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte sizes of one operand buffer (hi or lo) for a tile of R rows (MN extent) x BK
__host__ __device__ constexpr int kmajor_lbo(int R) { return R * 16 + 32; }          // stride between 16-byte k-chunks (+32: a quarter-warp = 2 rows x 4 chunks hits 8 distinct bank groups)
__host__ __device__ constexpr int kmajor_bytes(int R) { return kmajor_lbo(R) * (BK / 4); }
// MN-major tf32 operands must use the 128B-swizzle-with-32B-base layout (UMMA layout type 1): atoms of
// [4 k][32 consecutive row indices] = 4 rows of 128 B, the 32-byte chunk index XOR-ed with (k mod 4).
// A tile keeps the BK/4 atoms of one 32-row block contiguous: k-group stride (SBO) 512 B, block stride (LBO) 4 KiB.
constexpr int MN_SBO = 512;
constexpr int MN_LBO = (BK / 4) * MN_SBO;  // 2 KiB at BK = 16
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// hi = tf32(x) rounded to nearest; lo = x - hi is exact in fp32 (|lo| <= 2^-11 |x|) and is left unrounded: the
// tensor core reads only its top 19 bits, an error of 2^-10 |lo| <= 2^-21 |x|, the same order as the dropped lo*lo.
__device__ __forceinline__ void split4(float4 v, float4& hi, float4& lo) {
#ifdef PGNN_TRUNC_SPLIT
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
#else
  hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
#endif
  lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
}

// development trace: globaltimer stamps of CTA (0,0,0) at the phase boundaries (pgnn_debug_tc_trace reads it)
__device__ unsigned long long g_tc_trace[16];
#define TC_TRACE(slot)                                                                        \
  do {                                                                                        \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) g_tc_trace[slot] = globaltimer_ns(); \
  } while (0)

struct TcEpilogue {
  const float* bias;      // [N] or null
  int relu;
  const float* mask_src;  // [M,N] (ld = ldm): zero where mask_src <= 0
  int64_t ldm;
  int atomic;             // split-K: accumulate with atomics into a zeroed output
  PgnnGemmHooks hooks;    // fused column reductions over the final output tile (not with split-K)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Operand tile.  KC: source is contiguous along the reduction (element (r,k) at src[r*ld + k]); otherwise
// contiguous along the row index (element (r,k) at src[k*ld + r]).  R = rows (MN extent) of the tile.
// Each producer thread owns NV 16-byte pieces per block, the same ones in `issue` (cp.async raw -> smem) and in
// `make_lo` (re-read own pieces, write x - trunc_tf32(x)).  Everything that does not depend on the block index
// (global pointer of the piece in block 0, smem offset, row validity) is computed ONCE in `init`: the main
// loop is a pointer bump, one compare and the cp.async per piece — with only two producer warps per scheduler
// the per-block address arithmetic (div/mod by non-powers of two, 64-bit multiplies) was what bounded it.
template <bool KC, int R>
struct Operand {
  static constexpr int VEC = R * BK / 4;
  static constexpr int NV = (VEC + NPRODUCER - 1) / NPRODUCER;
  // rounded to 1 KiB so that every buffer (the swizzled MN-major ones need 512 B atoms) starts aligned
  static constexpr int BYTES = ((KC ? kmajor_bytes(R) : mnmajor_bytes(R)) + 1023) / 1024 * 1024;

  const float* gptr[NV];  // piece i of block 0
  int soff[NV];           // byte offset inside an operand buffer; -1: this thread has no i-th piece
  int kloc[NV];           // reduction index of the piece inside a block (K-major: first of 4; MN-major: the k row)
  bool rok[NV];           // row index in range
  int64_t step;           // pointer advance per block, in floats

  __device__ __forceinline__ static int smem_off(int f) {
    if (KC) {
      const int r = f / (BK / 4), kc = f % (BK / 4);
      return kc * kmajor_lbo(R) + (r >> 3) * 128 + (r & 7) * 16;
    }
    const int k = f / (R / 4), rc = f % (R / 4);  // rc: group of 4 consecutive row indices
    return (rc >> 3) * MN_LBO + (k >> 2) * MN_SBO + (k & 3) * 128 + ((((rc >> 1) & 3) ^ (k & 3)) << 5) + (rc & 1) * 16;
  }
  __device__ __forceinline__ void init(const float* __restrict__ src, int64_t ld, int r0, int rows, int k0) {
    step = KC ? (int64_t)BK : (int64_t)BK * ld;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int f = threadIdx.x + i * NPRODUCER;
      soff[i] = -1;
      gptr[i] = src;
      kloc[i] = 0;
      rok[i] = false;
      if (f < VEC) {
        soff[i] = smem_off(f);
        if (KC) {
          const int r = f / (BK / 4), kc = f % (BK / 4);
          kloc[i] = kc * 4;
          rok[i] = (r0 + r) < rows;
          gptr[i] = src + (int64_t)(rok[i] ? r0 + r : 0) * ld + k0 + kc * 4;
        } else {
          const int k = f / (R / 4), rc = f % (R / 4);
          kloc[i] = k;
          rok[i] = (r0 + rc * 4) < rows;                       // rows % 4 == 0
          gptr[i] = src + (int64_t)(k0 + k) * ld + (rok[i] ? r0 + rc * 4 : 0);
        }
      }
    }
  }
  // krem = reduction elements left from this block's start (kend - k0 - kb*BK); kend % 4 == 0 for K-major sources
  __device__ __forceinline__ void issue(int kb, int krem, uint8_t* raw) const {
    const uint32_t base = smem_u32(raw);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (soff[i] >= 0) {
        const bool ok = rok[i] && kloc[i] < krem;
        cp_async16(base + soff[i], ok ? gptr[i] + kb * step : gptr[i], ok ? 16u : 0u);  // 0 bytes: nothing is read, the 16 B are zero-filled
      }
    }
  }
  __device__ __forceinline__ void make_lo(const uint8_t* raw, uint8_t* lo) const {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (soff[i] >= 0) {
        const float4 v = *reinterpret_cast<const float4*>(raw + soff[i]);
        float4 l;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        *reinterpret_cast<float4*>(lo + soff[i]) = l;
      }
    }
  }
  // descriptor of k-step j (8 reduction elements) inside a staged buffer
  __device__ __forceinline__ static uint64_t desc(uint32_t base, int j) {
    if (KC) return umma_desc(base + 2 * j * kmajor_lbo(R), kmajor_lbo(R), 128);
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
__host__ __device__ constexpr int smem_bytes() { return (NRAW + NLO) * (Operand<A_KC, BM>::BYTES + Operand<B_KC, BN>::BYTES) + 1024; }

// C[m, n] = sum_r A(m, r) * B(n, r) over r in [kbeg, kend) of this split.
// Warp roles: warps 0..7 = producers (global -> registers -> hi/lo split -> smem stage, then the epilogue),
// warp 8 = MMA issuer.  Stages are handed over with mbarriers only (full[s]: 8 producer-warp arrivals,
// empty[s]: tcgen05.commit), so producers run up to NSTAGE blocks ahead of the tensor core and no block-wide
// barrier sits in the main loop.
template <bool A_KC, bool B_KC, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_3xtf32(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, float* __restrict__ C, int64_t ldc,
              int M, int N, int K, int k_per_split, TcEpilogue ep) {
  using OpA = Operand<A_KC, BM>;
  using OpB = Operand<B_KC, BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[NLO], raw_empty[NRAW], lo_empty[NLO], acc_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[BN];  // this tile's bias slice (zero past N), loaded once at the prologue

  uint8_t* bufs = smem;
  constexpr int STAGE_BYTES = OpA::BYTES + OpB::BYTES;
  auto a_hi = [&](int kb) { return bufs + (kb % NRAW) * STAGE_BYTES; };              // raw fp32 = hi operand
  auto b_hi = [&](int kb) { return a_hi(kb) + OpA::BYTES; };
  auto a_lo = [&](int kb) { return bufs + (NRAW + kb % NLO) * STAGE_BYTES; };
  auto b_lo = [&](int kb) { return a_lo(kb) + OpA::BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int nkb = (kend - kbeg + BK - 1) / BK;
  if (warp == 0) TC_TRACE(0);
  const bool s_bias_on = ep.bias != nullptr && blockIdx.z == 0;
  for (int i = threadIdx.x; i < BN; i += NTHREADS) s_bias[i] = (s_bias_on && n0 + i < N) ? ep.bias[n0 + i] : 0.f;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)tmem_cols<BN>())
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < NLO; ++s) {
      mbar_init(smem_u32(&full_bar[s]), NPRODUCER / 32);
      mbar_init(smem_u32(&lo_empty[s]), 1);
    }
    for (int s = 0; s < NRAW; ++s) mbar_init(smem_u32(&raw_empty[s]), 1);
    mbar_init(smem_u32(&acc_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_s;
  constexpr uint32_t idesc = umma_idesc(BM, BN, !A_KC, !B_KC);
  if (warp == 0) TC_TRACE(1);

  if (warp < NPRODUCER / 32) {
    // ---------------- producers ----------------
    // block kb: wait for its raw stage, cp.async the raw pieces (they double as the hi operand); then finish block
    // kb-AHEAD: its cp.async group has landed, derive lo from this thread's own pieces and publish it.
    OpA pa;
    OpB pb;
    pa.init(A, lda, m0, M, kbeg);
    pb.init(B, ldb, n0, N, kbeg);
    auto publish = [&](int j) {
      if (j >= NLO) mbar_wait(smem_u32(&lo_empty[j % NLO]), ((j / NLO) - 1) & 1);  // MMAs of block j-NLO have read the lo stage
      pa.make_lo(a_hi(j), a_lo(j));
      pb.make_lo(b_hi(j), b_lo(j));
      // No proxy fence here: on the producer side it drains the thread's whole memory pipeline, i.e. the cp.async
      // groups still in flight, and turns the main loop into one L2 round trip per block (measured).  The
      // mbarrier arrive (release) / wait (acquire) pair orders these generic-proxy writes before the MMA
      // warp's fence.proxy.async, which then orders them before its tcgen05.mma reads.
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&full_bar[j % NLO]));
    };
#pragma unroll 1
    for (int kb = 0; kb < nkb; ++kb) {
      if (kb >= NRAW) mbar_wait(smem_u32(&raw_empty[kb % NRAW]), ((kb / NRAW) - 1) & 1);  // MMAs of block kb-NRAW are done
      pa.issue(kb, kend - kbeg - kb * BK, a_hi(kb));
      pb.issue(kb, kend - kbeg - kb * BK, b_hi(kb));
      cp_async_commit();
      if (kb >= AHEAD) {
        cp_async_wait<AHEAD>();  // all but the newest AHEAD groups are complete -> block kb-AHEAD has landed
        publish(kb - AHEAD);
        if (warp == 0 && kb == AHEAD) TC_TRACE(2);  // first block published
      }
    }
    cp_async_wait<0>();
    for (int j = (nkb > AHEAD ? nkb - AHEAD : 0); j < nkb; ++j) publish(j);
    if (warp == 0) TC_TRACE(3);  // last block published
  } else {
    // ---------------- MMA issuer ----------------
#pragma unroll 1
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(smem_u32(&full_bar[kb % NLO]), (kb / NLO) & 1);
      fence_async_smem();  // generic-proxy writes observed through the barrier -> async proxy (tensor core reads)
      tc_fence_after();
      if (kb == 0) TC_TRACE(4);        // MMA warp: first stage ready
      if (kb == nkb - 1) TC_TRACE(5);  // MMA warp: last stage ready
      if (lane == 0) {
        const uint32_t ah = smem_u32(a_hi(kb)), al = smem_u32(a_lo(kb)), bh = smem_u32(b_hi(kb)), bl = smem_u32(b_lo(kb));
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          const uint32_t first = (kb == 0 && j == 0) ? 0u : 1u;
          umma_tf32(tmem_acc + BN, OpA::desc(al, j), OpB::desc(bh, j), idesc, first);  // cross terms -> second accumulator
          umma_tf32(tmem_acc + BN, OpA::desc(ah, j), OpB::desc(bl, j), idesc, 1u);
          umma_tf32(tmem_acc, OpA::desc(ah, j), OpB::desc(bh, j), idesc, first);
        }
        umma_commit(smem_u32(&lo_empty[kb % NLO]));     // both rings are released by the completion of these MMAs
        umma_commit(smem_u32(&raw_empty[kb % NRAW]));
        if (kb == nkb - 1) umma_commit(smem_u32(&acc_bar));  // accumulators complete
      }
      __syncwarp();
    }
  }
  if (warp >= NPRODUCER / 32) {
    // the MMA warp only has to stay alive until TMEM is released
    tc_fence_before();
    __syncthreads();
    return;
  }
  if (nkb > 0) mbar_wait(smem_u32(&acc_bar), 0);
  tc_fence_after();
  if (warp == 0) TC_TRACE(6);  // accumulators complete, epilogue starts

  // ---- epilogue ----
  // (1) TMEM -> registers -> smem staging tile [128][BN+4] (the operand stages are free: every MMA has
  //     completed).  Warp w owns TMEM lanes 32*(w%4).., column half w/4; both accumulators are summed here.
  // (2) the 8 warps write the tile out row-contiguously (a warp instruction covers 512 consecutive bytes of
  //     one output row), applying bias / ReLU / mask on the way.  Writing straight from the TMEM register
  //     layout (one row per lane) would issue 16-byte stores to 32 different rows per instruction.
  constexpr int SLD = BN + 4;  // staging row stride in floats: 16 B aligned, quarter-warps hit distinct banks
  float* stage = reinterpret_cast<float*>(smem);
  {
    const int row = (warp & 3) * 32 + lane;
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
      float* dst = stage + row * SLD + cbeg + c;
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
  if (warp == 0) TC_TRACE(9);
  // only the producer warps take part from here on (the MMA warp has left through its own barrier below)
  asm volatile("bar.sync 1, %0;" ::"n"(NPRODUCER) : "memory");
  {
    constexpr int C4 = BN / 4;
    constexpr int UNR = 4;  // independent row pieces per thread and trip: keeps the mask loads in flight together
    const int rows_here = min(BM, M - m0);
    const int total = rows_here * C4;
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    const bool mvec_ok = ep.mask_src && ((ep.ldm & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.mask_src) & 15) == 0);
    const bool want_hooks = ep.hooks.colsum || ep.hooks.stats || ep.hooks.S;
    for (int base = threadIdx.x; base < total; base += NPRODUCER * UNR) {
      float4 o[UNR], mk[UNR];
      int gm[UNR], gn[UNR];
      bool live[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int idx = base + u * NPRODUCER;
        const int r = idx / C4, c4 = idx - r * C4;
        gm[u] = m0 + r;
        gn[u] = n0 + c4 * 4;
        live[u] = idx < total && gn[u] < N;
        mk[u] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (live[u]) {
          o[u] = *reinterpret_cast<const float4*>(stage + r * SLD + c4 * 4);
          if (ep.mask_src) {
            const float* mp = ep.mask_src + (int64_t)gm[u] * ep.ldm + gn[u];
            if (mvec_ok && gn[u] + 3 < N) mk[u] = *reinterpret_cast<const float4*>(mp);
            else {
              mk[u].x = mp[0];
              mk[u].y = gn[u] + 1 < N ? mp[1] : 1.f;
              mk[u].z = gn[u] + 2 < N ? mp[2] : 1.f;
              mk[u].w = gn[u] + 3 < N ? mp[3] : 1.f;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (!live[u]) continue;
        float ov[4] = {o[u].x, o[u].y, o[u].z, o[u].w};
        const float mv[4] = {mk[u].x, mk[u].y, mk[u].z, mk[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (s_bias_on) ov[q] += s_bias[gn[u] - n0 + q];
          if (ep.relu) ov[q] = fmaxf(ov[q], 0.f);
          ov[q] = mv[q] > 0.f ? ov[q] : 0.f;
        }
        if (want_hooks)  // keep the final values in the staging tile for the column reductions below
          *reinterpret_cast<float4*>(stage + (gm[u] - m0) * SLD + (gn[u] - n0)) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        float* dst = C + (int64_t)gm[u] * ldc + gn[u];
        if (gn[u] + 3 < N && vec_ok) {
          if (ep.atomic) atomicAdd(reinterpret_cast<float4*>(dst), make_float4(ov[0], ov[1], ov[2], ov[3]));
          else *reinterpret_cast<float4*>(dst) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (gn[u] + q < N) {
              if (ep.atomic) atomicAdd(dst + q, ov[q]); else dst[q] = ov[q];
            }
        }
      }
    }
  }
  // (3) fused column reductions over the final tile: thread c owns output column n0 + c (conflict-free smem
  //     column walks), one atomic per (CTA, column[, q]).
  if (ep.hooks.colsum || ep.hooks.stats || ep.hooks.S) {
    float* sS = stage + BM * SLD;  // [128][Q] slice of the per-row weights, behind the staging tile
    const int rows_here = min(BM, M - m0);
    if (ep.hooks.S)
      for (int i = threadIdx.x; i < rows_here * ep.hooks.Q; i += NPRODUCER) sS[i] = ep.hooks.S[(int64_t)m0 * ep.hooks.Q + i];
    asm volatile("bar.sync 1, %0;" ::"n"(NPRODUCER) : "memory");
    const int c = threadIdx.x;
    if (c < BN && n0 + c < N) {
      const int Q = ep.hooks.Q;
      float s1 = 0.f;
      double d1 = 0.0, d2 = 0.0;
      float tq[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) tq[q] = 0.f;
      for (int r = 0; r < rows_here; ++r) {
        const float v = stage[r * SLD + c];
        s1 += v;
        if (ep.hooks.stats) { d1 += (double)v; d2 += (double)v * (double)v; }
        if (ep.hooks.S) {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q < Q) tq[q] = fmaf(sS[r * Q + q], v, tq[q]);
        }
      }
      if (ep.hooks.colsum) atomicAdd(&ep.hooks.colsum[n0 + c], s1);
      if (ep.hooks.stats) {
        atomicAdd(&ep.hooks.stats[n0 + c], d1);
        atomicAdd(&ep.hooks.stats[(int64_t)N + n0 + c], d2);
      }
      if (ep.hooks.S) {
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (q < Q)
            atomicAdd(q < ep.hooks.q_split ? &ep.hooks.gT[(int64_t)q * ep.hooks.ldt + n0 + c]
                                           : &ep.hooks.gT2[(int64_t)(q - ep.hooks.q_split) * ep.hooks.ldt + n0 + c], tq[q]);
      }
    }
  }
  if (warp == 0) TC_TRACE(7);  // warp 0 epilogue done
  tc_fence_before();
  __syncthreads();
  if (warp == 0) TC_TRACE(8);  // all epilogue warps done
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)tmem_cols<BN>()) : "memory");
  }
}

// gb[n] = sum_m gy[m][n]: 32 columns x 8 row-lanes per block, coalesced row sweeps, smem fold, one atomic per column
__global__ void __launch_bounds__(256)
k_colsum_tc(const float* __restrict__ gy, int64_t ld, int M, int N, int rows_per_block, float* __restrict__ gb) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a = 0.f;
  if (n < N) {
#pragma unroll 4
    for (int r = r0 + w; r < r1; r += 8) a += gy[(int64_t)r * ld + n];
  }
  red[w][lane] = a;
  __syncthreads();
  if (w == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    atomicAdd(&gb[n], t);
  }
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

// Tile width.  Measured (profiles/r01_gemm_phases.md): the main loop runs at the rate one SM can pull operands
// (~26 GB/s per SM at DRAM latency, all producer designs alike), so per-CTA time ~ (128 + BN) bytes per block.
//   * if some width covers the problem in ONE wave (<= 148 CTAs incl. splits), take the narrowest such width:
//     least work per CTA, most SMs busy;
//   * otherwise (many waves) operand re-reads dominate: take the widest width with <= 15% padding waste.
inline int pick_bn(int M, int N, int splits) {
  const int cand[4] = {64, 128, 160, 224};
  const int64_t mt = ceil_div(M, BM) * (splits > 0 ? splits : 1);
  for (int bn : cand)
    if (mt * ceil_div(N, bn) <= kNumSMs) return bn;
  for (int i = 3; i >= 0; --i) {
    const int bn = cand[i];
    const int padded = (int)ceil_div(N, bn) * bn;
    if (padded <= N + N * 15 / 100 || bn == 64) return bn;
  }
  return 64;
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

extern "C" __attribute__((visibility("default"))) int pgnn_debug_tc_trace(unsigned long long* host16) {
  return cudaMemcpyFromSymbol(host16, g_tc_trace, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -2;
}

// development entry: C[M,N] = sum_r A(m,r) B(n,r) with explicit operand majors (1 = reduction-contiguous)
extern "C" __attribute__((visibility("default"))) int pgnn_debug_tc_gemm(int a_kc, int b_kc, int bn, const float* A, int64_t lda,
                                                                        const float* B, int64_t ldb, float* C, int64_t ldc, int M,
                                                                        int N, int K, void* stream) {
  TcEpilogue ep{nullptr, 0, nullptr, 0, 0, PgnnGemmHooks{}};
  cudaStream_t st = as_stream(stream);
  if (a_kc && b_kc) return dispatch<true, true>(bn, A, lda, B, ldb, C, ldc, M, N, K, 1, K, ep, st);
  if (a_kc && !b_kc) return dispatch<true, false>(bn, A, lda, B, ldb, C, ldc, M, N, K, 1, K, ep, st);
  if (!a_kc && b_kc) return dispatch<false, true>(bn, A, lda, B, ldb, C, ldc, M, N, K, 1, K, ep, st);
  return dispatch<false, false>(bn, A, lda, B, ldb, C, ldc, M, N, K, 1, K, ep, st);
}

// y[M,N] = act(x[M,K] . w[N,K]^T + bias)
int pgnn_tc_linear_fwd(const float* x, int64_t ldx, const float* w, const float* bias, int64_t M, int64_t N, int64_t K, int relu,
                       float* y, int64_t ldy, cudaStream_t st, const PgnnGemmHooks* hooks) {
  if (K % 4 || ldx % 4 || !aligned16(x) || !aligned16(w) || !aligned16(y) || M < 1) return PGNN_EUNSUPPORTED;
  TcEpilogue ep{bias, relu, nullptr, 0, 0, hooks ? *hooks : PgnnGemmHooks{}};
  return dispatch<true, true>(pick_bn((int)M, (int)N, 1), x, ldx, w, K, y, ldy, (int)M, (int)N, (int)K, 1, (int)K, ep, st);
}

// gx[M,K] = (gy[M,N] . w[N,K]) masked by relu_src > 0
int pgnn_tc_linear_bwd_x(const float* gy, int64_t ldgy, const float* w, int64_t M, int64_t N, int64_t K, const float* relu_src,
                         int64_t ldr, float* gx, int64_t ldgx, cudaStream_t st, const PgnnGemmHooks* hooks) {
  if (N % 4 || K % 4 || ldgy % 4 || !aligned16(gy) || !aligned16(w) || !aligned16(gx) || M < 1) return PGNN_EUNSUPPORTED;
  TcEpilogue ep{nullptr, 0, relu_src, ldr, 0, hooks ? *hooks : PgnnGemmHooks{}};
  // output columns are K; the reduction runs over N; B(n_out = k, r = n) = w[r*K + k] is row-index contiguous
  return dispatch<true, false>(pick_bn((int)M, (int)K, 1), gy, ldgy, w, K, gx, ldgx, (int)M, (int)K, (int)N, 1, (int)N, ep, st);
}

// gw[N,K] = gy[M,N]^T . x[M,K]; gb[N] = column sums of gy
int pgnn_tc_linear_bwd_w(const float* gy, int64_t ldgy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K, float* gw,
                         float* gb, cudaStream_t st) {
  if (N % 4 || K % 4 || ldgy % 4 || ldx % 4 || !aligned16(gy) || !aligned16(x) || !aligned16(gw) || M < 1) return PGNN_EUNSUPPORTED;
  // output [N, K] (rows N = "M" of the MMA), reduction over the M node rows, split so the grid fills the chip
  int bn = 224;  // the reduction is long and both operands are re-read per tile: widest tile with <= 15% padding
  for (int c : {224, 160, 128, 64}) {
    bn = c;
    if ((int)ceil_div(K, c) * c <= K + K * 15 / 100) break;
  }
  const int tiles = (int)(ceil_div(N, BM) * ceil_div(K, bn));
  int splits = kNumSMs / tiles;  // floor: 150 CTAs on 148 SMs would run as two waves
  // the tensor core accumulates with truncation: keep each TMEM chain <= 1024 rows, fold the rest in fp32 atomics
  if (splits < (int)ceil_div(M, 1024)) splits = (int)ceil_div(M, 1024);
  const int max_splits = (int)ceil_div(M, 2 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int per = (int)align_up(ceil_div(M, splits), BK);
  splits = (int)ceil_div(M, per);
  if (splits > 1) PGNN_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * N * K, st));
  TcEpilogue ep{nullptr, 0, nullptr, 0, splits > 1, PgnnGemmHooks{}};
  int rc = dispatch<false, false>(bn, gy, ldgy, x, ldx, gw, K, (int)N, (int)K, (int)M, splits, per, ep, st);
  if (rc != PGNN_OK) return rc;
  if (gb) {
    PGNN_CUDA(cudaMemsetAsync(gb, 0, sizeof(float) * N, st));
    const int rows_per = 256;
    dim3 g2((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, rows_per));
    k_colsum_tc<<<g2, 256, 0, st>>>(gy, ldgy, (int)M, (int)N, rows_per, gb);
    PGNN_LAUNCH_CHECK();
  }
  return PGNN_OK;
}
