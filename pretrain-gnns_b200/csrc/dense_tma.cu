// 3xTF32 GEMM with TMA-staged operands.  Operands may be reduction-contiguous ("K-major": element (r,k) at
// src[r*ld + k]; one 128B-swizzled box [R x 32] per block) or row-index-contiguous ("MN-major": element (r,k) at
// src[k*ld + r]; R/32 boxes [32 k x 32 r] in the 128B-swizzle / 32B-atom mode, the only layout tf32 accepts transposed).
// fwd = K/K, dgrad = K/MN (or K/K with a transposed weight), wgrad = MN/MN with split-K.
//
// dense_tc.cu feeds the tensor core with cp.async, which one SM can only drive at ~26 GB/s (LDGSTS issue rate /
// requests in flight; profiles/r01_gemm_phases.md).  Here one thread issues two cp.async.bulk.tensor.2d per
// 32-deep reduction block (A box 128 x 32 fp32, B box BN x 32 fp32, 128B-swizzled, zero-filled out of bounds)
// and the copy engine completes them on an mbarrier:
//
//   warp 9   TMA producer: wait raw_empty[s] -> expect_tx -> 2 tensor copies into raw stage s
//   warps 0-7 converters: wait raw_full[s] (+ lo_empty[t]) -> lo = x - trunc_tf32(x), a LINEAR pass over the stage
//            (the swizzle is irrelevant for an element-wise map onto an identically laid out buffer) -> arrive lo_full[t]
//   warp 8   MMA issuer: wait lo_full[t] -> fence.proxy.async -> 4 k-steps x 3 tcgen05.mma (raw tile = hi operand,
//            SWIZZLE_128B K-major descriptors, start address advanced 32 B per k-step) -> commit lo_empty[t], raw_empty[s]
//   epilogue (warps 0-7): tc_common.cuh, identical to dense_tc.cu (two accumulators, staged coalesced stores, hooks).
//
// PAIR variant (K-major operands, no split-K): the grid is launched in clusters of two CTAs along x (row tiles) and the pair runs
// tcgen05.mma.cta_group::2 with M = 256: each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows), the
// leader CTA (rank 0) issues every MMA, and each CTA's TMEM receives the accumulator rows of its own 128-row tile.  The
// main loop is bound by shared-memory bandwidth (TMA writes + lo conversion + 3x operand reads); halving the B bytes per
// SM takes 264 KB per 32-deep block down to 180 KB at BN = 224.  Protocol differences: the converters of BOTH CTAs
// arrive on the LEADER's lo_full (remote mbarrier arrive, release.cluster, after a generic->async proxy fence), and the
// leader's commits are multicast to both CTAs' lo_empty / raw_empty / acc_bar.
//
// Decomposition builds (tools/ubench_gemm.py; never defined in the product build): PGNN_UB_NO_TMA (stages "land" without a
// copy), PGNN_UB_NO_CONVERT (lo stages are published unconverted), PGNN_UB_NO_MMA (stages are committed back without MMAs),
// PGNN_UB_NO_EPILOGUE.  Results are garbage; the timings isolate what each role costs under full-chip load.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace {

constexpr int TBK = 32;                         // fp32 per reduction block = one 128-byte swizzle row
constexpr int T_NTHREADS = NPRODUCER + 64;      // 8 converter/epilogue warps + MMA warp + TMA warp

template <int BN, bool PAIR>
struct TmaCfg {
  static constexpr int A_BYTES = BM * TBK * 4;  // 16 KiB (either major: R rows x 32 k x 4 B)
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * TBK * 4;  // per CTA
  static constexpr int STAGE = A_BYTES + B_BYTES;
  // Ring depths.  A raw stage is refilled only after the MMAs that read it complete (commit -> TMA issue -> TMA latency ->
  // conversion -> arrive -> MMA: ~3 us round trip), a lo stage likewise (~2 us), so a block costs at least
  // max(tensor time, 3 us / NRAW, 2 us / NLO).  Measured at BN = 224: 1.35 us per block with 2 + 2 stages, 1.16 us with 3 + 2,
  // against 0.67 us of tensor time.  As many stages as 220 KiB hold, at most 4 + 4.
  static constexpr int TOTAL = 225280 / STAGE;
  static constexpr int NLO = TOTAL >= 8 ? 4 : TOTAL >= 6 ? 3 : 2;
  static constexpr int NRAW = (TOTAL - NLO) < 4 ? (TOTAL - NLO) : 4;
  static constexpr int EPI = BM * (BN + 4) * 4 + BM * 16 * 4;  // epilogue staging tile + the hooks' [128][Q <= 16] slice
  static constexpr int RING = (NRAW + NLO) * STAGE;
  static constexpr int SMEM = (RING > EPI ? RING : EPI) + 1024;
};
constexpr int MNB_BYTES = 32 * TBK * 4;  // one MN-major box: 32 k rows x 128 B

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
// cta_group::2 forms: one MMA over the CTA pair (M = 256), commit delivered to the same barrier offset in both CTAs
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
// mbar_wait with cluster-scope acquire (the arrivals come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const uint64_t t0 = globaltimer_ns();
#pragma unroll 1
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
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
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// K-major operand, SWIZZLE_128B (layout type 2): 8-row groups are 1024 B apart; LBO is unused for swizzled K-major
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// descriptor of k-step j (8 reduction elements) of a staged operand
template <bool MN>
__device__ __forceinline__ uint64_t tma_desc(uint32_t base, int j) {
  if (MN) return umma_desc(base + j * 1024, MNB_BYTES, 512, 1);  // 32-row blocks 4 KiB apart, 4-deep k groups 512 B apart
  return umma_desc_sw128(base + j * 32);                         // 8 fp32 = 32 B inside the 128 B swizzle row
}

// Split-K fold inside the GEMM (TcEpilogue::fold_counter).  Called by the 256 epilogue threads after their partial tile is
// stored.  The wait is a spin on a global counter: legal only because the host launches this mode with at most one CTA per SM
// of the device (every CTA of the grid is resident, none can be waiting for a slot behind a spinning one); time-bounded.
template <int BN>
__device__ __forceinline__ void splitk_fold_in_kernel(const float* __restrict__ part, int64_t ldc, int m0, int n0, int M, int N,
                                                      const TcEpilogue& ep) {
  const unsigned splits = gridDim.z;
  unsigned int* ctr = ep.fold_counter + (blockIdx.y * gridDim.x + blockIdx.x);
  asm volatile("bar.sync 1, %0;" ::"n"(NPRODUCER) : "memory");   // every thread's stores of the partial tile are issued
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const uint64_t t0 = globaltimer_ns();
    unsigned seen = 0;
    for (uint32_t it = 0;; ++it) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
      if (seen >= splits) break;
      if ((it & 255u) == 255u && globaltimer_ns() - t0 > 2000000000ull) asm volatile("trap;");
    }
    __threadfence();
  }
  asm volatile("bar.sync 1, %0;" ::"n"(NPRODUCER) : "memory");
  const int rows_here = min(BM, M - m0);
  const int cols_here = min(BN, N - n0);         // N % 4 == 0 in this mode (checked by the host)
  const int per = (rows_here + (int)splits - 1) / (int)splits;
  const int r0 = (int)blockIdx.z * per, r1 = min(rows_here, r0 + per);
  const int C4 = cols_here >> 2;
  for (int idx = threadIdx.x; idx < (r1 - r0) * C4; idx += NPRODUCER) {
    const int r = r0 + idx / C4, c = (idx - (idx / C4) * C4) * 4;
    const int64_t off = (int64_t)(m0 + r) * ldc + n0 + c;
    float4 a = __ldcg(reinterpret_cast<const float4*>(part + off));
    for (unsigned sp = 1; sp < splits; ++sp) {   // fixed order 0, 1, 2, ...: the result does not depend on arrival order
      const float4 b = __ldcg(reinterpret_cast<const float4*>(part + (int64_t)sp * ep.split_stride + off));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4*>(ep.fold_out + off) = a;
  }
}

// ---- TMA-store epilogue -------------------------------------------------------------------------------------------------
// The tile leaves through the copy engine instead of through 256 threads' st.global: each warp turns its TMEM rows into FINAL
// values (both accumulators summed, bias / ReLU / ReLU-mask applied in registers) and writes them into a [128 rows x 16 cols]
// box in the SWIZZLE_64B layout (the 16-byte chunk index XORed with bits 1-2 of the row: the 8 lanes of a store phase hit 8
// distinct bank groups); the four warps of a column half meet on a named barrier and one thread issues
// cp.async.bulk.tensor (shared -> global) for the box while the next 16 columns are being read from TMEM.  Rows / columns beyond
// the tensor's extent are clipped by the copy engine (the map carries the true M, N), so ragged edges need no predication.
// The staged values stay in shared memory for the fused column reductions (which only read them).
constexpr int CBOX = 16;                       // columns per store box = one tcgen05.ld.x16
constexpr int CBOX_BYTES = BM * CBOX * 4;      // 8 KiB

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// wait for the issued TMEM loads; the "+r" operands tie the destination registers to the wait so that no use is scheduled above it
__device__ __forceinline__ void tmem_ld_wait16(uint32_t r[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1),
               "r"(c2)
               : "memory");
}
// float index of element (r, c) of the staged tile (boxes of 16 columns, 64-byte rows, SWIZZLE_64B)
__device__ __forceinline__ int stage_idx(int r, int c) {
  return (c >> 4) * (CBOX_BYTES / 4) + r * CBOX + ((((c >> 2) & 3) ^ ((r >> 1) & 3)) << 2) + (c & 3);
}

template <int BN>
__device__ __forceinline__ void tc_epilogue_tma(uint8_t* smem, const float* s_bias, bool s_bias_on, uint32_t tmem_acc, int nkb, int m0,
                                                int n0, int M, int N, const CUtensorMap* tmap_c, const TcEpilogue& ep) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* stage = reinterpret_cast<float*>(smem);
  constexpr int CSPLIT = (BN + 31) / 32 * 16;   // columns [0, CSPLIT): warps 0-3, [CSPLIT, BN): warps 4-7
  const int grp = warp >> 2, q = warp & 3;
  const int row = q * 32 + lane;
  const int cbeg = grp * CSPLIT;
  const int cnum = grp ? BN - CSPLIT : CSPLIT;
  const uint32_t trow = tmem_acc + ((uint32_t)(q * 32) << 16);
  const int sw = (row >> 1) & 3;
  // ReLU mask (dgrad through a ReLU: zero where mask_src <= 0), read in the TMEM register layout: one row per lane, 64 contiguous
  // bytes per chunk.  (Reading the whole row up front as bits, 28 loads in flight per thread, measured SLOWER: with 220 KB of the
  // SM's 256 KB carved out as shared memory the ~30 KB L1 thrashes on 32 rows x 7 lines per warp.)
  const bool row_ok = m0 + row < M;
  const float* mrow = (ep.mask_src && row_ok) ? ep.mask_src + (int64_t)(m0 + row) * ep.ldm + n0 : nullptr;
  const bool mvec_ok = ep.mask_src && ((ep.ldm & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.mask_src) & 15) == 0) && ((n0 & 3) == 0);
#pragma unroll 1
  for (int c = 0; c < cnum; c += CBOX) {
    const int col0 = cbeg + c;
    if (n0 + col0 >= N) break;                  // whole box beyond the last column (uniform over the four warps of the group)
    uint32_t a[16], b[16];
    if (nkb > 0) {
      tmem_ld16_issue(trow + (uint32_t)col0, a);
      tmem_ld16_issue(trow + (uint32_t)(BN + col0), b);
    }
    float4 mk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) mk[j] = make_float4(1.f, 1.f, 1.f, 1.f);
    if (mrow) {                                 // the mask reads fly under the TMEM loads
      if (mvec_ok && n0 + col0 + 15 < N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mk[j] = __ldg(reinterpret_cast<const float4*>(mrow + col0 + 4 * j));
      } else {
        float* mw = reinterpret_cast<float*>(mk);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (n0 + col0 + i < N) mw[i] = __ldg(mrow + col0 + i);
      }
    }
    const float* mf = reinterpret_cast<const float*>(mk);
    float v[16];
    if (nkb > 0) {
      tmem_ld_wait16(a);
      asm volatile("" : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]),
                        "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]));
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a[i]) + __uint_as_float(b[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    float* dst = stage + (col0 >> 4) * (CBOX_BYTES / 4) + row * CBOX;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float t = v[4 * j + e];
        if (s_bias_on) t += s_bias[col0 + 4 * j + e];   // one address per warp instruction: a broadcast
        if (ep.relu) t = fmaxf(t, 0.f);
        o[e] = mf[4 * j + e] > 0.f ? t : 0.f;
      }
      *reinterpret_cast<float4*>(dst + ((j ^ sw) << 2)) = make_float4(o[0], o[1], o[2], o[3]);
    }
    fence_async_smem();                          // this thread's generic-proxy stores -> visible to the copy engine
    asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
    if (q == 0 && lane == 0) {
      tma_store_3d(tmap_c, smem_u32(stage + (col0 >> 4) * (CBOX_BYTES / 4)), n0 + col0, m0, (int)blockIdx.z);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (warp == 0) TC_TRACE(9);
  // fused column reductions over the final tile (thread c owns output column n0 + c)
  if (ep.hooks.colsum || ep.hooks.stats || ep.hooks.S) {
    float* sS = stage + BM * BN;                 // [128][Q] slice of the per-row weights, behind the staged tile
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
      for (int qq = 0; qq < 16; ++qq) tq[qq] = 0.f;
      for (int r = 0; r < rows_here; ++r) {
        const float v = stage[stage_idx(r, c)];
        s1 += v;
        if (ep.hooks.stats) { d1 += (double)v; d2 += (double)v * (double)v; }
        if (ep.hooks.S) {
#pragma unroll
          for (int qq = 0; qq < 16; ++qq)
            if (qq < Q) tq[qq] = fmaf(sS[r * Q + qq], v, tq[qq]);
        }
      }
      if (ep.hooks.colsum) atomicAdd(&ep.hooks.colsum[n0 + c], s1);
      if (ep.hooks.stats) {
        atomicAdd(&ep.hooks.stats[n0 + c], d1);
        atomicAdd(&ep.hooks.stats[(int64_t)N + n0 + c], d2);
      }
      if (ep.hooks.S) {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq)
          if (qq < Q)
            atomicAdd(qq < ep.hooks.q_split ? &ep.hooks.gT[(int64_t)qq * ep.hooks.ldt + n0 + c]
                                            : &ep.hooks.gT2[(int64_t)(qq - ep.hooks.q_split) * ep.hooks.ldt + n0 + c], tq[qq]);
      }
    }
  }
  // shared memory must stay intact until the copy engine has read every box this thread submitted
  if (q == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <bool A_MN, bool B_MN, int BN, bool PAIR>
__global__ void __launch_bounds__(T_NTHREADS, 1)
k_gemm_3xtf32_tma(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, int tma_store, float* __restrict__ C,
                  int64_t ldc, int M, int N, int K, int k_per_split, TcEpilogue ep) {
  static_assert(!PAIR || (!A_MN && !B_MN && BN % 32 == 0), "the pair variant takes K-major operands; UMMA N % 16 == 0 at M = 256");
  static_assert(BN % 16 == 0 && BN <= 256 && (!B_MN || BN % 32 == 0), "UMMA N % 16 == 0 at M = 128; MN-major B boxes hold 32 rows");
  using Cfg = TmaCfg<BN, PAIR>;
  constexpr int NRAW = Cfg::NRAW, T_NLO = Cfg::NLO;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t raw_full[NRAW], raw_empty[NRAW], lo_full[T_NLO], lo_empty[T_NLO], acc_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[BN];

  auto raw = [&](int kb) { return smem + (kb % NRAW) * Cfg::STAGE; };             // [A raw | B raw]
  auto lo = [&](int kb) { return smem + (NRAW + kb % T_NLO) * Cfg::STAGE; };     // [A lo | B lo]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pair: the two row tiles of a pair are neighbours along grid.x (the cluster dimension cta_group::2 kernels must use)
  const int m0 = (PAIR ? blockIdx.x : blockIdx.y) * BM, n0 = (PAIR ? blockIdx.y : blockIdx.x) * BN;
  const int kbeg = blockIdx.z * k_per_split;  // k_per_split % 32 == 0 whenever there is more than one split
  const int kend = min(K, kbeg + k_per_split);
  const int nkb = (kend - kbeg + TBK - 1) / TBK;
  const bool s_bias_on = ep.bias != nullptr && blockIdx.z == 0;
  if (warp == 0) TC_TRACE(0);

  if (warp == 0) {  // pair: warp 0 of BOTH CTAs, same destination offset
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                   "r"((uint32_t)tmem_cols<BN>())
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                   "r"((uint32_t)tmem_cols<BN>())
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < NRAW; ++s) {
      mbar_init(smem_u32(&raw_full[s]), 1);
      mbar_init(smem_u32(&raw_empty[s]), 1);
    }
    for (int s = 0; s < T_NLO; ++s) {
      mbar_init(smem_u32(&lo_full[s]), (PAIR ? 2 : 1) * (NPRODUCER / 32));  // pair: the leader's copy collects both CTAs' converters
      mbar_init(smem_u32(&lo_empty[s]), 1);
    }
    mbar_init(smem_u32(&acc_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  if (PAIR) cluster_sync_all();  // the peer's barriers exist before anything arrives at them
  const uint32_t tmem_acc = tmem_base_s;
  constexpr uint32_t idesc = umma_idesc(PAIR ? 2 * BM : BM, BN, A_MN, B_MN);
  // everything above (TMEM allocation, barrier init) is independent of the previous kernel's output
  pdl_prologue();
  if (threadIdx.x < NPRODUCER)
    for (int i = threadIdx.x; i < BN; i += NPRODUCER) s_bias[i] = (s_bias_on && n0 + i < N) ? ep.bias[n0 + i] : 0.f;
  if (warp == 0) TC_TRACE(1);

  if (warp == 9) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NRAW;
        if (kb >= NRAW) mbar_wait(smem_u32(&raw_empty[s]), ((kb / NRAW) - 1) & 1);
#ifdef PGNN_UB_NO_TMA
        mbar_arrive(smem_u32(&raw_full[s]));
        continue;
#endif
        mbar_expect_tx(smem_u32(&raw_full[s]), Cfg::STAGE);  // a box is always written in full (zero-filled out of bounds)
        const int k0 = kbeg + kb * TBK;
        const uint32_t bar = smem_u32(&raw_full[s]), da = smem_u32(raw(kb)), db = da + Cfg::A_BYTES;
        if (A_MN) {
#pragma unroll
          for (int b = 0; b < BM / 32; ++b) tma_load_2d(da + b * MNB_BYTES, &tmap_a, bar, m0 + 32 * b, k0);
        } else {
          tma_load_2d(da, &tmap_a, bar, k0, m0);
        }
        if (PAIR) {  // this CTA's half of the column tile (the tensor map's box is BN/2 rows)
          tma_load_2d(db, &tmap_b, bar, k0, n0 + (int)crank * (BN / 2));
        } else if (B_MN) {
#pragma unroll
          for (int b = 0; b < BN / 32; ++b) tma_load_2d(db + b * MNB_BYTES, &tmap_b, bar, n0 + 32 * b, k0);
        } else {
          tma_load_2d(db, &tmap_b, bar, k0, n0);
        }
      }
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer (pair: the leader CTA only) ----------------
    if (!PAIR || crank == 0) {
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        if (PAIR) mbar_wait_cluster(smem_u32(&lo_full[kb % T_NLO]), (kb / T_NLO) & 1);
        else mbar_wait(smem_u32(&lo_full[kb % T_NLO]), (kb / T_NLO) & 1);
        fence_async_smem();  // the converters' generic-proxy lo stores, observed through the barrier -> async proxy
        tc_fence_after();
        if (kb == 0) TC_TRACE(4);
        if (kb == nkb - 1) TC_TRACE(5);
        if (lane == 0) {
          const uint32_t ah = smem_u32(raw(kb)), bh = ah + Cfg::A_BYTES, al = smem_u32(lo(kb)), bl = al + Cfg::A_BYTES;
#ifndef PGNN_UB_NO_MMA
#pragma unroll
          for (int j = 0; j < TBK / 8; ++j) {
            const uint32_t first = (kb == 0 && j == 0) ? 0u : 1u;
            if (PAIR) {
              umma_tf32_pair(tmem_acc + BN, tma_desc<A_MN>(al, j), tma_desc<B_MN>(bh, j), idesc, first);  // cross terms
              umma_tf32_pair(tmem_acc + BN, tma_desc<A_MN>(ah, j), tma_desc<B_MN>(bl, j), idesc, 1u);
              umma_tf32_pair(tmem_acc, tma_desc<A_MN>(ah, j), tma_desc<B_MN>(bh, j), idesc, first);
            } else {
              umma_tf32(tmem_acc + BN, tma_desc<A_MN>(al, j), tma_desc<B_MN>(bh, j), idesc, first);  // cross terms
              umma_tf32(tmem_acc + BN, tma_desc<A_MN>(ah, j), tma_desc<B_MN>(bl, j), idesc, 1u);
              umma_tf32(tmem_acc, tma_desc<A_MN>(ah, j), tma_desc<B_MN>(bh, j), idesc, first);
            }
          }
#else
          (void)ah; (void)bh; (void)al; (void)bl;
#endif
          if (PAIR) {  // both CTAs' stages were read by these MMAs: release them in both
            umma_commit_pair(smem_u32(&lo_empty[kb % T_NLO]));
            umma_commit_pair(smem_u32(&raw_empty[kb % NRAW]));
            if (kb == nkb - 1) umma_commit_pair(smem_u32(&acc_bar));
          } else {
            umma_commit(smem_u32(&lo_empty[kb % T_NLO]));
            umma_commit(smem_u32(&raw_empty[kb % NRAW]));
            if (kb == nkb - 1) umma_commit(smem_u32(&acc_bar));
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- converters ----------------
    constexpr int V4 = Cfg::STAGE / 16;
#pragma unroll 1
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(smem_u32(&raw_full[kb % NRAW]), (kb / NRAW) & 1);  // TMA bytes have landed
      if (kb >= T_NLO) mbar_wait(smem_u32(&lo_empty[kb % T_NLO]), ((kb / T_NLO) - 1) & 1);
      const float4* src = reinterpret_cast<const float4*>(raw(kb));
      float4* dst = reinterpret_cast<float4*>(lo(kb));
#ifndef PGNN_UB_NO_CONVERT
#pragma unroll 4
      for (int i = threadIdx.x; i < V4; i += NPRODUCER) {
        const float4 v = src[i];
        float4 l;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        dst[i] = l;
      }
#else
      (void)src; (void)dst; (void)V4;
#endif
      if (PAIR) fence_async_smem();  // writer-side proxy fence: the peer CTA's tensor-core reads are ordered through a remote arrive
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_remote(smem_u32(&lo_full[kb % T_NLO]), 0u);
        else mbar_arrive(smem_u32(&lo_full[kb % T_NLO]));
      }
      if (warp == 0 && kb == 0) TC_TRACE(2);        // first block converted
      if (warp == 0 && kb == nkb - 1) TC_TRACE(3);  // last block converted
    }
  }
  if (warp < NPRODUCER / 32) {
    if (nkb > 0) mbar_wait(smem_u32(&acc_bar), 0);
    tc_fence_after();
    if (warp == 0) TC_TRACE(6);
#ifndef PGNN_UB_NO_EPILOGUE
    if (tma_store) tc_epilogue_tma<BN>(smem, s_bias, s_bias_on, tmem_acc, nkb, m0, n0, M, N, &tmap_c, ep);
    else tc_epilogue<BN>(smem, s_bias, s_bias_on, tmem_acc, nkb, m0, n0, M, N, C + (int64_t)blockIdx.z * ep.split_stride, ldc, ep);
    if (ep.fold_counter) splitk_fold_in_kernel<BN>(C, ldc, m0, n0, M, N, ep);
#endif
    if (warp == 0) TC_TRACE(7);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // neither CTA leaves (or frees its TMEM half) while the pair's MMAs or barrier arrivals may still target it
  if (warp == 0) TC_TRACE(8);
  if (warp == 0) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)tmem_cols<BN>()) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)tmem_cols<BN>()) : "memory");
  }
}

// ---- fused GIN layer front half: neighbour gather + GEMM1 in one kernel ------------------------------------------------------
// chem/model.py:37-55: aggr[i] = sum_{k: target(k) = i} act(x[source(k)]) + act(x[i]) + S[i,:].T (edge-feature embedding as a
// 9-bin dot product, neighbour reduction over the target-bucketed edge list, self-loop last, the previous layer's BatchNorm +
// ReLU applied on load) and z1 = relu(aggr . W1^T + b1).
//
// The gridDim.x CTAs that own the column tiles of ONE 128-row tile work as a team.  Phase 1: they split the row tile's reduction
// blocks among themselves (block kb goes to CTA kb mod gridDim.x); the 8 producer warps of each gather their blocks — a thread
// owns four (row, 16-byte chunk) items of a [128 x 32] block, loads the self row and up to four neighbour rows per item with
// every load in flight, reduces in edge order (bit-identical to k_aggregate_fwd) — and store the rows to `aggr` (the backward's
// weight-gradient GEMM needs them anyway).  Meanwhile the TMA warp already streams the first weight tiles.  A per-row-tile
// arrival counter (red.release.gpu after a generic->async proxy fence; the TMA thread spins with ld.acquire.gpu, time-bounded)
// publishes the tile inside the team; phase 2 is the main loop of k_gemm_3xtf32_tma with the A boxes loaded from the
// just-written, L2-resident rows.
//
// Why not gather straight into the swizzled operand stage (the first version of this kernel did): each of the three column
// tiles then repeats the whole gather, and 3 x 30 MB of L2 reads per layer made the fused kernel SLOWER than gather + GEMM
// (58 vs 43 us, measured).  Splitting the gather over the team keeps the L2 traffic of the unfused pair and removes what
// separated them: a launch, a full-chip drain and refill, and the GEMM prologue now overlaps the gather.
// Why a counter and not a thread-block cluster (the second version): clusters are placed inside one GPC, and only 45 clusters of
// three 221 KB CTAs fit this chip at once (cudaOccupancyMaxActiveClusters) while the B = 256 batch needs 47 - a second wave.
// A team's CTAs are adjacent in launch order (x fastest), so they become resident together; a CTA whose team mates are not
// resident yet only delays its A loads (every wait is time-bounded and traps instead of hanging).
struct GinGather {
  const float* x; int64_t ldx;                     // input rows [M, K], pre-affine
  const float* in_scale; const float* in_shift;    // explicit affine (or null)
  PgnnBnFold fold;                                 // or: derive it from the producer's BatchNorm sums (fold.acc != null)
  int relu;
  const int* rowptr; const int* nbr;               // edges bucketed by target
  const float* S; int Q; const float* T; const float* T2; int q_split;   // per-node edge-feature summary and the stacked tables
  float* aggr; int64_t lda;                        // the aggregated rows (output of phase 1, A operand of phase 2)
};

constexpr int G_KPAD = 320;       // K <= 320 (affine / table rows staged in shared memory)
constexpr int G_MAXQ = 9;
constexpr int G_STAGING = (2 * G_KPAD + BM * G_MAXQ + G_MAXQ * G_KPAD) * 4;  // lives in the (idle) lo stages

__device__ __forceinline__ float4 g_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// 16-byte load that does not allocate in L1 (ld.global.cg), predicated without a branch.  This kernel carves 221 KB of the SM's
// 256 KB out as shared memory: ~30 KB of L1 = ~240 lines, and an L1-allocating load holds a line while it is in flight, so the
// 5120 row loads a CTA issues per reduction block were throttled to what 240 lines turn over (measured: 5 us per block, 1.5 TB/s
// chip-wide, the same data the stand-alone gather moves at 3.8 TB/s with the whole 256 KB as L1).
__device__ __forceinline__ float4 g_ldcg4(const float* p, bool pred) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ float4 g_act(float4 v, float4 sc, float4 sh, bool affine, bool relu) {
  if (affine) { v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w); }
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  return v;
}
__device__ __forceinline__ void g_add4(float4& a, float4 v) {
  a.x = __fadd_rn(a.x, v.x); a.y = __fadd_rn(a.y, v.y); a.z = __fadd_rn(a.z, v.z); a.w = __fadd_rn(a.w, v.w);
}

// development stamps of CTA (0,0) of the fused kernel (pgnn_debug_gather_trace reads them)
__device__ unsigned long long g_gather_trace[16];
#define G_TRACE(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0) g_gather_trace[slot] = globaltimer_ns(); } while (0)

template <int BN>
__global__ void __launch_bounds__(T_NTHREADS, 1)
k_gin_gather_gemm(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, int M, int N, int K, TcEpilogue ep, GinGather g, unsigned int* tile_ctr) {
  static_assert(BN % 16 == 0 && BN <= 256, "UMMA N % 16 == 0 at M = 128");
  using Cfg = TmaCfg<BN, false>;
  constexpr int NRAW = Cfg::NRAW, T_NLO = Cfg::NLO;
  static_assert(T_NLO * Cfg::STAGE >= G_STAGING, "the gather's staging area lives in the lo stages");
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t raw_full[NRAW], raw_empty[NRAW], lo_full[T_NLO], lo_empty[T_NLO], acc_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[BN];

  auto raw = [&](int kb) { return smem + (kb % NRAW) * Cfg::STAGE; };
  auto lo = [&](int kb) { return smem + (NRAW + kb % T_NLO) * Cfg::STAGE; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb = (K + TBK - 1) / TBK;
  const bool s_bias_on = ep.bias != nullptr;
  if (warp == 0) G_TRACE(0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)tmem_cols<BN>()) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < NRAW; ++s) {
      mbar_init(smem_u32(&raw_full[s]), 1);
      mbar_init(smem_u32(&raw_empty[s]), 1);
    }
    for (int s = 0; s < T_NLO; ++s) {
      mbar_init(smem_u32(&lo_full[s]), NPRODUCER / 32);
      mbar_init(smem_u32(&lo_empty[s]), 1);
    }
    mbar_init(smem_u32(&acc_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_s;
  constexpr uint32_t idesc = umma_idesc(BM, BN, false, false);
  const uint32_t crank = blockIdx.x, csize = gridDim.x;   // this CTA's place in the row tile's team
  pdl_prologue();
  if (warp == 0) G_TRACE(1);

  if (warp == 9) {
    // ---------------- TMA producer ----------------
    const int pre = nkb < NRAW ? nkb : NRAW;
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
      for (int kb = 0; kb < pre; ++kb) {   // the weight tiles do not depend on the gather: stream them under phase 1
        mbar_expect_tx(smem_u32(&raw_full[kb]), Cfg::STAGE);
        tma_load_2d(smem_u32(raw(kb)) + Cfg::A_BYTES, &tmap_b, smem_u32(&raw_full[kb]), kb * TBK, n0);
      }
    }
    __syncwarp();
    if (lane == 0) {
      // every CTA of the team has stored (and fenced) its share of the row tile
      const unsigned int* ctr = tile_ctr + blockIdx.y;
      const uint64_t t0 = globaltimer_ns();
      for (uint32_t it = 0;; ++it) {
        unsigned int seen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
        if (seen >= csize) break;
        if ((it & 255u) == 255u && globaltimer_ns() - t0 > 2000000000ull) asm volatile("trap;");
      }
      asm volatile("fence.proxy.async;" ::: "memory");   // the team's generic-proxy stores, acquired above -> this thread's TMA reads
      G_TRACE(6);
      for (int kb = 0; kb < pre; ++kb) tma_load_2d(smem_u32(raw(kb)), &tmap_a, smem_u32(&raw_full[kb]), kb * TBK, m0);
#pragma unroll 1
      for (int kb = pre; kb < nkb; ++kb) {
        const int s = kb % NRAW;
        mbar_wait(smem_u32(&raw_empty[s]), ((kb / NRAW) - 1) & 1);
        mbar_expect_tx(smem_u32(&raw_full[s]), Cfg::STAGE);
        const uint32_t bar = smem_u32(&raw_full[s]), da = smem_u32(raw(kb));
        tma_load_2d(da, &tmap_a, bar, kb * TBK, m0);
        tma_load_2d(da + Cfg::A_BYTES, &tmap_b, bar, kb * TBK, n0);
      }
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer ----------------
#pragma unroll 1
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(smem_u32(&lo_full[kb % T_NLO]), (kb / T_NLO) & 1);
      fence_async_smem();
      tc_fence_after();
      if (kb == 0) G_TRACE(8);
      if (lane == 0) {
        const uint32_t ah = smem_u32(raw(kb)), bh = ah + Cfg::A_BYTES, al = smem_u32(lo(kb)), bl = al + Cfg::A_BYTES;
#pragma unroll
        for (int j = 0; j < TBK / 8; ++j) {
          const uint32_t first = (kb == 0 && j == 0) ? 0u : 1u;
          umma_tf32(tmem_acc + BN, tma_desc<false>(al, j), tma_desc<false>(bh, j), idesc, first);  // cross terms
          umma_tf32(tmem_acc + BN, tma_desc<false>(ah, j), tma_desc<false>(bl, j), idesc, 1u);
          umma_tf32(tmem_acc, tma_desc<false>(ah, j), tma_desc<false>(bh, j), idesc, first);
        }
        umma_commit(smem_u32(&lo_empty[kb % T_NLO]));
        umma_commit(smem_u32(&raw_empty[kb % NRAW]));
        if (kb == nkb - 1) umma_commit(smem_u32(&acc_bar));
      }
      __syncwarp();
    }
  } else {
    // ---------------- phase 1: gather this CTA's share of the row tile ----------------
    // Staging: every global load of the prologue is issued before the first shared-memory store (one memory round trip instead of
    // one per table: the first version spent 5.7 us here), and only for the columns of THIS CTA's reduction blocks.
    const int tid = threadIdx.x;
    float* s_aff = reinterpret_cast<float*>(lo(0));                  // [2][G_KPAD]   (the lo stages are idle until phase 2)
    float* s_S = s_aff + 2 * G_KPAD;                                // [BM][G_MAXQ]
    float* s_T = s_S + BM * G_MAXQ;                                 // [G_MAXQ][G_KPAD]
    const bool affine = g.fold.acc != nullptr || g.in_scale != nullptr;
    const bool relu = g.relu != 0;
    const int j = tid & 7;      // 16-byte chunk of the 128-byte block row
    const int rb = tid >> 3;    // rows rb, rb + 32, rb + 64, rb + 96
    // (a) this thread's four rows: in-degree and the first four neighbour ids, loop-invariant over the reduction blocks
    int lo_[4], deg[4], src[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = m0 + rb + 32 * i;
      const bool ok = gi < M;
      lo_[i] = ok ? __ldg(g.rowptr + gi) : 0;
      deg[i] = ok ? __ldg(g.rowptr + gi + 1) : -1;
    }
    // (b) tables: my own columns are those of blocks kb = crank, crank + csize, ...: column c is mine iff (c / 32) % csize == crank
    const int my_blocks = (nkb - (int)crank + (int)csize - 1) / (int)csize;
    const int my_cols = my_blocks * TBK;                       // <= 128 for K = 300 and three column tiles
    // own column index (0 .. my_cols) -> global column
    auto own_col = [&](int oc) { return ((oc / TBK) * (int)csize + (int)crank) * TBK + (oc % TBK); };
    float bias_v = 0.f;
    if (tid < BN) bias_v = (s_bias_on && n0 + tid < N) ? __ldg(ep.bias + n0 + tid) : 0.f;
    float sv[5];   // BM * G_MAXQ = 1152 = 4.5 x 256
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int i = tid + u * NPRODUCER;
      const int r = i / G_MAXQ, q = i - r * G_MAXQ;
      sv[u] = (g.S && i < BM * G_MAXQ && m0 + r < M && q < g.Q) ? __ldg(g.S + (int64_t)(m0 + r) * g.Q + q) : 0.f;
    }
    float4 tv[5];  // G_MAXQ x my_cols / 4 <= 9 x 32 = 288 float4... per 32 columns; up to 4 blocks: 1152 = 4.5 x 256
    const int tcnt = G_MAXQ * (my_cols / 4);
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int i = tid + u * NPRODUCER;
      tv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g.S && i < tcnt) {
        const int q = i / (my_cols / 4), c = own_col((i - q * (my_cols / 4)) * 4);
        if (q < g.Q && c < K)
          tv[u] = __ldg(reinterpret_cast<const float4*>(q < g.q_split ? g.T + (int64_t)q * K + c : g.T2 + (int64_t)(q - g.q_split) * K + c));
      }
    }
    // (c) the affine of my columns (one column per thread: my_cols <= 128 < 256; wider shares loop)
    for (int oc = tid; oc < my_cols; oc += NPRODUCER) {
      const int c = own_col(oc);
      if (c < K) {
        if (g.fold.acc) bn_fold_column(g.fold, K, c, blockIdx.y == 0, s_aff[c], s_aff[G_KPAD + c]);   // row tile 0's team performs the module-state updates, each column once
        else if (g.in_scale) { s_aff[c] = __ldg(g.in_scale + c); s_aff[G_KPAD + c] = __ldg(g.in_shift + c); }
      }
    }
    // second round trip: neighbour ids (needs the row pointers)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (deg[i] >= 0) deg[i] -= lo_[i];
#pragma unroll
      for (int k = 0; k < 4; ++k) src[i][k] = (k < deg[i]) ? __ldg(g.nbr + lo_[i] + k) : 0;
    }
    if (tid < BN) s_bias[tid] = bias_v;
    if (g.S) {
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int i = tid + u * NPRODUCER;
        if (i < BM * G_MAXQ) s_S[i] = sv[u];
        if (i < tcnt) {
          const int q = i / (my_cols / 4), c = own_col((i - q * (my_cols / 4)) * 4);
          if (c < G_KPAD) *reinterpret_cast<float4*>(s_T + q * G_KPAD + c) = tv[u];
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NPRODUCER) : "memory");
    if (warp == 0) G_TRACE(3);
#pragma unroll 1
    for (int kb = (int)crank; kb < nkb; kb += (int)csize) {
      const int col = kb * TBK + j * 4;
      if (col >= K) continue;   // K % 4 == 0: whole chunks
      // every row load of the block is issued before the first use; absent neighbours are predicated off inside the load
      // instruction (with an `if (k < deg)` branch around the loads the compiler serialised them: 10 us per block)
      float4 vs[4], vn[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gi = m0 + rb + 32 * i;
        vs[i] = g_ldcg4(g.x + (int64_t)(deg[i] >= 0 ? gi : 0) * g.ldx + col, deg[i] >= 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) vn[i][k] = g_ldcg4(g.x + (int64_t)src[i][k] * g.ldx + col, k < deg[i]);
      }
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (affine) { sc = g_ld4(s_aff + col); sh = g_ld4(s_aff + G_KPAD + col); }
      // the edge term does not depend on the gathered rows: it is computed while they are in flight
      float4 ev[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // all Q bins unconditionally: fmaf(0, t, e) == e, so this equals k_aggregate_fwd's skip of the empty bins bit for bit
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.S) {
          const float* sr = s_S + (rb + 32 * i) * G_MAXQ;
#pragma unroll
          for (int q = 0; q < G_MAXQ; ++q) {   // rows of s_S / s_T beyond Q are zero-filled
            const float w = sr[q];
            const float4 t = g_ld4(s_T + q * G_KPAD + col);
            e.x = fmaf(w, t.x, e.x); e.y = fmaf(w, t.y, e.y); e.z = fmaf(w, t.z, e.z); e.w = fmaf(w, t.w, e.w);
          }
        }
        ev[i] = e;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gi = m0 + rb + 32 * i;
        if (deg[i] < 0) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float4 v = g_act(vn[i][k], sc, sh, affine, relu);
          if (k >= deg[i]) v = make_float4(0.f, 0.f, 0.f, 0.f);   // acc + 0 == acc exactly (acc starts at +0 and never becomes -0)
          g_add4(acc, v);
        }
        if (deg[i] > 4) {   // rare for molecules (a fifth bond): a plain serial loop, deliberately not unrolled (code size)
#pragma unroll 1
          for (int k = 4; k < deg[i]; ++k)
            g_add4(acc, g_act(g_ldcg4(g.x + (int64_t)__ldg(g.nbr + lo_[i] + k) * g.ldx + col, true), sc, sh, affine, relu));
        }
        g_add4(acc, g_act(vs[i], sc, sh, affine, relu));   // self-loop last (chem/model.py:39)
        if (g.S) g_add4(acc, ev[i]);
        *reinterpret_cast<float4*>(g.aggr + (int64_t)gi * g.lda + col) = acc;
      }
    }
    if (warp == 0) G_TRACE(4);
    __threadfence();
    if (warp == 0) G_TRACE(5);
    asm volatile("fence.proxy.async;" ::: "memory");   // this thread's generic-proxy stores of aggr -> visible to the team's TMA reads
    asm volatile("bar.sync 1, %0;" ::"n"(NPRODUCER) : "memory");
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(tile_ctr + blockIdx.y) : "memory");
    if (warp == 0) G_TRACE(7);
    // ---------------- phase 2: converters, exactly k_gemm_3xtf32_tma's ----------------
    constexpr int V4 = Cfg::STAGE / 16;
#pragma unroll 1
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(smem_u32(&raw_full[kb % NRAW]), (kb / NRAW) & 1);
      if (kb >= T_NLO) mbar_wait(smem_u32(&lo_empty[kb % T_NLO]), ((kb / T_NLO) - 1) & 1);
      const float4* src = reinterpret_cast<const float4*>(raw(kb));
      float4* dst = reinterpret_cast<float4*>(lo(kb));
#pragma unroll 4
      for (int i = threadIdx.x; i < V4; i += NPRODUCER) {
        const float4 v = src[i];
        float4 l;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        dst[i] = l;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&lo_full[kb % T_NLO]));
    }
  }
  if (warp < NPRODUCER / 32) {
    if (nkb > 0) mbar_wait(smem_u32(&acc_bar), 0);
    tc_fence_after();
    if (warp == 0) G_TRACE(9);
    tc_epilogue_tma<BN>(smem, s_bias, s_bias_on, tmem_acc, nkb, m0, n0, M, N, &tmap_c, ep);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) G_TRACE(10);
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)tmem_cols<BN>()) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map of one operand of extent R (row index) x Kred (reduction):
//   K-major  (element (r,k) at base[r*ld + k]): global tensor [R][Kred], box [box_r x 32 k], SWIZZLE_128B
//   MN-major (element (r,k) at base[k*ld + r]): global tensor [Kred][R], box [32 k x 32 r], SWIZZLE_128B_ATOM_32B
bool make_map(CUtensorMap* map, const float* base, bool mn, int64_t R, int64_t Kred, int64_t ld, int box_r) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)(mn ? R : Kred), (cuuint64_t)(mn ? Kred : R)};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)(mn ? TBK : box_r)};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Encoding a tensor map costs a few microseconds on the host and a training step needs ~60 of them (2 per GEMM), while
// the operands recur: weights keep their addresses and the caching allocator hands the same workspace blocks back for
// recurring batch shapes.  Memoise by (address, extents, stride, major, box).
struct MapKey {
  const void* base; int64_t R, K, ld; int mn, box;
  bool operator==(const MapKey& o) const { return base == o.base && R == o.R && K == o.K && ld == o.ld && mn == o.mn && box == o.box; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    for (int64_t v : {k.R, k.K, k.ld, (int64_t)k.mn, (int64_t)k.box}) h = h * 1000003u ^ (size_t)v;
    return h;
  }
};
bool cached_map(CUtensorMap* out, const float* base, bool mn, int64_t R, int64_t Kred, int64_t ld, int box_r) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  const MapKey key{base, R, Kred, ld, mn ? 1 : 0, box_r};
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  alignas(64) CUtensorMap m;
  if (!make_map(&m, base, mn, R, Kred, ld, box_r)) return false;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return true;
}

// Output tensor map of the TMA-store epilogue: global tensor [splits][M][N] fp32 (row stride ldc, split stride split_stride
// floats), box [1][128 rows][16 cols], SWIZZLE_64B.  The extents are the true ones: the copy engine clips ragged edge tiles.
struct MapCKey {
  const void* base; int64_t M, N, ldc, splits, stride;
  bool operator==(const MapCKey& o) const {
    return base == o.base && M == o.M && N == o.N && ldc == o.ldc && splits == o.splits && stride == o.stride;
  }
};
struct MapCKeyHash {
  size_t operator()(const MapCKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    for (int64_t v : {k.M, k.N, k.ldc, k.splits, k.stride}) h = h * 1000003u ^ (size_t)v;
    return h;
  }
};
bool cached_map_c(CUtensorMap* out, const float* base, int64_t M, int64_t N, int64_t ldc, int64_t splits, int64_t split_stride) {
  static std::unordered_map<MapCKey, CUtensorMap, MapCKeyHash> cache;
  static std::mutex mu;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  if (splits <= 1) { splits = 1; split_stride = M * ldc; }
  const MapCKey key{base, M, N, ldc, splits, split_stride};
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  alignas(64) CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)splits};
  cuuint64_t strides[2] = {(cuuint64_t)ldc * 4, (cuuint64_t)split_stride * 4};
  cuuint32_t box[3] = {(cuuint32_t)CBOX, (cuuint32_t)BM, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  if (fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return true;
}
// PGNN_TMA_STORE=0: the st.global epilogue of tc_common.cuh (development switch / A-B measurement)
bool tma_store_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PGNN_TMA_STORE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// CTA pairs (cta_group::2) for the K-major x K-major GEMMs: opt-in with PGNN_PAIR=1.  Measured on B200 (tools/check_tc.py,
// tools/trace_tc.py): numerically identical to the single-CTA kernel, but a 32-deep block costs 1.14-1.17 us whatever the tile
// width (BN = 224 and 128 alike, 1.0 us for a lone cluster on an idle GPU, with or without the proxy fence / cluster-scope
// wait), i.e. ~95 ns per UTCHMMA.2CTA with K = 8, against 1.16 us (BN 224) / 0.83 us (BN 128) for the single-CTA kernel.
// GEMM1 31.0 vs 27.6 us, GEMM2 38.7 vs 29.2 us.  Kept as the starting point for round 2.
// (An earlier 2-CTA variant that only TMA-multicast the B tile into both CTAs also measured no gain.)
bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PGNN_PAIR");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <bool A_MN, bool B_MN, int BN>
int launch_tma(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N, int K, int splits,
               int k_per_split, const TcEpilogue& ep, cudaStream_t st) {
  const int mt = (int)ceil_div(M, BM), nt = (int)ceil_div(N, BN);
  constexpr bool kPairable = !A_MN && !B_MN && (BN % 32 == 0);
  // pairs of row tiles share the column tile; an odd tile count gets an all-padding partner, which must not cost a wave
  const int64_t ctas = (int64_t)mt * nt, ctas_pair = (int64_t)(mt + (mt & 1)) * nt;
  const bool pair = kPairable && pair_enabled() && splits == 1 && mt >= 2 && ceil_div(ctas_pair, kNumSMs) <= ceil_div(ctas, kNumSMs);
  alignas(64) CUtensorMap ma, mb, mc;
  if (!cached_map(&ma, A, A_MN, M, K, lda, BM) || !cached_map(&mb, B, B_MN, N, K, ldb, pair ? BN / 2 : BN)) return PGNN_EUNSUPPORTED;
  // TMA-store epilogue whenever the output is a plain (non-atomic) store with 16-byte aligned rows
  int tma_store = 0;
  if (tma_store_enabled() && !ep.atomic && !ep.fold_counter && (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
      (splits == 1 || (ep.split_stride & 3) == 0) && cached_map_c(&mc, C, M, N, ldc, splits, ep.split_stride))
    tma_store = 1;
  if (!tma_store) mc = ma;  // unused by the kernel, but a kernel parameter must be a valid object
  static bool configured = false;
  if (!configured) {
    PGNN_CUDA(cudaFuncSetAttribute(k_gemm_3xtf32_tma<A_MN, B_MN, BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   TmaCfg<BN, false>::SMEM));
    if constexpr (kPairable)
      PGNN_CUDA(cudaFuncSetAttribute(k_gemm_3xtf32_tma<A_MN, B_MN, BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TmaCfg<BN, true>::SMEM));
    configured = true;
  }
  if constexpr (kPairable) {
    if (pair) {
      dim3 grid((unsigned)(mt + (mt & 1)), (unsigned)nt, 1u);
      PGNN_CUDA(pgnn_launch_cluster_x(k_gemm_3xtf32_tma<A_MN, B_MN, BN, true>, dim3(grid), dim3(T_NTHREADS), TmaCfg<BN, true>::SMEM, st, 2u,
                                      ma, mb, mc, tma_store, C, ldc, M, N, K, k_per_split, ep));
      PGNN_LAUNCH_CHECK();
      return PGNN_OK;
    }
  }
  dim3 grid((unsigned)nt, (unsigned)mt, (unsigned)splits);
  PGNN_CUDA(pgnn_launch(k_gemm_3xtf32_tma<A_MN, B_MN, BN, false>, dim3(grid), dim3(T_NTHREADS), TmaCfg<BN, false>::SMEM, st, ma, mb, mc, tma_store, C, ldc, M, N,
                        K, k_per_split, ep));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

template <bool A_MN, bool B_MN>
int dispatch_tma(int bn, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N, int K,
                 int splits, int k_per_split, const TcEpilogue& ep, cudaStream_t st) {
  if constexpr (!A_MN && !B_MN) {  // K-major boxes take any BN % 16 == 0
    if (bn == 112) return launch_tma<A_MN, B_MN, 112>(A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
    if (bn == 208) return launch_tma<A_MN, B_MN, 208>(A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
  }
  switch (bn) {
    case 64: return launch_tma<A_MN, B_MN, 64>(A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
    case 128: return launch_tma<A_MN, B_MN, 128>(A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
    case 160: return launch_tma<A_MN, B_MN, 160>(A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
    default: return launch_tma<A_MN, B_MN, 224>(A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
  }
}

// Opt-in (PGNN_FUSED_GATHER=1 or pgnn_debug_set_fused_gather(1)).  Measured on the masking step (B = 256, one B200, same box):
// fused 44.5 us per layer against 20.1 + 23 us for k_aggregate_fwd + GEMM1 -- the step is 1.154 ms fused, 1.120 ms unfused.  The
// gather phase moves its 30 MB at the same ~3 TB/s of L2 gather bandwidth as the stand-alone kernel (11.8 us for a CTA's four
// blocks), and what fusion removes (one launch, one drain) is paid back by the team hand-over (fence + counter + first A load:
// 3.6 us) and the prologue's two dependent round trips (4.1 us) on the critical path of EVERY CTA; a 128-row tile cannot hide a
// memory-bound phase under a tensor-bound one inside one CTA that holds 221 KB of shared memory (no co-resident partner).
std::atomic<int> g_fused_gather{-1};
bool fused_gather_enabled() {
  int v = g_fused_gather.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("PGNN_FUSED_GATHER");
    v = (e && e[0] == '1') ? 1 : 0;
    g_fused_gather.store(v, std::memory_order_relaxed);
  }
  return v == 1;
}

template <int BN>
int launch_gin_gather(const GinGather& g, const float* W, int64_t ldw, float* C, int64_t ldc, int M, int N, int K, const TcEpilogue& ep,
                      unsigned int* tile_ctr, cudaStream_t st) {
  const int nt = (int)ceil_div(N, BN), mt = (int)ceil_div(M, BM);
  alignas(64) CUtensorMap ma, mb, mc;
  if (!cached_map(&ma, g.aggr, false, M, K, g.lda, BM) || !cached_map(&mb, W, false, N, K, ldw, BN) || !cached_map_c(&mc, C, M, N, ldc, 1, 0))
    return PGNN_EUNSUPPORTED;
  static bool configured = false;
  if (!configured) {
    PGNN_CUDA(cudaFuncSetAttribute(k_gin_gather_gemm<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TmaCfg<BN, false>::SMEM));
    configured = true;
  }
  dim3 grid((unsigned)nt, (unsigned)mt, 1u);   // x fastest: the CTAs of a team are adjacent in launch order
  PGNN_CUDA(pgnn_launch(k_gin_gather_gemm<BN>, dim3(grid), dim3(T_NTHREADS), TmaCfg<BN, false>::SMEM, st, ma, mb, mc, M, N, K, ep, g, tile_ctr));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

}  // namespace

// C[M,N] = sum_k A(m,k) B(n,k) (+ epilogue) with TMA-staged operands; a_mn / b_mn select the operand major (see the top
// of the file).  splits > 1 requires k_per_split % 32 == 0 and an atomic epilogue.  PGNN_EUNSUPPORTED when the layout does
// not meet the TMA constraints (16-byte aligned base, row stride multiple of 16 bytes) or the driver entry is missing.
int pgnn_tma_gemm(bool a_mn, bool b_mn, int bn, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M,
                  int N, int K, int splits, int k_per_split, const TcEpilogue& ep, cudaStream_t st) {
  if ((lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return PGNN_EUNSUPPORTED;
  if (splits > 1 && (k_per_split % TBK)) return PGNN_EUNSUPPORTED;
  if (!a_mn && !b_mn) return dispatch_tma<false, false>(bn, A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
  if (!a_mn && b_mn) return dispatch_tma<false, true>(bn, A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
  if (a_mn && b_mn) return dispatch_tma<true, true>(bn, A, lda, B, ldb, C, ldc, M, N, K, splits, k_per_split, ep, st);
  return PGNN_EUNSUPPORTED;
}

// aggr = gather(x) and z1[M,N] = relu?(aggr . W^T + bias) in one kernel (k_gin_gather_gemm; see GinGather).  PGNN_EUNSUPPORTED when the
// shape / alignment does not fit (the caller then runs pgnn_internal_aggregate_fwd + the plain GEMM).
int pgnn_tma_gin_gather_gemm(const float* x, int64_t ldx, const float* in_scale, const float* in_shift, const PgnnBnFold* fold, int in_relu,
                             const int32_t* rowptr_t, const int32_t* nbr_t, const float* S, int Q, const float* T, const float* T2,
                             int q_split, float* aggr, int64_t lda, const float* W, const float* bias, int relu, float* z1, int64_t ldz,
                             int64_t M, int64_t N, int64_t K, unsigned int* tile_ctr, cudaStream_t st) {
  // tile_ctr: ceil(M / 128) zeroed uint32 (arrival counters of the row-tile teams; consumed by this call)
  if (!tile_ctr || !fused_gather_enabled() || !tma_store_enabled()) return PGNN_EUNSUPPORTED;
  if (K % 4 || K > G_KPAD || ldx % 4 || !aggr || lda % 4 || ldz % 4 || (S && (Q < 1 || Q > G_MAXQ)) || M < 1) return PGNN_EUNSUPPORTED;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(x) || !al(W) || !al(z1) || (aggr && !al(aggr)) || (T && !al(T)) || (T2 && !al(T2))) return PGNN_EUNSUPPORTED;
  GinGather g{};
  g.x = x; g.ldx = ldx; g.in_scale = in_scale; g.in_shift = in_shift;
  if (fold) g.fold = *fold;
  g.relu = in_relu; g.rowptr = rowptr_t; g.nbr = nbr_t; g.S = S; g.Q = Q; g.T = T; g.T2 = T2 ? T2 : T; g.q_split = T2 ? q_split : Q;
  g.aggr = aggr; g.lda = lda;
  TcEpilogue ep{bias, relu, nullptr, 0, 0, PgnnGemmHooks{}};
  if (N <= 112) return launch_gin_gather<112>(g, W, K, z1, ldz, (int)M, (int)N, (int)K, ep, tile_ctr, st);
  return launch_gin_gather<208>(g, W, K, z1, ldz, (int)M, (int)N, (int)K, ep, tile_ctr, st);
}

extern "C" __attribute__((visibility("default"))) int pgnn_debug_tma_trace(unsigned long long* host16) {
  return cudaMemcpyFromSymbol(host16, g_tc_trace, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -2;
}

// development / tests: 1 = fused gather + GEMM1 kernel on the chem GIN encoder path, 0 = separate kernels, -1 = re-read PGNN_FUSED_GATHER
extern "C" __attribute__((visibility("default"))) int pgnn_debug_set_fused_gather(int on) {
  g_fused_gather.store(on < 0 ? -1 : (on ? 1 : 0), std::memory_order_relaxed);
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pgnn_debug_gather_trace(unsigned long long* host16) {
  return cudaMemcpyFromSymbol(host16, g_gather_trace, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -2;
}

#ifdef PGNN_TRACE_ALL
// reset != 0: prime the envelope (min = ~0, max = 0) before the launch of interest; else read [min 16 | max 16]
extern "C" __attribute__((visibility("default"))) int pgnn_debug_tma_trace_env(unsigned long long* host32, int reset) {
  if (reset) {
    unsigned long long init[32];
    for (int i = 0; i < 16; ++i) { init[i] = ~0ull; init[16 + i] = 0ull; }
    return cudaMemcpyToSymbol(g_tc_trace_env, init, sizeof init) == cudaSuccess ? 0 : -2;
  }
  return cudaMemcpyFromSymbol(host32, g_tc_trace_env, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -2;
}
#endif

int pgnn_tma_gemm_kk(int bn, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N, int K,
                     const TcEpilogue& ep, cudaStream_t st) {
  return pgnn_tma_gemm(false, false, bn, A, lda, B, ldb, C, ldc, M, N, K, 1, K, ep, st);
}
