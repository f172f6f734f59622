// Shared device helpers of the tensor-core GEMM kernels (dense_tc.cu: cp.async producers, any operand major;
// dense_tma.cu: TMA producers, K-major operands): UMMA descriptors, tcgen05 / mbarrier wrappers and the epilogue.
#pragma once
#include "common.cuh"

namespace {

constexpr int BM = 128;       // UMMA M (TMEM lanes)
constexpr int NPRODUCER = 256;  // 8 producer warps (they also run the epilogue)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
#ifdef PGNN_TRACE_ALL
// development builds only (tools/ubench_gemm.py): the envelope of every CTA's stamps, [0,16) earliest, [16,32) latest
__device__ unsigned long long g_tc_trace_env[32];
#define TC_TRACE(slot)                                                                        \
  do {                                                                                        \
    if (lane == 0) {                                                                          \
      const unsigned long long t__ = globaltimer_ns();                                        \
      if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_trace[slot] = t__;      \
      atomicMin(&g_tc_trace_env[slot], t__);                                                  \
      atomicMax(&g_tc_trace_env[16 + slot], t__);                                             \
    }                                                                                         \
  } while (0)
#else
#define TC_TRACE(slot)                                                                        \
  do {                                                                                        \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) g_tc_trace[slot] = globaltimer_ns(); \
  } while (0)
#endif


// Two fp32 accumulators per tile: hi*hi in columns [0,BN), the two cross terms in [BN,2BN).  The tensor
// core adds each k-step into the accumulator with truncation, so the error of a chain grows with its
// length; keeping the (2^-11 times smaller) cross terms out of the main chain cuts its length by 3x and
// brings the GEMM to plain-fp32 accuracy (measured, tools/check_tc.py).
template <int BN>
__host__ __device__ constexpr int tmem_cols() { return 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512; }


// Epilogue shared by the GEMM kernels.  Called by the NPRODUCER (256) non-MMA threads once every MMA has completed
// (the operand stages in `smem` are free).  `nkb == 0` means the accumulators were never written (treated as zero).
template <int BN>
__device__ __forceinline__ void tc_epilogue(uint8_t* smem, const float* s_bias, bool s_bias_on, uint32_t tmem_acc, int nkb, int m0,
                                            int n0, int M, int N, float* __restrict__ C, int64_t ldc, const TcEpilogue& ep) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- epilogue ----
  // (1) TMEM -> registers -> smem staging tile [128][BN+4] (the operand stages are free: every MMA has
  //     completed).  Warp w owns TMEM lanes 32*(w%4).., column half w/4; both accumulators are summed here.
  // (2) the 8 warps write the tile out row-contiguously (a warp instruction covers 512 consecutive bytes of
  //     one output row), applying bias / ReLU / mask on the way.  Writing straight from the TMEM register
  //     layout (one row per lane) would issue 16-byte stores to 32 different rows per instruction.
  constexpr int SLD = BN + 4;  // staging row stride in floats: 16 B aligned, quarter-warps hit distinct banks
  float* stage = reinterpret_cast<float*>(smem);
  {
    // columns [0, CSPLIT) go to warps 0-3, [CSPLIT, BN) to warps 4-7; both extents are multiples of the 16-column TMEM load
    constexpr int CSPLIT = (BN + 31) / 32 * 16;
    static_assert(BN % 16 == 0, "the epilogue reads TMEM 16 columns at a time");
    const int row = (warp & 3) * 32 + lane;
    const int cbeg = (warp >> 2) * CSPLIT;
    const int cnum = (warp >> 2) ? BN - CSPLIT : CSPLIT;
#pragma unroll 1
    for (int c = 0; c < cnum; c += 16) {
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
        // one 16-byte read: four scalar reads at a lane stride of 4 floats are 4-way bank conflicts each
        const float4 b4 = s_bias_on ? *reinterpret_cast<const float4*>(s_bias + (gn[u] - n0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (s_bias_on) ov[q] += bv[q];
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
}

}  // namespace
