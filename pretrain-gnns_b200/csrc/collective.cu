// Gradient all-reduce over NVLink peer memory (one all-reduce of the flat gradient buffer per step).
//
// The flat gradient buffer of every rank lives at the same offset of a symmetric allocation, so each rank holds the
// peer-mapped address of every other rank's buffer and of a small flag array.  Two-shot, in place, no library collective:
//
//   barrier  (every rank's backward has finished and is visible)
//   reduce-scatter: rank r sums slice r over all ranks' buffers (peer loads over NVLink, fixed rank order) -> local scratch
//   barrier  (every rank has finished READING every buffer, so slices may be overwritten)
//   all-gather: rank r stores its reduced slice into slice r of every rank's buffer (peer stores)
//   barrier  (all stores have landed)
//
// Each element is summed by exactly one rank in the order 0..G-1, so all ranks end with bit-identical gradients and the
// result does not depend on timing.  A cross-GPU barrier is a one-CTA kernel: thread t publishes a monotonically increasing
// sequence number to rank t's flag word [rank] (st.release.sys after a system fence) and spins on its own flag word [t]
// (ld.acquire.sys, time-bounded: a lost peer traps after ~60 s instead of hanging the GPU; the bound is generous because ranks
// may reach their first exchange seconds apart -- module loading, first-touch allocations).  Kernel boundaries order the
// phases inside a GPU.  For the 7.4 MB chem-GIN buffer on 2 GPUs this replaces a ~180 us NCCL call.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t gtimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(32) k_xgpu_barrier(uint32_t* const* __restrict__ flags, int rank, int world, uint32_t seq) {
  pdl_prologue();  // the previous kernel of this stream has completed and flushed
  const int t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    uint32_t* remote = flags[t] + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(seq) : "memory");
    const uint32_t* mine = flags[rank] + t;
    const uint64_t t0 = gtimer_ns();
    for (uint32_t it = 0;; ++it) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - seq) >= 0) break;
      if ((it & 255u) == 255u && gtimer_ns() - t0 > 60000000000ull) asm volatile("trap;");
    }
  }
}

__device__ __forceinline__ float4 ld_peer4(const float* p) { return __ldcv(reinterpret_cast<const float4*>(p)); }

// scratch[i - lo] = scale * sum_k bufs[k][i] for i in [lo, hi)
__global__ void __launch_bounds__(256)
k_reduce_slice(float* const* __restrict__ bufs, int world, int64_t lo, int64_t hi, float scale, float* __restrict__ scratch) {
  pdl_prologue();
  const int64_t n = hi - lo;
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(bufs[0] + lo) & 15) == 0) ? n / 4 : 0;  // symmetric offsets: same alignment on every rank
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = ld_peer4(bufs[0] + lo + 4 * i);
    for (int k = 1; k < world; ++k) {
      const float4 b = ld_peer4(bufs[k] + lo + 4 * i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(scratch)[i] = make_float4(a.x * scale, a.y * scale, a.z * scale, a.w * scale);
  }
  for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a = __ldcv(bufs[0] + lo + i);
    for (int k = 1; k < world; ++k) a += __ldcv(bufs[k] + lo + i);
    scratch[i] = a * scale;
  }
}

// bufs[k][i] = scratch[i - lo] for every rank k, i in [lo, hi)
__global__ void __launch_bounds__(256)
k_broadcast_slice(float* const* __restrict__ bufs, int world, int64_t lo, int64_t hi, const float* __restrict__ scratch) {
  pdl_prologue();
  const int64_t n = hi - lo;
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(bufs[0] + lo) & 15) == 0) ? n / 4 : 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(scratch)[i];
    for (int k = 0; k < world; ++k) reinterpret_cast<float4*>(bufs[k] + lo)[i] = v;
  }
  for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = scratch[i];
    for (int k = 0; k < world; ++k) bufs[k][lo + i] = v;
  }
}

// NVLS form of the two middle phases: one multimem.ld_reduce per 16 bytes fetches the sum over all ranks from the switch,
// one multimem.st stores it (scaled) into every rank's buffer.  Rank r is the only one that reads or writes slice r, so no
// barrier is needed between the two.  `mc` is the multicast address of the symmetric buffer.
__global__ void __launch_bounds__(256) k_nvls_reduce_bcast_slice(float* __restrict__ mc, int64_t lo, int64_t hi, float scale) {
  pdl_prologue();
  const int64_t n4 = (hi - lo) / 4;  // the caller guarantees 16-byte aligned slices of whole float4s
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float* p = mc + lo + 4 * i;
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
}


// ---- one-kernel all-reduce ---------------------------------------------------------------------------------------------
// The two-shot exchange above costs five launches and three cross-GPU barriers AFTER the backward (measured 68-97 us exposed per
// step at 2-8 GPUs).  This form is one launch and two flag exchanges, out of place:
//
//   arrive : CTA 0 publishes seq to every rank's arrive word [rank]; every CTA polls the LOCAL arrive words until all ranks have
//            arrived (a rank's arrival = its backward has completed and is visible: the store is a system-scope release
//            behind the kernel boundary)
//   reduce + scatter, fused: rank r sums slice r over all ranks' INPUT buffers (peer loads, fixed rank order -> bit-identical
//            results everywhere) and stores the scaled sum into slice r of every rank's OUTPUT buffer (peer stores).  Nobody reads
//            an output buffer during the exchange, so no barrier separates the loads from the stores.
//   done   : the last CTA of the grid to finish (device counter) publishes seq to every rank's done word [rank]; CTA 0 polls the
//            local done words, so kernel completion implies every rank's slice has landed in the local output buffer.
//
// With `mc_in` / `mc_out` (multicast mappings of the same buffers) the middle phase is one multimem.ld_reduce + one multimem.st
// per 16 bytes: the NVSwitch sums and broadcasts.  Hazards across steps: a rank reads peers' input buffers only before its own
// `done`, and every rank waits for all `done`s before its kernel completes, so the next backward may overwrite the input buffer;
// a rank writes peers' output buffers only after their next `arrive`, which they send after everything that read the previous
// result (stream order).  No CTA waits for another CTA of its own grid, so the grid need not be co-resident.
constexpr int kArriveWord = 64, kDoneWord = 96;   // flag words [64, 96) and [96, 128) of every rank's flag array

__device__ __forceinline__ void spin_until(const uint32_t* p, uint32_t seq) {
  const uint64_t t0 = gtimer_ns();
  for (uint32_t it = 0;; ++it) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if ((int32_t)(v - seq) >= 0) return;
    if ((it & 255u) == 255u && gtimer_ns() - t0 > 60000000000ull) asm volatile("trap;");
  }
}

__global__ void __launch_bounds__(512)
k_allreduce_fused(float* const* __restrict__ in, float* const* __restrict__ out, uint32_t* const* __restrict__ flags, float* mc_in,
                  float* mc_out, unsigned int* counter, int rank, int world, int64_t n, int64_t chunk, float scale, uint32_t seq) {
  pdl_prologue();
  const int t = threadIdx.x;
  if (blockIdx.x == 0 && t < world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags[t] + kArriveWord + rank), "r"(seq) : "memory");
  }
  if (t < world) spin_until(flags[rank] + kArriveWord + t, seq);
  __syncthreads();
  const int64_t lo = chunk * rank < n ? chunk * rank : n;
  const int64_t hi = lo + chunk < n ? lo + chunk : n;
  const int64_t cnt = hi - lo;
  const bool al = (reinterpret_cast<uintptr_t>(in[rank] + lo) & 15) == 0 && (reinterpret_cast<uintptr_t>(out[rank] + lo) & 15) == 0;
  const int64_t n4 = al ? cnt / 4 : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (mc_in && mc_out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + t; i < n4; i += stride) {
      float4 v;
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "l"(mc_in + lo + 4 * i)
                   : "memory");
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_out + lo + 4 * i), "f"(v.x), "f"(v.y),
                   "f"(v.z), "f"(v.w)
                   : "memory");
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + t; i < n4; i += stride) {
      float4 a = ld_peer4(in[0] + lo + 4 * i);
      for (int k = 1; k < world; ++k) {
        const float4 b = ld_peer4(in[k] + lo + 4 * i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
      for (int k = 0; k < world; ++k) reinterpret_cast<float4*>(out[k] + lo)[i] = a;
    }
  }
  for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + t; i < cnt; i += stride) {  // unaligned slices / the ragged tail
    float a = __ldcv(in[0] + lo + i);
    for (int k = 1; k < world; ++k) a += __ldcv(in[k] + lo + i);
    a *= scale;
    for (int k = 0; k < world; ++k) out[k][lo + i] = a;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned int s_last;
  if (t == 0) {
    const unsigned int prev = atomicAdd(counter, 1u);
    s_last = (prev == gridDim.x - 1) ? 1u : 0u;
    if (s_last) *counter = 0u;   // every CTA has passed its increment: ready for the next call
  }
  __syncthreads();
  if (s_last && t < world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags[t] + kDoneWord + rank), "r"(seq) : "memory");
  }
  if (blockIdx.x == 0 && t < world) spin_until(flags[rank] + kDoneWord + t, seq);
}

}  // namespace

extern "C" {

// EXPERIMENTAL (not yet measured on hardware; nothing calls it by default): the all-reduce above with the switch doing the
// reduction.  n must be a multiple of 4 * world and the buffer 16-byte aligned; `mc_buf` is the multicast mapping of the
// same symmetric buffer, `flags` as for pgnn_allreduce_p2p (two barriers per call: pass epoch counting calls of THIS function
// and do not mix the two functions on one flag array).
int pgnn_allreduce_nvls(float* mc_buf, void* const* flags, int rank, int world, int64_t n, float scale, int64_t epoch, void* stream) {
  PGNN_CHECK_ARG(mc_buf && flags && world >= 1 && world <= 32 && rank >= 0 && rank < world && n >= 0 && epoch >= 0);
  PGNN_CHECK_ARG(n % (4 * (int64_t)world) == 0 && (reinterpret_cast<uintptr_t>(mc_buf) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  uint32_t* const* f = reinterpret_cast<uint32_t* const*>(flags);
  const uint32_t seq = (uint32_t)(2 * epoch);
  const int64_t chunk = n / world, lo = chunk * rank, hi = lo + chunk;
  const int64_t work = ceil_div(chunk, 4 * 256);
  const unsigned blocks = (unsigned)(work < 1 ? 1 : work > 2 * kNumSMs ? 2 * kNumSMs : work);
  PGNN_CUDA(pgnn_launch(k_xgpu_barrier, dim3(1), dim3(32), 0, st, f, rank, world, seq + 1));
  PGNN_LAUNCH_CHECK();
  PGNN_CUDA(pgnn_launch(k_nvls_reduce_bcast_slice, dim3(blocks), dim3(256), 0, st, mc_buf, lo, hi, scale));
  PGNN_LAUNCH_CHECK();
  PGNN_CUDA(pgnn_launch(k_xgpu_barrier, dim3(1), dim3(32), 0, st, f, rank, world, seq + 2));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

// One-kernel, out-of-place all-reduce (k_allreduce_fused): out[k][i] = scale * sum_r in[r][i] on every rank k.
//   in, out : DEVICE arrays of `world` peer-mapped pointers (input / output buffer of every rank, n floats each, distinct)
//   flags   : DEVICE array of `world` pointers to every rank's flag words (uint32[128], zero before the first call; this entry
//             uses words [64, 128) and may share the array with pgnn_allreduce_p2p / _nvls, which use [0, 64))
//   mc_in, mc_out : multicast mappings of the input / output buffers (both non-null -> the NVSwitch reduces and broadcasts;
//             then n % (4 * world) == 0 and 16-byte aligned buffers are required), or null
//   counter : LOCAL device uint32, zero before the first call
//   epoch   : 0, 1, 2, ... identical on all ranks
int pgnn_allreduce_fused(void* const* in, void* const* out, void* const* flags, float* mc_in, float* mc_out, unsigned int* counter,
                         int rank, int world, int64_t n, float scale, int64_t epoch, void* stream) {
  PGNN_CHECK_ARG(in && out && flags && counter && world >= 1 && world <= 32 && rank >= 0 && rank < world && n >= 0 && epoch >= 0);
  const bool nvls = mc_in && mc_out;
  if (nvls) PGNN_CHECK_ARG(n % (4 * (int64_t)world) == 0 && (reinterpret_cast<uintptr_t>(mc_in) & 15) == 0 && (reinterpret_cast<uintptr_t>(mc_out) & 15) == 0);
  const int64_t chunk = align_up(ceil_div(n, world), 4);
  const int64_t work = ceil_div(chunk, 4 * 512);
  // enough CTAs to keep the NVLink loads in flight, few enough that the kernel does not queue behind its own waves
  const unsigned blocks = (unsigned)(work < 1 ? 1 : work > 96 ? 96 : work);
  PGNN_CUDA(pgnn_launch(k_allreduce_fused, dim3(blocks), dim3(512), 0, as_stream(stream), reinterpret_cast<float* const*>(in),
                        reinterpret_cast<float* const*>(out), reinterpret_cast<uint32_t* const*>(flags), nvls ? mc_in : nullptr,
                        nvls ? mc_out : nullptr, counter, rank, world, n, chunk, scale, (uint32_t)(epoch + 1)));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

int64_t pgnn_allreduce_p2p_scratch_floats(int64_t n, int world) {
  if (n < 0 || world < 1) return PGNN_EINVAL;
  return align_up(ceil_div(n, world), 4);
}

int pgnn_allreduce_p2p(void* const* bufs, void* const* flags, int rank, int world, int64_t n, float scale, float* scratch,
                       int64_t scratch_floats, int64_t epoch, void* stream) {
  PGNN_CHECK_ARG(bufs && flags && world >= 1 && world <= 32 && rank >= 0 && rank < world && n >= 0 && epoch >= 0);
  if (world == 1 && scale == 1.f) return PGNN_OK;
  const int64_t chunk = pgnn_allreduce_p2p_scratch_floats(n, world);  // multiple of 4: slice starts keep the buffer's alignment
  PGNN_CHECK_ARG(scratch || n == 0);
  if (scratch_floats < chunk) return PGNN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  float* const* b = reinterpret_cast<float* const*>(bufs);
  uint32_t* const* f = reinterpret_cast<uint32_t* const*>(flags);
  const uint32_t seq = (uint32_t)(3 * epoch);  // three barriers per call; the caller passes epoch = 0, 1, 2, ...
  const int64_t lo = chunk * rank < n ? chunk * rank : n;
  const int64_t hi = lo + chunk < n ? lo + chunk : n;
  const int64_t work = ceil_div(hi - lo, 4 * 256);
  const unsigned blocks = (unsigned)(work < 1 ? 1 : work > 2 * kNumSMs ? 2 * kNumSMs : work);
  PGNN_CUDA(pgnn_launch(k_xgpu_barrier, dim3(1), dim3(32), 0, st, f, rank, world, seq + 1));
  PGNN_LAUNCH_CHECK();
  PGNN_CUDA(pgnn_launch(k_reduce_slice, dim3(blocks), dim3(256), 0, st, b, world, lo, hi, scale, scratch));
  PGNN_LAUNCH_CHECK();
  PGNN_CUDA(pgnn_launch(k_xgpu_barrier, dim3(1), dim3(32), 0, st, f, rank, world, seq + 2));
  PGNN_LAUNCH_CHECK();
  PGNN_CUDA(pgnn_launch(k_broadcast_slice, dim3(blocks), dim3(256), 0, st, b, world, lo, hi, (const float*)scratch));
  PGNN_LAUNCH_CHECK();
  PGNN_CUDA(pgnn_launch(k_xgpu_barrier, dim3(1), dim3(32), 0, st, f, rank, world, seq + 3));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

}  // extern "C"
