// Shared helpers for libpgnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pgnn_b200.h"

#include <atomic>
#include <utility>
extern thread_local int g_pgnn_last_cuda_error;
extern std::atomic<long long> g_pgnn_kernel_launches;  // every kernel this library enqueues (pgnn_kernel_launch_count)
// Per-kernel timing mode (pgnn_profile_enable): every launch is bracketed by a CUDA event pair on its own stream, so that
// a caller can read each kernel's duration inside the real step (warm L2, true operand residency) without a profiler.
extern std::atomic<int> g_pgnn_profile_on;
void pgnn_profile_mark(const void* kernel, cudaStream_t st, bool after);

// Device error word of the current device (lazily allocated, zeroed): kernels that consume indices OR a PGNN_DEVERR_* bit into
// it instead of reading or writing out of range; pgnn_device_error_flags() reads it back.  May return null (then unchecked).
unsigned int* pgnn_error_flag_ptr();

#define PGNN_CHECK_ARG(cond)            \
  do {                                  \
    if (!(cond)) return PGNN_EINVAL;    \
  } while (0)

#define PGNN_CUDA(call)                       \
  do {                                        \
    cudaError_t e__ = (call);                 \
    if (e__ != cudaSuccess) {                 \
      g_pgnn_last_cuda_error = (int)e__;      \
      return PGNN_ECUDA;                      \
    }                                         \
  } while (0)

#define PGNN_LAUNCH_CHECK()                                        \
  do {                                                             \
    g_pgnn_kernel_launches.fetch_add(1, std::memory_order_relaxed); \
    PGNN_CUDA(cudaGetLastError());                                 \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Optional column reductions fused into the tensor-core GEMM epilogue (the output tile is already staged in shared
// memory there).  All outputs must be zeroed by the caller; every CTA adds its tile's contribution with atomics.
struct PgnnGemmHooks {
  float* colsum = nullptr;   // [N]     += sum over rows of the (final) output              -> bias gradients
  double* stats = nullptr;   // [2][N]  += sum, sum of squares (fp64)                        -> BatchNorm batch statistics
  const float* S = nullptr;  // [M][Q]  per-row weights: gT[q][n] += sum_m S[m][q] out[m][n] -> bond-table gradients
  int Q = 0;
  float* gT = nullptr;       // rows [0, q_split)
  float* gT2 = nullptr;      // rows [q_split, Q)
  int q_split = 0;
  int64_t ldt = 0;
  bool any() const { return colsum || stats || S; }
};

// BatchNorm finalisation folded into the consumer kernel: from the fp64 column sums `acc` ([2][C]: sum, sum of squares
// over M rows) every CTA derives scale/shift itself; CTA 0 also performs the module-state updates of torch.nn.BatchNorm1d.
struct PgnnBnFold {
  const double* acc = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float* running_mean = nullptr;
  float* running_var = nullptr;
  int64_t* nbt = nullptr;
  float* save_mean = nullptr;
  float* save_invstd = nullptr;
  float momentum = 0.1f, eps = 1e-5f;
  int M = 0;
  double inv_m = 0.0, unbias = 1.0;  // 1 / M and M / (M - 1), formed on the host (set_rows)
  void set_rows(int rows) {
    M = rows;
    inv_m = rows > 0 ? 1.0 / (double)rows : 0.0;
    unbias = rows > 1 ? (double)rows / (double)(rows - 1) : 1.0;
  }
};

// per-column BatchNorm constants from the accumulated sums; `leader` performs the running-statistics side effects
// Every CTA of the consumer runs this for its columns, so it must be cheap: the fp64 part is two multiplies and one FMA (the
// sums are fp64 because E[x^2] - E[x]^2 cancels); the reciprocal square root is taken in fp32 like torch's own BatchNorm
// kernels do.  (With fp64 division and square root — software sequences of ~100 instructions each — this prologue was most
// of the 450 instructions per thread ncu counted in k_aggregate_fwd: profiles/r02_kernels_masking.md.)
__device__ __forceinline__ void bn_fold_column(const PgnnBnFold& f, int C, int c, bool leader, float& scale, float& shift) {
  const double s = f.acc[c], ss = f.acc[(int64_t)C + c];
  const double mean = s * f.inv_m;
  double var = fma(ss, f.inv_m, -mean * mean);
  var = var < 0.0 ? 0.0 : var;
  const float invstd = 1.0f / sqrtf((float)var + f.eps);
  const float meanf = (float)mean;
  scale = f.gamma[c] * invstd;
  shift = fmaf(-meanf, scale, f.beta[c]);
  if (leader) {
    if (f.save_mean) f.save_mean[c] = meanf;
    if (f.save_invstd) f.save_invstd[c] = invstd;
    if (f.running_mean) f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * meanf;
    if (f.running_var) {
      f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)(var * f.unbias);
    }
    if (c == 0 && f.nbt) *f.nbt += 1;
  }
}

// Epilogue of the tensor-core GEMM kernels (dense_tc.cu, dense_tma.cu)
struct TcEpilogue {
  const float* bias;      // [N] or null
  int relu;
  const float* mask_src;  // [M,N] (ld = ldm): zero where mask_src <= 0
  int64_t ldm;
  int atomic;             // split-K: accumulate with atomics into a zeroed output
  PgnnGemmHooks hooks;    // fused column reductions over the final output tile (not with split-K)
  int64_t split_stride = 0;  // split-K without atomics: split z stores its tile at C + z * split_stride (folded afterwards)
  // In-kernel fold of those partial tiles (TMA kernel only; requires every CTA of the grid to be co-resident): after storing
  // its partial tile a CTA bumps fold_counter[tile], waits until all `gridDim.z` splits of the tile have arrived, and sums a
  // 1/gridDim.z row slice of the tile over the splits IN SPLIT ORDER (bit-reproducible) into fold_out (leading dimension ldc).
  unsigned int* fold_counter = nullptr;  // [tiles], zero before the launch
  float* fold_out = nullptr;
};

// B200: 148 SMs.  Grids for grid-stride kernels are sized as a multiple of this.
constexpr int kNumSMs = 148;

// Programmatic dependent launch (PDL).  Every kernel of this library starts with pdl_prologue(): wait until the
// previous kernel in the stream has completed and flushed (griddepcontrol.wait), then allow the NEXT kernel's CTAs to
// become resident (griddepcontrol.launch_dependents) so that its launch latency and prologue overlap this kernel's
// execution; they park on their own wait.  All launches go through pgnn_launch(), which sets the
// programmatic-stream-serialization attribute.  A step is ~85 short kernels: the ~2-3 us of drain + launch between two
// dependent kernels was ~15% of the step.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_wait();
  pdl_trigger();
}

// same, for thread-block clusters of `cluster` CTAs along grid.x (kernels that use tcgen05 cta_group::2 are rejected with
// cudaErrorInvalidClusterSize unless the pair lies along x)
template <typename... KArgs, typename... Args>
inline cudaError_t pgnn_launch_cluster_x(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster,
                                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cluster;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  const bool prof = g_pgnn_profile_on.load(std::memory_order_relaxed) != 0;
  if (prof) pgnn_profile_mark(reinterpret_cast<const void*>(kernel), st, false);
  const cudaError_t err = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
  if (prof) pgnn_profile_mark(reinterpret_cast<const void*>(kernel), st, true);
  return err;
}

template <typename... KArgs, typename... Args>
inline cudaError_t pgnn_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const bool prof = g_pgnn_profile_on.load(std::memory_order_relaxed) != 0;
  if (prof) pgnn_profile_mark(reinterpret_cast<const void*>(kernel), st, false);
  const cudaError_t err = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
  if (prof) pgnn_profile_mark(reinterpret_cast<const void*>(kernel), st, true);
  return err;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// edge weight of the aggregation modes (PGNN_AGG_*), for target t with in-degree deg_t (real edges)
__device__ __forceinline__ float agg_weight(int mode, const float* __restrict__ dinv, int t, int s, int deg_t) {
  if (mode == PGNN_AGG_SUM) return 1.0f;
  if (mode == PGNN_AGG_MEAN) return __frcp_rn((float)(deg_t + 1));
  return __fmul_rn(dinv[t], dinv[s]);  // chem/model.py:82
}
