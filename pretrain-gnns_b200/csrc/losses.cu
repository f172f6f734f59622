// Binary cross-entropy with logits, evaluated in fp64 on fp32 logits, mean-reduced, with its gradient in the same pass:
//   bio/pretrain_supervised.py:22,36      criterion(pred.double(), y)                      y in {0,1}            (kind 1)
//   chem/pretrain_contextpred.py:40,86-87 criterion(pred_pos.double(), ones) / (pred_neg.double(), zeros)        (kind 0)
//   chem/finetune.py:25,33-43             BCE(reduction=none) on (y+1)/2, entries with y == 0 dropped,
//                                         loss = sum / number of valid entries               y in {-1,0,+1}      (kind 2)
// The reference casts the logits to double, so the loss value and d loss / d logits carry fp64 rounding only; the
// gradient is handed back as fp32 (what autograd's .double() backward produces).  The sum is deterministic: one fp64
// partial per CTA, folded in CTA order by the last CTA to finish (ticket), no floating-point atomics.
#include "common.cuh"

namespace {

constexpr int kBceThreads = 256;
constexpr int kBceMaxBlocks = kNumSMs * 4;

struct BceWs {
  double partial[kBceMaxBlocks];
  long long valid;      // kind 2: number of entries with y != 0
  unsigned int ticket;
  unsigned int pad;
};

__global__ void __launch_bounds__(kBceThreads)
k_bce_count_valid(const int64_t* __restrict__ target, int64_t ldt, int64_t M, int64_t N, BceWs* __restrict__ ws) {
  pdl_prologue();
  long long c = 0;
  const int64_t total = M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / N;
    c += target[r * ldt + (i - r * N)] != 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(reinterpret_cast<unsigned long long*>(&ws->valid), (unsigned long long)c);
}

__global__ void __launch_bounds__(kBceThreads)
k_bce_logits(const float* __restrict__ logits, int64_t ld, int64_t M, int64_t N, const int64_t* __restrict__ target, int64_t ldt,
             int kind, double tconst, BceWs* __restrict__ ws, double* __restrict__ loss, float* __restrict__ dlogits, int64_t lddl) {
  pdl_prologue();
  __shared__ double s_part[kBceThreads / 32];
  __shared__ bool s_last;
  const int64_t total = M * N;
  const double count = kind == 2 ? (double)ws->valid : (double)total;
  const double inv = count > 0.0 ? 1.0 / count : 0.0;
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / N, c = i - r * N;
    const double x = (double)logits[r * ld + c];
    double t = tconst;
    bool valid = true;
    if (kind == 1) {
      t = (double)target[r * ldt + c];
    } else if (kind == 2) {
      const int64_t y = target[r * ldt + c];
      valid = y != 0;
      t = 0.5 * ((double)y + 1.0);
    }
    const double e = exp(-fabs(x));
    const double sig = x >= 0.0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
    if (valid) acc += fmax(x, 0.0) - x * t + log1p(e);
    dlogits[r * lddl + c] = valid ? (float)((sig - t) * inv) : 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int w = 0; w < kBceThreads / 32; ++w) b += s_part[w];
    ws->partial[blockIdx.x] = b;
    __threadfence();
    s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += reinterpret_cast<volatile double*>(ws->partial)[b];
    *loss = s * inv;  // mean over the (valid) entries; 0 when there are none
  }
}

}  // namespace

extern "C" {

int64_t pgnn_bce_logits_workspace_bytes(void) { return (int64_t)sizeof(BceWs); }

int pgnn_bce_logits_fwd(const float* logits, int64_t ld, int64_t M, int64_t N, const int64_t* target, int64_t ldt, int target_kind,
                        double const_target, double* loss_mean, float* dlogits, int64_t lddl, void* workspace, int64_t workspace_bytes,
                        void* stream) {
  PGNN_CHECK_ARG(M >= 0 && N >= 0 && loss_mean && target_kind >= 0 && target_kind <= 2 && workspace);
  if (workspace_bytes < (int64_t)sizeof(BceWs)) return PGNN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (M * N == 0) {
    PGNN_CUDA(cudaMemsetAsync(loss_mean, 0, sizeof(double), st));
    return PGNN_OK;
  }
  PGNN_CHECK_ARG(logits && dlogits && ld >= N && lddl >= N && (target_kind == 0 || (target && ldt >= N)));
  BceWs* ws = reinterpret_cast<BceWs*>(workspace);
  PGNN_CUDA(cudaMemsetAsync(&ws->valid, 0, sizeof(long long) + 2 * sizeof(unsigned int), st));
  int64_t blocks = ceil_div(M * N, (int64_t)kBceThreads * 4);
  if (blocks > kBceMaxBlocks) blocks = kBceMaxBlocks;
  if (blocks < 1) blocks = 1;
  if (target_kind == 2) {
    PGNN_CUDA(pgnn_launch(k_bce_count_valid, dim3((unsigned)blocks), dim3(kBceThreads), 0, st, target, ldt, M, N, ws));
    PGNN_LAUNCH_CHECK();
  }
  PGNN_CUDA(pgnn_launch(k_bce_logits, dim3((unsigned)blocks), dim3(kBceThreads), 0, st, logits, ld, M, N, target, ldt, target_kind, const_target,
                        ws, loss_mean, dlogits, lddl));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

}  // extern "C"
