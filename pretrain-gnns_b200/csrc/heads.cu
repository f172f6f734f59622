// Graph-level and self-supervised heads: global_mean_pool (chem/model.py:326,369; bio/model.py:342),
// row gathers for masked atoms / bonds / centre nodes (chem/pretrain_masking.py:51,58-59;
// chem/pretrain_contextpred.py:54,57; bio/model.py:343) and the cyclic-shift negative-sampling dot
// products (chem/pretrain_contextpred.py:36-39,64-67).
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// One CTA per (segment, chunk of 32 float4 columns): 8 row-lanes (warps) walk the segment's rows eight apart, four independent
// row loads in flight each, and the eight partial sums are folded through shared memory in a fixed order (deterministic).  The
// first version walked a segment's rows serially in one thread per float4 column: a chain of dependent-latency loads that took
// 258 us for the 64 PPI ego graphs of ~500 nodes of the bio supervised step (a 38 MB read).
__global__ void __launch_bounds__(256)
k_segment_mean_fwd(const float* __restrict__ x, int64_t ldx, const int* __restrict__ seg_ptr, const int* __restrict__ seg_order,
                   int64_t num_seg, int C4, float* __restrict__ out, int64_t ldo) {
  pdl_prologue();
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  const int c4 = blockIdx.y * 32 + lane;
  const int lo = seg_ptr[b], hi = seg_ptr[b + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < C4) {
    const float* xc = x + 4 * c4;
    int k = lo + w;
    for (; k + 24 < hi; k += 32) {
      const float4 v0 = ld4(xc + (int64_t)seg_order[k] * ldx), v1 = ld4(xc + (int64_t)seg_order[k + 8] * ldx);
      const float4 v2 = ld4(xc + (int64_t)seg_order[k + 16] * ldx), v3 = ld4(xc + (int64_t)seg_order[k + 24] * ldx);
      acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
      acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
      acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
      acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
    }
    for (; k < hi; k += 8) {
      const float4 v = ld4(xc + (int64_t)seg_order[k] * ldx);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[w][lane] = acc;
  __syncthreads();
  if (w == 0 && c4 < C4) {
    float4 t = red[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = red[k][lane];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    const float cnt = (float)max(hi - lo, 1);  // count.clamp(min=1)
    st4(out + b * ldo + 4 * c4, make_float4(t.x / cnt, t.y / cnt, t.z / cnt, t.w / cnt));
  }
}

__global__ void __launch_bounds__(256)
k_segment_mean_bwd(const float* __restrict__ g, int64_t ldg, const int64_t* __restrict__ seg, const int* __restrict__ seg_ptr,
                   int64_t n, int C4, float* __restrict__ gx, int64_t ldgx) {
  pdl_prologue();
  const int64_t total = n * C4;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / C4;
    const int c = (int)(idx - r * C4) * 4;
    const int64_t b = seg[r];
    const float cnt = (float)max(seg_ptr[b + 1] - seg_ptr[b], 1);
    const float4 v = ld4(g + b * ldg + c);
    st4(gx + r * ldgx + c, make_float4(v.x / cnt, v.y / cnt, v.z / cnt, v.w / cnt));
  }
}

__global__ void __launch_bounds__(256)
k_row_gather_fwd(const float* __restrict__ x, int64_t ldx, int64_t rows, const int64_t* __restrict__ idx1, const int64_t* __restrict__ idx2,
                 int64_t m, int C4, float* __restrict__ out, int64_t ldo, unsigned int* __restrict__ err) {
  pdl_prologue();
  const int64_t total = m * C4;
  bool bad = false;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / C4;
    const int c = (int)(idx - r * C4) * 4;
    const int64_t i1 = idx1[r], i2 = idx2 ? idx2[r] : 0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // an out-of-range index (torch would assert) contributes zeros and is flagged
    if (i1 >= 0 && i1 < rows) v = ld4(x + i1 * ldx + c); else bad = true;
    if (idx2) {
      if (i2 >= 0 && i2 < rows) {
        const float4 u = ld4(x + i2 * ldx + c);
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
      } else {
        bad = true;
      }
    }
    st4(out + r * ldo + c, v);
  }
  if (bad && err) atomicOr(err, (unsigned)PGNN_DEVERR_GATHER);
}

// index_put_(accumulate=True): duplicates are legal (two masked bonds may share an atom), so rows are
// accumulated with vector atomics (red.global.add.v4.f32).
__global__ void __launch_bounds__(256)
k_row_gather_bwd(const float* __restrict__ g, int64_t ldg, const int64_t* __restrict__ idx1, const int64_t* __restrict__ idx2,
                 int64_t m, int C4, float* __restrict__ gx, int64_t ldgx, int64_t rows) {
  pdl_prologue();
  const int64_t total = m * C4;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / C4;
    const int c = (int)(idx - r * C4) * 4;
    const float4 v = ld4(g + r * ldg + c);
    const int64_t i1 = idx1[r];
    if (i1 >= 0 && i1 < rows) atomicAdd(reinterpret_cast<float4*>(gx + i1 * ldgx + c), v);
    if (idx2) {
      const int64_t i2 = idx2[r];
      if (i2 >= 0 && i2 < rows) atomicAdd(reinterpret_cast<float4*>(gx + i2 * ldgx + c), v);
    }
  }
}

// one warp per row
__global__ void __launch_bounds__(256)
k_shifted_rowdot_fwd(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb, int64_t B, int C,
                     int64_t shift, float* __restrict__ out) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < B; r += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t rb = (r + shift) % B;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(a[r * lda + c], b[rb * ldb + c], s);
    s = warp_sum(s);
    if (lane == 0) out[r] = s;
  }
}

__global__ void __launch_bounds__(256)
k_shifted_rowdot_bwd(const float* __restrict__ g, const float* __restrict__ a, int64_t lda, const float* __restrict__ b,
                     int64_t ldb, int64_t B, int C, int64_t shift, int accumulate, float* __restrict__ ga, int64_t ldga,
                     float* __restrict__ gb, int64_t ldgb) {
  pdl_prologue();
  const int64_t total = B * C;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / C;
    const int c = (int)(idx - r * C);
    const int64_t rf = (r + shift) % B;            // partner row of a[r]
    const int64_t rbk = ((r - shift) % B + B) % B;  // row of a whose partner is b[r]
    const float va = g[r] * b[rf * ldb + c];
    const float vb = g[rbk] * a[rbk * lda + c];
    if (accumulate) {
      ga[r * ldga + c] += va;
      gb[r * ldgb + c] += vb;
    } else {
      ga[r * ldga + c] = va;
      gb[r * ldgb + c] = vb;
    }
  }
}

// Mean cross-entropy over fp32 logits evaluated in fp64, as `criterion(pred_node.double(), labels)` does
// (chem/pretrain_masking.py:26,52).  One warp per row: max, log-sum-exp, -log p[label]; the same pass writes
// dlogits = (softmax - onehot) / M, so the backward of the loss needs no kernel of its own.
__global__ void __launch_bounds__(256)
k_softmax_ce(const float* __restrict__ logits, int64_t ld, int64_t M, int V, const int64_t* __restrict__ labels,
             double* __restrict__ loss_mean, float* __restrict__ dlogits, int64_t lddl, unsigned int* __restrict__ err) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < M; r += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const float* row = logits + r * ld;
    double mx = -1e300;
    for (int v = lane; v < V; v += 32) mx = fmax(mx, (double)row[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double se = 0.0;
    for (int v = lane; v < V; v += 32) se += exp((double)row[v] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const double lse = mx + log(se);
    int64_t y = labels[r];
    if (y < 0 || y >= V) {  // nll_loss would assert: flag, and let the row contribute log-sum-exp only
      if (lane == 0 && err) atomicOr(err, (unsigned)PGNN_DEVERR_LABEL);
      y = -1;
    }
    const double inv_m = 1.0 / (double)M;
    for (int v = lane; v < V; v += 32) {
      const double p = exp((double)row[v] - lse);
      dlogits[r * lddl + v] = (float)((p - (v == y ? 1.0 : 0.0)) * inv_m);
    }
    for (int v = V + lane; v < lddl; v += 32) dlogits[r * lddl + v] = 0.f;  // padding columns of the 16-byte-aligned row
    if (lane == 0) atomicAdd(loss_mean, (lse - (y >= 0 ? (double)row[y] : 0.0)) * inv_m);
  }
}

inline int grid_items(int64_t items, int threads) {
  int64_t b = ceil_div(items, threads);
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

int pgnn_segment_mean_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, const int32_t* seg_order, int64_t num_seg,
                          int64_t C, float* out, int64_t ldo, void* stream) {
  PGNN_CHECK_ARG(num_seg >= 0 && C > 0);
  if (num_seg == 0) return PGNN_OK;
  PGNN_CHECK_ARG(seg_ptr && out);
  if (C % 4 || ldx % 4 || ldo % 4 || !aligned16(x) || !aligned16(out)) return PGNN_EUNSUPPORTED;
  const int C4 = (int)(C / 4);
  PGNN_CUDA(pgnn_launch(k_segment_mean_fwd, dim3((unsigned)num_seg, (unsigned)ceil_div(C4, 32)), dim3(256), 0, as_stream(stream), x, ldx, seg_ptr, seg_order,
                        num_seg, C4, out, ldo));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

int pgnn_segment_mean_bwd(const float* g, int64_t ldg, const int64_t* seg, const int32_t* seg_ptr, int64_t num_rows, int64_t C,
                          float* gx, int64_t ldgx, void* stream) {
  PGNN_CHECK_ARG(num_rows >= 0 && C > 0);
  if (num_rows == 0) return PGNN_OK;
  PGNN_CHECK_ARG(g && seg && seg_ptr && gx);
  if (C % 4 || ldg % 4 || ldgx % 4 || !aligned16(g) || !aligned16(gx)) return PGNN_EUNSUPPORTED;
  const int C4 = (int)(C / 4);
  PGNN_CUDA(pgnn_launch(k_segment_mean_bwd, dim3(grid_items(num_rows * C4, 256)), dim3(256), 0, as_stream(stream), g, ldg, seg, seg_ptr, num_rows, C4, gx, ldgx));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

int pgnn_row_gather_fwd(const float* x, int64_t ldx, int64_t num_rows, const int64_t* idx, const int64_t* idx2, int64_t num_idx, int64_t C,
                        float* out, int64_t ldo, void* stream) {
  PGNN_CHECK_ARG(num_idx >= 0 && C > 0 && num_rows >= 0);
  if (num_idx == 0) return PGNN_OK;
  PGNN_CHECK_ARG(x && idx && out);
  if (C % 4 || ldx % 4 || ldo % 4 || !aligned16(x) || !aligned16(out)) return PGNN_EUNSUPPORTED;
  const int C4 = (int)(C / 4);
  PGNN_CUDA(pgnn_launch(k_row_gather_fwd, dim3(grid_items(num_idx * C4, 256)), dim3(256), 0, as_stream(stream), x, ldx, num_rows, idx, idx2, num_idx, C4, out, ldo,
                        pgnn_error_flag_ptr()));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

int pgnn_row_gather_bwd(const float* g, int64_t ldg, const int64_t* idx, const int64_t* idx2, int64_t num_idx, int64_t C,
                        float* gx, int64_t ldgx, int64_t num_rows, void* stream) {
  PGNN_CHECK_ARG(num_idx >= 0 && C > 0);
  if (num_idx == 0) return PGNN_OK;
  PGNN_CHECK_ARG(g && idx && gx);
  if (C % 4 || ldg % 4 || ldgx % 4 || !aligned16(g) || !aligned16(gx)) return PGNN_EUNSUPPORTED;
  const int C4 = (int)(C / 4);
  PGNN_CUDA(pgnn_launch(k_row_gather_bwd, dim3(grid_items(num_idx * C4, 256)), dim3(256), 0, as_stream(stream), g, ldg, idx, idx2, num_idx, C4, gx, ldgx, num_rows));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

int pgnn_softmax_ce_fwd(const float* logits, int64_t ld, int64_t M, int64_t V, const int64_t* labels, double* loss_mean,
                        float* dlogits, int64_t lddl, void* stream) {
  PGNN_CHECK_ARG(M >= 0 && V > 0 && loss_mean && lddl >= V && ld >= V);
  cudaStream_t st = as_stream(stream);
  PGNN_CUDA(cudaMemsetAsync(loss_mean, 0, sizeof(double), st));
  if (M == 0) return PGNN_OK;
  PGNN_CHECK_ARG(logits && labels && dlogits);
  PGNN_CUDA(pgnn_launch(k_softmax_ce, dim3(grid_items(M * 32, 256)), dim3(256), 0, st, logits, ld, M, (int)V, labels, loss_mean, dlogits, lddl, pgnn_error_flag_ptr()));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

int pgnn_shifted_rowdot_fwd(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t B, int64_t C, int64_t shift,
                            float* out, void* stream) {
  PGNN_CHECK_ARG(B >= 0 && C > 0 && shift >= 0);
  if (B == 0) return PGNN_OK;
  PGNN_CHECK_ARG(a && b && out);
  PGNN_CUDA(pgnn_launch(k_shifted_rowdot_fwd, dim3(grid_items(B * 32, 256)), dim3(256), 0, as_stream(stream), a, lda, b, ldb, B, (int)C, shift, out));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

int pgnn_shifted_rowdot_bwd(const float* g, const float* a, int64_t lda, const float* b, int64_t ldb, int64_t B, int64_t C,
                            int64_t shift, int accumulate, float* ga, int64_t ldga, float* gb, int64_t ldgb, void* stream) {
  PGNN_CHECK_ARG(B >= 0 && C > 0 && shift >= 0);
  if (B == 0) return PGNN_OK;
  PGNN_CHECK_ARG(g && a && b && ga && gb);
  PGNN_CUDA(pgnn_launch(k_shifted_rowdot_bwd, dim3(grid_items(B * C, 256)), dim3(256), 0, as_stream(stream), g, a, lda, b, ldb, B, (int)C, shift, accumulate, ga,
                                                                             ldga, gb, ldgb));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

}  // extern "C"
