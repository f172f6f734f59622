// Graph preparation: stable bucketing of the COO edge list by target / by source, GCN degree
// normalisation and the per-node edge-feature summaries.  Integer results are bit-exact against
// oracle/graph_prep_oracle.py (stable argsort).
//
// Replaces the per-edge addressing that torch_geometric 1.0.3's MessagePassing.propagate performs for
// every layer (reference call sites chem/model.py:49,101,148,196; bio/model.py:52,111,163,218).
#include "common.cuh"

namespace {

__global__ void k_histogram(const int64_t* __restrict__ keys, int64_t stride, int64_t n, int* __restrict__ counts) {
  pdl_prologue();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&counts[keys[i * stride]], 1);
}

// Exclusive scan of counts[0..n) into rowptr[0..n] (rowptr[n] = total) and cursor (a copy of rowptr).
// Single CTA of 1024 threads walking the array in 1024-element tiles with a running carry: the arrays
// here are node / graph / vocabulary counts (<= a few 10^5), so one CTA is latency- not throughput-bound.
__global__ void __launch_bounds__(1024) k_exclusive_scan(const int* counts, int64_t n,
                                                         int* rowptr, int* cursor) {
  pdl_prologue();
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += 1024) {
    int64_t i = base + threadIdx.x;
    int v = (i < n) ? counts[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + (wid ? warp_tot[wid - 1] : 0) + inc - v;
    if (i < n) {
      rowptr[i] = excl;
      cursor[i] = excl;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) rowptr[n] = carry_s;
}

__global__ void k_place(const int64_t* __restrict__ keys, int64_t stride, int64_t n, int* __restrict__ cursor,
                        int* __restrict__ tmp) {
  pdl_prologue();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int pos = atomicAdd(&cursor[keys[i * stride]], 1);
    tmp[pos] = (int)i;
  }
}

// The atomic placement leaves every bucket holding the right SET in arbitrary order; rank each element
// among its bucket-mates by original position to obtain the stable order (rank by counting: bucket sizes
// are in-degrees / graph sizes / vocabulary hits, so the quadratic term stays tiny).
__global__ void k_rank_in_bucket(const int64_t* __restrict__ keys, int64_t stride, int64_t n,
                                 const int* __restrict__ rowptr, const int* __restrict__ tmp,
                                 const int64_t* __restrict__ vals, int64_t val_stride, int* __restrict__ order,
                                 int* __restrict__ vals_out) {
  pdl_prologue();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int me = tmp[p];
    const int64_t b = keys[(int64_t)me * stride];
    const int lo = rowptr[b], hi = rowptr[b + 1];
    int rank = 0;
    for (int q = lo; q < hi; ++q) rank += (tmp[q] < me);
    order[lo + rank] = me;
    if (vals_out) vals_out[lo + rank] = (int)vals[(int64_t)me * val_stride];
  }
}

__global__ void k_gcn_dinv(const int* __restrict__ rowptr, int64_t n, float* __restrict__ dinv) {
  pdl_prologue();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // deg.pow(-0.5) over the loop-augmented edges: in-degree + 1 >= 1, never inf (chem/model.py:78-80)
    float deg = (float)(rowptr[i + 1] - rowptr[i] + 1);
    dinv[i] = __frcp_rn(__fsqrt_rn(deg));
  }
}

__global__ void k_chem_edge_summary(const int64_t* __restrict__ edge_attr, const int* __restrict__ rowptr,
                                    const int* __restrict__ nbr, const int* __restrict__ eid, int64_t n, int mode,
                                    const float* __restrict__ dinv, float* __restrict__ S) {
  pdl_prologue();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) s[q] = 0.f;
    const int lo = rowptr[i], hi = rowptr[i + 1];
    for (int k = lo; k < hi; ++k) {
      const int64_t e = eid[k];
      const int a0 = (int)edge_attr[2 * e], a1 = (int)edge_attr[2 * e + 1];
      const float w = agg_weight(mode, dinv, (int)i, nbr[k], hi - lo);
#pragma unroll
      for (int q = 0; q < 6; ++q) s[q] += (a0 == q) ? w : 0.f;
#pragma unroll
      for (int q = 0; q < 3; ++q) s[6 + q] += (a1 == q) ? w : 0.f;
    }
    const float wl = agg_weight(mode, dinv, (int)i, (int)i, hi - lo);
    s[4] += wl;  // self-loop: bond type 4, direction 0 (chem/model.py:42-45)
    s[6] += wl;
#pragma unroll
    for (int q = 0; q < 9; ++q) S[i * 9 + q] = s[q];
  }
}

__global__ void k_bio_edge_summary(const float* __restrict__ edge_attr, const int* __restrict__ rowptr,
                                   const int* __restrict__ nbr, const int* __restrict__ eid, int64_t n, int mode,
                                   const float* __restrict__ dinv, float* __restrict__ S) {
  pdl_prologue();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) s[q] = 0.f;
    const int lo = rowptr[i], hi = rowptr[i + 1];
    for (int k = lo; k < hi; ++k) {
      const float* a = edge_attr + (int64_t)eid[k] * 9;
      const float w = agg_weight(mode, dinv, (int)i, nbr[k], hi - lo);
#pragma unroll
      for (int q = 0; q < 9; ++q) s[q] += w * a[q];
      s[9] += w;
    }
    const float wl = agg_weight(mode, dinv, (int)i, (int)i, hi - lo);
    s[7] += wl;  // self-loop row is one-hot at column 7 (bio/model.py:42-43)
    s[9] += wl;
#pragma unroll
    for (int q = 0; q < 10; ++q) S[i * 10 + q] = s[q];
  }
}

inline int grid_for(int64_t n, int threads) {
  int64_t b = ceil_div(n, threads);
  int64_t cap = (int64_t)kNumSMs * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int64_t pgnn_bucket_workspace_bytes(int64_t num_keys, int64_t num_buckets) {
  if (num_keys < 0 || num_buckets < 0) return PGNN_EINVAL;
  return align_up((num_buckets + 1) * 4, 256) + align_up((num_keys > 0 ? num_keys : 1) * 4, 256);
}

int pgnn_bucket(const int64_t* keys, int64_t key_stride, int64_t num_keys, int64_t num_buckets, const int64_t* vals,
                int64_t val_stride, int32_t* rowptr, int32_t* order, int32_t* vals_out, void* workspace,
                int64_t workspace_bytes, void* stream) {
  PGNN_CHECK_ARG(num_keys >= 0 && num_buckets >= 0 && rowptr && workspace);
  PGNN_CHECK_ARG(num_keys == 0 || (keys && order));
  PGNN_CHECK_ARG((vals == nullptr) == (vals_out == nullptr) || num_keys == 0);
  PGNN_CHECK_ARG(num_keys < (int64_t)1 << 31 && num_buckets < (int64_t)1 << 31);
  if (workspace_bytes < pgnn_bucket_workspace_bytes(num_keys, num_buckets)) return PGNN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  int* counts = reinterpret_cast<int*>(workspace);  // becomes the placement cursor after the scan
  int* tmp = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + align_up((num_buckets + 1) * 4, 256));
  PGNN_CUDA(cudaMemsetAsync(counts, 0, (num_buckets + 1) * 4, st));
  if (num_keys > 0) {
    PGNN_CUDA(pgnn_launch(k_histogram, dim3(grid_for(num_keys, 256)), dim3(256), 0, st, keys, key_stride, num_keys, counts));
    PGNN_LAUNCH_CHECK();
  }
  PGNN_CUDA(pgnn_launch(k_exclusive_scan, dim3(1), dim3(1024), 0, st, counts, num_buckets, rowptr, counts));
  PGNN_LAUNCH_CHECK();
  if (num_keys > 0) {
    PGNN_CUDA(pgnn_launch(k_place, dim3(grid_for(num_keys, 256)), dim3(256), 0, st, keys, key_stride, num_keys, counts, tmp));
    PGNN_LAUNCH_CHECK();
    PGNN_CUDA(pgnn_launch(k_rank_in_bucket, dim3(grid_for(num_keys, 256)), dim3(256), 0, st, keys, key_stride, num_keys, rowptr, tmp, vals,
                                                              val_stride, order, vals_out));
    PGNN_LAUNCH_CHECK();
  }
  return PGNN_OK;
}

int64_t pgnn_graph_prep_workspace_bytes(int64_t num_nodes, int64_t num_edges) {
  return pgnn_bucket_workspace_bytes(num_edges, num_nodes);
}

int pgnn_graph_prep(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, int32_t* rowptr_t,
                    int32_t* nbr_t, int32_t* eid_t, int32_t* rowptr_s, int32_t* nbr_s, int32_t* eid_s, void* workspace,
                    int64_t workspace_bytes, void* stream) {
  PGNN_CHECK_ARG(num_edges >= 0 && num_nodes >= 0 && rowptr_t && rowptr_s);
  PGNN_CHECK_ARG(num_edges == 0 || (edge_index && nbr_t && eid_t && nbr_s && eid_s));
  const int64_t* tgt = edge_index;
  const int64_t* src = edge_index ? edge_index + num_edges : nullptr;
  int rc = pgnn_bucket(tgt, 1, num_edges, num_nodes, src, 1, rowptr_t, eid_t, nbr_t, workspace, workspace_bytes, stream);
  if (rc != PGNN_OK) return rc;
  return pgnn_bucket(src, 1, num_edges, num_nodes, tgt, 1, rowptr_s, eid_s, nbr_s, workspace, workspace_bytes, stream);
}

int pgnn_gcn_dinv(const int32_t* rowptr_t, int64_t num_nodes, float* dinv, void* stream) {
  PGNN_CHECK_ARG(num_nodes >= 0 && (num_nodes == 0 || (rowptr_t && dinv)));
  if (num_nodes == 0) return PGNN_OK;
  PGNN_CUDA(pgnn_launch(k_gcn_dinv, dim3(grid_for(num_nodes, 256)), dim3(256), 0, as_stream(stream), rowptr_t, num_nodes, dinv));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

int pgnn_chem_edge_summary(const int64_t* edge_attr, const int32_t* rowptr_t, const int32_t* nbr_t, const int32_t* eid_t,
                           int64_t num_nodes, int mode, const float* dinv, float* S, void* stream) {
  PGNN_CHECK_ARG(num_nodes >= 0 && mode >= 0 && mode <= 2 && (mode != PGNN_AGG_GCN || dinv));
  if (num_nodes == 0) return PGNN_OK;
  PGNN_CHECK_ARG(rowptr_t && S);
  PGNN_CUDA(pgnn_launch(k_chem_edge_summary, dim3(grid_for(num_nodes, 128)), dim3(128), 0, as_stream(stream), edge_attr, rowptr_t, nbr_t, eid_t, num_nodes,
                                                                               mode, dinv, S));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

int pgnn_bio_edge_summary(const float* edge_attr, const int32_t* rowptr_t, const int32_t* nbr_t, const int32_t* eid_t,
                          int64_t num_nodes, int mode, const float* dinv, float* S, void* stream) {
  PGNN_CHECK_ARG(num_nodes >= 0 && mode >= 0 && mode <= 2 && (mode != PGNN_AGG_GCN || dinv));
  if (num_nodes == 0) return PGNN_OK;
  PGNN_CHECK_ARG(rowptr_t && S);
  PGNN_CUDA(pgnn_launch(k_bio_edge_summary, dim3(grid_for(num_nodes, 128)), dim3(128), 0, as_stream(stream), edge_attr, rowptr_t, nbr_t, eid_t, num_nodes,
                                                                              mode, dinv, S));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

}  // extern "C"
