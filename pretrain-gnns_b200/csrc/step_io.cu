// The two neighbours of the hot path inside a training step:
//   * batch collation on the device: chem/batch.py:17-52 (BatchMasking.from_data_list) concatenates the per-graph
//     tensors in a Python loop and offsets edge_index by the running node count; here the molecule store lives in HBM
//     and one launch builds the batch tensors for a list of graph ids.
//   * the optimizer step: torch.optim.Adam(model.parameters(), lr, weight_decay) (chem/pretrain_masking.py:134-136,
//     72-74) walks 40-odd small tensors; here one launch updates every tensor from a chunk table.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// collation
// ---------------------------------------------------------------------------------------------------------------------

// exclusive scans of the node and edge counts of the selected graphs; one CTA, any B
__global__ void __launch_bounds__(1024)
k_collate_scan(const int64_t* __restrict__ node_ptr, const int64_t* __restrict__ edge_ptr, const int64_t* __restrict__ ids, int64_t B,
               int64_t* __restrict__ node_off, int64_t* __restrict__ edge_off) {
  pdl_prologue();
  __shared__ int64_t wn[32], we[32];
  __shared__ int64_t carry_n, carry_e;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_n = carry_e = 0;
  __syncthreads();
  for (int64_t base = 0; base < B; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    int64_t n = 0, e = 0;
    if (i < B) {
      const int64_t g = ids[i];
      n = node_ptr[g + 1] - node_ptr[g];
      e = edge_ptr[g + 1] - edge_ptr[g];
    }
    int64_t sn = n, se = e;  // inclusive warp scans
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int64_t tn = __shfl_up_sync(0xffffffffu, sn, d), te = __shfl_up_sync(0xffffffffu, se, d);
      if (lane >= d) { sn += tn; se += te; }
    }
    if (lane == 31) { wn[warp] = sn; we[warp] = se; }
    __syncthreads();
    if (warp == 0) {
      int64_t a = wn[lane], b = we[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int64_t ta = __shfl_up_sync(0xffffffffu, a, d), tb = __shfl_up_sync(0xffffffffu, b, d);
        if (lane >= d) { a += ta; b += tb; }
      }
      wn[lane] = a; we[lane] = b;
    }
    __syncthreads();
    const int64_t pn = carry_n + (warp ? wn[warp - 1] : 0), pe = carry_e + (warp ? we[warp - 1] : 0);
    if (i < B) { node_off[i] = pn + sn - n; edge_off[i] = pe + se - e; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) { carry_n = pn + sn; carry_e = pe + se; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { node_off[B] = carry_n; edge_off[B] = carry_e; }
}

// one warp per selected graph: widen the store's bytes to the int64 tensors GNN.forward takes, add the node offset
__global__ void __launch_bounds__(256)
k_collate_fill(const int64_t* __restrict__ node_ptr, const int64_t* __restrict__ edge_ptr, const uint8_t* __restrict__ sx,
               const int32_t* __restrict__ sei, int64_t store_edges, const uint8_t* __restrict__ sea, const int64_t* __restrict__ ids,
               int64_t B, const int64_t* __restrict__ node_off, const int64_t* __restrict__ edge_off, int64_t* __restrict__ x,
               int64_t* __restrict__ edge_index, int64_t* __restrict__ edge_attr, int64_t* __restrict__ batch) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t E = edge_off[B];
  for (int64_t i = wid; i < B; i += nw) {
    const int64_t g = ids[i];
    const int64_t n0 = node_ptr[g], n = node_ptr[g + 1] - n0, e0 = edge_ptr[g], e = edge_ptr[g + 1] - e0;
    const int64_t no = node_off[i], eo = edge_off[i];
    for (int64_t k = lane; k < 2 * n; k += 32) x[2 * no + k] = sx[2 * n0 + k];
    for (int64_t k = lane; k < n; k += 32) batch[no + k] = i;
    for (int64_t k = lane; k < e; k += 32) {
      edge_index[eo + k] = no + sei[e0 + k];
      edge_index[E + eo + k] = no + sei[store_edges + e0 + k];
    }
    for (int64_t k = lane; k < 2 * e; k += 32) edge_attr[2 * eo + k] = sea[2 * e0 + k];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Adam
// ---------------------------------------------------------------------------------------------------------------------
struct AdamConsts {
  float beta1, beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay, grad_scale;
  float step_size;  // torch: lr / bc1            | legacy (1.0.1): lr * sqrt(bc2) / bc1
  float sqrt_bc2;   // torch: sqrt(bc2)             | legacy: 1
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamConsts& c) {
  g *= c.grad_scale;
  g = fmaf(c.weight_decay, p, g);                    // grad.add(param, alpha=weight_decay)
  m = fmaf(g - m, c.one_minus_beta1, m);             // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(c.one_minus_beta2 * g, g, c.beta2 * v);   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / c.sqrt_bc2 + c.eps;
  p -= c.step_size * (m / denom);                    // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256)
k_adam(const PgnnAdamChunk* __restrict__ chunks, AdamConsts c) {
  pdl_prologue();
  const PgnnAdamChunk ch = chunks[blockIdx.x];
  float* __restrict__ p = ch.param;
  const float* __restrict__ g = ch.grad;
  float* __restrict__ m = ch.exp_avg;
  float* __restrict__ v = ch.exp_avg_sq;
  const int64_t n = ch.n;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  int64_t done = 0;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
      const float4 G = reinterpret_cast<const float4*>(g)[i];
      adam_one(P.x, G.x, M.x, V.x, c);
      adam_one(P.y, G.y, M.y, V.y, c);
      adam_one(P.z, G.z, M.z, V.z, c);
      adam_one(P.w, G.w, M.w, V.w, c);
      reinterpret_cast<float4*>(p)[i] = P;
      reinterpret_cast<float4*>(m)[i] = M;
      reinterpret_cast<float4*>(v)[i] = V;
    }
    done = n4 << 2;
  }
  for (int64_t i = done + threadIdx.x; i < n; i += blockDim.x) {
    float P = p[i], M = m[i], V = v[i];
    adam_one(P, g[i], M, V, c);
    p[i] = P; m[i] = M; v[i] = V;
  }
}

}  // namespace

extern "C" {

int pgnn_collate_chem(const int64_t* node_ptr, const int64_t* edge_ptr, const uint8_t* store_x, const int32_t* store_edge_index,
                      int64_t store_num_edges, const uint8_t* store_edge_attr, const int64_t* graph_ids, int64_t B,
                      int64_t* node_off, int64_t* edge_off, int64_t* x, int64_t* edge_index, int64_t* edge_attr, int64_t* batch,
                      void* stream) {
  PGNN_CHECK_ARG(B >= 0 && store_num_edges >= 0);
  PGNN_CHECK_ARG(node_ptr && edge_ptr && node_off && edge_off);
  if (B > 0) PGNN_CHECK_ARG(graph_ids && store_x && x && batch);
  cudaStream_t st = as_stream(stream);
  PGNN_CUDA(pgnn_launch(k_collate_scan, dim3(1), dim3(1024), 0, st, node_ptr, edge_ptr, graph_ids, B, node_off, edge_off));
  PGNN_LAUNCH_CHECK();
  if (B == 0) return PGNN_OK;
  const int64_t blocks = ceil_div(B, 8);
  PGNN_CUDA(pgnn_launch(k_collate_fill, dim3((unsigned)(blocks < 8 * kNumSMs ? blocks : 8 * kNumSMs)), dim3(256), 0, st, node_ptr, edge_ptr,
                        store_x, store_edge_index, store_num_edges, store_edge_attr, graph_ids, B, (const int64_t*)node_off,
                        (const int64_t*)edge_off, x, edge_index, edge_attr, batch));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

int pgnn_adam_step(const PgnnAdamChunk* chunks, int64_t num_chunks, double lr, double beta1, double beta2, double eps, double weight_decay,
                   double grad_scale, int64_t step, int legacy_eps, void* stream) {
  PGNN_CHECK_ARG(num_chunks >= 0 && step >= 1);
  if (num_chunks == 0) return PGNN_OK;
  PGNN_CHECK_ARG(chunks != nullptr);
  // bias corrections in double on the host, as torch does with the Python-float step count
  // (hyper-parameters arrive as doubles for the same reason: 1 - beta2 formed in fp32 is off by 1.3e-5 relative)
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  AdamConsts c;
  c.beta1 = (float)beta1; c.beta2 = (float)beta2;
  c.one_minus_beta1 = (float)(1.0 - beta1); c.one_minus_beta2 = (float)(1.0 - beta2);
  c.eps = (float)eps; c.weight_decay = (float)weight_decay; c.grad_scale = (float)grad_scale;
  if (legacy_eps) {  // torch 1.0.1 (requirements.txt:2): denom = sqrt(v) + eps; step = lr * sqrt(bc2) / bc1
    c.step_size = (float)(lr * sqrt(bc2) / bc1);
    c.sqrt_bc2 = 1.f;
  } else {           // torch >= 1.5: denom = sqrt(v) / sqrt(bc2) + eps; step = lr / bc1
    c.step_size = (float)(lr / bc1);
    c.sqrt_bc2 = (float)sqrt(bc2);
  }
  PGNN_CUDA(pgnn_launch(k_adam, dim3((unsigned)num_chunks), dim3(256), 0, as_stream(stream), chunks, c));
  PGNN_LAUNCH_CHECK();
  return PGNN_OK;
}

}  // extern "C"
