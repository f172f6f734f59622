// Library-level entry points: version, error strings, device query.
#include "common.cuh"

thread_local int g_pgnn_last_cuda_error = 0;
std::atomic<long long> g_pgnn_kernel_launches{0};

extern "C" {

int pgnn_version(void) { return 100; }

const char* pgnn_error_string(int code) {
  switch (code) {
    case PGNN_OK: return "ok";
    case PGNN_EINVAL: return "invalid argument";
    case PGNN_ECUDA: return "CUDA error (see pgnn_last_cuda_error)";
    case PGNN_EWORKSPACE: return "workspace too small";
    case PGNN_EUNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

int pgnn_last_cuda_error(void) { return g_pgnn_last_cuda_error; }

int64_t pgnn_kernel_launch_count(void) { return (int64_t)g_pgnn_kernel_launches.load(); }

int pgnn_device_sm_count(int device) {
  int n = 0;
  cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) {
    g_pgnn_last_cuda_error = (int)e;
    return PGNN_ECUDA;
  }
  return n;
}

}  // extern "C"
