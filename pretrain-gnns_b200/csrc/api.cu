// Library-level entry points: version, error strings, device query, per-kernel timing mode.
#include "common.cuh"

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

thread_local int g_pgnn_last_cuda_error = 0;
std::atomic<long long> g_pgnn_kernel_launches{0};
std::atomic<int> g_pgnn_profile_on{0};

namespace {
struct ProfRec {
  const void* fn;
  cudaEvent_t a, b;
  bool closed;
};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
constexpr size_t kProfMax = 1 << 16;  // stop recording beyond this many launches (the mode is meant for a few steps)
}  // namespace

// before: open a record and stamp its first event; after: stamp the second event of the newest open record of this kernel
void pgnn_profile_mark(const void* kernel, cudaStream_t st, bool after) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (!after) {
    if (g_prof.size() >= kProfMax) return;
    ProfRec r{kernel, nullptr, nullptr, false};
    if (cudaEventCreate(&r.a) != cudaSuccess) return;
    if (cudaEventCreate(&r.b) != cudaSuccess) { cudaEventDestroy(r.a); return; }
    if (cudaEventRecord(r.a, st) != cudaSuccess) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); return; }
    g_prof.push_back(r);
  } else {
    for (size_t i = g_prof.size(); i-- > 0;) {
      if (g_prof[i].fn == kernel && !g_prof[i].closed) {
        if (cudaEventRecord(g_prof[i].b, st) == cudaSuccess) g_prof[i].closed = true;
        return;
      }
    }
  }
}

unsigned int* pgnn_error_flag_ptr() {
  static std::mutex mu;
  static unsigned int* ptr[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> g(mu);
  if (!ptr[dev]) {
    unsigned int* p = nullptr;
    if (cudaMalloc(&p, sizeof(unsigned int)) != cudaSuccess || cudaMemset(p, 0, sizeof(unsigned int)) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    ptr[dev] = p;
  }
  return ptr[dev];
}

extern "C" {

int pgnn_device_error_flags(int clear) {
  unsigned int* p = pgnn_error_flag_ptr();
  if (!p) return PGNN_ECUDA;
  unsigned int v = 0;
  PGNN_CUDA(cudaDeviceSynchronize());
  PGNN_CUDA(cudaMemcpy(&v, p, sizeof v, cudaMemcpyDeviceToHost));
  if (clear && v) PGNN_CUDA(cudaMemset(p, 0, sizeof v));
  return (int)(v & 0x7fffffffu);
}

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

int pgnn_profile_enable(int on) {
  g_pgnn_profile_on.store(on ? 1 : 0);
  return PGNN_OK;
}

// Waits for the recorded events, writes one line per kernel "name<TAB>launches<TAB>total_us\n" (NUL-terminated, truncated
// to buflen), clears the records and returns the number of launches they covered (negative on error).
int64_t pgnn_profile_read(char* buf, int64_t buflen) {
  std::vector<ProfRec> recs;
  {
    std::lock_guard<std::mutex> g(g_prof_mu);
    recs.swap(g_prof);
  }
  std::map<const void*, std::pair<long long, double>> agg;
  long long n = 0;
  for (ProfRec& r : recs) {
    if (r.closed && cudaEventSynchronize(r.b) == cudaSuccess) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
        auto& e = agg[r.fn];
        e.first += 1;
        e.second += (double)ms * 1e3;
        ++n;
      }
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  cudaGetLastError();  // a failed query must not poison the next launch check
  std::string out;
  for (auto& kv : agg) {
    const char* name = nullptr;
    if (cudaFuncGetName(&name, kv.first) != cudaSuccess || !name) name = "?";
    char line[640];
    snprintf(line, sizeof line, "%.500s\t%lld\t%.3f\n", name, kv.second.first, kv.second.second);
    out += line;
  }
  cudaGetLastError();
  if (buf && buflen > 0) {
    const size_t m = out.size() < (size_t)(buflen - 1) ? out.size() : (size_t)(buflen - 1);
    memcpy(buf, out.data(), m);
    buf[m] = 0;
  }
  return (int64_t)n;
}

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
