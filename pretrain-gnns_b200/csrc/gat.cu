// GAT attention aggregation (chem/model.py:134-165, bio/model.py:147-180) — kernels pending; the entry
// points exist so the ABI is complete and report PGNN_EUNSUPPORTED until they land.
#include "common.cuh"

extern "C" {

int pgnn_gat_fwd(const float*, int64_t, int64_t, int64_t, const float*, const float*, int, const void*, const int32_t*,
                 const int32_t*, const int32_t*, int64_t, const float*, float, float*, float*, int64_t, void*) {
  return PGNN_EUNSUPPORTED;
}
int64_t pgnn_gat_bwd_workspace_bytes(int64_t, int64_t, int64_t, int64_t) { return 0; }
int pgnn_gat_bwd(const float*, int64_t, const float*, int64_t, int64_t, int64_t, const float*, const float*, int, const void*,
                 const int32_t*, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const int32_t*, int64_t, float,
                 const float*, float*, float*, float*, float*, void*, int64_t, void*) {
  return PGNN_EUNSUPPORTED;
}

}  // extern "C"
