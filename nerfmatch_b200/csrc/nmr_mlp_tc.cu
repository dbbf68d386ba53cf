// bf16 tcgen05 / TMEM MLP path -- placeholder translation unit (filled in next).
#include "nmr_tc.cuh"

namespace nmr {
int tc_pack_net(const Dims&, int, const HostNet&, TcNet* out, std::string*) { out->set = false; return 0; }
void tc_free_net(TcNet* n) { if (n->pool) cudaFree(n->pool); n->pool = nullptr; n->set = false; }
int64_t tc_round_chunk_rays(const Dims&, int64_t chunk_rays, int64_t) { return chunk_rays; }
int tc_ensure_workspace(const Dims&, int64_t, void**, size_t*, std::string* err) { *err = "bf16 path not built"; return -1; }
int tc_run(const Dims&, const TcNet&, const TcRun&, float**, const void**, int*, cudaStream_t, std::string* err) { *err = "bf16 path not built"; return -1; }
int tc_run_rows(const Dims&, const TcNet&, const float*, int64_t, int, float*, float*, void**, size_t*, int*, cudaStream_t, std::string* err) { *err = "bf16 path not built"; return -1; }
}  // namespace nmr
