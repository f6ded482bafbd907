// placeholder until the tcgen05 kernels land
#include "tt_common.cuh"
namespace tt {
bool index_tc_supported(int, int, int, int, int64_t, const void*, const void*) { return false; }
size_t index_tc_workspace(int, int64_t, int, int) { return 0; }
int index_tc(const float*, int, const float*, int, int, int64_t, int, int, int64_t, float*, int32_t*, void*, size_t, cudaStream_t) { set_error("tc path not built"); return TT_ERR_UNSUPPORTED; }
}
