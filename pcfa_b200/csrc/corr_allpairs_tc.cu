// tcgen05/TMEM all-pairs correlation with the pyramid pooling fused in the epilogue.
// (placeholder until the tensor-core kernel lands: reports "unsupported" so the dispatcher uses SIMT)
#include "common.cuh"
namespace pcfa {
bool corr_pyramid_tc_supported(int, int, int, int, int) { return false; }
int64_t corr_pyramid_tc_workspace_bytes(int, int, int, int, int) { return 0; }
int corr_pyramid_forward_tc(const float*, const float*, float*, void*, int64_t, int, int, int, int,
                            int, cudaStream_t) { return PCFA_E_BADARG; }
}  // namespace pcfa
