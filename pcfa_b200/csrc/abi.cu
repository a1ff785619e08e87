// Library-level entry points of the C ABI (include/pcfa_b200.h).
#include "common.cuh"

namespace pcfa { unsigned long long g_launch_count = 0; }

extern "C" int pcfa_abi_version(void) { return PCFA_ABI_VERSION; }

extern "C" int64_t pcfa_launch_count(void) { return (int64_t)pcfa::g_launch_count; }

extern "C" const char* pcfa_status_string(int status) {
    switch (status) {
        case PCFA_OK:          return "ok";
        case PCFA_E_BADARG:    return "pcfa: bad argument (null pointer, non-positive size or unsupported parameter)";
        case PCFA_E_TOOLARGE:  return "pcfa: dimension too large for the int32 index maps";
        case PCFA_E_NODEVICE:  return "pcfa: no sm_100 device or driver entry point unavailable";
        case PCFA_E_WORKSPACE: return "pcfa: workspace missing or too small";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "pcfa: unknown status";
}

// Developer hook (not part of the ABI header): per-CTA globaltimer stamps of the backward tensor-core kernels are
// written to `dev_ptr` (2 passes x 1024 CTAs x 8 u64) while it is non-null.  scripts/bwd_timeline.py reads them.
namespace pcfa { void corr_pyramid_bwd_set_trace(void* p); }
extern "C" void pcfa_debug_set_bwd_trace(void* dev_ptr) { pcfa::corr_pyramid_bwd_set_trace(dev_ptr); }
