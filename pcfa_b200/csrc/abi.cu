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
