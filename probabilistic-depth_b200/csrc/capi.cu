// Library-level entry points of the C ABI (include/dpv_b200.h).
#include "dpv_common.cuh"

namespace dpv {
std::atomic<long long> g_launch_count{0};   // the one piece of mutable library state: a relaxed counter
}

extern "C" int dpv_abi_version(void) { return 2; }

extern "C" long long dpv_launch_count(void) { return dpv::g_launch_count.load(std::memory_order_relaxed); }

extern "C" const char* dpv_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case DPV_E_BADARG: return "dpv: bad argument (null pointer or non-positive dimension)";
        case DPV_E_UNSUPP: return "dpv: dimension or option not supported by the sm_100a kernels";
        case DPV_E_NODEVICE: return "dpv: no usable sm_100 device";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "dpv: unknown error";
}
