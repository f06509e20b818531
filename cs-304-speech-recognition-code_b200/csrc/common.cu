// Error reporting and small queries of the loe_b200 C ABI.
#include "common.cuh"
#include <stdarg.h>

namespace loe {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace loe

extern "C" int loe_abi_version(void) { return LOE_ABI_VERSION; }
extern "C" const char* loe_last_error(void) { return loe::g_err; }
extern "C" int loe_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        loe::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return -LOE_ERR_CUDA;
    }
    return n;
}
