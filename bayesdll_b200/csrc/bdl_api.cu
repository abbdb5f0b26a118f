// bdl_api.cu -- ABI version, thread-local error string, device queries.
#include <stdarg.h>
#include <string.h>

#include "bdl_common.cuh"

namespace bdl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return BDL_OK;
    set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
    return BDL_ERR_CUDA;
}

int num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = kSMs;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kSMs;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) {
            cached_sms = sms;
            cached_dev = dev;
        }
    }
    return cached_sms;
}

}  // namespace bdl

extern "C" int bdl_abi_version(void) { return BDL_ABI_VERSION; }
extern "C" const char* bdl_last_error(void) { return bdl::g_err; }
