// libvaeq: error reporting and device queries of the C ABI (include/vaeq.h).
#include <stdarg.h>
#include "common.cuh"

namespace vaeq {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace vaeq

extern "C" int vaeq_abi_version(void) { return VAEQ_ABI_VERSION; }
extern "C" const char *vaeq_last_error(void) { return vaeq::g_err; }
extern "C" int vaeq_sm_count(void) { return vaeq::sm_count(); }
