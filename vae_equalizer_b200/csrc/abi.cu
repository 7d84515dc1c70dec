// libvaeq: error reporting and device queries of the C ABI (include/vaeq.h).
#include <stdarg.h>
#include <vector>
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

// ---- optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg) ----
struct KRec { int kind; cudaEvent_t a, b; };
static bool g_ktime_on = false;
static std::vector<KRec> g_krecs;
static std::vector<cudaEvent_t> g_evpool;
static long long g_launches[VAEQ_NKINDS] = {0};

static cudaEvent_t ev_get() {
    if (!g_evpool.empty()) { cudaEvent_t e = g_evpool.back(); g_evpool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void ktime_begin(int kind, cudaStream_t st) {
    if (kind >= 0 && kind < VAEQ_NKINDS) g_launches[kind]++;
    if (!g_ktime_on) return;
    KRec r; r.kind = kind; r.a = ev_get(); r.b = ev_get();
    cudaEventRecord(r.a, st);
    g_krecs.push_back(r);
}
void ktime_end(int kind, cudaStream_t st) {
    (void)kind;
    if (!g_ktime_on || g_krecs.empty()) return;
    cudaEventRecord(g_krecs.back().b, st);
}

}  // namespace vaeq

extern "C" int vaeq_kernel_timing(int32_t enable) {
    using namespace vaeq;
    for (auto &r : g_krecs) { g_evpool.push_back(r.a); g_evpool.push_back(r.b); }
    g_krecs.clear();
    g_ktime_on = enable != 0;
    return VAEQ_OK;
}
extern "C" int vaeq_kernel_timing_read(float *ms_sum, int32_t *count) {
    using namespace vaeq;
    for (int k = 0; k < VAEQ_NKINDS; ++k) { ms_sum[k] = 0.f; count[k] = 0; }
    for (auto &r : g_krecs) {
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e != cudaSuccess) { set_error("cudaEventSynchronize: %s", cudaGetErrorString(e)); return (int)e; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        if (r.kind >= 0 && r.kind < VAEQ_NKINDS) { ms_sum[r.kind] += ms; count[r.kind]++; }
    }
    return VAEQ_OK;
}
extern "C" int64_t vaeq_launch_count(int32_t kind) {
    using namespace vaeq;
    if (kind >= 0 && kind < VAEQ_NKINDS) return g_launches[kind];
    long long t = 0;
    for (int k = 0; k < VAEQ_NKINDS; ++k) t += g_launches[k];
    return t;
}

extern "C" int vaeq_abi_version(void) { return VAEQ_ABI_VERSION; }
extern "C" const char *vaeq_last_error(void) { return vaeq::g_err; }
extern "C" int vaeq_sm_count(void) { return vaeq::sm_count(); }
