// Shared host/device helpers of libvaeq (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/vaeq.h"

namespace vaeq {

void set_error(const char *fmt, ...);
int sm_count();
void ktime_begin(int kind, cudaStream_t st);   // counts the launch; records an event pair when timing is on
void ktime_end(int kind, cudaStream_t st);

#define VAEQ_CHECK_ARG(cond, ...)               \
    do {                                        \
        if (!(cond)) {                          \
            vaeq::set_error(__VA_ARGS__);       \
            return VAEQ_EINVAL;                 \
        }                                       \
    } while (0)

#define VAEQ_CUDA(call)                                                            \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) {                                                   \
            vaeq::set_error("%s failed: %s", #call, cudaGetErrorString(_e));       \
            return (int)_e;                                                        \
        }                                                                          \
    } while (0)

#define VAEQ_LAUNCH_CHECK(name)                                                    \
    do {                                                                           \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess) {                                                   \
            vaeq::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
            return (int)_e;                                                        \
        }                                                                          \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Function attributes (cudaFuncAttributeMaxDynamicSharedMemorySize) and occupancy-derived grids are PER DEVICE: every cache of
// them is indexed by the calling thread's current device, like sm_count().
constexpr int VAEQ_MAX_DEVICES = 64;
static inline int cur_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= VAEQ_MAX_DEVICES) return 0;
    return dev;
}
struct SmemAttrCache {
    size_t set[VAEQ_MAX_DEVICES] = {0};
};
template <typename K>
static inline int ensure_dyn_smem(K kern, size_t smem, SmemAttrCache &c) {
    const int dev = cur_device();
    if (smem > c.set[dev]) {
        VAEQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c.set[dev] = smem;
    }
    return VAEQ_OK;
}

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV values per thread; result valid in thread 0.  scratch: NV * 32 floats.
template <int NV, typename T>
__device__ __forceinline__ void block_sum(T (&v)[NV], T *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[i * 32 + wid] = v[i];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            T x = lane < nw ? scratch[i * 32 + lane] : T(0);
            v[i] = warp_sum(x);
        }
    }
}

// streaming (read-once / write-once) global accesses: keep them out of L1
__device__ __forceinline__ float ldg_stream(const float *p) { return __ldcs(p); }
__device__ __forceinline__ float4 ldg_stream(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ void stg_stream(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void stg_stream(float4 *p, float4 v) { __stcs(p, v); }

// float16 bit pattern -> float (tx symbols, sf:89 stores them as float16)
__device__ __forceinline__ float half_bits_to_float(uint16_t b) { return __half2float(__ushort_as_half(b)); }

}  // namespace vaeq
