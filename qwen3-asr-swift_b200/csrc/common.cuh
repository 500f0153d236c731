// common.cuh — shared helpers for the q3asr sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <stdexcept>
#include <string>
#include <utility>

namespace q3 {

typedef __nv_bfloat16 bf16;

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define Q3_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            throw ::q3::Error(3, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +  \
                                     __FILE__ + ":" + std::to_string(__LINE__) + ")");      \
    } while (0)

#define Q3_CHECK(cond, code, msg)                        \
    do {                                                 \
        if (!(cond)) throw ::q3::Error((code), (msg));   \
    } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute (the > 48 KB dynamic shared memory opt-in) is per device: a process that drives several GPUs
// (q3asr_pool, one worker thread per device) has to opt in on each of them.  One flag bit per device ordinal; the action is
// idempotent, so two threads racing on the same device are harmless.
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0};
    template <typename F>
    void operator()(F&& f) {
        int dev = 0;
        Q3_CUDA(cudaGetDevice(&dev));
        const unsigned long long bit = 1ull << (dev & 63);
        if (done.load(std::memory_order_acquire) & bit) return;
        f();
        done.fetch_or(bit, std::memory_order_release);
    }
};

// Kernel launch with optional programmatic stream serialization (PDL): the kernel may start while its predecessor
// in the stream is still draining; it must call ptx::grid_dep_wait() before touching the predecessor's outputs.
// The flag is per host thread (one handle = one caller thread); forward.cu raises it around the decode step.
inline bool& pdl_enabled() {
    static thread_local bool on = false;
    return on;
}
template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (pdl_enabled()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    Q3_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

// ---- device helpers -----------------------------------------------------------------
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact-erf GELU (MLX `gelu`): x * 0.5 * (1 + erf(x / sqrt(2)))
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// SiLU: x * sigmoid(x)
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

}  // namespace q3
