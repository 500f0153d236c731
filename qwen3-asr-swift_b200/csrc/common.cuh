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

// Synchronous host -> device copy that is really complete on return.  cudaMemcpy from PAGEABLE memory returns once the source has been
// staged; the DMA to the device may still be in flight in the legacy default stream, and the handle's streams are non-blocking, so a
// kernel launched on them right afterwards is not ordered behind it (seen as stale operands under stress in the debug hooks).
// Synchronising the legacy stream closes the gap.  (The hot path does not come through here: it copies from pinned staging buffers
// on its own streams.)
#define Q3_H2D_SYNC(dst, src, bytes)                                                \
    do {                                                                            \
        Q3_CUDA(cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice));         \
        Q3_CUDA(cudaStreamSynchronize(cudaStreamLegacy));                           \
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

// ---- Blackwell's packed fp32 pipe: two fp32 lanes per issue slot (FADD2 / FMUL2 / FFMA2), IEEE round-to-nearest per lane ----
__device__ __forceinline__ unsigned long long& f2_bits(float2& v) { return reinterpret_cast<unsigned long long&>(v); }
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
    float2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(f2_bits(r)) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return r;
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
    float2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(f2_bits(r)) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return r;
}
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f2_bits(r)) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return r;
}
__device__ __forceinline__ float2 f2_splat(float c) { return make_float2(c, c); }

// The same GELU for two values at once on the packed pipe.  erf by the two minimax polynomials of N. Juffa's erff (< 1 ulp each:
// |z| <= 0.927734375: z + z p(z^2); above: 1 - exp(q(|z|)), sign restored): 10 packed instructions per PAIR, + 11 packed and ~8
// scalar ones when a lane of the warp is in the tail, against ~28 per value for erff.  Differs from erff by <= 3 fp32 ulps,
// i.e. from gelu_erf after the bf16 store in about one value in 10^4 (the exact-erf GELU is 70 % of conv1's instructions and the
// reason the fc1 / conv epilogues are slower than their MMAs).
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
    const float2 z = f2_mul(x, f2_splat(0.70710678118654752440f));
    const float2 t = make_float2(fabsf(z.x), fabsf(z.y));
    const float2 s = f2_mul(z, z);
    // |z| <= 0.927734375
    float2 p = f2_fma(f2_splat(-5.96761703e-4f), s, f2_splat(4.99119423e-3f));
    p = f2_fma(p, s, f2_splat(-2.67681349e-2f));
    p = f2_fma(p, s, f2_splat(1.12819925e-1f));
    p = f2_fma(p, s, f2_splat(-3.76125336e-1f));
    p = f2_fma(p, s, f2_splat(1.28379166e-1f));
    float2 e = f2_fma(p, z, z);
    // |z| > 0.927734375: evaluated (for both values, branch-free) only when some lane of the warp needs it — one warp-uniform
    // branch instead of a divergent expf call per value.  exp(r) = 2^(r log2 e) on the SFU (ex2.approx: 2 ulps; r < 0 and the
    // result is subtracted from 1).  The vote runs over the lanes that happen to be converged here (__activemask); that is only
    // an optimisation hint: whichever way a lane gets here, it computes the same value.
    const bool tail_x = t.x > 0.927734375f, tail_y = t.y > 0.927734375f;
    if (__any_sync(__activemask(), tail_x || tail_y)) {
        float2 r = f2_fma(f2_splat(-1.72853470e-5f), t, f2_splat(3.83197126e-4f));
        const float2 u = f2_fma(f2_splat(-3.88396438e-3f), t, f2_splat(2.42546219e-2f));
        r = f2_fma(r, s, u);
        r = f2_fma(r, t, f2_splat(-1.06777877e-1f));
        r = f2_fma(r, t, f2_splat(-6.34846687e-1f));
        r = f2_fma(r, t, f2_splat(-1.28717512e-1f));
        r = f2_fma(r, t, make_float2(-t.x, -t.y));
        const float2 rl = f2_mul(r, f2_splat(1.4426950408889634f));
        float2 ex;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(rl.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(rl.y));
        const float2 om = f2_fma(ex, f2_splat(-1.0f), f2_splat(1.0f));
        e.x = tail_x ? copysignf(om.x, z.x) : e.x;
        e.y = tail_y ? copysignf(om.y, z.y) : e.y;
    }
    const float2 h = f2_mul(x, f2_splat(0.5f));
    return f2_fma(h, e, h);
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  The GEMM / attention epilogues hold one output ROW per thread, so a warp's
// 32 lanes touch 32 different lines per instruction: halving the instructions per row halves the L1 tag-stage work of an epilogue.
// 32-byte aligned addresses only.
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g,
                                             uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h)
                 : "memory");
}
struct U8 {
    uint32_t v[8];
};
__device__ __forceinline__ U8 ld_global_nc_v8(const void* p) {
    U8 r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}

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
