// skinny.cuh — weight-streaming tcgen05 GEMM for the decode step (few token rows, big weight matrix).
//
//   Y[m, n] = sum_k X[m, k] W[n, k]        m < Mtok <= NB <= 128 token rows, n < N
//
// The decode-step products of FloatTextDecoder.swift:77-79, 107, 128-132 have M = batch (<= 64 per GPU in
// the reference configs) and are pure weight streaming.  The general kernel (gemm.cuh) would waste half of
// its 128-row M tile on padding and run N/BN CTAs with a serial K loop each; here the roles are swapped:
// 128 weight rows are the UMMA M operand, the token rows are the UMMA N operand (NB columns of TMEM), and K
// is split across CTAs (grid = N/128 x splits) so that every SM streams a slice of W.  Split-K partials are
// fp32 in a workspace and are summed in a fixed order by the consumer kernel (fused with the residual add +
// RMSNorm, or with the q/k-norm + RoPE of the attention kernel), so results are deterministic.
//
// CTA = 256 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
#pragma once
#include "gemm.cuh"

namespace q3 {

struct SkinnyDev {
    int N, Mtok;
    int num_kb, kb_per_split;
    void* out;
    int ldo;                // bf16 outputs: row pitch in elements
    long long split_stride; // SK_PARTIAL: elements between split slabs
    int gu_half;            // SK_SWIGLU: rows per gate (and per up) block
    float* fix_ws;          // SK_SWIGLU with split K: fp32 slabs [tile][split][NB][128]
    int* tickets;           // SK_SWIGLU with split K: one arrival counter per tile (zero on entry, left zero)
    int splits;
};

constexpr int SK_BM = 128;  // weight rows per CTA
constexpr int SK_BK = 64;
__host__ __device__ constexpr int sk_stage_bytes(int NB) { return SK_BM * 128 + NB * 128; }
// ring depth: 96 KB of tiles in flight per CTA keeps an SM's share of HBM busy and leaves room for the next kernel's
// CTAs to become resident early (programmatic dependent launch)
__host__ __device__ constexpr int sk_stages(int NB) { return (96 * 1024) / sk_stage_bytes(NB) > 6 ? 6 : (96 * 1024) / sk_stage_bytes(NB); }
__host__ __device__ constexpr int sk_tmem_cols(int NB) { return NB <= 32 ? 32 : NB <= 64 ? 64 : NB <= 128 ? 128 : 256; }
__host__ __device__ constexpr int sk_smem_bytes(int NB, int epi) {
    return sk_stages(NB) * sk_stage_bytes(NB) + 1024 + 256 + (epi == SK_SWIGLU ? 64 * (NB + 1) * 4 : 0);
}

__device__ __forceinline__ float sk_swiglu(float g_acc, float u_acc) {
    const float g = bf16_round(g_acc);
    const float s = bf16_round(silu(g));
    return s * bf16_round(u_acc);
}

template <int NB, int EPI>
__global__ void __launch_bounds__(256, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const SkinnyDev p) {
    constexpr int STAGES = sk_stages(NB);
    constexpr int STAGE_BYTES = sk_stage_bytes(NB);
    constexpr int TMEM_COLS = sk_tmem_cols(NB);
    constexpr uint32_t IDESC = ptx::umma_idesc_bf16(SK_BM, NB);
    constexpr int CH = NB < 32 ? NB : 32;  // TMEM columns per epilogue chunk
    static_assert(NB % 16 == 0 && NB >= 16 && NB <= 128, "NB");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
    float* s_u = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);  // SK_SWIGLU: [64][NB + 1]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * SK_BM;
    const int split = blockIdx.y;
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmW);
        ptx::prefetch_tmap(&tmX);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(tfull_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            ptx::grid_dep_launch();
            // The weight tiles do not depend on the previous kernel: fill the ring with them first, then wait for the
            // producer of X (programmatic dependent launch), then add the X tiles of those stages.
            const int pre = min(STAGES, kb1 - kb0);
            for (int i = 0; i < pre; i++) {
                ptx::mbar_arrive_expect_tx(&full_bar[i], (uint32_t)STAGE_BYTES);
                ptx::tma_load_2d(smem + i * STAGE_BYTES, &tmW, (kb0 + i) * SK_BK, n0, &full_bar[i]);
            }
            ptx::grid_dep_wait();
            for (int i = 0; i < pre; i++)
                ptx::tma_load_2d(smem + i * STAGE_BYTES + SK_BM * 128, &tmX, (kb0 + i) * SK_BK, 0, &full_bar[i]);
            int stage = 0;
            uint32_t phase = 1;  // the ring has been filled once
            for (int kb = kb0 + pre; kb < kb1; kb++) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                ptx::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)STAGE_BYTES);
                ptx::tma_load_2d(sa, &tmW, kb * SK_BK, n0, &full_bar[stage]);
                ptx::tma_load_2d(sa + SK_BM * 128, &tmX, kb * SK_BK, 0, &full_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; kb++) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t sb = sa + SK_BM * 128;
#pragma unroll
                for (int k = 0; k < SK_BK / 16; k++)
                    ptx::mma_bf16_ss(tmem_base, ptx::umma_desc_sw128(sa + k * 32), ptx::umma_desc_sw128(sb + k * 32), IDESC,
                                     (kb > kb0 || k > 0) ? 1u : 0u);
                ptx::mma_commit(&empty_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            ptx::mma_commit(tfull_bar);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int r = q * 32 + lane;  // weight row within the tile = TMEM lane
        ptx::mbar_wait(tfull_bar, 0);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
        const bool have = kb1 > kb0;  // an empty K slice contributes zeros (its accumulator was never written)

        if constexpr (EPI == SK_SWIGLU) {
            const int blk = r / p.gu_half;
            const bool is_up = blk & 1;
            const int jl = (blk >> 1) * p.gu_half + r % p.gu_half;  // output column within the tile, 0..63
            bf16* out = reinterpret_cast<bf16*>(p.out) + (size_t)blockIdx.x * 64 + jl;
            if (p.splits == 1) {
                // pass 1: up lanes publish their values
#pragma unroll 1
                for (int c = 0; c < NB / CH; c++) {
                    uint32_t v[CH];
                    if constexpr (CH == 32) ptx::tmem_ld_32x32(t_row + c * CH, v); else ptx::tmem_ld_32x16(t_row + c * CH, reinterpret_cast<uint32_t(&)[16]>(v));
                    ptx::tmem_ld_wait();
                    if (is_up) {
#pragma unroll
                        for (int j = 0; j < CH; j++) s_u[jl * (NB + 1) + c * CH + j] = __uint_as_float(v[j]);
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
#pragma unroll 1
                for (int c = 0; c < NB / CH; c++) {
                    uint32_t v[CH];
                    if constexpr (CH == 32) ptx::tmem_ld_32x32(t_row + c * CH, v); else ptx::tmem_ld_32x16(t_row + c * CH, reinterpret_cast<uint32_t(&)[16]>(v));
                    ptx::tmem_ld_wait();
                    if (!is_up) {
#pragma unroll
                        for (int j = 0; j < CH; j++) {
                            const int m = c * CH + j;
                            if (m < p.Mtok)
                                out[(size_t)m * p.ldo] = __float2bfloat16_rn(sk_swiglu(__uint_as_float(v[j]), s_u[jl * (NB + 1) + m]));
                        }
                    }
                }
            } else {
                // split K: every CTA of the tile parks its fp32 accumulator in a slab; the last one to arrive sums the slabs
                // in split order (deterministic) and applies SwiGLU.
                float* slab0 = p.fix_ws + (size_t)blockIdx.x * p.splits * NB * SK_BM;
                float* mine = slab0 + (size_t)split * NB * SK_BM + r;
#pragma unroll 1
                for (int c = 0; c < NB / CH; c++) {
                    uint32_t v[CH];
                    if constexpr (CH == 32) ptx::tmem_ld_32x32(t_row + c * CH, v); else ptx::tmem_ld_32x16(t_row + c * CH, reinterpret_cast<uint32_t(&)[16]>(v));
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < CH; j++)
                        if (c * CH + j < p.Mtok) mine[(size_t)(c * CH + j) * SK_BM] = __uint_as_float(v[j]);
                }
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);
                if (r == 0) *s_flag = atomicAdd(p.tickets + blockIdx.x, 1);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (*s_flag == p.splits - 1) {
                    __threadfence();
                    if (r == 0) p.tickets[blockIdx.x] = 0;  // ready for the next launch
                    // thread t owns outputs [4 (t & 15), +4) of the tile for token rows m = (t >> 4) + 8 i: float4 loads of the gate
                    // rows and of their up partners, two token rows (12 loads for three splits) in flight at a time
                    const int t = q * 32 + lane;
                    const int j4 = (t & 15) * 4;
                    const int rg = (j4 / p.gu_half) * 2 * p.gu_half + j4 % p.gu_half;  // gate row; the up row is rg + gu_half
                    bf16* o4 = reinterpret_cast<bf16*>(p.out) + (size_t)blockIdx.x * 64 + j4;
                    for (int m0 = t >> 4; m0 < p.Mtok; m0 += 16) {
                        float4 g[2], u[2];
#pragma unroll
                        for (int i = 0; i < 2; i++) g[i] = u[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int sp = 0; sp < p.splits; sp++) {
                            float4 a[2], b[2];
#pragma unroll
                            for (int i = 0; i < 2; i++) {
                                const int m = m0 + 8 * i;
                                const float* rowp = slab0 + ((size_t)sp * NB + (m < p.Mtok ? m : m0)) * SK_BM;
                                a[i] = __ldcg(reinterpret_cast<const float4*>(rowp + rg));
                                b[i] = __ldcg(reinterpret_cast<const float4*>(rowp + rg + p.gu_half));
                            }
#pragma unroll
                            for (int i = 0; i < 2; i++) {
                                g[i].x += a[i].x; g[i].y += a[i].y; g[i].z += a[i].z; g[i].w += a[i].w;
                                u[i].x += b[i].x; u[i].y += b[i].y; u[i].z += b[i].z; u[i].w += b[i].w;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const int m = m0 + 8 * i;
                            if (m < p.Mtok)
                                *reinterpret_cast<uint2*>(o4 + (size_t)m * p.ldo) =
                                    make_uint2(pack_bf16x2(sk_swiglu(g[i].x, u[i].x), sk_swiglu(g[i].y, u[i].y)),
                                               pack_bf16x2(sk_swiglu(g[i].z, u[i].z), sk_swiglu(g[i].w, u[i].w)));
                        }
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int c = 0; c < NB / CH; c++) {
                uint32_t v[CH];
                if constexpr (CH == 32) ptx::tmem_ld_32x32(t_row + c * CH, v); else ptx::tmem_ld_32x16(t_row + c * CH, reinterpret_cast<uint32_t(&)[16]>(v));
                ptx::tmem_ld_wait();
                if (n0 + r < p.N) {
#pragma unroll
                    for (int j = 0; j < CH; j++) {
                        const int m = c * CH + j;
                        if (m < p.Mtok) {
                            const float f = have ? __uint_as_float(v[j]) : 0.f;
                            if constexpr (EPI == SK_PARTIAL)
                                reinterpret_cast<float*>(p.out)[(size_t)split * p.split_stride + (size_t)m * p.N + n0 + r] = f;
                            else
                                reinterpret_cast<bf16*>(p.out)[(size_t)m * p.ldo + n0 + r] = __float2bfloat16_rn(f);
                        }
                    }
                }
            }
        }
        ptx::tc_fence_before();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

}  // namespace q3
