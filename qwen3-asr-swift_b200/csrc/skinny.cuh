// skinny.cuh — weight-streaming tcgen05 GEMM for the decode step (few token rows, big weight matrix).
//
//   Y[m, n] = sum_k X[m, k] W[n, k]        m < Mtok <= NB <= 256 token rows (the widest UMMA N), n < N
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
};

constexpr int SK_BM = 128;  // weight rows per CTA
constexpr int SK_BK = 64;
__host__ __device__ constexpr int sk_stage_bytes(int NB) { return SK_BM * 128 + NB * 128; }
// Ring depth: bytes in flight per SM set the streaming rate (Little's law: ~2 us loaded latency), but a CTA that holds more than
// ~96 KB keeps the next kernel's CTAs from becoming resident early (programmatic dependent launch).  Short K slices (0.6B:
// latency-bound phases) use 96 KB; long ones (DEEP; 1.7B: bandwidth-bound phases) half as much again.  A compile-time constant:
// a run-time depth cost 5 % of the 0.6B decode step (measured).
__host__ __device__ constexpr int sk_stages(int NB, bool deep = false) {
    const int s = (96 * 1024) / sk_stage_bytes(NB) > 6 ? 6 : (96 * 1024) / sk_stage_bytes(NB);
    return deep ? s + s / 2 : s;
}
__host__ __device__ constexpr int sk_tmem_cols(int NB) { return NB <= 32 ? 32 : NB <= 64 ? 64 : NB <= 128 ? 128 : 256; }
__host__ __device__ constexpr int sk_smem_bytes(int NB, bool deep = false) { return sk_stages(NB, deep) * sk_stage_bytes(NB) + 1024 + 256; }

template <int NB, int EPI, bool DEEP>
__global__ void __launch_bounds__(256, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const SkinnyDev p) {
    constexpr int STAGES = sk_stages(NB, DEEP);
    constexpr int STAGE_BYTES = sk_stage_bytes(NB);
    constexpr int TMEM_COLS = sk_tmem_cols(NB);
    constexpr uint32_t IDESC = ptx::umma_idesc_bf16(SK_BM, NB);
    constexpr int CH = NB < 32 ? NB : 32;  // TMEM columns per epilogue chunk
    static_assert(NB % 16 == 0 && NB >= 16 && NB <= 256, "NB");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

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

        {
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
