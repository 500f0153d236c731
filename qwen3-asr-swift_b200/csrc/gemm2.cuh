// gemm2.cuh — the CTA-pair (cta_group::2) variant of the tcgen05 GEMM / implicit-GEMM convolution in gemm.cuh.
//
// Why: with one CTA per tile, a 128 x 256 tile moves 48 KB through shared memory per K block twice (TMA writes, UMMA reads):
// ~192 B/clk against the ~128 B/clk an SM's shared memory sustains, so the tensor pipe tops out near 70 % (measured 45-72 %,
// profiles/r1c_*).  A CTA pair on the two SMs of a TPC computes a 256 x BN tile with ONE tcgen05.mma.cta_group::2: each CTA
// stages its own 128 rows of A but only HALF of the weight tile (BN/2 rows); the hardware shares the halves between the pair.
// Per CTA and K block: 16 KB + BN * 64 B instead of 16 KB + BN * 128 B.
//
// Structure per CTA (same roles as gemm.cuh): warp 0 TMA producer (both CTAs, signalling the LEADER's full barrier), warp 1
// MMA issuer (leader CTA only; commits multicast to both CTAs' empty / accumulator-full barriers), warp 2 TMEM allocator
// (cta_group::2, both CTAs), warps 4-11 epilogue (each CTA drains its own 128 TMEM lanes; one lane per warp reports the
// drained accumulator to the leader's barrier).  The epilogue body is gemm.cuh's gemm_epilogue_tile.
#pragma once
#include "gemm.cuh"

namespace q3 {

namespace ptx {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* smem_ptr, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(smem_ptr)), "r"(rank));
    return out;
}
// Relaxed: this arrival only hands a drained TMEM accumulator back to the leader's MMA warp (the tcgen05.ld were waited for and
// fenced with tcgen05.fence::before_thread_sync); the epilogue's global stores need not be visible to the peer.  With
// .release.cluster the compiler emits MEMBAR.ALL.GPU + ERRBAR before every arrival, i.e. each epilogue warp waits for all of its
// outstanding stores once per tile (ncu: 11 % of the stall samples of the fused q|k|v product).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in this CTA's shared memory, the bytes are counted on the barrier at `bar_cluster_addr`
// (the leader's)
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* m, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* smem_dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem of both CTAs] (+)= A[each CTA's smem] * B[halves in the two CTAs' smem]; issued by the leader CTA only
__device__ __forceinline__ void mma2_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void mma2_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

}  // namespace ptx

__host__ __device__ constexpr int gemm2_stage_bytes(int BN) { return GEMM_BM * 128 + (BN / 2) * 128; }
__host__ __device__ constexpr int gemm2_stages(int BN) { return (196 * 1024) / gemm2_stage_bytes(BN) > 8 ? 8 : (196 * 1024) / gemm2_stage_bytes(BN); }
__host__ __device__ constexpr int gemm2_smem_bytes(int BN) { return gemm2_stages(BN) * gemm2_stage_bytes(BN) + EPI_STAGE_BYTES + 1024 + 256; }

// this CTA's M tile (index mt over tiles_w x tiles_h x tiles_b) of column tile tn
__device__ __forceinline__ TileCoord gemm2_tile_coord(const GemmDev& p, int tn, int mt) {
    TileCoord t;
    t.tn = tn;
    int r = mt;
    t.w0 = (r % p.tiles_w) * p.Wb;
    r /= p.tiles_w;
    t.h0 = (r % p.tiles_h) * p.Hb;
    t.b0 = (r / p.tiles_h) * p.Bb;  // mt past the last tile gives b0 >= OB: TMA zero-fills, the epilogue stores nothing
    return t;
}

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                const GemmDev p) {
    constexpr int STAGES = gemm2_stages(BN);
    constexpr int STAGE_BYTES = gemm2_stage_bytes(BN);
    constexpr int ACC_STRIDE = gemm_acc_stride(BN);
    constexpr int TMEM_COLS = 2 * ACC_STRIDE;
    constexpr uint32_t IDESC = ptx::umma_idesc_bf16(2 * GEMM_BM, BN);
    static_assert(BN % 32 == 0 && BN <= 256 && (BN / 2) % 8 == 0, "BN");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_stage = smem + STAGES * STAGE_BYTES;                                // output staging of the plain epilogue (gemm.cuh)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + (p.tma_out ? EPI_STAGE_BYTES : 0));  // used in the leader only
    uint64_t* empty_bar = full_bar + STAGES;                                         // per CTA
    uint64_t* tfull_bar = empty_bar + STAGES;                                        // [2] per CTA
    uint64_t* tempty_bar = tfull_bar + 2;                                            // [2] used in the leader only
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    const int m_pairs = (m_tiles + 1) >> 1;
    const int num_super = p.tiles_n * m_pairs;  // 256-row x BN super tiles
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tile_rows = p.Wb * p.Hb * p.Bb;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        if (p.tma_out) ptx::prefetch_tmap(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], 16);  // one lane of each of the 8 epilogue warps of both CTAs
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc2<TMEM_COLS>(tmem_slot);
    ptx::tc_fence_before();
    ptx::cluster_sync();  // both CTAs' barriers are initialised before either signals the other
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            ptx::grid_dep_launch();
            ptx::grid_dep_wait();
            int stage = 0;
            uint32_t phase = 0;
            for (int st = pair; st < num_super; st += n_pairs) {
                const TileCoord tc = gemm2_tile_coord(p, st % p.tiles_n, 2 * (st / p.tiles_n) + (int)rank);
                const int nrow = tc.tn * BN + (int)rank * (BN / 2);  // this CTA's half of the weight tile
                int tap = 0, cc = 0;
                for (int kb = 0; kb < p.num_kb; kb++) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + GEMM_BM * 128;
                    const uint32_t fb = ptx::map_to_cta(&full_bar[stage], 0);
                    // the leader arms its barrier with the bytes of BOTH CTAs; the peer's copies are counted on the same barrier
                    if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(2 * (tile_rows * 128 + (BN / 2) * 128)));
                    ptx::tma2_load_4d(sa, &tmA, cc * GEMM_BK, tc.w0 * p.sw + p.tap_dw[tap], tc.h0 * p.sh + p.tap_dh[tap], tc.b0, fb);
                    ptx::tma2_load_2d(sb, &tmB, tap * p.C + cc * GEMM_BK, nrow, fb);
                    if (++cc == p.kb_per_tap) { cc = 0; tap++; }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader && lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            const int last_steps = (p.C - (p.kb_per_tap - 1) * GEMM_BK + 15) >> 4;
            for (int st = pair; st < num_super; st += n_pairs, it++) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
                int cc = 0;
                for (int kb = 0; kb < p.num_kb; kb++) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sb = sa + GEMM_BM * 128;
                    const int steps = (cc == p.kb_per_tap - 1) ? last_steps : GEMM_BK / 16;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; k++) {
                        if (k < steps)
                            ptx::mma2_bf16_ss(d_tmem, ptx::umma_desc_sw128(sa + k * 32), ptx::umma_desc_sw128(sb + k * 32), IDESC,
                                              (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::mma2_commit_both(&empty_bar[stage]);  // frees the slot in both CTAs
                    if (++cc == p.kb_per_tap) cc = 0;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::mma2_commit_both(&tfull_bar[as]);  // accumulator complete -> both CTAs' epilogues
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (each CTA: its own 128 rows) =====
        const int q = warp & 3;
        const int chalf = (warp - 4) >> 2;
        int it = 0;
        ptx::grid_dep_wait();  // gemm_epilogue_tile prefetches residual rows (the previous kernel's output) ahead of the accumulator
        EpiStage es{&tmC, epi_stage + chalf * 2 * EPI_STAGE_BUF, 0u};
        for (int st = pair; st < num_super; st += n_pairs, it++) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const TileCoord tc = gemm2_tile_coord(p, st % p.tiles_n, 2 * (st / p.tiles_n) + (int)rank);
            gemm_epilogue_tile<BN, EPI>(p, tc, tmem_base + as * ACC_STRIDE, &tfull_bar[as], aphase, q, lane, chalf, tile_rows, es);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::map_to_cta(&tempty_bar[as], 0));
        }
        if (p.tma_out && q == 0 && lane == 0) ptx::bulk_wait_group<0>();  // the staged tiles have reached memory before the CTA leaves
    }

    ptx::tc_fence_before();
    ptx::cluster_sync();  // the leader's MMAs read the peer's shared memory and write its TMEM: leave together
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2<TMEM_COLS>(tmem_base);
    }
}

}  // namespace q3
