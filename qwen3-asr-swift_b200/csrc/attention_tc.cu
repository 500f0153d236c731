// attention_tc.cu — segment-packed multi-head attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Same contract as flash_attn_launch (attention.cu): replaces MLXFast.scaledDotProductAttention as used through
// SDPA.multiHead / attendAndMerge (/root/reference/Sources/MLXCommon/SDPA.swift:18-101) with the block-diagonal window mask
// of AudioEncoder.swift:337-357, 463-489 (encoder, head_dim 64) and the causal mask of FloatTextDecoder.swift:200-213
// (decoder prefill, head_dim 128, GQA); every window / prompt is an independent segment of the packed row list.
//
// One CTA = 128 queries of one (segment, head).  256 threads:
//   warp 0    TMA producer   Q tile once, then K / V blocks of 128 keys (128-byte swizzle, mbarrier ring)
//   warp 1    MMA issuer     S = Q K^T (M 128, N 128, K = head_dim; both operands K-major) into a TMEM score buffer,
//                            then PV = P V (M 128, N = head_dim, K 128; P from shared memory, V as an MN-major operand, so
//                            no transpose of V is ever materialised) into a second TMEM region
//   warp 2    TMEM allocator
//   warps 4-7 softmax        one thread per query row: tcgen05.ld the scores, online softmax in the exp2 domain with fp32
//                            statistics, P rounded to bf16 and written to shared memory in the swizzled K-major layout the
//                            MMA reads; the running output lives in registers (O = O * alpha + PV), so TMEM is never rescaled
// The score buffer is double-buffered for head_dim 128, so the tensor core computes block j+1's scores while the softmax
// warps work on block j.  Rounding points are those of attention.cu (P in bf16 for the PV product, row sums unrounded).
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "ops.cuh"
#include "ptx.cuh"

namespace q3 {

void make_tmap_2d_bf16(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t row_stride_elems, uint32_t box_cols,
                       uint32_t box_rows);  // gemm.cu

namespace {

constexpr int FA_BQ = 128, FA_BKV = 128;

struct FaParams {
    const int* row0;
    const int* len;
    bf16* o;
    int ldo;
    int group;  // query heads per kv head
    int o32;    // output rows are 32-byte aligned: 256-bit stores
    float scale_log2;
};

// NQ = query tiles per CTA that share the K / V blocks: the two query heads of a kv head (GQA group 2) are processed
// side by side, each by its own group of four softmax warps with its own score / output regions in TMEM, so the tensor
// core works on one tile while the softmax warps work on the other (and K / V are fetched once for both).
template <int HD, int NQ>
struct FaCfg {
    static constexpr int SUBS = HD / 64;                   // 64-column (128-byte) sub-tiles along head_dim
    static constexpr int KV_STAGES = (HD == 128 && NQ == 1) ? 2 : 1;
    static constexpr int S_BUFS = (HD == 128 && NQ == 1) ? 2 : 1;  // score buffers per tile
    static constexpr int Q_BYTES = FA_BQ * HD * 2;
    static constexpr int KV_BYTES = FA_BKV * HD * 2;       // one K (or V) block
    static constexpr int P_BYTES = FA_BQ * FA_BKV * 2;
    static constexpr int TMEM_NEED = NQ * (S_BUFS * FA_BKV + HD);
    static constexpr int TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
    static constexpr int THREADS = 128 + 128 * NQ;
    static constexpr int SMEM = NQ * Q_BYTES + 2 * KV_STAGES * KV_BYTES + NQ * P_BYTES + 1024 + 256;
};

// MN-major, 128-byte-swizzled operand descriptor: rows of 64 contiguous MN elements (128 B) per K index, 8 K indices per
// 1024-byte swizzle atom (SBO), the next 64 MN elements `lbo_bytes` further on.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// 2^x on the SFU (ex2.approx: 2 ulp; the result is rounded to bf16 or only scales fp32 partial sums)
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD, bool CAUSAL, int NQ>
__global__ void __launch_bounds__(FaCfg<HD, NQ>::THREADS, (HD == 64 && NQ == 1) ? 2 : 1)
fa_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
             const FaParams p) {
    using C = FaCfg<HD, NQ>;
    // NQ == 1: blockIdx.y = query head.  NQ == 2: blockIdx.y = kv head, tiles = its two query heads.
    const int seg = blockIdx.z;
    const int head0 = NQ == 1 ? blockIdx.y : blockIdx.y * NQ;
    const int len = p.len[seg];
    // causal: the last query tile of a segment sees the most key blocks, so tiles are handed out heaviest first
    const int q0 = (CAUSAL ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * FA_BQ;
    if (q0 >= len) return;
    const int row0 = p.row0[seg];
    const int kvh = head0 / p.group;
    const int kv_end = CAUSAL ? min(len, q0 + FA_BQ) : len;
    const int nb = (kv_end + FA_BKV - 1) / FA_BKV;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                   // [NQ] tiles
    uint8_t* sK = sQ + NQ * C::Q_BYTES;
    uint8_t* sV = sK + C::KV_STAGES * C::KV_BYTES;
    uint8_t* sP = sV + C::KV_STAGES * C::KV_BYTES;        // [NQ] tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NQ * C::P_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = q_full + 1;
    uint64_t* k_empty = k_full + C::KV_STAGES;
    uint64_t* v_full = k_empty + C::KV_STAGES;
    uint64_t* v_empty = v_full + C::KV_STAGES;
    uint64_t* s_full = v_empty + C::KV_STAGES;            // [NQ][S_BUFS]
    uint64_t* s_free = s_full + NQ * C::S_BUFS;           // [NQ][S_BUFS]
    uint64_t* p_full = s_free + NQ * C::S_BUFS;           // [NQ]
    uint64_t* pv_full = p_full + NQ;                      // [NQ]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + NQ);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        for (int s = 0; s < C::KV_STAGES; s++) {
            ptx::mbar_init(&k_full[s], 1);
            ptx::mbar_init(&k_empty[s], 1);
            ptx::mbar_init(&v_full[s], 1);
            ptx::mbar_init(&v_empty[s], 1);
        }
        for (int s = 0; s < NQ * C::S_BUFS; s++) {
            ptx::mbar_init(&s_full[s], 1);
            ptx::mbar_init(&s_free[s], 128);
        }
        for (int t = 0; t < NQ; t++) {
            ptx::mbar_init(&p_full[t], 128);
            ptx::mbar_init(&pv_full[t], 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<C::TMEM_COLS>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: tile t has S_BUFS score buffers at t * (S_BUFS * 128) and its P V region after all score buffers
    auto tmem_s = [&](int t, int sb) { return tmem_base + (uint32_t)((t * C::S_BUFS + sb) * FA_BKV); };
    auto tmem_o = [&](int t) { return tmem_base + (uint32_t)(NQ * C::S_BUFS * FA_BKV + t * HD); };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(q_full, NQ * C::Q_BYTES);
            for (int t = 0; t < NQ; t++)
                for (int s = 0; s < C::SUBS; s++)
                    ptx::tma_load_2d(sQ + t * C::Q_BYTES + s * FA_BQ * 128, &tmQ, (head0 + t) * HD + s * 64, row0 + q0, q_full);
            for (int j = 0; j < nb; j++) {
                const int st = j % C::KV_STAGES;
                const uint32_t ph = (j / C::KV_STAGES) & 1;
                ptx::mbar_wait(&k_empty[st], ph ^ 1);
                ptx::mbar_arrive_expect_tx(&k_full[st], C::KV_BYTES);
                for (int s = 0; s < C::SUBS; s++)
                    ptx::tma_load_2d(sK + st * C::KV_BYTES + s * FA_BKV * 128, &tmK, kvh * HD + s * 64, row0 + j * FA_BKV, &k_full[st]);
                ptx::mbar_wait(&v_empty[st], ph ^ 1);
                ptx::mbar_arrive_expect_tx(&v_full[st], C::KV_BYTES);
                for (int s = 0; s < C::SUBS; s++)
                    ptx::tma_load_2d(sV + st * C::KV_BYTES + s * FA_BKV * 128, &tmV, kvh * HD + s * 64, row0 + j * FA_BKV, &v_full[st]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t IDESC_S = ptx::umma_idesc_bf16(FA_BQ, FA_BKV);
            constexpr uint32_t IDESC_PV = ptx::umma_idesc_bf16(FA_BQ, HD) | (1u << 16);  // B (= V) is MN-major
            const uint32_t aQ = ptx::smem_u32(sQ), aP = ptx::smem_u32(sP);
            auto issue_s = [&](int t, int j) {  // scores of tile t against key block j (K_j already waited for)
                const int st = j % C::KV_STAGES, sb = j % C::S_BUFS;
                ptx::mbar_wait(&s_free[t * C::S_BUFS + sb], ((j / C::S_BUFS) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t aK = ptx::smem_u32(sK + st * C::KV_BYTES);
#pragma unroll
                for (int ks = 0; ks < HD / 16; ks++) {
                    const uint32_t off = (ks >> 2) * (128 * 128) + (ks & 3) * 32;
                    ptx::mma_bf16_ss(tmem_s(t, sb), ptx::umma_desc_sw128(aQ + t * C::Q_BYTES + off), ptx::umma_desc_sw128(aK + off), IDESC_S,
                                     ks > 0 ? 1u : 0u);
                }
                ptx::mma_commit(&s_full[t * C::S_BUFS + sb]);
            };
            auto issue_pv = [&](int t, int j) {  // P_t V_j (V_j already waited for)
                const int st = j % C::KV_STAGES;
                ptx::mbar_wait(&p_full[t], j & 1);
                ptx::tc_fence_after();
                const uint32_t aV = ptx::smem_u32(sV + st * C::KV_BYTES);
#pragma unroll
                for (int ks = 0; ks < FA_BKV / 16; ks++) {
                    const uint64_t da = ptx::umma_desc_sw128(aP + t * C::P_BYTES + (ks >> 2) * (128 * 128) + (ks & 3) * 32);
                    const uint64_t db = umma_desc_mn_sw128(aV + ks * 16 * 128, FA_BKV * 128);
                    ptx::mma_bf16_ss(tmem_o(t), da, db, IDESC_PV, ks > 0 ? 1u : 0u);
                }
                ptx::mma_commit(&pv_full[t]);
            };
            ptx::mbar_wait(q_full, 0);
            ptx::mbar_wait(&k_full[0], 0);
            for (int t = 0; t < NQ; t++) issue_s(t, 0);
            ptx::mma_commit(&k_empty[0]);
            for (int j = 0; j < nb; j++) {
                const int st = j % C::KV_STAGES;
                const bool more = j + 1 < nb;
                if (more) ptx::mbar_wait(&k_full[(j + 1) % C::KV_STAGES], ((j + 1) / C::KV_STAGES) & 1);
                ptx::mbar_wait(&v_full[st], (j / C::KV_STAGES) & 1);
                // per tile: this block's P V, then the next block's scores, so each softmax group always has work queued
                for (int t = 0; t < NQ; t++) {
                    if (NQ == 1 && more) issue_s(t, j + 1);  // double-buffered scores: run ahead of the softmax
                    issue_pv(t, j);
                    if (NQ > 1 && more) issue_s(t, j + 1);
                }
                ptx::mma_commit(&v_empty[st]);
                if (more) ptx::mma_commit(&k_empty[(j + 1) % C::KV_STAGES]);
            }
        }
    } else if (warp >= 4) {
        // ===== softmax / output: one thread per query row =====
        const int q = warp & 3;
        const int tq = (warp - 4) >> 2;  // which query tile this softmax group owns
        const int head = head0 + tq;
        const int r = q * 32 + lane;
        const int q_abs = q0 + r;
        const uint32_t lane_addr = uint32_t(q * 32) << 16;
        const uint32_t tmem_pv = tmem_o(tq);
        // the running output as (even, odd) pairs: the rescale-and-add of a block and the score scaling are packed fp32 operations
        // (FADD2 / FMUL2 / FFMA2: the softmax warps are bound by their own instruction stream, ~10 per score)
        float2 o_acc[HD / 2];
#pragma unroll
        for (int i = 0; i < HD / 2; i++) o_acc[i] = make_float2(0.f, 0.f);
        float m_run = -INFINITY, l_run = 0.f;
        uint8_t* p_row = sP + tq * C::P_BYTES + r * 128;
        for (int j = 0; j < nb; j++) {
            const int sb = j % C::S_BUFS;
            const int k0 = j * FA_BKV;
            ptx::mbar_wait(&s_full[tq * C::S_BUFS + sb], (j / C::S_BUFS) & 1);
            ptx::tc_fence_after();
            const uint32_t t_s = tmem_s(tq, sb) + lane_addr;
            // columns this row may see in this block: keys < len, and <= its own position when causal
            int vis = len - k0;
            if (CAUSAL) vis = min(vis, q_abs - k0 + 1);
            vis = min(vis, FA_BKV);  // rows past the segment end (garbage rows) still see >= 1 key, so the max stays finite
            // TMEM loads are software-pipelined when this group is alone on the tensor core's results (NQ == 1): the load of
            // chunk c + 1 is in flight while chunk c is processed.  With two groups the other group hides the latency.
            constexpr bool PIPE = NQ == 1 && HD == 128;  // head_dim 64 runs two CTAs per SM at 128 registers: no room, no need
            constexpr int NC = FA_BKV / 32;
            uint32_t v[PIPE ? 2 : 1][32];
            // pass 1: row maximum (only the chunk that straddles `vis` needs per-element masking)
            float mx = -INFINITY;
            if (PIPE) ptx::tmem_ld_32x32(t_s, v[0]);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                uint32_t(&cur)[32] = v[PIPE ? (c & 1) : 0];
                if (!PIPE) ptx::tmem_ld_32x32(t_s + c * 32, cur);
                ptx::tmem_ld_wait();
                if (PIPE && c + 1 < NC) ptx::tmem_ld_32x32(t_s + (c + 1) * 32, v[(c + 1) & 1]);
                if (c * 32 + 32 <= vis) {
#pragma unroll
                    for (int i = 0; i < 32; i++) mx = fmaxf(mx, __uint_as_float(cur[i]));
                } else if (c * 32 < vis) {
#pragma unroll
                    for (int i = 0; i < 32; i++)
                        if (c * 32 + i < vis) mx = fmaxf(mx, __uint_as_float(cur[i]));
                }
            }
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float alpha = fast_exp2(m_run - m_new);
            m_run = m_new;
            // fold in the previous block's P V (this also guarantees the tensor core is done reading the P buffer)
            if (j > 0) {
                ptx::mbar_wait(&pv_full[tq], (j - 1) & 1);
                ptx::tc_fence_after();
                if (PIPE) ptx::tmem_ld_32x32(tmem_pv + lane_addr, v[0]);
#pragma unroll
                for (int c = 0; c < HD / 32; c++) {
                    uint32_t(&cur)[32] = v[PIPE ? (c & 1) : 0];
                    if (!PIPE) ptx::tmem_ld_32x32(tmem_pv + lane_addr + c * 32, cur);
                    ptx::tmem_ld_wait();
                    if (PIPE && c + 1 < HD / 32) ptx::tmem_ld_32x32(tmem_pv + lane_addr + (c + 1) * 32, v[(c + 1) & 1]);
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        o_acc[c * 16 + i] = f2_mul(f2_add(o_acc[c * 16 + i], make_float2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1]))),
                                                   make_float2(alpha, alpha));
                }
            }
            // pass 2: P = exp2(s * scale - m), rounded to bf16 into the swizzled A-operand tile; row sum unrounded
            float rs = 0.f;
            if (PIPE) ptx::tmem_ld_32x32(t_s, v[0]);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                uint32_t(&cur)[32] = v[PIPE ? (c & 1) : 0];
                if (!PIPE) ptx::tmem_ld_32x32(t_s + c * 32, cur);
                ptx::tmem_ld_wait();
                if (PIPE && c + 1 < NC) ptx::tmem_ld_32x32(t_s + (c + 1) * 32, v[(c + 1) & 1]);
                uint32_t pk[16];
                if (c * 32 + 32 <= vis) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float2 sc = f2_fma(make_float2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])),
                                                 make_float2(p.scale_log2, p.scale_log2), make_float2(-m_new, -m_new));
                        const float a = fast_exp2(sc.x), b = fast_exp2(sc.y);
                        rs += a + b;
                        pk[i] = pack_bf16x2(a, b);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float a = c * 32 + 2 * i < vis ? fast_exp2(fmaf(__uint_as_float(cur[2 * i]), p.scale_log2, -m_new)) : 0.f;
                        const float b = c * 32 + 2 * i + 1 < vis ? fast_exp2(fmaf(__uint_as_float(cur[2 * i + 1]), p.scale_log2, -m_new)) : 0.f;
                        rs += a + b;
                        pk[i] = pack_bf16x2(a, b);
                    }
                }
                // 32 keys = four 16-byte chunks; chunk index within the 64-key sub-tile is XOR-swizzled with the row
                uint8_t* sub = p_row + (c >> 1) * (128 * 128);
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    const int c8 = (c & 1) * 4 + ch;
                    *reinterpret_cast<uint4*>(sub + ((c8 ^ (r & 7)) << 4)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
                }
            }
            l_run = l_run * alpha + rs;
            ptx::tc_fence_before();
            ptx::mbar_arrive(&s_free[tq * C::S_BUFS + sb]);  // score buffer may be overwritten
            ptx::fence_proxy_async();                        // P stores visible to the tensor core
            ptx::mbar_arrive(&p_full[tq]);
        }
        ptx::mbar_wait(&pv_full[tq], (nb - 1) & 1);
        ptx::tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        bf16* orow = p.o + (size_t)(row0 + q_abs) * p.ldo + head * HD;
#pragma unroll
        for (int c = 0; c < HD / 16; c++) {
            uint32_t v[16];
            ptx::tmem_ld_32x16(tmem_pv + lane_addr + c * 16, v);
            ptx::tmem_ld_wait();
            if (q_abs < len) {
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; e++)
                    w[e] = pack_bf16x2((o_acc[c * 8 + e].x + __uint_as_float(v[2 * e])) * inv, (o_acc[c * 8 + e].y + __uint_as_float(v[2 * e + 1])) * inv);
                if (p.o32) {  // one 256-bit store per 16 dims (a thread owns a row: 32 lines per warp instruction either way)
                    st_global_v8(orow + c * 16, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
                } else {
                    *reinterpret_cast<uint4*>(orow + c * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(orow + c * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                }
            }
        }
        ptx::tc_fence_before();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

template <int HD, bool CAUSAL, int NQ>
void launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const FaParams& p, const AttnSegs& segs, int heads,
            cudaStream_t st) {
    using C = FaCfg<HD, NQ>;
    static PerDeviceOnce attr_once;
    attr_once([] {
        Q3_CUDA(cudaFuncSetAttribute(fa_tc_kernel<HD, CAUSAL, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    });
    dim3 grid((segs.max_len + FA_BQ - 1) / FA_BQ, heads / NQ, segs.n_segs);
    fa_tc_kernel<HD, CAUSAL, NQ><<<grid, C::THREADS, C::SMEM, st>>>(tq, tk, tv, p);
    Q3_CUDA(cudaGetLastError());
}

}  // namespace

void flash_attn_tc_launch(const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* o, int ldo, const AttnSegs& segs,
                          int total_rows, int heads, int group, int head_dim, bool causal, float scale, cudaStream_t st) {
    if (segs.n_segs <= 0 || segs.max_len <= 0) return;
    Q3_CHECK(segs.n_segs <= 65535 && heads <= 65535, 1, "attention: too many segments for one launch");
    Q3_CHECK(head_dim == 64 || head_dim == 128, 1, "attention: head_dim must be 64 or 128");
    Q3_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, 1, "attention: leading dimensions must be multiples of 8");
    const int kv_heads = heads / group;
    CUtensorMap tq, tk, tv;
    make_tmap_2d_bf16(&tq, q, (uint64_t)heads * head_dim, total_rows, ldq, 64, FA_BQ);
    make_tmap_2d_bf16(&tk, k, (uint64_t)kv_heads * head_dim, total_rows, ldk, 64, FA_BKV);
    make_tmap_2d_bf16(&tv, v, (uint64_t)kv_heads * head_dim, total_rows, ldv, 64, FA_BKV);
    FaParams p;
    p.row0 = segs.row0;
    p.len = segs.len;
    p.o = o;
    p.ldo = ldo;
    p.group = group;
    p.o32 = (reinterpret_cast<uintptr_t>(o) & 31) == 0 && ldo % 16 == 0 && head_dim % 16 == 0;
    p.scale_log2 = scale * 1.4426950408889634f;
    static const bool no_pair = getenv("Q3ASR_ATTN_NO_PAIR") != nullptr && atoi(getenv("Q3ASR_ATTN_NO_PAIR")) != 0;
    const bool pair = group == 2 && head_dim == 128 && !no_pair;  // GQA: both query heads of a kv head in one CTA
    if (head_dim == 64 && !causal) launch<64, false, 1>(tq, tk, tv, p, segs, heads, st);
    else if (head_dim == 64 && causal) launch<64, true, 1>(tq, tk, tv, p, segs, heads, st);
    else if (!causal && pair) launch<128, false, 2>(tq, tk, tv, p, segs, heads, st);
    else if (!causal) launch<128, false, 1>(tq, tk, tv, p, segs, heads, st);
    else if (pair) launch<128, true, 2>(tq, tk, tv, p, segs, heads, st);
    else launch<128, true, 1>(tq, tk, tv, p, segs, heads, st);
}

}  // namespace q3
