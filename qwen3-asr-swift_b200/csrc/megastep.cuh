// megastep.cuh — the decoder layers of ONE decode step as a single persistent cooperative kernel.
//
// Replaces, for the decode step (FloatTextDecoder.swift:35-226 with seqLen 1; loop Qwen3ASR.swift:344-389), the chain of
// 7 dependent launches per layer (skinny.cuh / ops.cu decode_attn_mma_kernel / reduce_resid_rmsnorm_kernel / gemm.cuh SwiGLU)
// by one grid of num_sms CTAs that walks a host-built table of phases:
//
//     per layer:  QKV -> ATTN -> O -> NORM1 -> GATE|UP -> DOWN -> NORM2        (x n_sub sub-batches, interleaved)
//
// Why: at 64 sequences every phase of the chain is latency-bound (launch hand-over, one L2 round trip for the activations, TMA,
// MMA, TMEM read-back, store: ~2.7 us fixed each), so the step ran at 0.46 of its HBM floor.  Here
//   * the batch is split into two sub-batches whose phases alternate (A.qkv, B.qkv, A.attn, B.attn, ...): a phase only depends
//     on the SAME sub-batch's previous phase, which finished a whole phase earlier, so the grid-wide hand-over (an arrival
//     counter per phase in global memory) is almost never waited on — the dependent latency of one sub-batch is hidden under
//     the other's work;
//   * weight tiles do not depend on activations: the TMA weight producer runs ahead of the phase it feeds (through the barrier
//     waits and the norm phases), so the ring is full when a phase's activations arrive; the second sub-batch re-reads the same
//     tiles from L2;
//   * there is no launch inside a step's layers any more: 1 + 28 * 7 + 3 launches become 5.
// Arithmetic is IDENTICAL to the multi-kernel path (same split-K partition, same MMA k order, same fixed-order reductions, the
// same canonical key streams in the attention), so the two paths produce bit-identical ids; tests/test_gpu_model.py checks it.
//
// CTA = 384 threads: warp 0 weight producer (TMA), warp 1 activation producer (TMA, after the dependency), warp 2 MMA issuer,
// warp 3 TMEM allocator, warps 4-11 workers (4-7 are also the GEMM epilogue: one TMEM lane quadrant each).  Attention and the
// norms run on the 8 worker warps.  The GEMM stage ring and the attention's per-warp cp.async rings share one 192 KB region
// (a CTA is in one kind of phase at a time; the producers are gated on the attention phases of their CTA).
#pragma once
#include "gemm.cuh"
#include "megastep_params.h"
#include "ops.cuh"
#include "skinny.cuh"

namespace q3 {

constexpr int MEGA_THREADS = 384;
constexpr int MEGA_RING_BYTES = 8 * 3 * 2 * 16 * 128 * 2;  // 8 worker warps x 3 stages x (K + V) x 16 keys x 128 dims x bf16 = 192 KB
__host__ __device__ constexpr int mega_stage_bytes(int NB, int GU_BN) { return 16384 + (NB > GU_BN ? NB : GU_BN) * 128; }
__host__ __device__ constexpr int mega_stages(int NB, int GU_BN) { return MEGA_RING_BYTES / mega_stage_bytes(NB, GU_BN); }
constexpr int MEGA_SMALL_BYTES = 2 * 8 * 128 * 4 /*s_acc*/ + 4 * 2 * 128 * 2 * 2 /*s_qb, s_new*/ + 2 * 2 * 8 * 4 /*s_m, s_l*/ + 64 * 4 /*s_red*/ + 512 /*barriers*/;
constexpr int mega_smem_bytes() { return MEGA_RING_BYTES + MEGA_SMALL_BYTES + 1024; }

namespace mega {

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* smem_ptr) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* smem_ptr) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int dm_off(int key, int seg) { return key * 128 + ((seg ^ (key & 7)) << 3); }
__device__ __forceinline__ float merge_scales(float am, float bm, float* fa, float* fb) {
    const float m = fmaxf(am, bm);
    *fa = am == -INFINITY ? 0.f : exp2f(am - m);
    *fb = bm == -INFINITY ? 0.f : exp2f(bm - m);
    return m;
}

// Waits until every CTA has arrived at phase `dep` (its outputs are then visible: the arrival is a release after a CTA-wide
// barrier, this is the acquire).  One polling thread per warp; the others are released by __syncwarp.
__device__ __forceinline__ void wait_phase(const MegaParams& P, int dep, int lane) {
    if (dep >= 0) {
        if (lane == 0) {
            unsigned spins = 0;
            while (ld_acquire(P.cnt + dep) < (unsigned)P.G) {
                __nanosleep(40);
                if (++spins > (1u << 24)) {  // ~1 s: a broken phase table; fail loudly instead of hanging the GPU
                    atomicExch(P.err, dep + 1);
                    __trap();
                }
            }
        }
        __syncwarp();
    }
}

struct AttnShared {
    bf16 (*qb)[128];   // [2] query heads of this group
    bf16 (*nw)[128];   // [2] the new token's k and v rows
    float* m;          // s_m[head][8 worker warps]
    float* l;
    float* acc;        // s_acc[head][8 worker warps][128]
};

constexpr int DM_CHUNK = 16, DM_STAGES = 3, DM_STREAMS = 8;

// One (sequence, kv head) item on a group of `nw` warps (2, 4 or 8): the body of decode_attn_mma_kernel (ops.cu) with the CTA
// replaced by the warp group.  gw: warp within the group; ww: worker warp index in the CTA (0..7) — it selects the cp.async ring
// and the slots of the merge arrays.  The key partition is the canonical one (chunk c -> stream c % 8, streams merged by a fixed
// tree), so the result does not depend on nw.
__device__ __forceinline__ void attention_item(const MegaParams& P, int layer, int sub, int seq_in_sub, int kvh, int gw, int nw, int ww0, int lane,
                                               bf16* ring, const AttnShared& S, int bar_id) {
    constexpr int GROUP = 2;
    const int gthreads = nw * 32, gt = gw * 32 + lane;
    const int seq = P.sub[sub].row0 + seq_in_sub;
    const KvCache& cache = P.cache;
    const int* pt = cache.page_table + (size_t)seq * cache.max_pages;
    const int len = P.kv_len[seq];
    const int heads = P.heads, nqkv = P.nqkv;
    const size_t head_off = (((size_t)layer * 2) * cache.kv_heads + kvh) * (KV_PAGE * 128);
    const size_t page_elems = (size_t)cache.layers * 2 * cache.kv_heads * (KV_PAGE * 128);
    const size_t v_off = (size_t)cache.kv_heads * (KV_PAGE * 128);
    const int n_chunks = (len + DM_CHUNK - 1) / DM_CHUNK;
    auto issue = [&](int chunk, int stage) {
        if (chunk < n_chunks) {
            const int j0 = chunk * DM_CHUNK;
            const bf16* kb = cache.pool + (size_t)pt[j0 / KV_PAGE] * page_elems + head_off + (j0 % KV_PAGE) * 128;
            bf16* sk = ring + (size_t)stage * 2 * DM_CHUNK * 128;
#pragma unroll
            for (int i = 0; i < DM_CHUNK * 16 / 32; i++) {
                const int lin = i * 32 + lane, key = lin >> 4, seg = lin & 15;
                const uint32_t n = j0 + key < len ? 16u : 0u;
                cp_async16(sk + dm_off(key, seg), kb + lin * 8, n);
                cp_async16(sk + DM_CHUNK * 128 + dm_off(key, seg), kb + v_off + lin * 8, n);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto seq_next = [&](int& sj, int& sc) {
        sc += DM_STREAMS;
        while (sc >= n_chunks && sj + nw < DM_STREAMS) {
            sj += nw;
            sc = sj;
        }
    };
    int ij = gw, ic = gw;
    while (ic >= n_chunks && ij + nw < DM_STREAMS) { ij += nw; ic = ij; }
#pragma unroll
    for (int s = 0; s < DM_STAGES - 1; s++) {
        issue(ic, s);
        seq_next(ij, ic);
    }

    // ---- 1. new-token q / k / v from the split-K partials of this sub-batch's QKV phase (written by other CTAs: L2 loads) ----
    if (gt < 16 * (GROUP + 2)) {
        const int slot = gt >> 4, hl = gt & 15;
        const unsigned half_mask = 0xffffu << (gt & 16);
        const int d0 = hl * 8;
        const int col = slot < GROUP ? (kvh * GROUP + slot) * 128
                                     : slot == GROUP ? heads * 128 + kvh * 128 : (heads + cache.kv_heads) * 128 + kvh * 128;
        const float* src = P.ws[sub] + (size_t)seq_in_sub * nqkv + col + d0;
        const long long split_stride = (long long)P.sub[sub].rows * nqkv;
        const int splits = P.g[0].splits;
        float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int s0 = 0; s0 < splits; s0 += 4) {
            float4 b[4][2];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const bool on = s0 + i < splits;
                const float4* sp = reinterpret_cast<const float4*>(src + (size_t)(s0 + i) * split_stride);
                b[i][0] = on ? __ldcg(sp) : make_float4(0.f, 0.f, 0.f, 0.f);
                b[i][1] = on ? __ldcg(sp + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                x[0] += b[i][0].x; x[1] += b[i][0].y; x[2] += b[i][0].z; x[3] += b[i][0].w;
                x[4] += b[i][1].x; x[5] += b[i][1].y; x[6] += b[i][1].z; x[7] += b[i][1].w;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = bf16_round(x[j]);
        const int p = P.pos[seq];
        if (slot <= GROUP) {
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) q = fmaf(x[j], x[j], q);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(half_mask, q, o);
            const float r = rsqrtf(q * (1.0f / 128.0f) + P.eps);
            const bf16* nwp = P.norm_w[layer * 4 + (slot < GROUP ? 0 : 1)];
            const uint4 wu = *reinterpret_cast<const uint4*>(nwp + d0);
            const uint32_t ww[4] = {wu.x, wu.y, wu.z, wu.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 wf = unpack_bf16x2(ww[j]);
                x[2 * j] = bf16_round(x[2 * j] * r * wf.x);
                x[2 * j + 1] = bf16_round(x[2 * j + 1] * r * wf.y);
            }
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; j++) y[j] = __shfl_xor_sync(half_mask, x[j], 8);
            const float sgn = hl < 8 ? -1.f : 1.f;
            const float4* tp = reinterpret_cast<const float4*>(P.rope_tab + (size_t)p * 64 + (d0 & 63));
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 t = __ldg(tp + j);
                x[2 * j] = fmaf(x[2 * j], t.x, sgn * y[2 * j] * t.y);
                x[2 * j + 1] = fmaf(x[2 * j + 1], t.z, sgn * y[2 * j + 1] * t.w);
            }
        }
        const uint4 packed = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
        if (slot < GROUP) {
            *reinterpret_cast<uint4*>(&S.qb[slot][d0]) = packed;
        } else {
            const int page = pt[p / KV_PAGE];
            bf16* dst = cache.pool + ((((size_t)page * cache.layers + layer) * 2 + (slot == GROUP ? 0 : 1)) * cache.kv_heads + kvh) * (KV_PAGE * 128) +
                        (p % KV_PAGE) * 128 + d0;
            *reinterpret_cast<uint4*>(dst) = packed;
            *reinterpret_cast<uint4*>(&S.nw[slot - GROUP][d0]) = packed;
        }
    }
    bar_sync(bar_id, gthreads);

    // ---- 2. attention on the tensor cores (mma.sync: two query rows do not fill a UMMA tile) ----
    const int g = lane >> 2, t = lane & 3;
    uint32_t qa[8][2];
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
        qa[ks][0] = lane < 8 ? *reinterpret_cast<const uint32_t*>(&S.qb[g & 1][ks * 16 + t * 2]) : 0u;
        qa[ks][1] = lane < 8 ? *reinterpret_cast<const uint32_t*>(&S.qb[g & 1][ks * 16 + 8 + t * 2]) : 0u;
    }
    float o[16][4];
    float m_run = -INFINITY, l_run = 0.f;
    const int mi = lane >> 3, r8 = lane & 7;
    int stage = 0;
    int cj = gw, chunk = gw;
    while (chunk >= n_chunks && cj + nw < DM_STREAMS) { cj += nw; chunk = cj; }
    int cur_stream = -1;
    bool first_stream = true;
    const int ww = ww0 + gw;  // slot of this warp in the merge arrays
    float* s_m0 = S.m;        // [head][8]
    float* s_l0 = S.l;
    float* s_acc0 = S.acc;    // [head][8][128]
    auto flush_stream = [&]() {
        float l = l_run;
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        if (lane < 8) {
            const float am = first_stream ? -INFINITY : s_m0[g * 8 + ww], al = first_stream ? 0.f : s_l0[g * 8 + ww];
            float fa, fb;
            const float mm = merge_scales(am, m_run, &fa, &fb);
#pragma unroll
            for (int nt = 0; nt < 16; nt++) {
                float* dst = &s_acc0[(g * 8 + ww) * 128 + nt * 8 + t * 2];
                const float a0 = first_stream ? 0.f : dst[0], a1 = first_stream ? 0.f : dst[1];
                dst[0] = fmaf(fb, o[nt][0], fa * a0);
                dst[1] = fmaf(fb, o[nt][1], fa * a1);
            }
            __syncwarp(0xffu);
            if (t == 0) { s_m0[g * 8 + ww] = mm; s_l0[g * 8 + ww] = fmaf(fb, l, fa * al); }
        }
        first_stream = false;
    };
    for (; chunk < n_chunks; seq_next(cj, chunk)) {
        if (cj != cur_stream) {
            if (cur_stream >= 0) flush_stream();
            cur_stream = cj;
#pragma unroll
            for (int i = 0; i < 16; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
            m_run = -INFINITY;
            l_run = 0.f;
        }
        issue(ic, (stage + DM_STAGES - 1) % DM_STAGES);
        seq_next(ij, ic);
        asm volatile("cp.async.wait_group %0;" ::"n"(DM_STAGES - 1) : "memory");
        __syncwarp();
        bf16* sk = ring + (size_t)stage * 2 * DM_CHUNK * 128;
        bf16* sv = sk + DM_CHUNK * 128;
        const int j0 = chunk * DM_CHUNK;
        if (chunk == n_chunks - 1) {  // the prefetch may have read the new token's row before it was written
            const int row = (len - 1) - j0;
            *reinterpret_cast<uint4*>((lane < 16 ? sk : sv) + dm_off(row, lane & 15)) = reinterpret_cast<const uint4*>(S.nw[lane >> 4])[lane & 15];
            __syncwarp();
        }
        float s[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            uint32_t b[4];
            const int key = (mi >> 1) * 8 + r8;
            ldsm_x4(b, sk + dm_off(key, ks * 2 + (mi & 1)));
            const uint32_t a[4] = {qa[ks][0], 0u, qa[ks][1], 0u};
            mma16816(s[0], a, b[0], b[1]);
            mma16816(s[1], a, b[2], b[3]);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const bool ok = j0 + nt * 8 + t * 2 + e < len;
                s[nt][e] = ok ? s[nt][e] * P.scale_log2 : -INFINITY;
                mx = fmaxf(mx, s[nt][e]);
            }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float mn = fmaxf(m_run, mx);
        const float alpha = exp2f(m_run - mn);
        m_run = mn;
        const float p00 = exp2f(s[0][0] - mn), p01 = exp2f(s[0][1] - mn), p10 = exp2f(s[1][0] - mn), p11 = exp2f(s[1][1] - mn);
        l_run = l_run * alpha + ((p00 + p01) + (p10 + p11));
        const uint32_t pa[4] = {pack_bf16x2(p00, p01), 0u, pack_bf16x2(p10, p11), 0u};
#pragma unroll
        for (int i = 0; i < 16; i++) { o[i][0] *= alpha; o[i][1] *= alpha; }
#pragma unroll
        for (int np = 0; np < 8; np++) {
            uint32_t b[4];
            const int key = r8 + (mi & 1) * 8;
            ldsm_x4_t(b, sv + dm_off(key, np * 2 + (mi >> 1)));
            mma16816(o[2 * np], pa, b[0], b[1]);
            mma16816(o[2 * np + 1], pa, b[2], b[3]);
        }
        __syncwarp();
        stage = (stage + 1) % DM_STAGES;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    // ---- 3. merge the streams: per warp in order, across warps by the fixed tree (even chain, odd chain, even + odd) ----
    if (cur_stream >= 0) {
        flush_stream();
    } else if (lane < 8) {
        if (t == 0) { s_m0[g * 8 + ww] = -INFINITY; s_l0[g * 8 + ww] = 0.f; }
#pragma unroll
        for (int nt = 0; nt < 16; nt++) s_acc0[(g * 8 + ww) * 128 + nt * 8 + t * 2] = s_acc0[(g * 8 + ww) * 128 + nt * 8 + t * 2 + 1] = 0.f;
    }
    bar_sync(bar_id, gthreads);
    for (int i = gt; i < GROUP * 128; i += gthreads) {
        const int gg = i >> 7, d = i & 127;
        float cm[2] = {-INFINITY, -INFINITY}, cl[2] = {0.f, 0.f}, ca[2] = {0.f, 0.f};
#pragma unroll 1
        for (int w = 0; w < nw; w++) {
            float fa, fb;
            const float mm = merge_scales(cm[w & 1], s_m0[gg * 8 + ww0 + w], &fa, &fb);
            ca[w & 1] = fmaf(fb, s_acc0[(gg * 8 + ww0 + w) * 128 + d], fa * ca[w & 1]);
            cl[w & 1] = fmaf(fb, s_l0[gg * 8 + ww0 + w], fa * cl[w & 1]);
            cm[w & 1] = mm;
        }
        float fa, fb;
        merge_scales(cm[0], cm[1], &fa, &fb);
        const float num = fmaf(fb, ca[1], fa * ca[0]), den = fmaf(fb, cl[1], fa * cl[0]);
        P.att[((size_t)seq * heads + kvh * GROUP + gg) * 128 + d] = __float2bfloat16_rn(num / den);
    }
    bar_sync(bar_id, gthreads);  // the merge arrays and s_qb / s_new are reused by the group's next item
}

// One row of a reduce + residual + RMSNorm phase on the 256 worker threads: the arithmetic of reduce_resid_rmsnorm_kernel
// (ops.cu) in the same order — SG partial sums of the splits (4 at a time), added in order; x += bf16(sum); y = x * r * w.
__device__ __forceinline__ void norm_row(const float* part, int splits, long long split_stride, int sg_count, bf16* x, const bf16* w, bf16* y,
                                         int d, float eps, int wt, float* s_red) {
    const int ncg = d >> 2, per = ncg >> 8;  // column groups of 4; 256 threads take `per` of them each (d = 1024: 1, 2048: 2)
    float v[2][4];
    for (int j = 0; j < per; j++) {
        const int cg = wt + 256 * j, c0 = cg * 4;
        const float* src = part + c0;
        float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sg = 0; sg < sg_count; sg++) {
            float4 b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int sp = sg + i * sg_count;
                b[i] = sp < splits ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)sp * split_stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 a = b[0];
#pragma unroll
            for (int i = 1; i < 4; i++) { a.x += b[i].x; a.y += b[i].y; a.z += b[i].z; a.w += b[i].w; }
            for (int sp = sg + 4 * sg_count; sp < splits; sp += sg_count) {
                const float4 t = __ldcg(reinterpret_cast<const float4*>(src + (size_t)sp * split_stride));
                a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
            }
            if (sg == 0) tot = a;
            else { tot.x += a.x; tot.y += a.y; tot.z += a.z; tot.w += a.w; }
        }
        const uint2 u = __ldcg(reinterpret_cast<const uint2*>(x + c0));
        const float2 x0 = unpack_bf16x2(u.x), x1 = unpack_bf16x2(u.y);
        v[j][0] = bf16_round(x0.x + bf16_round(tot.x));
        v[j][1] = bf16_round(x0.y + bf16_round(tot.y));
        v[j][2] = bf16_round(x1.x + bf16_round(tot.z));
        v[j][3] = bf16_round(x1.y + bf16_round(tot.w));
        *reinterpret_cast<uint2*>(x + c0) = make_uint2(pack_bf16x2(v[j][0], v[j][1]), pack_bf16x2(v[j][2], v[j][3]));
        float q = fmaf(v[j][0], v[j][0], fmaf(v[j][1], v[j][1], fmaf(v[j][2], v[j][2], v[j][3] * v[j][3])));
        q = warp_sum(q);
        if ((wt & 31) == 0) s_red[(wt >> 5) + 8 * j] = q;
    }
    bar_sync(1, 256);
    float tot = 0.f;
    for (int i = 0; i < (ncg >> 5); i++) tot += s_red[i];
    const float r = rsqrtf(tot / (float)d + eps);
    for (int j = 0; j < per; j++) {
        const int c0 = (wt + 256 * j) * 4;
        const uint2 wu = *reinterpret_cast<const uint2*>(w + c0);
        const float2 w0 = unpack_bf16x2(wu.x), w1 = unpack_bf16x2(wu.y);
        *reinterpret_cast<uint2*>(y + c0) =
            make_uint2(pack_bf16x2(v[j][0] * r * w0.x, v[j][1] * r * w0.y), pack_bf16x2(v[j][2] * r * w1.x, v[j][3] * r * w1.y));
    }
    bar_sync(1, 256);  // s_red is reused by the next row
}

}  // namespace mega

template <int NB, int GU_BN>
__global__ void __launch_bounds__(MEGA_THREADS, 1) megastep_kernel(const __grid_constant__ MegaParams P) {
    using namespace mega;
    constexpr int STAGE_BYTES = mega_stage_bytes(NB, GU_BN);
    constexpr int STAGES = mega_stages(NB, GU_BN);
    constexpr int ACC_COLS = 128;  // TMEM columns per accumulator (two accumulators)
    constexpr uint32_t IDESC_SK = ptx::umma_idesc_bf16(128, NB);
    constexpr uint32_t IDESC_GU = ptx::umma_idesc_bf16(128, GU_BN);
    static_assert(NB == 16 || NB == 32 || NB == 64 || NB == 128, "NB");
    static_assert(GU_BN == 64 || GU_BN == 128, "GU_BN");
    static_assert(STAGES >= 4 && STAGES <= 8, "ring depth");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* small = smem + MEGA_RING_BYTES;
    float* s_acc = reinterpret_cast<float*>(small);                              // [2][8][128]
    bf16(*s_qb)[2][128] = reinterpret_cast<bf16(*)[2][128]>(small + 8192);        // [4 groups][2][128]
    bf16(*s_new)[2][128] = reinterpret_cast<bf16(*)[2][128]>(small + 8192 + 2048);
    float* s_m = reinterpret_cast<float*>(small + 8192 + 4096);                   // [2][8]
    float* s_l = s_m + 16;
    float* s_red = s_l + 16;                                                      // [64]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(small + 8192 + 4096 + 128 + 256);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tfull_bar = empty_bar + 8;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    volatile int* attn_done = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x, G = P.G;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 2);   // weight producer + activation producer, each with its byte count
            ptx::mbar_init(&empty_bar[s], 1);  // the MMA warp's commit
        }
        for (int a = 0; a < 2; a++) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&tempty_bar[a], 4);  // one arrival per epilogue warp
        }
        *attn_done = 0;
        ptx::fence_barrier_init();
    }
    if (warp == 3) ptx::tmem_alloc<2 * ACC_COLS>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_wait();  // the step's first kernels (embedding gather, first RMSNorm) wrote x and xn

    auto first_unit = [&](int rot) { return ((cta - rot) % G + G) % G; };
    auto gemm_of = [&](int kind) -> const MegaGemm& { return P.g[kind == MK_QKV ? 0 : kind == MK_O ? 1 : kind == MK_GU ? 2 : 3]; };
    auto is_gemm = [](int kind) { return kind == MK_QKV || kind == MK_O || kind == MK_GU || kind == MK_DOWN; };

    if (warp == 0) {
        // ================= weight producer: runs ahead of the phases (weights depend on nothing) =================
        if (lane == 0) {
            int stage = 0;
            uint32_t parity = 0;
            for (int p = 0; p < P.n_phases; p++) {
                const MegaPhase ph = P.phases[p];
                if (!is_gemm(ph.kind)) continue;
                while (*attn_done < ph.attn_before) __nanosleep(20);  // the ring belongs to the attention until then
                const MegaGemm& gm = gemm_of(ph.kind);
                const int wj = ph.kind == MK_QKV ? 0 : ph.kind == MK_O ? 1 : ph.kind == MK_GU ? 2 : 3;
                const CUtensorMap* tm = P.maps + ph.layer * 4 + wj;
                for (int u = first_unit(ph.rot); u < gm.units; u += G) {
                    const int tile = u % gm.tiles_n, split = u / gm.tiles_n;
                    const int kb0 = split * gm.kb_per_split, kb1 = min(gm.num_kb, kb0 + gm.kb_per_split);
                    for (int kb = kb0; kb < kb1; kb++) {
                        ptx::mbar_wait(&empty_bar[stage], parity ^ 1);
                        uint8_t* st = smem + stage * STAGE_BYTES;
                        if (ph.kind == MK_GU) {
                            ptx::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)GU_BN * 128);
                            ptx::tma_load_2d(st + 16384, tm, kb * 64, tile * GU_BN, &full_bar[stage]);
                        } else {
                            ptx::mbar_arrive_expect_tx(&full_bar[stage], 16384u);
                            ptx::tma_load_2d(st, tm, kb * 64, tile * 128, &full_bar[stage]);
                        }
                        if (++stage == STAGES) { stage = 0; parity ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= activation producer: waits for the phase's dependency, then loads its tiles =================
        if (lane == 0) {
            int stage = 0;
            uint32_t parity = 0;
            for (int p = 0; p < P.n_phases; p++) {
                const MegaPhase ph = P.phases[p];
                if (!is_gemm(ph.kind)) continue;
                const MegaGemm& gm = gemm_of(ph.kind);
                if (first_unit(ph.rot) >= gm.units) continue;  // no unit of this phase on this CTA
                while (*attn_done < ph.attn_before) __nanosleep(20);
                if (ph.dep >= 0) {
                    unsigned spins = 0;
                    while (ld_acquire(P.cnt + ph.dep) < (unsigned)G) {
                        __nanosleep(40);
                        if (++spins > (1u << 24)) { atomicExch(P.err, ph.dep + 1); __trap(); }
                    }
                }
                fence_proxy_async_all();  // the tiles were written through the generic proxy of other SMs; TMA reads through the async proxy
                const int xj = ph.kind == MK_QKV ? 0 : ph.kind == MK_O ? 1 : ph.kind == MK_GU ? 2 : 3;
                const CUtensorMap* tm = P.maps + P.layers * 4 + ph.sub * 4 + xj;
                for (int u = first_unit(ph.rot); u < gm.units; u += G) {
                    const int split = u / gm.tiles_n;
                    const int kb0 = split * gm.kb_per_split, kb1 = min(gm.num_kb, kb0 + gm.kb_per_split);
                    for (int kb = kb0; kb < kb1; kb++) {
                        ptx::mbar_wait(&empty_bar[stage], parity ^ 1);
                        uint8_t* st = smem + stage * STAGE_BYTES;
                        if (ph.kind == MK_GU) {
                            ptx::mbar_arrive_expect_tx(&full_bar[stage], 16384u);
                            ptx::tma_load_2d(st, tm, kb * 64, 0, &full_bar[stage]);
                        } else {
                            ptx::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)NB * 128);
                            ptx::tma_load_2d(st + 16384, tm, kb * 64, 0, &full_bar[stage]);
                        }
                        if (++stage == STAGES) { stage = 0; parity ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t parity = 0, aparity[2] = {0, 0};
            for (int p = 0; p < P.n_phases; p++) {
                const MegaPhase ph = P.phases[p];
                if (!is_gemm(ph.kind)) continue;
                const MegaGemm& gm = gemm_of(ph.kind);
                const uint32_t idesc = ph.kind == MK_GU ? IDESC_GU : IDESC_SK;
                for (int u = first_unit(ph.rot); u < gm.units; u += G) {
                    const int split = u / gm.tiles_n;
                    const int kb0 = split * gm.kb_per_split, kb1 = min(gm.num_kb, kb0 + gm.kb_per_split);
                    ptx::mbar_wait(&tempty_bar[acc], aparity[acc] ^ 1);  // the epilogue has drained this accumulator
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                    for (int kb = kb0; kb < kb1; kb++) {
                        ptx::mbar_wait(&full_bar[stage], parity);
                        ptx::tc_fence_after();
                        const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
                        const uint32_t sb = sa + 16384;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            ptx::mma_bf16_ss(d_tmem, ptx::umma_desc_sw128(sa + k * 32), ptx::umma_desc_sw128(sb + k * 32), idesc,
                                             (kb > kb0 || k > 0) ? 1u : 0u);
                        ptx::mma_commit(&empty_bar[stage]);
                        if (++stage == STAGES) { stage = 0; parity ^= 1; }
                    }
                    ptx::mma_commit(&tfull_bar[acc]);
                    aparity[acc] ^= 1;
                    acc ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ================= workers: GEMM epilogues (warps 4-7), attention and norms (warps 4-11) =================
        const int ww = warp - 4;               // worker warp 0..7
        const int wt = threadIdx.x - 128;      // worker thread 0..255
        const int q = ww & 3;                  // TMEM lane quadrant of an epilogue warp
        int acc = 0;
        uint32_t aparity[2] = {0, 0};
        int attn_seen = 0;
        for (int p = 0; p < P.n_phases; p++) {
            const MegaPhase ph = P.phases[p];
            const MegaSub sb = P.sub[ph.sub];
            if (is_gemm(ph.kind)) {
                if (ww >= 4) continue;  // warps 8-11 have no part in the GEMM phases
                const MegaGemm& gm = gemm_of(ph.kind);
                if (P.trace && wt == 0) P.trace[((size_t)p * G + cta) * 2] = globaltimer();
                for (int u = first_unit(ph.rot); u < gm.units; u += G) {
                    const int tile = u % gm.tiles_n, split = u / gm.tiles_n;
                    const int kb0 = split * gm.kb_per_split, kb1 = min(gm.num_kb, kb0 + gm.kb_per_split);
                    ptx::mbar_wait(&tfull_bar[acc], aparity[acc]);
                    ptx::tc_fence_after();
                    const uint32_t t_row = tmem_base + acc * ACC_COLS + (uint32_t(q * 32) << 16);
                    const int r = q * 32 + lane;
                    if (ph.kind == MK_GU) {
                        // token row r of the sub-batch; columns alternate 32 gate / 32 up (gemm.cuh EPI_SWIGLU)
                        const bool row_ok = r < sb.rows;
                        bf16* out = P.act + (size_t)(sb.row0 + r) * P.inter + tile * (GU_BN / 2);
#pragma unroll 1
                        for (int c = 0; c < GU_BN / 32; c++) {
                            const int col = (c >> 1) * (2 * GU_UNIT) + (c & 1) * 16;
                            uint32_t gv[16], uv[16];
                            ptx::tmem_ld_32x16(t_row + col, gv);
                            ptx::tmem_ld_32x16(t_row + col + GU_UNIT, uv);
                            ptx::tmem_ld_wait();
                            if (row_ok) {
                                uint32_t pk[8];
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const float a = epi_swiglu(__uint_as_float(gv[2 * j]), __uint_as_float(uv[2 * j]));
                                    const float bb = epi_swiglu(__uint_as_float(gv[2 * j + 1]), __uint_as_float(uv[2 * j + 1]));
                                    pk[j] = pack_bf16x2(a, bb);
                                }
                                uint4* dst = reinterpret_cast<uint4*>(out + c * 16);
                                dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                            }
                        }
                    } else {
                        // weight row r of the tile; fp32 split-K partial [split][token][N] (skinny.cuh SK_PARTIAL)
                        const bool have = kb1 > kb0;
                        float* out = P.ws[ph.sub] + (size_t)split * ((size_t)sb.rows * gm.N) + tile * 128 + r;
                        constexpr int CH = NB < 32 ? NB : 32;
#pragma unroll 1
                        for (int c = 0; c < NB / CH; c++) {
                            uint32_t v[CH];
                            if constexpr (CH == 32) ptx::tmem_ld_32x32(t_row + c * CH, v);
                            else ptx::tmem_ld_32x16(t_row + c * CH, reinterpret_cast<uint32_t(&)[16]>(v));
                            ptx::tmem_ld_wait();
                            if (tile * 128 + r < gm.N) {
#pragma unroll
                                for (int j = 0; j < CH; j++) {
                                    const int m = c * CH + j;
                                    if (m < sb.rows) out[(size_t)m * gm.N] = have ? __uint_as_float(v[j]) : 0.f;
                                }
                            }
                        }
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
                    aparity[acc] ^= 1;
                    acc ^= 1;
                }
                // this CTA's share of the phase is written: CTA-wide barrier of the epilogue warps, then one release arrival
                fence_proxy_async_all();
                bar_sync(2, 128);
                if (wt == 0) {
                    __threadfence();
                    atomicAdd(P.cnt + p, 1u);
                    if (P.trace) P.trace[((size_t)p * G + cta) * 2 + 1] = globaltimer();
                }
            } else if (ph.kind == MK_ATTN) {
                if (P.trace && wt == 0) P.trace[((size_t)p * G + cta) * 2] = globaltimer();
                wait_phase(P, ph.dep, lane);
                bar_sync(1, 256);  // every worker is past its earlier phases: the GEMM ring is drained and the producers are gated
                const int nw = P.nw_attn, groups = 8 / nw, grp = ww / nw, gw = ww % nw;
                const int items = sb.rows * P.kv_heads;
                AttnShared S;
                S.qb = s_qb[grp];
                S.nw = s_new[grp];
                S.m = s_m;
                S.l = s_l;
                S.acc = s_acc;
                bf16* ring = reinterpret_cast<bf16*>(smem) + (size_t)ww * DM_STAGES * 2 * DM_CHUNK * 128;
                // item i -> CTA (i + rot) % G, group (i / G) % groups, round i / (G * groups)
                for (int i = first_unit(ph.rot) + grp * G; i < items; i += G * groups)
                    attention_item(P, ph.layer, ph.sub, i / P.kv_heads, i % P.kv_heads, gw, nw, grp * nw, lane, ring, S, 3 + grp);
                fence_proxy_async_all();  // generic-proxy writes to the ring region before TMA reuses it
                bar_sync(1, 256);
                attn_seen++;
                if (wt == 0) {
                    __threadfence();
                    atomicAdd(P.cnt + p, 1u);
                    *attn_done = attn_seen;  // the producers of the next GEMM phases may use the ring again
                    if (P.trace) P.trace[((size_t)p * G + cta) * 2 + 1] = globaltimer();
                }
            } else {
                // MK_NORM1 / MK_NORM2: rows of the sub-batch, one at a time on the 256 worker threads
                if (P.trace && wt == 0) P.trace[((size_t)p * G + cta) * 2] = globaltimer();
                wait_phase(P, ph.dep, lane);
                const bool second = ph.kind == MK_NORM2;
                const MegaGemm& gm = P.g[second ? 3 : 1];
                const bool last = second && ph.layer + 1 == P.layers;
                const bf16* w = P.norm_w[ph.layer * 4 + (second ? 3 : 2)];
                bf16* ybase = last ? P.dlast : P.xn;
                for (int r = first_unit(ph.rot); r < sb.rows; r += G) {
                    const int row = sb.row0 + r;
                    norm_row(P.ws[ph.sub] + (size_t)r * P.H, gm.splits, (long long)sb.rows * P.H, second ? P.sg2 : P.sg1,
                             P.x + (size_t)row * P.H, w, ybase + (size_t)row * P.H, P.H, P.eps, wt, s_red);
                }
                fence_proxy_async_all();
                bar_sync(1, 256);
                if (wt == 0) {
                    __threadfence();
                    atomicAdd(P.cnt + p, 1u);
                    if (P.trace) P.trace[((size_t)p * G + cta) * 2 + 1] = globaltimer();
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<2 * ACC_COLS>(tmem_base);
    }
}

}  // namespace q3
