// attention.cu — segment-packed multi-head attention (encoder windows, decoder prefill).
//
// Replaces MLXFast.scaledDotProductAttention as used through SDPA.multiHead / attendAndMerge
// (/root/reference/Sources/MLXCommon/SDPA.swift:18-101) with the block-diagonal window mask of
// AudioEncoder.swift:337-357, 463-489 and the causal mask of FloatTextDecoder.swift:200-213: instead
// of materialising a [T,T] additive mask, every window / prompt is an independent segment.
//
// One CTA = 64 queries of one (segment, head); 4 warps x 16 query rows; keys stream through shared
// memory in tiles of 64 with an online softmax (exp2 domain, fp32 statistics); both products run on
// the tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate; P is rounded to bf16 for the PV product,
// the row sum uses the unrounded values).  [round-1 kernel: the tcgen05/TMEM rewrite is future work.]
#include "ops.cuh"

namespace q3 {

namespace {

constexpr int BQ = 64, BKV = 64;

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

struct AttnParams {
    const bf16 *q, *k, *v;
    bf16* o;
    int ldq, ldk, ldv, ldo;
    const int* row0;
    const int* len;
    int group;
    float scale_log2;
};

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(128) flash_attn_kernel(const AttnParams p) {
    constexpr int LDS = HD + 8;  // padded smem row (bf16): conflict-free ldmatrix
    extern __shared__ uint4 smem_u4[];
    bf16* sQ = reinterpret_cast<bf16*>(smem_u4);
    bf16* sK = sQ + BQ * LDS;
    bf16* sV = sK + BKV * LDS;

    const int seg = blockIdx.z, head = blockIdx.y;
    const int len = p.len[seg];
    const int q0 = blockIdx.x * BQ;
    if (q0 >= len) return;
    const int row0 = p.row0[seg];
    const int kvh = head / p.group;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int VPR = HD / 8;  // 16-byte vectors per row

    // stage Q (rows beyond the segment are zero)
    for (int i = tid; i < BQ * VPR; i += 128) {
        const int r = i / VPR, c = i % VPR;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q0 + r < len) v = *reinterpret_cast<const uint4*>(p.q + (size_t)(row0 + q0 + r) * p.ldq + head * HD + c * 8);
        *reinterpret_cast<uint4*>(sQ + r * LDS + c * 8) = v;
    }

    float o_acc[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; i++) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    const int kv_end = CAUSAL ? min(len, q0 + BQ) : len;
    const int qr = q0 + warp * 16 + (lane >> 2);  // this thread's first query row (second is +8)

    for (int k0 = 0; k0 < kv_end; k0 += BKV) {
        __syncthreads();  // previous tile fully consumed (also orders the Q staging before first use)
        for (int i = tid; i < BKV * VPR; i += 128) {
            const int r = i / VPR, c = i % VPR;
            uint4 kv4 = make_uint4(0, 0, 0, 0), vv4 = make_uint4(0, 0, 0, 0);
            if (k0 + r < len) {
                kv4 = *reinterpret_cast<const uint4*>(p.k + (size_t)(row0 + k0 + r) * p.ldk + kvh * HD + c * 8);
                vv4 = *reinterpret_cast<const uint4*>(p.v + (size_t)(row0 + k0 + r) * p.ldv + kvh * HD + c * 8);
            }
            *reinterpret_cast<uint4*>(sK + r * LDS + c * 8) = kv4;
            *reinterpret_cast<uint4*>(sV + r * LDS + c * 8) = vv4;
        }
        __syncthreads();

        // ---- S = Q K^T (16 x 64 per warp) ----
        float s[BKV / 8][4];
#pragma unroll
        for (int i = 0; i < BKV / 8; i++) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ks++) {
            uint32_t a[4];
            ldsm_x4(a, sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int np = 0; np < BKV / 16; np++) {
                uint32_t b[4];
                const int mi = lane >> 3;
                ldsm_x4(b, sK + (np * 16 + (lane & 7) + (mi >> 1) * 8) * LDS + ks * 16 + (mi & 1) * 8);
                mma16816(s[2 * np], a, b[0], b[1]);
                mma16816(s[2 * np + 1], a, b[2], b[3]);
            }
        }
        // ---- mask + online softmax ----
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < BKV / 8; nt++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int key = k0 + nt * 8 + (lane & 3) * 2 + (e & 1);
                const int qrow = qr + (e >> 1) * 8;
                const bool ok = key < len && (!CAUSAL || key <= qrow);
                const float val = ok ? s[nt][e] * p.scale_log2 : -INFINITY;
                s[nt][e] = val;
                mx[e >> 1] = fmaxf(mx[e >> 1], val);
            }
        }
        float alpha[2], mnew[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            mnew[h] = fmaxf(m_run[h], mx[h]);
            alpha[h] = mnew[h] == -INFINITY ? 1.f : exp2f(m_run[h] - mnew[h]);
            m_run[h] = mnew[h];
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pa[BKV / 16][4];
#pragma unroll
        for (int nt = 0; nt < BKV / 8; nt++) {
            float pv[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float mm = mnew[e >> 1];
                pv[e] = mm == -INFINITY ? 0.f : exp2f(s[nt][e] - mm);
                rs[e >> 1] += pv[e];
            }
            pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(pv[0], pv[1]);
            pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) l_run[h] = l_run[h] * alpha[h] + rs[h];
#pragma unroll
        for (int i = 0; i < HD / 8; i++) {
            o_acc[i][0] *= alpha[0]; o_acc[i][1] *= alpha[0];
            o_acc[i][2] *= alpha[1]; o_acc[i][3] *= alpha[1];
        }
        // ---- O += P V ----
#pragma unroll
        for (int kk = 0; kk < BKV / 16; kk++) {
#pragma unroll
            for (int np = 0; np < HD / 16; np++) {
                uint32_t b[4];
                const int mi = lane >> 3;
                ldsm_x4_t(b, sV + (kk * 16 + (lane & 7) + (mi & 1) * 8) * LDS + np * 16 + (mi >> 1) * 8);
                mma16816(o_acc[2 * np], pa[kk], b[0], b[1]);
                mma16816(o_acc[2 * np + 1], pa[kk], b[2], b[3]);
            }
        }
    }

    // ---- normalise and store ----
#pragma unroll
    for (int h = 0; h < 2; h++) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
    }
    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f, inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
#pragma unroll
    for (int i = 0; i < HD / 8; i++) {
        const int col = head * HD + i * 8 + (lane & 3) * 2;
        if (qr < len) *reinterpret_cast<uint32_t*>(p.o + (size_t)(row0 + qr) * p.ldo + col) = pack_bf16x2(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
        if (qr + 8 < len)
            *reinterpret_cast<uint32_t*>(p.o + (size_t)(row0 + qr + 8) * p.ldo + col) = pack_bf16x2(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
    }
}

template <int HD, bool CAUSAL>
void launch(const AttnParams& p, const AttnSegs& segs, int heads, cudaStream_t st) {
    constexpr int smem = (BQ + 2 * BKV) * (HD + 8) * 2;
    static PerDeviceOnce attr_once;
    attr_once([] {
        Q3_CUDA(cudaFuncSetAttribute(flash_attn_kernel<HD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    });
    dim3 grid((segs.max_len + BQ - 1) / BQ, heads, segs.n_segs);
    flash_attn_kernel<HD, CAUSAL><<<grid, 128, smem, st>>>(p);
    Q3_CUDA(cudaGetLastError());
}

}  // namespace

void flash_attn_launch(const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* o, int ldo, const AttnSegs& segs,
                       int heads, int group, int head_dim, bool causal, float scale, cudaStream_t st) {
    if (segs.n_segs <= 0 || segs.max_len <= 0) return;
    Q3_CHECK(segs.n_segs <= 65535 && heads <= 65535, 1, "attention: too many segments for one launch");
    AttnParams p;
    p.q = q; p.k = k; p.v = v; p.o = o;
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
    p.row0 = segs.row0; p.len = segs.len;
    p.group = group;
    p.scale_log2 = scale * 1.4426950408889634f;
    if (head_dim == 64 && !causal) launch<64, false>(p, segs, heads, st);
    else if (head_dim == 64 && causal) launch<64, true>(p, segs, heads, st);
    else if (head_dim == 128 && !causal) launch<128, false>(p, segs, heads, st);
    else if (head_dim == 128 && causal) launch<128, true>(p, segs, heads, st);
    else throw Error(1, "attention: head_dim must be 64 or 128");
}

}  // namespace q3
