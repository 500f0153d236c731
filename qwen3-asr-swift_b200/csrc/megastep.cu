// megastep.cu — host side of the persistent decode-step kernel (megastep.cuh): phase table, tensor maps, launch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "megastep.cuh"
#include "model.h"

namespace q3 {

void make_tmap_2d_bf16(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t row_stride_elems, uint32_t box_cols,
                       uint32_t box_rows);

namespace {

int nb_for(int rows) { return rows <= 16 ? 16 : rows <= 32 ? 32 : rows <= 64 ? 64 : 128; }

int gu_bn_for(const q3asr_config& c, int num_sms) {
    // the tile width decoder_layers_decode picks for the gate|up product (forward.cu): the narrowest 64-multiple with at most one
    // tile per SM
    int bn = 64;
    while ((2 * c.dec_inter) / bn > num_sms && bn < 256 && (2 * c.dec_inter) % (2 * bn) == 0) bn *= 2;
    return bn;
}

int reduce_sg(int d, int splits) {  // the split-group count reduce_resid_rmsnorm_launch uses (ops.cu): it fixes the summation order
    const int ncg = d / 4;
    int sg = std::max(1, 1024 / ncg);
    while (sg > 1 && sg > splits) sg >>= 1;
    return sg;
}

template <int NB, int GU_BN>
void launch(const MegaParams& P, cudaStream_t st, bool cooperative) {
    static PerDeviceOnce attr_once;
    attr_once([] {
        Q3_CUDA(cudaFuncSetAttribute(megastep_kernel<NB, GU_BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, mega_smem_bytes()));
    });
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)P.G);
    cfg.blockDim = dim3(MEGA_THREADS);
    cfg.dynamicSmemBytes = mega_smem_bytes();
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (cooperative) {  // every CTA must be resident (they wait on each other): two such grids on one GPU must not interleave
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    Q3_CUDA(cudaLaunchKernelEx(&cfg, megastep_kernel<NB, GU_BN>, P));
}

template <int NB>
void launch_nb(int gu_bn, const MegaParams& P, cudaStream_t st, bool coop) {
    if (gu_bn == 64) launch<NB, 64>(P, st, coop);
    else launch<NB, 128>(P, st, coop);
}

}  // namespace

bool megastep_supported(const Handle* h, const BatchState* bs) {
    const q3asr_config& c = h->cfg;
    // Opt-in (Q3ASR_MEGA=1): measured on B200 the persistent kernel is bit-identical to the multi-kernel decode step but slower
    // (every phase keeps its ~3 us dependent pipeline latency, and splitting the batch to hide the
    // hand-over doubles the phases), so the chain of PDL launches stays the default.
    const char* env = getenv("Q3ASR_MEGA");
    if (!(env && atoi(env) != 0)) return false;
    const int gu = gu_bn_for(c, h->num_sms);
    return bs->dec_rows >= 1 && bs->dec_rows <= 128 /* the persistent kernel's sub-batch tiles */ && c.dec_head_dim == 128 && c.dec_heads == 2 * c.dec_kv_heads &&
           c.dec_hidden % 1024 == 0 && c.dec_hidden <= 2048 && c.dec_inter % 64 == 0 && (gu == 64 || gu == 128) &&
           (2 * c.dec_inter) % gu == 0 && h->num_sms >= 64;
}

// Builds the phase table, the tensor maps and the pointer tables of one resident batch.  Called once per batch (run_decode):
// the activation buffers may have moved since the last one.
void megastep_prepare(Handle* h, BatchState* bs) {
    const q3asr_config& c = h->cfg;
    const Model& m = *h->model;
    const int B = bs->dec_rows, H = c.dec_hidden, hd = c.dec_head_dim, nq = c.dec_heads * hd, nkv = c.dec_kv_heads * hd, nqkv = nq + 2 * nkv;
    const int L = c.dec_layers, G = h->num_sms;
    MegaParams& P = bs->mega;
    memset(&P, 0, sizeof(P));
    P.G = G;
    P.layers = L;
    P.n_sub = B >= 2 ? 2 : 1;
    if (const char* e = getenv("Q3ASR_MEGA_SUBS")) P.n_sub = atoi(e) >= 2 && B >= 2 ? 2 : 1;  // experiment switch
    P.sub[0] = P.n_sub == 2 ? MegaSub{0, (B + 1) / 2} : MegaSub{0, B};
    P.sub[1] = P.n_sub == 2 ? MegaSub{(B + 1) / 2, B - (B + 1) / 2} : MegaSub{B, 0};
    P.n_phases = L * 7 * P.n_sub;
    P.H = H; P.nq = nq; P.nkv = nkv; P.nqkv = nqkv; P.inter = c.dec_inter; P.heads = c.dec_heads; P.kv_heads = c.dec_kv_heads;
    const int items = P.sub[0].rows * c.dec_kv_heads;
    // 2 or 8 warps per (sequence, kv head) only: those are the groupings whose fold order equals the canonical stream tree of
    // decode_attn_mma_kernel (four warps would fold s0+s4 before s2: a different fp32 order, i.e. ids that depend on the batch size)
    P.nw_attn = items <= G ? 8 : 2;
    P.eps = c.dec_rms_eps;
    P.scale_log2 = (1.0f / sqrtf((float)hd)) * 1.4426950408889634f;
    bs->mega_nb = nb_for(P.sub[0].rows);
    bs->mega_gu = gu_bn_for(c, G);
    auto skinny = [&](int N, int K) {
        MegaGemm g;
        g.N = N;
        g.num_kb = cdiv(K, SK_BK);
        g.splits = gemm_skinny_splits(N, K, SK_PARTIAL);
        g.kb_per_split = cdiv(g.num_kb, g.splits);
        g.tiles_n = cdiv(N, SK_BM);
        g.units = g.tiles_n * g.splits;
        return g;
    };
    P.g[0] = skinny(nqkv, H);
    P.g[1] = skinny(H, nq);
    P.g[3] = skinny(H, c.dec_inter);
    P.g[2].N = 2 * c.dec_inter;
    P.g[2].num_kb = H / GEMM_BK;
    P.g[2].kb_per_split = P.g[2].num_kb;
    P.g[2].splits = 1;
    P.g[2].tiles_n = (2 * c.dec_inter) / bs->mega_gu;
    P.g[2].units = P.g[2].tiles_n;
    P.sg1 = reduce_sg(H, P.g[1].splits);
    P.sg2 = reduce_sg(H, P.g[3].splits);

    // ---- host-side tables in one pinned staging buffer -> one device buffer ----
    const size_t n_maps = (size_t)L * 4 + (size_t)P.n_sub * 4;
    const size_t off_maps = 0;
    const size_t off_phases = off_maps + n_maps * sizeof(CUtensorMap);
    const size_t off_normw = off_phases + (size_t)P.n_phases * sizeof(MegaPhase);
    const size_t off_cnt = (off_normw + (size_t)L * 4 * sizeof(void*) + 255) & ~size_t(255);
    const size_t total = off_cnt + (size_t)P.n_phases * sizeof(unsigned) + 256;
    bs->mega_tab.reserve(total);
    bs->h_mega.reserve(total);
    uint8_t* hp = bs->h_mega.as<uint8_t>();
    uint8_t* dp = reinterpret_cast<uint8_t*>(bs->mega_tab.p);
    memset(hp, 0, total);
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(hp + off_maps);
    for (int l = 0; l < L; l++) {
        const DecLayerW& w = m.dec[l];
        make_tmap_2d_bf16(&maps[l * 4 + 0], w.qkv_w, H, nqkv, H, SK_BK, SK_BM);
        make_tmap_2d_bf16(&maps[l * 4 + 1], w.o_w, nq, H, nq, SK_BK, SK_BM);
        make_tmap_2d_bf16(&maps[l * 4 + 2], w.gu_w, H, 2 * c.dec_inter, H, GEMM_BK, bs->mega_gu);
        make_tmap_2d_bf16(&maps[l * 4 + 3], w.down_w, c.dec_inter, H, c.dec_inter, SK_BK, SK_BM);
    }
    bs->dws2.reserve(bs->dws.cap);  // the second sub-batch's split-K partials
    for (int s = 0; s < P.n_sub; s++) {
        const int r0 = P.sub[s].row0, rows = P.sub[s].rows;
        CUtensorMap* xm = maps + (size_t)L * 4 + (size_t)s * 4;
        make_tmap_2d_bf16(&xm[0], bs->dxn.as<bf16>() + (size_t)r0 * H, H, rows, H, SK_BK, bs->mega_nb);
        make_tmap_2d_bf16(&xm[1], bs->datt.as<bf16>() + (size_t)r0 * nq, nq, rows, nq, SK_BK, bs->mega_nb);
        make_tmap_2d_bf16(&xm[2], bs->dxn.as<bf16>() + (size_t)r0 * H, H, rows, H, GEMM_BK, GEMM_BM);
        make_tmap_2d_bf16(&xm[3], bs->dact.as<bf16>() + (size_t)r0 * c.dec_inter, c.dec_inter, rows, c.dec_inter, SK_BK, bs->mega_nb);
    }
    MegaPhase* ph = reinterpret_cast<MegaPhase*>(hp + off_phases);
    auto idx = [&](int l, int kind, int s) { return (l * 7 + kind) * P.n_sub + s; };
    for (int l = 0; l < L; l++)
        for (int kind = 0; kind < 7; kind++)
            for (int s = 0; s < P.n_sub; s++) {
                MegaPhase& x = ph[idx(l, kind, s)];
                x.kind = kind;
                x.layer = l;
                x.sub = s;
                x.dep = kind == MK_QKV ? (l == 0 ? -1 : idx(l - 1, MK_NORM2, s)) : idx(l, kind - 1, s);
                x.attn_before = (kind > MK_ATTN ? l + 1 : l) * P.n_sub;
                x.rot = (int)(((long)idx(l, kind, s) * 37) % G);
            }
    const bf16** nw = reinterpret_cast<const bf16**>(hp + off_normw);
    for (int l = 0; l < L; l++) {
        nw[l * 4 + 0] = m.dec[l].q_norm;
        nw[l * 4 + 1] = m.dec[l].k_norm;
        nw[l * 4 + 2] = m.dec[l].post_ln;
        nw[l * 4 + 3] = l + 1 < L ? m.dec[l + 1].in_ln : m.final_norm;
    }
    Q3_CUDA(cudaMemcpyAsync(dp, hp, total, cudaMemcpyHostToDevice, h->stream));
    Q3_CUDA(cudaStreamSynchronize(h->stream));  // h_mega is rewritten by the next batch
    P.maps = reinterpret_cast<const CUtensorMap*>(dp + off_maps);
    P.phases = reinterpret_cast<const MegaPhase*>(dp + off_phases);
    P.norm_w = reinterpret_cast<const bf16* const*>(dp + off_normw);
    P.cnt = reinterpret_cast<unsigned int*>(dp + off_cnt);
    P.err = reinterpret_cast<int*>(dp + off_cnt + (size_t)P.n_phases * sizeof(unsigned));
    P.x = bs->dx.as<bf16>();
    P.xn = bs->dxn.as<bf16>();
    P.att = bs->datt.as<bf16>();
    P.act = bs->dact.as<bf16>();
    P.dlast = bs->dlast.as<bf16>();
    P.ws[0] = bs->dws.as<float>();
    P.ws[1] = bs->dws2.as<float>();
    P.pos = bs->st_pos.as<int>();
    P.kv_len = bs->st_kv_len.as<int>();
    P.rope_tab = bs->rope_tab.as<float2>();
    P.cache.pool = bs->kv_pool.as<bf16>();
    P.cache.page_table = (bs->page_cur ? bs->page_tab2 : bs->page_tab).as<int>();
    P.cache.max_pages = bs->pages_per_seq;
    P.cache.layers = c.dec_layers;
    P.cache.kv_heads = c.dec_kv_heads;
    P.cache.head_dim = c.dec_head_dim;
    P.trace = nullptr;
    if (getenv("Q3ASR_MEGA_TRACE")) {  // debug: per-CTA phase timestamps of the next eager launch, written to that file
        bs->mega_trace.reserve((size_t)P.n_phases * G * 2 * sizeof(unsigned long long));
        P.trace = bs->mega_trace.as<unsigned long long>();
    }
    bs->mega_ready = true;
}

// One launch = the 28 decoder layers of one decode step.  On entry xn holds RMSNorm(x) under the first block's input norm; on
// exit dlast holds the final-norm hidden states (the LM-head input), x the residual stream.
void megastep_launch(Handle* h, BatchState* bs) {
    const MegaParams& P = bs->mega;
    Q3_CHECK(bs->mega_ready, Q3ASR_ERR_STATE, "megastep: not prepared");
    cudaStream_t st = h->stream;
    Q3_CUDA(cudaMemsetAsync(P.cnt, 0, (size_t)P.n_phases * sizeof(unsigned) + sizeof(int), st));
    const bool coop = getenv("Q3ASR_MEGA_COOP") == nullptr || atoi(getenv("Q3ASR_MEGA_COOP")) != 0;
    switch (bs->mega_nb) {
        case 16: launch_nb<16>(bs->mega_gu, P, st, coop); break;
        case 32: launch_nb<32>(bs->mega_gu, P, st, coop); break;
        default: launch_nb<64>(bs->mega_gu, P, st, coop); break;
    }
    h->launches++;
    if (P.trace) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs == cudaStreamCaptureStatusNone) {
            Q3_CUDA(cudaStreamSynchronize(st));
            std::vector<unsigned long long> t((size_t)P.n_phases * P.G * 2);
            Q3_CUDA(cudaMemcpy(t.data(), P.trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            if (FILE* f = fopen(getenv("Q3ASR_MEGA_TRACE"), "wb")) {
                const int hdr[4] = {P.n_phases, P.G, P.n_sub, 7};
                fwrite(hdr, sizeof(hdr), 1, f);
                fwrite(t.data(), sizeof(unsigned long long), t.size(), f);
                fclose(f);
            }
        }
    }
}

}  // namespace q3
