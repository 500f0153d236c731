// megastep_params.h — plain parameter structs of the persistent decode-step kernel (megastep.cuh), shared with the batch state.
#pragma once
#include <cuda.h>

#include "ops.cuh"

namespace q3 {

enum MegaKind : int { MK_QKV = 0, MK_ATTN = 1, MK_O = 2, MK_NORM1 = 3, MK_GU = 4, MK_DOWN = 5, MK_NORM2 = 6 };

struct MegaPhase {  // 32 bytes, host-built
    int kind, layer, sub;
    int dep;          // phase that must be complete grid-wide before this one reads its inputs (-1: none)
    int attn_before;  // GEMM phases: attention phases that precede this one in program order (the ring is theirs until they end)
    int rot;          // unit -> CTA rotation
    int pad0, pad1;
};

struct MegaGemm {
    int N, num_kb, kb_per_split, tiles_n, splits, units;
};
struct MegaSub {
    int row0, rows;
};

struct MegaParams {
    int G, n_phases, layers, n_sub;
    int H, nq, nkv, nqkv, inter, heads, kv_heads;
    int nw_attn;   // warps per attention item: 2 or 8 (the groupings that reproduce the canonical stream merge order)
    int sg1, sg2;  // split groups of the two reduce phases (the summation order of reduce_resid_rmsnorm_kernel)
    float eps, scale_log2;
    MegaGemm g[4];  // qkv, o, gate|up, down
    MegaSub sub[2];
    const MegaPhase* phases;
    const CUtensorMap* maps;  // [layers][4] weights (qkv, o, gate|up, down), then [n_sub][4] activations (xn->qkv, att->o, xn->gu, act->down)
    unsigned int* cnt;        // [n_phases] arrival counters, zeroed before the launch
    bf16 *x, *xn, *att, *act, *dlast;
    float* ws[2];                // split-K partials of each sub-batch
    const bf16* const* norm_w;   // [layers][4]: q_norm, k_norm, post_attention_layernorm, the NEXT block's input norm (last: final norm)
    const int *pos, *kv_len;
    const float2* rope_tab;
    KvCache cache;
    unsigned long long* trace;  // debug (Q3ASR_MEGA_TRACE): [n_phases][G][2] globaltimer at phase begin / end of every CTA, or null
    int* err;  // set to the phase index + 1 when a wait times out (debug aid: a wrong table would otherwise hang the GPU)
};

}  // namespace q3
