// mel.cu — K1: fused log-mel frontend for sm_100a.
//
// Replaces WhisperFeatureExtractor.extractFeatures
// (/root/reference/Sources/Qwen3ASR/AudioPreprocessing.swift:169-317): reflect padding (:174-192),
// Hann window + zero-pad 400->512 + real FFT with vDSP's 2x scale + power (:209-250), the 257->128
// slaney filterbank product (:263-268), clip/log10/clip-to-(max-8)/scale (:275-293), drop of the
// last frame and the 120000-frame cap (:296-313), [128, T] mel-major output (:315-316).
//
// B200 design.  The stage moves 7.2 B per sample, but a 512-point FFT per 160 new samples makes it a shared-memory kernel: ncu
// shows the LSU data pipe (one 128-byte wavefront per clock per SM) ~80 % busy while the SM is active, so
// the design goal is wavefronts per frame, then issue slots:
//   * Blackwell's packed fp32 pipe (FADD2 / FMUL2 / FFMA2) processes two fp32 values per instruction.  Every value in the
//     transform is held as a PAIR (frame A, frame B) of two neighbouring frames, so that EVERY arithmetic instruction of the FFT,
//     the real-split step, the power spectrum and the filterbank is a packed one, with no shuffling inside a pair: half the
//     arithmetic instructions per frame, and the shared-memory transpose moves both frames with one 128-bit access.
//   * Sixteen lanes share one frame pair (two pairs per warp, every synchronisation inside the transform is a __syncwarp); a CTA
//     of 8 * MEL_TILE threads transforms the frames of a tile in one pass, four (16-frame tiles) CTAs per SM.
//   * The 512-point real FFT is the 256-point complex FFT of the even/odd packed frame, done as 16 x 16 (two in-register
//     radix-16 passes with ONE shared-memory transpose), followed by the real-split step that yields bins k and 256-k together;
//     the partner value Z[256 - k] comes from lane 16 - t by warp shuffle, not through shared memory.  The vDSP factor 2 cancels
//     against the 1/2 of the split step (|2X|^2 = |e + w o|^2).  The filterbank is an ELL-packed sparse product (504 non-zeros
//     instead of 32896 MACs per frame).  (Measured and dropped in round 2: the filterbank as a banded bf16 hi/lo-split product on
//     mma.sync — 20 % fewer wavefronts, but a latency-bound phase behind a third barrier: 188-200 us against 160.)
//   * 0.25*log10(mel)+1 is written unclamped; the values are parked in shared memory and leave as whole mel rows (one 128-byte
//     line per warp instruction instead of 16 scattered lines), with an ordered-int atomicMax per clip and a per-tile minimum;
//     the second kernel applies the max-8 clamp only to tiles whose minimum is below it (exact, because clamp and the monotone
//     affine map commute), so in the common case the features are written once and never re-read.
//   * The tile loop is software-pipelined: the next tile's samples travel (cp.async) while this tile is transformed, the
//     previous tile's rows are stored at the top of the next iteration; clip descriptors are looked up two tiles ahead by one
//     thread and kept in shared memory.
// The arithmetic per frame is operation for operation that of the scalar kernel this replaces (same products, same fused
// multiply-adds, same order), so the results are bit-identical to it.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <vector>

#include "mel.cuh"
#include "ptx.cuh"

namespace q3 {

namespace {

constexpr int TPF = 16;                                // lanes per frame pair
constexpr int MEL_THREADS = 8 * MEL_TILE;               // threads that transform one tile together (a group)
// A CTA is GROUPS independent groups of MEL_THREADS threads: each walks its own tiles with its own sample buffer, scratch and named
// barrier, and they share the constant tables.  One CTA per SM: 5 groups of 16-frame tiles (20 warps at <= 102 registers, 224 KB
// of shared memory — both budgets nearly full) where four separate CTAs, each with its own copy of the tables, were the limit.
constexpr int GROUPS = MEL_TILE == 16 ? 5 : 2;
constexpr int PAIRS = MEL_THREADS / TPF;               // frame pairs in flight per group = the frames of a tile
static_assert(MEL_TILE == 2 * PAIRS, "a group transforms a whole tile in one pass");
constexpr int TILE_SAMPLES = (MEL_TILE - 1) * MEL_HOP + MEL_NFFT;  // 2800 (16-frame tiles)
constexpr int SX_FLOATS = TILE_SAMPLES + 16;
constexpr int SCR_F4 = 256;                            // float4s of scratch per frame pair: the 16x16 transpose (XOR-swizzled columns,
                                                       // no padding) and then the flat spectrum
constexpr int PARK_F2 = 384;                          // float2 offset of the parked features inside a pair's scratch block (the spectrum ends at 257)
constexpr int HANN_PAD = 416;                          // window taps, zero beyond 400

struct MelParams {
    const float* hann;
    const float2* tw256;
    const float2* tw512;
    const float* fbw;
    const int* fb_start;
    int fb_round_off[MEL_ROUNDS + 1];
    int fb_rows;
    const float* pcm;
    float* out;
    const MelClip* clips;
    int batch;
    int tile_lo;      // tiles [tile_lo, total_tiles) of the clips [0, batch) are transformed by this launch
    int total_tiles;
    int* gmax;
    float* tmin;
    int* tclip;  // clip of every tile (for the clamp pass)
};

// Filter weights in shared memory: one row of FBW_STRIDE floats per lane (its eight filters' taps, round after round), read with
// 128-bit loads: 13 per frame pair instead of 40 scalar ones.  52 = 20 mod 32 puts the eight lanes of a quarter-warp in different
// bank quads.
constexpr int FBW_STRIDE = 52;
constexpr int FBW_FLOATS = 16 * FBW_STRIDE;
// hann | tw256 | tw512 | filterbank weights | filter starts | per group: samples | scratch | reductions.  The tables are NOT stored as
// (w, w) pairs: the kernel is bound by shared-memory wavefronts (ncu: LSU data pipe 80 % busy), not by issue slots, so a pair is
// formed with a register move after an 8-byte load
__host__ __device__ constexpr int mel_smem_bytes(int fb_rows) {
    return (HANN_PAD + 2 * 256 + 2 * 130 + FBW_FLOATS + 128 + GROUPS * (SX_FLOATS + 4 * PAIRS * SCR_F4 + 64)) * 4;  // the last 64 words: reductions [32], next clips' indices [2], clips [2]
}

// Widths (taps) of the eight filterbank rounds: a property of the slaney filterbank at 16 kHz / 512 points / 128 bins, checked
// against the computed table in mel_tables_create.  Compile-time constants let the tap loops unroll completely (the loop control
// was 15 % of the kernel's instructions).
template <int J> struct FbRound;  // W taps; POFF: offset of the round in a lane's weight row (rounds padded to multiples of four taps)
template <> struct FbRound<0> { static constexpr int W = 2, POFF = 0; };
template <> struct FbRound<1> { static constexpr int W = 2, POFF = 4; };
template <> struct FbRound<2> { static constexpr int W = 2, POFF = 8; };
template <> struct FbRound<3> { static constexpr int W = 3, POFF = 12; };
template <> struct FbRound<4> { static constexpr int W = 4, POFF = 16; };
template <> struct FbRound<5> { static constexpr int W = 6, POFF = 20; };
template <> struct FbRound<6> { static constexpr int W = 9, POFF = 28; };
template <> struct FbRound<7> { static constexpr int W = 12, POFF = 40; };
constexpr int FB_POFF[MEL_ROUNDS] = {0, 4, 8, 12, 16, 20, 28, 40};
constexpr int FB_ROWS = 40;
constexpr int FB_WIDTHS[MEL_ROUNDS] = {2, 2, 2, 3, 4, 6, 9, 12};

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c);
template <int J>
__device__ __forceinline__ float2 fb_round(const float2* __restrict__ scr, const float* __restrict__ s_fbw, const int* __restrict__ s_fbstart, int t) {
    constexpr int W = FbRound<J>::W, Q = (W + 3) / 4;
    const float2* pp = scr + s_fbstart[t + 16 * J];
    const float4* wq = reinterpret_cast<const float4*>(s_fbw + t * FBW_STRIDE + FbRound<J>::POFF);
    float wv[4 * Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const float4 v = wq[q];
        wv[4 * q] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w;
    }
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int w = 0; w < W; w++) acc = fma2(pp[w], make_float2(wv[w], wv[w]), acc);
    return acc;
}

// Blackwell's packed fp32 pipe (FADD2 / FMUL2 / FFMA2: two fp32 lanes per issue slot; add.rn.f32x2 etc. in PTX): a complex
// add / subtract, or an element-wise product of two float2, is ONE instruction.  The kernel is bound by issue slots, not by bytes
// (DESIGN.md section 4.1), so this is where its time goes down.  Same IEEE results as the scalar forms (round-to-nearest each).
__device__ __forceinline__ unsigned long long& as_u64(float2& v) { return reinterpret_cast<unsigned long long&>(v); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    float2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(r)) : "l"(as_u64(a)), "l"(as_u64(b)));
    return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    float2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(r)) : "l"(as_u64(a)), "l"(as_u64(b)));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {  // element-wise
    float2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(r)) : "l"(as_u64(a)), "l"(as_u64(b)));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {  // element-wise a * b + c
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(as_u64(r)) : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
    return r;
}
// A complex value of both frames of a pair: re = (re_A, re_B), im = (im_A, im_B).
struct C2 {
    float2 re, im;
};
__device__ __forceinline__ float2 splat(float c) { return make_float2(c, c); }
// x * (wr + i wi) with the products and fused multiply-adds of the scalar form (fmaf(x.re, wr, -x.im * wi), fmaf(x.re, wi, x.im * wr));
// nwi = -wi (an exact negation, so x.im * nwi = -(x.im * wi))
__device__ __forceinline__ C2 cmul(C2 x, float2 wr, float2 wi, float2 nwi) {
    C2 r;
    r.re = fma2(x.re, wr, mul2(x.im, nwi));
    r.im = fma2(x.re, wi, mul2(x.im, wr));
    return r;
}
__device__ __forceinline__ C2 cmulc(C2 x, float wr, float wi) { return cmul(x, splat(wr), splat(wi), splat(-wi)); }

// forward radix-4 butterfly, in place: (a,b,c,d) <- DFT4
__device__ __forceinline__ void bfly4(C2& a, C2& b, C2& c, C2& d) {
    const float2 t0r = cadd(a.re, c.re), t0i = cadd(a.im, c.im), t1r = csub(a.re, c.re), t1i = csub(a.im, c.im);
    const float2 t2r = cadd(b.re, d.re), t2i = cadd(b.im, d.im), t3r = csub(b.re, d.re), t3i = csub(b.im, d.im);
    a.re = cadd(t0r, t2r); a.im = cadd(t0i, t2i);
    c.re = csub(t0r, t2r); c.im = csub(t0i, t2i);
    b.re = cadd(t1r, t3i); b.im = csub(t1i, t3r);
    d.re = csub(t1r, t3i); d.im = cadd(t1i, t3r);
}

// forward 16-point DFT in registers.  Input x[n] natural order; on return X[k] is at x[4*(k&3) + (k>>2)].
__device__ __forceinline__ void fft16(C2 (&x)[16]) {
    constexpr float C = 0.92387953251128674f, S = 0.38268343236508977f, R = 0.70710678118654752f;
#pragma unroll
    for (int n1 = 0; n1 < 4; n1++) bfly4(x[n1], x[n1 + 4], x[n1 + 8], x[n1 + 12]);
    // x[n1 + 4*k2] *= W16^(n1*k2)
    x[1 + 4] = cmulc(x[1 + 4], C, -S);    // W^1
    x[1 + 8] = cmulc(x[1 + 8], R, -R);    // W^2
    x[1 + 12] = cmulc(x[1 + 12], S, -C);  // W^3
    x[2 + 4] = cmulc(x[2 + 4], R, -R);    // W^2
    {                                     // W^4 = -i: (re, im) <- (im, -re)
        const float2 r = x[2 + 8].re;
        x[2 + 8].re = x[2 + 8].im;
        x[2 + 8].im = csub(make_float2(0.f, 0.f), r);
    }
    x[2 + 12] = cmulc(x[2 + 12], -R, -R); // W^6
    x[3 + 4] = cmulc(x[3 + 4], S, -C);    // W^3
    x[3 + 8] = cmulc(x[3 + 8], -R, -R);   // W^6
    x[3 + 12] = cmulc(x[3 + 12], -C, S);  // W^9
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) bfly4(x[4 * k2], x[4 * k2 + 1], x[4 * k2 + 2], x[4 * k2 + 3]);
}
__device__ __forceinline__ constexpr int rev16(int k) { return 4 * (k & 3) + (k >> 2); }

__device__ __forceinline__ int enc_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}

__device__ __forceinline__ int find_clip(const MelClip* clips, int batch, int tile) {
    int lo = 0, hi = batch - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&clips[mid].tile0) <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}


// Stages one tile's padded samples (pad index p0 + i, i < SX_FLOATS) into dst.  Interior tiles: ONE bulk copy (cp.async.bulk, the TMA
// unit writes shared memory through the async proxy: no LSU wavefronts and one instruction instead of 704 LDGSTS per tile — the
// staging writes were 9 % of the kernel's shared-memory wavefronts, the unit that bounds it), completion on the group's mbarrier.
// Tiles that touch a reflected edge: plain loads and stores by every thread, and a plain arrival so that the barrier's phase
// advances once per staged tile either way (their visibility comes from the group barrier that follows the wait).
__device__ __forceinline__ void stage_tile(float* dst, const MelClip& c, const float* pcm, int f0, int tid, uint64_t* bar) {
    const float* x = pcm + c.in_off;
    const int n = c.n;
    const int p0 = f0 * MEL_HOP;
    if (p0 >= MEL_NFFT / 2 && p0 + SX_FLOATS <= MEL_NFFT / 2 + n) {
        if (tid == 0) {
            ptx::mbar_arrive_expect_tx(bar, SX_FLOATS * 4);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ptx::smem_u32(dst)),
                         "l"(x + (p0 - MEL_NFFT / 2)), "n"(SX_FLOATS * 4), "r"(ptx::smem_u32(bar))
                         : "memory");
        }
    } else {
        for (int i = tid; i < SX_FLOATS; i += MEL_THREADS) {
            const int pp = p0 + i;
            float v = 0.f;
            if (pp < MEL_NFFT / 2) {
                int s = MEL_NFFT / 2 - pp;
                s = s > n - 1 ? n - 1 : s;
                v = __ldg(x + s);
            } else if (pp < MEL_NFFT / 2 + n) {
                v = __ldg(x + (pp - MEL_NFFT / 2));
            } else if (pp < MEL_NFFT + n) {
                int s = n - 2 - (pp - MEL_NFFT / 2 - n);
                s = s < 0 ? 0 : s;
                v = __ldg(x + s);
            }
            dst[i] = v;
        }
        if (tid == 0) ptx::mbar_arrive(bar);
    }
}
static_assert((SX_FLOATS * 4) % 16 == 0 && (MEL_HOP * 4) % 16 == 0 && (MEL_NFFT / 2 * 4) % 16 == 0, "bulk copies need 16-byte aligned tiles");

__global__ void __launch_bounds__(GROUPS * MEL_THREADS, 1) mel_kernel(const MelParams p) {
    ptx::grid_dep_launch();  // the clamp pass may be scheduled while this grid drains (it waits for the grid's completion itself)
    extern __shared__ float4 smem4[];
    float* s_hann = reinterpret_cast<float*>(smem4);                      // [416], zero beyond 400
    float2* s_tw256 = reinterpret_cast<float2*>(s_hann + HANN_PAD);       // [256] (wr, wi)
    float2* s_tw512 = s_tw256 + 256;                                      // [130]
    float* s_fbw = reinterpret_cast<float*>(s_tw512 + 130);               // [16 lanes][FBW_STRIDE]
    int* s_fbstart = reinterpret_cast<int*>(s_fbw + FBW_FLOATS);          // [128]
    const int group = threadIdx.x / MEL_THREADS;                          // this thread's group and its private buffers
    float* s_x = reinterpret_cast<float*>(s_fbstart + 128) + (size_t)group * (SX_FLOATS + 4 * PAIRS * SCR_F4 + 64);
    float4* s_scr = reinterpret_cast<float4*>(s_x + SX_FLOATS);           // [PAIRS][SCR_F4]
    float* s_red = reinterpret_cast<float*>(s_scr + PAIRS * SCR_F4);      // [32]
    int* s_next = reinterpret_cast<int*>(s_red + 32);  // clip index of the tiles this group takes next (found two tiles ahead) ...
    MelClip* s_clip = reinterpret_cast<MelClip*>(s_red + 36);  // ... and the clips themselves
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + 56);  // completion of the sample tile in flight
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(MEL_THREADS) : "memory"); };

    const int tid = threadIdx.x % MEL_THREADS;  // within the group
    const int lane = tid & 31, warp = tid >> 5;
    const int t = tid & 15, pr = tid >> 4;  // lane within the frame pair, frame pair within the tile

    for (int i = threadIdx.x; i < HANN_PAD; i += blockDim.x) s_hann[i] = i < MEL_NFFT ? __ldg(&p.hann[i]) : 0.f;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tw256[i] = __ldg(&p.tw256[i]);
    for (int i = threadIdx.x; i < 129; i += blockDim.x) s_tw512[i] = __ldg(&p.tw512[i]);
    for (int i = threadIdx.x; i < FBW_FLOATS; i += blockDim.x) s_fbw[i] = __ldg(&p.fbw[i]);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_fbstart[i] = __ldg(&p.fb_start[i]);

    float4* scr4 = s_scr + pr * SCR_F4;                    // complex pairs: (re_A, re_B, im_A, im_B)
    float2* scrp = reinterpret_cast<float2*>(scr4);        // power spectrum pairs (P_A, P_B)
    // Software pipeline over the CTA's tiles.  Per tile: (A) the previous tile's features leave as whole rows, (B) this tile's
    // windowed samples are read into registers, barrier, (C) the NEXT tile's samples start travelling into the same buffer
    // (cp.async, landing during the transforms), (D) transforms, filterbank, parking of the features, barrier.  Two CTA barriers
    // per tile, none of them waiting for global memory.
    const int tile_first = p.tile_lo + GROUPS * blockIdx.x + group, tile_step = GROUPS * gridDim.x;
    constexpr int ROWS_PER_INSTR = 32 / MEL_TILE, ROWS_PER_WARP = MEL_BINS / (MEL_THREADS / 32);
    const int row_fr = lane % MEL_TILE, row_pq = row_fr >> 1;
    const float* parked = reinterpret_cast<const float*>(s_scr) + (size_t)row_pq * (4 * SCR_F4) + 2 * PARK_F2 + (row_fr & 1);
    float* prev_o = nullptr;  // row stores of the previous tile: first element of this lane's column, row stride, in-range flag
    int prev_T = 0;
    bool prev_on = false;
    auto store_rows = [&]() {  // a warp instruction writes 32 / MEL_TILE whole mel rows of the tile, lane -> (row, frame)
#pragma unroll
        for (int i = 0; i < ROWS_PER_WARP / ROWS_PER_INSTR; i++) {
            // two rows per instruction (16-frame tiles): rows m and m + 8, whose parked words fall in different banks
            const int m = ROWS_PER_INSTR == 2 ? ROWS_PER_WARP * warp + (i & 7) + 16 * (i >> 3) + 8 * (lane / MEL_TILE) : ROWS_PER_WARP * warp + i;
            const float val = parked[2 * ((m + row_pq) & 127)];
            if (prev_on) prev_o[(size_t)m * prev_T] = val;
        }
    };
    if (tid == 0) {
        ptx::mbar_init(s_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {  // clips of the first two tiles (later ones are looked up two tiles ahead): one round of loads for the whole warp
                      // instead of two binary searches by one thread (13 dependent L2 round trips, ~9 us before the first tile moved)
        const int ta = min(tile_first, p.total_tiles - 1), tb = min(tile_first + tile_step, p.total_tiles - 1);
        int ca, cb;
        if (p.batch <= 32 * 4) {
            int na = 0, nb = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int first = lane + 32 * k < p.batch ? __ldg(&p.clips[lane + 32 * k].tile0) : 0x7fffffff;
                na += first <= ta ? 1 : 0;
                nb += first <= tb ? 1 : 0;
            }
            ca = __reduce_add_sync(0xffffffffu, na) - 1;
            cb = __reduce_add_sync(0xffffffffu, nb) - 1;
        } else {
            ca = __shfl_sync(0xffffffffu, lane == 0 ? find_clip(p.clips, p.batch, ta) : 0, 0);
            cb = __shfl_sync(0xffffffffu, lane == 1 ? find_clip(p.clips, p.batch, tb) : 0, 1);
        }
        if (lane == 0) { s_next[0] = ca; s_next[1] = cb; }
        constexpr int CW = (int)(sizeof(MelClip) / 4);
        if (lane < 2 * CW)
            reinterpret_cast<int*>(&s_clip[lane / CW])[lane % CW] = __ldg(reinterpret_cast<const int*>(&p.clips[lane < CW ? ca : cb]) + lane % CW);
    }
    __syncthreads();
    if (tile_first < p.total_tiles) {
        const MelClip c0 = s_clip[0];
        stage_tile(s_x, c0, p.pcm, (tile_first - c0.tile0) * MEL_TILE, tid, s_bar);
        ptx::mbar_wait(s_bar, 0);
    }
    uint32_t stage_phase = 1;  // parity of the next staged tile's completion
    __syncthreads();
    // Clip of the tile two steps ahead, found by the last warp without anybody waiting for it: the candidates' first-tile numbers
    // are loaded (one per lane and 32 clips) while the transforms run, counted at the end of the tile, the clip's descriptor is
    // fetched then and put into shared memory at the top of the next iteration.  (One thread's binary search, six dependent L2
    // round trips per tile, made its warp the last at every tile barrier.)
    constexpr int LOOKUP_WARP = MEL_THREADS / 32 - 1, LOOKUP_SLOTS = 4;
    int pend_cj = -1, pend_word = 0;
    int it = 0;
    for (int tile = tile_first; tile < p.total_tiles; tile += tile_step, it++) {
        const int ci = s_next[it & 1];
        const MelClip c = s_clip[it & 1];
        if (warp == LOOKUP_WARP && pend_cj >= 0) {  // found during the previous tile: the clip of tile + tile_step
            if (lane == 0) s_next[(it + 1) & 1] = pend_cj;
            if (lane < (int)(sizeof(MelClip) / 4)) reinterpret_cast<int*>(&s_clip[(it + 1) & 1])[lane] = pend_word;
            pend_cj = -1;
        }
        const int f0 = (tile - c.tile0) * MEL_TILE;
        const int nF = c.n / MEL_HOP + 1;  // frames incl. the one that is dropped (max runs over it, Q3)
        if (prev_o != nullptr) store_rows();

        const int fl = 2 * pr;  // local frames fl (A) and fl + 1 (B)
        const float* xa = s_x + fl * MEL_HOP + 2 * t;
        const float* xb = xa + MEL_HOP;

        // ---- pass 1: 16-point DFTs over n2 for n1 = t  (z[n] = y[2n] + i y[2n+1], n = t + 16 n2) ----
        C2 v[16];
#pragma unroll
        for (int n2 = 0; n2 < 13; n2++) {
            const float2 a = *reinterpret_cast<const float2*>(xa + 32 * n2);
            const float2 b = *reinterpret_cast<const float2*>(xb + 32 * n2);
            const float2 hw = *reinterpret_cast<const float2*>(s_hann + 2 * t + 32 * n2);  // taps of samples 2n and 2n + 1
            v[n2].re = mul2(make_float2(a.x, b.x), splat(hw.x));
            v[n2].im = mul2(make_float2(a.y, b.y), splat(hw.y));
        }
        v[13].re = v[13].im = v[14].re = v[14].im = v[15].re = v[15].im = make_float2(0.f, 0.f);
        group_sync();  // every lane has its samples (the buffer is free) and has read the previous tile's parked features
        {
            const int ntile = tile + tile_step;
            if (ntile < p.total_tiles) {
                const MelClip cn = s_clip[(it + 1) & 1];
                stage_tile(s_x, cn, p.pcm, (ntile - cn.tile0) * MEL_TILE, tid, s_bar);
            }
        }
        const int tile2 = tile + 2 * tile_step;  // the tile whose clip is looked up during this one
        const bool lookup = warp == LOOKUP_WARP && tile2 < p.total_tiles;
        int first_tile[LOOKUP_SLOTS];  // candidates' first tiles, or (very large batches: one lane's dependent search after all) the answer
        if (lookup) {
            if (p.batch <= 32 * LOOKUP_SLOTS) {
#pragma unroll
                for (int k = 0; k < LOOKUP_SLOTS; k++)
                    first_tile[k] = lane + 32 * k < p.batch ? __ldg(&p.clips[lane + 32 * k].tile0) : 0x7fffffff;
            } else {
                first_tile[0] = lane == 0 ? find_clip(p.clips, p.batch, tile2) : 0;
            }
        }
        fft16(v);
        // row t, column k2 lives at float4 index 16 t + (k2 ^ t): conflict-free for the row-wise stores and the column-wise loads
        scr4[t * 16 + t] = make_float4(v[0].re.x, v[0].re.y, v[0].im.x, v[0].im.y);
#pragma unroll
        for (int k2 = 1; k2 < 16; k2++) {
            const float2 tw = s_tw256[k2 * 16 + t];
            const C2 z = cmulc(v[rev16(k2)], tw.x, tw.y);
            scr4[t * 16 + (k2 ^ t)] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
        }
        __syncwarp();
        // ---- pass 2: 16-point DFTs over n1 for k2 = t ----
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const float4 q = scr4[n1 * 16 + (t ^ n1)];
            v[n1].re = make_float2(q.x, q.y);
            v[n1].im = make_float2(q.z, q.w);
        }
        __syncwarp();
        fft16(v);  // Z[16 k1 + t] is now v[rev16(k1)]
        // ---- real split: bins k and 256-k from Z[k], Z[256-k];  P = |2 X|^2.  Lane t owns k = t + 16 j (j < 8): Z[k] is its own
        //      v[rev16(j)]; the partner Z[256 - k] = Z[(16 - t) + 16 (15 - j)] lives in lane 16 - t at k1 = 15 - j (lane 0 pairs with
        //      itself, at k1 = 16 - j) and comes over with warp shuffles instead of a second trip through shared memory (a 4 KB
        //      store and 4 KB of gathers per frame pair: the kernel is bound by shared-memory wavefronts) ----
        // The spectrum goes straight into the pair's scratch block (every lane of the pair has read its transposed values before
        // the __syncwarp above), pair by pair of bins, so that the 32 power values are never all live in registers.
        const int partner = (lane & 16) | ((16 - t) & 15);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = t + 16 * j;
            const C2 mine = v[rev16(j)];
            const C2 give0 = v[rev16((16 - j) & 15)], give = v[rev16(15 - j)];  // what this lane hands out: lane 0 / the others
            float2 zmr, zmi;
            zmr.x = __shfl_sync(0xffffffffu, t == 0 ? give0.re.x : give.re.x, partner);
            zmr.y = __shfl_sync(0xffffffffu, t == 0 ? give0.re.y : give.re.y, partner);
            zmi.x = __shfl_sync(0xffffffffu, t == 0 ? give0.im.x : give.im.x, partner);
            zmi.y = __shfl_sync(0xffffffffu, t == 0 ? give0.im.y : give.im.y, partner);
            const float2 w = s_tw512[k];
            const float2 zkr = mine.re, zki = mine.im;
            const float2 er = cadd(zkr, zmr), ei = csub(zki, zmi);
            C2 o;
            o.re = cadd(zki, zmi);
            o.im = csub(zmr, zkr);
            const C2 tw = cmulc(o, w.x, w.y);
            const float2 ax = cadd(er, tw.re), ay = cadd(ei, tw.im), bx = csub(er, tw.re), by = csub(ei, tw.im);
            float2 pk = fma2(ax, ax, mul2(ay, ay)), pm = fma2(bx, bx, mul2(by, by));
            if (j == 0 && t == 0) {  // bins 0, 256 (from Z[0]) and 128 (from Z[128]) are real-valued special cases
                const float2 z0r = v[rev16(0)].re, z0i = v[rev16(0)].im;
                const float2 dc = mul2(splat(2.f), cadd(z0r, z0i)), ny = mul2(splat(2.f), csub(z0r, z0i));
                pk = mul2(dc, dc);
                pm = mul2(ny, ny);
                const float2 zhr = v[rev16(8)].re, zhi = v[rev16(8)].im;
                scrp[128] = mul2(splat(4.f), fma2(zhr, zhr, mul2(zhi, zhi)));
            }
            scrp[k] = pk;
            scrp[256 - k] = pm;
        }
        __syncwarp();
        // ---- sparse filterbank, log10, scale; lane t owns mel bins t, t+16, ... of both frames ----
        const bool in_max_a = (f0 + fl) < nF, in_max_b = (f0 + fl + 1) < nF;
        const bool in_out_a = (f0 + fl) < c.frames, in_out_b = (f0 + fl + 1) < c.frames;
        float2 accs[MEL_ROUNDS];
        accs[0] = fb_round<0>(scrp, s_fbw, s_fbstart, t);
        accs[1] = fb_round<1>(scrp, s_fbw, s_fbstart, t);
        accs[2] = fb_round<2>(scrp, s_fbw, s_fbstart, t);
        accs[3] = fb_round<3>(scrp, s_fbw, s_fbstart, t);
        accs[4] = fb_round<4>(scrp, s_fbw, s_fbstart, t);
        accs[5] = fb_round<5>(scrp, s_fbw, s_fbstart, t);
        accs[6] = fb_round<6>(scrp, s_fbw, s_fbstart, t);
        accs[7] = fb_round<7>(scrp, s_fbw, s_fbstart, t);
        float lmax = -INFINITY, lmin = INFINITY;
        // The features leave through shared memory: lane t holds mel rows t, t + 16, ... of two frames, and storing them from here
        // touches 16 different 128-byte lines per warp instruction (16 wavefronts each, 64 per frame: a quarter of the kernel's
        // shared-memory / L1 wavefronts).  They are parked in the unused tail of the pair's scratch block instead (row m of pair pr
        // at word 2 ((m + pr) & 127), which makes both the parking stores and the row-wise reads below conflict-free) and written
        // out after the tile barrier as whole rows: one 128-byte line per warp instruction.
        float2* park = reinterpret_cast<float2*>(scr4) + PARK_F2;
#pragma unroll
        for (int j = 0; j < MEL_ROUNDS; j++) {
            const int m = t + 16 * j;
            const float La = 0.30102999566398120f * __log2f(fmaxf(accs[j].x, 1e-10f));
            const float Lb = 0.30102999566398120f * __log2f(fmaxf(accs[j].y, 1e-10f));
            park[(m + pr) & 127] = make_float2(fmaf(0.25f, La, 1.0f), fmaf(0.25f, Lb, 1.0f));
            if (in_max_a) lmax = fmaxf(lmax, La);
            if (in_max_b) lmax = fmaxf(lmax, Lb);
            if (in_out_a) lmin = fminf(lmin, La);
            if (in_out_b) lmin = fminf(lmin, Lb);
        }

        // ---- tile reductions ----
        lmax = warp_max(lmax);
        lmin = -warp_max(-lmin);
        if (lane == 0) { s_red[warp] = lmax; s_red[16 + warp] = lmin; }
        if (lookup) {  // clips are ordered by first tile: the clip of tile2 is the last one that starts at or before it
            int cj;
            if (p.batch <= 32 * LOOKUP_SLOTS) {
                int cnt = 0;
#pragma unroll
                for (int k = 0; k < LOOKUP_SLOTS; k++) cnt += first_tile[k] <= tile2 ? 1 : 0;
                cj = __reduce_add_sync(0xffffffffu, cnt) - 1;
            } else {
                cj = __shfl_sync(0xffffffffu, first_tile[0], 0);
            }
            pend_cj = cj;
            pend_word = lane < (int)(sizeof(MelClip) / 4) ? __ldg(reinterpret_cast<const int*>(&p.clips[cj]) + lane) : 0;
        }
        if (tile + tile_step < p.total_tiles) {  // the next tile's samples have landed (long ago)
            ptx::mbar_wait(s_bar, stage_phase);
            stage_phase ^= 1;
        }
        group_sync();  // the tile's features are parked, its reductions written, the next tile's samples visible to every lane
        if (tid == 0) {
            float gm = s_red[0], tm = s_red[16];
#pragma unroll
            for (int w = 1; w < MEL_THREADS / 32; w++) { gm = fmaxf(gm, s_red[w]); tm = fminf(tm, s_red[16 + w]); }
            atomicMax(p.gmax + ci, enc_ordered(gm));
            p.tmin[tile] = tm;
            p.tclip[tile] = ci;
        }
        prev_o = p.out + c.out_off + f0 + row_fr;
        prev_T = c.frames;
        prev_on = f0 + row_fr < c.frames;
    }
    if (prev_o != nullptr) store_rows();
}

// Second pass: clip to (clip max - 8) where a tile needs it (AudioPreprocessing.swift:281-293).  One THREAD per tile decides from
// three coalesced loads (the tile's clip, recorded by the first pass; the clip's maximum; the tile's minimum); almost every tile is
// already above the floor (1-2 % of the tiles of the bench clips are not).  The tiles that need the clamp are collected in shared
// memory and then handled by the whole CTA, one tile at a time: warp w takes 16 mel rows, lane f frame f, all 16 rows in flight.
constexpr int CLAMP_TILES = 64;
__global__ void __launch_bounds__(256) mel_clamp_kernel(float* out, const MelClip* clips, int tile_lo, int total_tiles, const int* gmax,
                                                        const float* tmin, const int* tclip) {
    __shared__ int s_tile[CLAMP_TILES], s_clip[CLAMP_TILES];
    __shared__ float s_lo[CLAMP_TILES];
    __shared__ int s_n;
    ptx::grid_dep_wait();  // launched with programmatic stream serialization behind mel_kernel: its maxima, minima and features are complete
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int tile = tile_lo + blockIdx.x * CLAMP_TILES + threadIdx.x;  // CLAMP_TILES deciders per CTA: about one tile to clamp per CTA
    if (threadIdx.x < CLAMP_TILES && tile < total_tiles) {
        const int ci = tclip[tile];
        const float lo = mel_decode_max(gmax[ci]) - 8.0f;
        if (tmin[tile] < lo) {
            const int k = atomicAdd(&s_n, 1);
            s_tile[k] = tile;
            s_clip[k] = ci;
            s_lo[k] = lo;
        }
    }
    __syncthreads();
    const int n = s_n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < n; k++) {
        const MelClip c = clips[s_clip[k]];
        const int f0 = (s_tile[k] - c.tile0) * MEL_TILE;
        const float lo_s = fmaf(0.25f, s_lo[k], 1.0f);
        for (int fr = lane; fr < MEL_TILE; fr += 32) {
            if (f0 + fr >= c.frames) break;
            float* o = out + c.out_off + f0 + fr + (size_t)(warp * 16) * c.frames;
            float v[16];
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = o[(size_t)m * c.frames];
#pragma unroll
            for (int m = 0; m < 16; m++) o[(size_t)m * c.frames] = fmaxf(v[m], lo_s);
        }
    }
}

// slaney mel scale helpers, Float arithmetic like the reference (AudioPreprocessing.swift:61-164)
float hz_to_mel(float hz) {
    const float min_log_hz = 1000.0f, min_log_mel = 15.0f, logstep = 27.0f / logf(6.4f);
    return hz >= min_log_hz ? min_log_mel + logf(hz / min_log_hz) * logstep : 3.0f * hz / 200.0f;
}
float mel_to_hz(float mel) {
    const float min_log_hz = 1000.0f, min_log_mel = 15.0f, logstep = logf(6.4f) / 27.0f;
    return mel >= min_log_mel ? min_log_hz * expf((mel - min_log_mel) * logstep) : 200.0f * mel / 3.0f;
}

}  // namespace

void mel_filterbank_host(float* fb) {
    const float mlo = hz_to_mel(0.0f), mhi = hz_to_mel(8000.0f);
    float edge[MEL_BINS + 2];
    for (int i = 0; i < MEL_BINS + 2; i++) edge[i] = mel_to_hz(mlo + (float)i * (mhi - mlo) / (float)(MEL_BINS + 1));
    for (int m = 0; m < MEL_BINS; m++) {
        const float lo = edge[m], ce = edge[m + 1], hi = edge[m + 2];
        const float norm = 2.0f / (hi - lo);
        for (int k = 0; k < MEL_NFREQ; k++) {
            const float f = (float)k * 16000.0f / (float)MEL_FFT;
            const float up = (f - lo) / (ce - lo), dn = (hi - f) / (hi - ce);
            const float tri = fmaxf(0.0f, fminf(up, dn));
            fb[m * MEL_NFREQ + k] = tri * norm;
        }
    }
}

void mel_tables_create(MelTables* t) {
    memset(t, 0, sizeof(*t));
    std::vector<float> hann(MEL_NFFT);
    const float pi = 3.14159265358979323846f;
    for (int i = 0; i < MEL_NFFT; i++) hann[i] = 0.5f * (1.0f - cosf(2.0f * pi * (float)i / (float)MEL_NFFT));
    std::vector<float2> tw256(256), tw512(129);
    for (int k2 = 0; k2 < 16; k2++)
        for (int tt = 0; tt < 16; tt++) {
            const double a = -2.0 * M_PI * (double)(tt * k2) / 256.0;
            tw256[k2 * 16 + tt] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int k = 0; k <= 128; k++) {
        const double a = -2.0 * M_PI * (double)k / 512.0;
        tw512[k] = make_float2((float)cos(a), (float)sin(a));
    }
    std::vector<float> fb((size_t)MEL_BINS * MEL_NFREQ);
    mel_filterbank_host(fb.data());
    // ELL packing: round j holds filters 16j..16j+15, padded to the widest of the round
    int start[MEL_BINS], width[MEL_BINS];
    for (int m = 0; m < MEL_BINS; m++) {
        int a = -1, b = -1;
        for (int k = 0; k < MEL_NFREQ; k++)
            if (fb[m * MEL_NFREQ + k] != 0.0f) { if (a < 0) a = k; b = k; }
        if (a < 0) { a = 0; b = 0; }
        start[m] = a;
        width[m] = b - a + 1;
    }
    t->fb_round_off[0] = 0;
    for (int j = 0; j < MEL_ROUNDS; j++) {
        int mw = 1;
        for (int i = 0; i < 16; i++) mw = std::max(mw, width[16 * j + i]);
        // A filter narrower than the round's width may start up to (mw - width) bins early (zero weights in front): the slack is
        // used to give the 16 lanes of a round start bins that are distinct modulo 16 (or equal), so that the 8-byte gathers of a
        // frame pair's spectrum touch each bank pair once: the gathers took 92 wavefronts per frame pair instead of 40 (ncu: 6 M
        // bank conflicts, 20 % of the kernel's shared-memory wavefronts).  Leading zero taps leave the sums bit-identical.
        int lo[16], hi[16], pick[16];
        for (int i = 0; i < 16; i++) {
            hi[i] = std::min(start[16 * j + i], MEL_NFREQ - mw);  // keep padded taps in range
            lo[i] = std::max(0, start[16 * j + i] - (mw - width[16 * j + i]));
            lo[i] = std::min(lo[i], hi[i]);
        }
        int owner[16];  // start bin that holds each residue, -1 = free
        for (int r = 0; r < 16; r++) owner[r] = -1;
        std::function<bool(int)> place = [&](int i) -> bool {
            if (i == 16) return true;
            for (int s0 = hi[i]; s0 >= lo[i]; s0--) {
                const int r = s0 & 15;
                if (owner[r] >= 0 && owner[r] != s0) continue;
                const int was = owner[r];
                owner[r] = s0;
                pick[i] = s0;
                if (place(i + 1)) return true;
                owner[r] = was;
            }
            return false;
        };
        if (place(0)) for (int i = 0; i < 16; i++) start[16 * j + i] = pick[i];
        else for (int i = 0; i < 16; i++) start[16 * j + i] = hi[i];
        t->fb_round_off[j + 1] = t->fb_round_off[j] + mw;
    }
    t->fb_rows = t->fb_round_off[MEL_ROUNDS];
    for (int j = 0; j < MEL_ROUNDS; j++)  // the kernel's tap loops are unrolled for exactly these widths
        Q3_CHECK(t->fb_round_off[j + 1] - t->fb_round_off[j] == FB_WIDTHS[j] && t->fb_rows == FB_ROWS, 2,
                 "mel filterbank round widths differ from the kernel's compile-time table");
    std::vector<float> fbw((size_t)FBW_FLOATS, 0.0f);  // [lane i][round j: FB_POFF[j] + w]
    for (int j = 0; j < MEL_ROUNDS; j++) {
        const int mw = t->fb_round_off[j + 1] - t->fb_round_off[j];
        for (int i = 0; i < 16; i++) {
            const int m = 16 * j + i;
            for (int w = 0; w < mw; w++) fbw[(size_t)i * FBW_STRIDE + FB_POFF[j] + w] = fb[m * MEL_NFREQ + start[m] + w];
        }
    }
    Q3_CUDA(cudaMalloc(&t->hann, sizeof(float) * MEL_NFFT));
    Q3_CUDA(cudaMalloc(&t->tw256, sizeof(float2) * 256));
    Q3_CUDA(cudaMalloc(&t->tw512, sizeof(float2) * 129));
    Q3_CUDA(cudaMalloc(&t->fbw, sizeof(float) * fbw.size()));
    Q3_CUDA(cudaMalloc(&t->fb_start, sizeof(int) * MEL_BINS));
    Q3_H2D_SYNC(t->hann, hann.data(), sizeof(float) * MEL_NFFT);
    Q3_H2D_SYNC(t->tw256, tw256.data(), sizeof(float2) * 256);
    Q3_H2D_SYNC(t->tw512, tw512.data(), sizeof(float2) * 129);
    Q3_H2D_SYNC(t->fbw, fbw.data(), sizeof(float) * fbw.size());
    Q3_H2D_SYNC(t->fb_start, start, sizeof(int) * MEL_BINS);
    Q3_CUDA(cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mel_smem_bytes(t->fb_rows)));
}

void mel_tables_destroy(MelTables* t) {
    cudaFree(t->hann);
    cudaFree(t->tw256);
    cudaFree(t->tw512);
    cudaFree(t->fbw);
    cudaFree(t->fb_start);
    memset(t, 0, sizeof(*t));
}

void mel_launch(const MelTables& t, const float* d_pcm, float* d_out, const MelClip* d_clips, int batch, int total_tiles,
                int* d_gmax, float* d_tmin, int num_sms, cudaStream_t st) {
    mel_launch_range(t, d_pcm, d_out, d_clips, 0, batch, 0, total_tiles, total_tiles, d_gmax, d_tmin, num_sms, st);
}

// Clips [clip0, clip1) only: their tiles are [tile_lo, tile_hi) of the batch's `total_tiles` (MelClip::tile0 stays batch-wide).
void mel_launch_range(const MelTables& t, const float* d_pcm, float* d_out, const MelClip* d_clips, int clip0, int clip1, int tile_lo,
                      int tile_hi, int total_tiles, int* d_gmax, float* d_tmin, int num_sms, cudaStream_t st) {
    const int batch = clip1 - clip0;
    if (batch <= 0 || tile_hi <= tile_lo) return;
    d_clips += clip0;  // clip indices inside the kernels are relative to the range, and so is the maximum array
    d_gmax += clip0;
    MelParams p;
    p.hann = t.hann;
    p.tw256 = t.tw256;
    p.tw512 = t.tw512;
    p.fbw = t.fbw;
    p.fb_start = t.fb_start;
    for (int j = 0; j <= MEL_ROUNDS; j++) p.fb_round_off[j] = t.fb_round_off[j];
    p.fb_rows = t.fb_rows;
    p.pcm = d_pcm;
    p.out = d_out;
    p.clips = d_clips;
    p.batch = batch;
    p.tile_lo = tile_lo;
    p.total_tiles = tile_hi;
    p.gmax = d_gmax;
    p.tmin = d_tmin;
    p.tclip = reinterpret_cast<int*>(d_tmin + total_tiles);  // d_tmin holds 2 * total_tiles words: [minimum | clip]
    Q3_CUDA(cudaMemsetAsync(d_gmax, 0x80, sizeof(int) * batch, st));
    const int n_tiles = tile_hi - tile_lo;
    const int grid = std::min((n_tiles + GROUPS - 1) / GROUPS, num_sms);  // one CTA of GROUPS groups per SM
    mel_kernel<<<grid, GROUPS * MEL_THREADS, mel_smem_bytes(t.fb_rows), st>>>(p);
    Q3_CUDA(cudaGetLastError());
    {   // the clamp pass as a programmatic dependent: its launch latency hides behind the main kernel's tail
        const bool was = pdl_enabled();
        pdl_enabled() = true;
        try {
            launch_kernel(mel_clamp_kernel, dim3((n_tiles + CLAMP_TILES - 1) / CLAMP_TILES), dim3(256), 0, st, d_out, d_clips, tile_lo, tile_hi,
                          (const int*)d_gmax, (const float*)d_tmin, (const int*)p.tclip);
        } catch (...) {
            pdl_enabled() = was;
            throw;
        }
        pdl_enabled() = was;
    }
}

}  // namespace q3
