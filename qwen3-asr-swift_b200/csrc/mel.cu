// mel.cu — K1: fused log-mel frontend for sm_100a.
//
// Replaces WhisperFeatureExtractor.extractFeatures
// (/root/reference/Sources/Qwen3ASR/AudioPreprocessing.swift:169-317): reflect padding (:174-192),
// Hann window + zero-pad 400->512 + real FFT with vDSP's 2x scale + power (:209-250), the 257->128
// slaney filterbank product (:263-268), clip/log10/clip-to-(max-8)/scale (:275-293), drop of the
// last frame and the 120000-frame cap (:296-313), [128, T] mel-major output (:315-316).
//
// B200 design.  One persistent CTA (128 threads, 4 per SM) walks 24-frame tiles.  A tile's samples
// (4080 floats, reflect padding resolved while loading) are staged in shared memory with 128-bit
// coalesced loads; sixteen lanes share one frame and two frames share a warp, so every
// synchronisation inside the transform is a __syncwarp.  The 512-point real FFT is the 256-point
// complex FFT of the even/odd packed frame, done as 16x16 (two in-register radix-16 passes with one
// shared-memory transpose), followed by the real-split step that yields bins k and 256-k together.
// The vDSP factor 2 cancels against the 1/2 of the split step (|2X|^2 = |e + w o|^2).  The filterbank
// is applied as an ELL-packed sparse product (504 non-zeros instead of 32896 MACs per frame).  The
// kernel writes 0.25*log10(mel)+1 unclamped through a shared-memory transpose tile (128-byte row
// stores), an ordered-int atomicMax per clip and a per-tile minimum; the second kernel applies the
// max-8 clamp only to tiles whose minimum is below it (exact, because clamp and the monotone affine
// map commute), so in the common case the features are written once and never re-read.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "mel.cuh"

namespace q3 {

namespace {

constexpr int TPF = 16;                                // lanes per frame
constexpr int MEL_THREADS = 128;
constexpr int FR_PER_IT = (MEL_THREADS / TPF);         // 8 frames in flight per CTA
constexpr int TILE_SAMPLES = (MEL_TILE - 1) * MEL_HOP + MEL_NFFT;  // 5360
constexpr int SX_FLOATS = TILE_SAMPLES + 16;
constexpr int XROW = 17;                               // padded row (float2) of the 16x16 transpose
constexpr int SCR_FLOATS = 2 * TPF * XROW + 16;        // 544 floats per frame + 16: the two frames of a warp sit 16 banks apart
constexpr int OUT_STRIDE = MEL_TILE + 1;

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
    int total_tiles;
    int* gmax;
    float* tmin;
    int* tclip;  // clip of every tile (for the clamp pass)
};

__host__ __device__ constexpr int mel_smem_bytes(int fb_rows) {
    return (SX_FLOATS + 2 * 256 + 2 * 130 + fb_rows * 16 + 128 + FR_PER_IT * SCR_FLOATS + MEL_BINS * OUT_STRIDE + 32) * 4;
}

// Widths (taps) of the eight filterbank rounds: a property of the slaney filterbank at 16 kHz / 512 points / 128 bins, checked
// against the computed table in mel_tables_create.  Compile-time constants let the tap loops unroll completely (the loop control
// was 15 % of the kernel's instructions).
template <int J> struct FbRound;
template <> struct FbRound<0> { static constexpr int W = 2, OFF = 0; };
template <> struct FbRound<1> { static constexpr int W = 2, OFF = 2; };
template <> struct FbRound<2> { static constexpr int W = 2, OFF = 4; };
template <> struct FbRound<3> { static constexpr int W = 3, OFF = 6; };
template <> struct FbRound<4> { static constexpr int W = 4, OFF = 9; };
template <> struct FbRound<5> { static constexpr int W = 6, OFF = 13; };
template <> struct FbRound<6> { static constexpr int W = 9, OFF = 19; };
template <> struct FbRound<7> { static constexpr int W = 12, OFF = 28; };
constexpr int FB_ROWS = 40;
constexpr int FB_WIDTHS[MEL_ROUNDS] = {2, 2, 2, 3, 4, 6, 9, 12};

template <int J>
__device__ __forceinline__ float fb_round(const float* __restrict__ scr, const float* __restrict__ s_fbw, const int* __restrict__ s_fbstart, int t) {
    const float* pp = scr + s_fbstart[t + 16 * J];
    const float* wp = s_fbw + FbRound<J>::OFF * 16 + t;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < FbRound<J>::W; w++) acc = fmaf(pp[w], wp[w * 16], acc);
    return acc;
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}

// forward radix-4 butterfly, in place: (a,b,c,d) <- DFT4
__device__ __forceinline__ void bfly4(float2& a, float2& b, float2& c, float2& d) {
    const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
    a = cadd(t0, t2);
    c = csub(t0, t2);
    b = make_float2(t1.x + t3.y, t1.y - t3.x);
    d = make_float2(t1.x - t3.y, t1.y + t3.x);
}

// forward 16-point DFT in registers.  Input x[n] natural order; on return X[k] is at x[4*(k&3) + (k>>2)].
__device__ __forceinline__ void fft16(float2 (&x)[16]) {
    constexpr float C = 0.92387953251128674f, S = 0.38268343236508977f, R = 0.70710678118654752f;
#pragma unroll
    for (int n1 = 0; n1 < 4; n1++) bfly4(x[n1], x[n1 + 4], x[n1 + 8], x[n1 + 12]);
    // x[n1 + 4*k2] *= W16^(n1*k2)
    x[1 + 4] = cmul(x[1 + 4], make_float2(C, -S));    // W^1
    x[1 + 8] = cmul(x[1 + 8], make_float2(R, -R));    // W^2
    x[1 + 12] = cmul(x[1 + 12], make_float2(S, -C));  // W^3
    x[2 + 4] = cmul(x[2 + 4], make_float2(R, -R));    // W^2
    x[2 + 8] = make_float2(x[2 + 8].y, -x[2 + 8].x);  // W^4 = -i
    x[2 + 12] = cmul(x[2 + 12], make_float2(-R, -R)); // W^6
    x[3 + 4] = cmul(x[3 + 4], make_float2(S, -C));    // W^3
    x[3 + 8] = cmul(x[3 + 8], make_float2(-R, -R));   // W^6
    x[3 + 12] = cmul(x[3 + 12], make_float2(-C, S));  // W^9
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

__global__ void __launch_bounds__(MEL_THREADS, 4) mel_kernel(const MelParams p) {
    extern __shared__ float4 smem4[];
    float* s_x = reinterpret_cast<float*>(smem4);
    float2* s_tw256 = reinterpret_cast<float2*>(s_x + SX_FLOATS);
    float2* s_tw512 = s_tw256 + 256;
    float* s_fbw = reinterpret_cast<float*>(s_tw512 + 130);
    int* s_fbstart = reinterpret_cast<int*>(s_fbw + p.fb_rows * 16);
    float* s_scr = reinterpret_cast<float*>(s_fbstart + 128);
    float* s_out = s_scr + FR_PER_IT * SCR_FLOATS;
    float* s_red = s_out + MEL_BINS * OUT_STRIDE;
    int* s_next = reinterpret_cast<int*>(s_red + 8);  // clip of the tile this CTA takes next (searched one tile ahead)

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int t = lane & 15, half = lane >> 4;

    for (int i = tid; i < 256; i += MEL_THREADS) s_tw256[i] = __ldg(&p.tw256[i]);
    for (int i = tid; i < 129; i += MEL_THREADS) s_tw512[i] = __ldg(&p.tw512[i]);
    for (int i = tid; i < p.fb_rows * 16; i += MEL_THREADS) s_fbw[i] = __ldg(&p.fbw[i]);
    for (int i = tid; i < 128; i += MEL_THREADS) s_fbstart[i] = __ldg(&p.fb_start[i]);

    // this lane's window taps: samples 2t + 32*n2 (+1), n2 = 0..12
    float2 hw[13];
#pragma unroll
    for (int n2 = 0; n2 < 13; n2++) {
        const int idx = 2 * t + 32 * n2;
        hw[n2] = idx < MEL_NFFT ? make_float2(__ldg(&p.hann[idx]), __ldg(&p.hann[idx + 1])) : make_float2(0.f, 0.f);
    }
    float* scr = s_scr + (warp * 2 + half) * SCR_FLOATS;
    float2* scr2 = reinterpret_cast<float2*>(scr);
    __syncthreads();

    if (tid == 0) *s_next = find_clip(p.clips, p.batch, min((int)blockIdx.x, p.total_tiles - 1));
    __syncthreads();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int ci = *s_next;
        const MelClip c = p.clips[ci];
        const int f0 = (tile - c.tile0) * MEL_TILE;
        const int nF = c.n / MEL_HOP + 1;  // frames incl. the one that is dropped (max runs over it, Q3)
        const float* x = p.pcm + c.in_off;
        const int n = c.n;

        // ---- stage the tile's padded samples: pad index p0 + i, i < TILE_SAMPLES ----
        const int p0 = f0 * MEL_HOP;
        if (p0 >= MEL_NFFT / 2 && p0 + SX_FLOATS <= MEL_NFFT / 2 + n) {
            const float4* src = reinterpret_cast<const float4*>(x + (p0 - MEL_NFFT / 2));
            float4* dst = reinterpret_cast<float4*>(s_x);
#pragma unroll 4
            for (int i = tid; i < SX_FLOATS / 4; i += MEL_THREADS) dst[i] = __ldg(src + i);
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
                s_x[i] = v;
            }
        }
        __syncthreads();
        // the binary search for the next tile's clip (dependent L2 loads) overlaps this tile's transforms
        if (tid == MEL_THREADS - 1 && tile + (int)gridDim.x < p.total_tiles) *s_next = find_clip(p.clips, p.batch, tile + gridDim.x);

        float lmax = -INFINITY, lmin = INFINITY;
#pragma unroll 1
        for (int it = 0; it < MEL_TILE / FR_PER_IT; it++) {
            const int fl = it * (FR_PER_IT / 2) + warp + half * (MEL_TILE / 2);  // local frame
            const float* xs = s_x + fl * MEL_HOP + 2 * t;

            // ---- pass 1: 16-point DFTs over n2 for n1 = t  (z[n] = y[2n] + i y[2n+1], n = t + 16 n2) ----
            float2 v[16];
#pragma unroll
            for (int n2 = 0; n2 < 13; n2++) {
                const float2 a = *reinterpret_cast<const float2*>(xs + 32 * n2);
                v[n2] = make_float2(a.x * hw[n2].x, a.y * hw[n2].y);
            }
            v[13] = v[14] = v[15] = make_float2(0.f, 0.f);
            fft16(v);
            scr2[t * XROW] = v[0];
#pragma unroll
            for (int k2 = 1; k2 < 16; k2++) scr2[t * XROW + k2] = cmul(v[rev16(k2)], s_tw256[k2 * 16 + t]);
            __syncwarp();
            // ---- pass 2: 16-point DFTs over n1 for k2 = t ----
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) v[n1] = scr2[n1 * XROW + t];
            __syncwarp();
            fft16(v);
#pragma unroll
            for (int k1 = 0; k1 < 16; k1++) scr2[16 * k1 + t] = v[rev16(k1)];  // Z[16 k1 + t]
            __syncwarp();
            // ---- real split: bins k and 256-k from Z[k], Z[256-k];  P = |2 X|^2 ----
            float pk[8], pm[8];
            float p128 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = t + 16 * j;
                const float2 zk = scr2[k];
                const float2 zm = scr2[(256 - k) & 255];
                const float2 w = s_tw512[k];
                const float2 e = make_float2(zk.x + zm.x, zk.y - zm.y);
                const float2 o = make_float2(zk.y + zm.y, zm.x - zk.x);
                const float2 tw = cmul(o, w);
                const float ax = e.x + tw.x, ay = e.y + tw.y, bx = e.x - tw.x, by = e.y - tw.y;
                pk[j] = fmaf(ax, ax, ay * ay);
                pm[j] = fmaf(bx, bx, by * by);
            }
            if (t == 0) {
                const float2 z0 = scr2[0];
                const float dc = 2.f * (z0.x + z0.y), ny = 2.f * (z0.x - z0.y);
                pk[0] = dc * dc;
                pm[0] = ny * ny;
                const float2 zh = scr2[128];
                p128 = 4.f * fmaf(zh.x, zh.x, zh.y * zh.y);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = t + 16 * j;
                scr[k] = pk[j];
                scr[256 - k] = pm[j];
            }
            if (t == 0) scr[128] = p128;
            __syncwarp();
            // ---- sparse filterbank, log10, scale; lane t owns mel bins t, t+16, ... ----
            const bool in_max = (f0 + fl) < nF;
            const bool in_out = (f0 + fl) < c.frames;
            float accs[MEL_ROUNDS];
            accs[0] = fb_round<0>(scr, s_fbw, s_fbstart, t);
            accs[1] = fb_round<1>(scr, s_fbw, s_fbstart, t);
            accs[2] = fb_round<2>(scr, s_fbw, s_fbstart, t);
            accs[3] = fb_round<3>(scr, s_fbw, s_fbstart, t);
            accs[4] = fb_round<4>(scr, s_fbw, s_fbstart, t);
            accs[5] = fb_round<5>(scr, s_fbw, s_fbstart, t);
            accs[6] = fb_round<6>(scr, s_fbw, s_fbstart, t);
            accs[7] = fb_round<7>(scr, s_fbw, s_fbstart, t);
#pragma unroll
            for (int j = 0; j < MEL_ROUNDS; j++) {
                const int m = t + 16 * j;
                const float acc = accs[j];
                const float L = 0.30102999566398120f * __log2f(fmaxf(acc, 1e-10f));
                s_out[m * OUT_STRIDE + fl] = fmaf(0.25f, L, 1.0f);
                if (in_max) lmax = fmaxf(lmax, L);
                if (in_out) lmin = fminf(lmin, L);
            }
            __syncwarp();
        }

        // ---- tile reductions + transposed store ----
        lmax = warp_max(lmax);
        lmin = -warp_max(-lmin);
        if (lane == 0) { s_red[warp] = lmax; s_red[4 + warp] = lmin; }
        __syncthreads();
        if (tid == 0) {
            const float gm = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
            const float tm = fminf(fminf(s_red[4], s_red[5]), fminf(s_red[6], s_red[7]));
            atomicMax(p.gmax + ci, enc_ordered(gm));
            p.tmin[tile] = tm;
            p.tclip[tile] = ci;
        }
        const int T = c.frames;
        if (lane < MEL_TILE && f0 + lane < T) {
            float* o = p.out + c.out_off + f0 + lane;
#pragma unroll 4
            for (int m = warp; m < MEL_BINS; m += MEL_THREADS / 32) o[(size_t)m * T] = s_out[m * OUT_STRIDE + lane];
        }
        __syncthreads();
    }
}

// Second pass: clip to (clip max - 8) where a tile needs it (AudioPreprocessing.swift:281-293).  One THREAD per tile decides from
// three coalesced loads (the tile's clip, recorded by the first pass; the clip's maximum; the tile's minimum); almost every tile is
// already above the floor (1-2 % of the tiles of the bench clips are not).  The tiles that need the clamp are collected in shared
// memory and then handled by the whole CTA, one tile at a time: warp w takes 16 mel rows, lane f frame f, all 16 rows in flight.
constexpr int CLAMP_TILES = 64;
__global__ void __launch_bounds__(256) mel_clamp_kernel(float* out, const MelClip* clips, int total_tiles, const int* gmax,
                                                        const float* tmin, const int* tclip) {
    __shared__ int s_tile[CLAMP_TILES], s_clip[CLAMP_TILES];
    __shared__ float s_lo[CLAMP_TILES];
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int tile = blockIdx.x * CLAMP_TILES + threadIdx.x;  // CLAMP_TILES deciders per CTA: about one tile to clamp per CTA
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
        if (lane < MEL_TILE && f0 + lane < c.frames) {
            float* o = out + c.out_off + f0 + lane + (size_t)(warp * 16) * c.frames;
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
        for (int i = 0; i < 16; i++)
            if (start[16 * j + i] + mw > MEL_NFREQ) start[16 * j + i] = MEL_NFREQ - mw;  // keep padded taps in range
        t->fb_round_off[j + 1] = t->fb_round_off[j] + mw;
    }
    t->fb_rows = t->fb_round_off[MEL_ROUNDS];
    for (int j = 0; j < MEL_ROUNDS; j++)  // the kernel's tap loops are unrolled for exactly these widths
        Q3_CHECK(t->fb_round_off[j + 1] - t->fb_round_off[j] == FB_WIDTHS[j] && t->fb_rows == FB_ROWS, 2,
                 "mel filterbank round widths differ from the kernel's compile-time table");
    std::vector<float> fbw((size_t)t->fb_rows * 16, 0.0f);
    for (int j = 0; j < MEL_ROUNDS; j++) {
        const int mw = t->fb_round_off[j + 1] - t->fb_round_off[j];
        for (int i = 0; i < 16; i++) {
            const int m = 16 * j + i;
            for (int w = 0; w < mw; w++) fbw[(size_t)(t->fb_round_off[j] + w) * 16 + i] = fb[m * MEL_NFREQ + start[m] + w];
        }
    }
    Q3_CUDA(cudaMalloc(&t->hann, sizeof(float) * MEL_NFFT));
    Q3_CUDA(cudaMalloc(&t->tw256, sizeof(float2) * 256));
    Q3_CUDA(cudaMalloc(&t->tw512, sizeof(float2) * 129));
    Q3_CUDA(cudaMalloc(&t->fbw, sizeof(float) * fbw.size()));
    Q3_CUDA(cudaMalloc(&t->fb_start, sizeof(int) * MEL_BINS));
    Q3_CUDA(cudaMemcpy(t->hann, hann.data(), sizeof(float) * MEL_NFFT, cudaMemcpyHostToDevice));
    Q3_CUDA(cudaMemcpy(t->tw256, tw256.data(), sizeof(float2) * 256, cudaMemcpyHostToDevice));
    Q3_CUDA(cudaMemcpy(t->tw512, tw512.data(), sizeof(float2) * 129, cudaMemcpyHostToDevice));
    Q3_CUDA(cudaMemcpy(t->fbw, fbw.data(), sizeof(float) * fbw.size(), cudaMemcpyHostToDevice));
    Q3_CUDA(cudaMemcpy(t->fb_start, start, sizeof(int) * MEL_BINS, cudaMemcpyHostToDevice));
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
    if (batch <= 0 || total_tiles <= 0) return;
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
    p.total_tiles = total_tiles;
    p.gmax = d_gmax;
    p.tmin = d_tmin;
    p.tclip = reinterpret_cast<int*>(d_tmin + total_tiles);  // d_tmin holds 2 * total_tiles words: [minimum | clip]
    Q3_CUDA(cudaMemsetAsync(d_gmax, 0x80, sizeof(int) * batch, st));
    const int grid = std::min(total_tiles, num_sms * 4);
    mel_kernel<<<grid, MEL_THREADS, mel_smem_bytes(t.fb_rows), st>>>(p);
    mel_clamp_kernel<<<(total_tiles + CLAMP_TILES - 1) / CLAMP_TILES, 256, 0, st>>>(d_out, d_clips, total_tiles, d_gmax, d_tmin, p.tclip);
    Q3_CUDA(cudaGetLastError());
}

}  // namespace q3
