// mel.cuh — interface of the log-mel frontend (K1).
//
// Replaces WhisperFeatureExtractor.extractFeatures / extractFeaturesRaw
// (/root/reference/Sources/Qwen3ASR/AudioPreprocessing.swift:169-317, 347-470).
#pragma once
#include "common.cuh"

namespace q3 {

constexpr int MEL_NFFT = 400;
constexpr int MEL_HOP = 160;
constexpr int MEL_BINS = 128;
constexpr int MEL_FFT = 512;       // zero-padded FFT length (reference quirk Q1)
constexpr int MEL_NFREQ = 257;
constexpr int MEL_MAX_FRAMES = 120000;
constexpr int MEL_TILE = 16;       // frames per tile: one pass of a CTA of 8 * MEL_TILE threads (8 frame pairs, 16 lanes each); 16 or 32
constexpr int MEL_ROUNDS = 8;      // 128 mel bins / 16 lanes per frame

// One clip of a batch, all offsets in elements.
struct MelClip {
    long long in_off;    // start of the clip's samples in the packed pcm buffer (multiple of 4)
    long long out_off;   // start of the clip's [128, frames] block in the output buffer
    int n;               // samples
    int frames;          // frames kept: min(n/160, 120000)
    int tile0;           // first tile id of this clip
    int ntiles;          // ceil((n/160 + 1) / MEL_TILE)
};

// Constant tables living in device memory (built once per handle by mel_tables_create).
struct MelTables {
    float* hann;        // [400]
    float2* tw256;      // [16 k2][16 t]: exp(-2 pi i t k2 / 256)
    float2* tw512;      // [129]: exp(-2 pi i k / 512)
    float* fbw;         // ELL filterbank weights [sum(maxw)][16]
    int* fb_start;      // [128] first non-zero FFT bin of each filter
    int fb_round_off[MEL_ROUNDS + 1];  // offsets (in 16-wide rows) of each round in fbw
    int fb_rows;        // sum(maxw)
};

void mel_tables_create(MelTables* t);   // host: builds window/twiddles/filterbank, uploads
void mel_tables_destroy(MelTables* t);
// host copy of the dense [128,257] filterbank (debug / tests)
void mel_filterbank_host(float* fb);

// Launches the log-mel kernels for `batch` clips described by d_clips (device copy of MelClip[batch]).
//   d_pcm    packed samples
//   d_out    packed [128, frames_b] fp32 blocks
//   d_gmax   [batch] int scratch (ordered-int encoded clip maxima)
//   d_tmin   [2 * total_tiles] words of scratch: per-tile minima (float), then per-tile clip indices (int)
// total_tiles = sum of ntiles.  Two launches: main kernel, clamp pass (exits early per tile).
void mel_launch(const MelTables& t, const float* d_pcm, float* d_out, const MelClip* d_clips, int batch,
                int total_tiles, int* d_gmax, float* d_tmin, int num_sms, cudaStream_t st);
// the same for the clips [clip0, clip1) of the batch only (their tiles are [tile_lo, tile_hi) of total_tiles): the upload overlaps
// the first groups' transforms with the remaining copies
void mel_launch_range(const MelTables& t, const float* d_pcm, float* d_out, const MelClip* d_clips, int clip0, int clip1, int tile_lo,
                      int tile_hi, int total_tiles, int* d_gmax, float* d_tmin, int num_sms, cudaStream_t st);

// Decoded clip maximum of log10(mel) (after mel_launch): used by consumers that fuse the clamp.
__host__ __device__ inline float mel_decode_max(int key) {
    int i = key >= 0 ? key : key ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f;
    memcpy(&f, &i, 4);
    return f;
#endif
}

}  // namespace q3
