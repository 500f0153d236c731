// frontend.cu — host-side planning for the mel stage: packs a ragged batch of clips into one buffer and
// enumerates the MEL_TILE-frame tiles the kernel walks.
#include "model.h"

namespace q3 {

// AudioPreprocessing.swift:195 (nFrames = n/160 + 1), :296 (drop last), :304 (cap 120000)
int mel_frames_for(size_t n) {
    const size_t t = n / MEL_HOP;
    return (int)(t > (size_t)MEL_MAX_FRAMES ? MEL_MAX_FRAMES : t);
}

MelPlan mel_plan(const size_t* n_samples, int batch) {
    MelPlan p;
    p.clips.resize(batch);
    long long in_off = 0, out_off = 0;
    int tile = 0;
    for (int b = 0; b < batch; b++) {
        MelClip& c = p.clips[b];
        c.n = (int)n_samples[b];
        c.frames = mel_frames_for(n_samples[b]);
        c.in_off = in_off;
        c.out_off = out_off;
        c.tile0 = tile;
        const int nF = c.n / MEL_HOP + 1;
        c.ntiles = (nF + MEL_TILE - 1) / MEL_TILE;
        tile += c.ntiles;
        in_off += ((long long)c.n + 31) & ~31LL;  // 128-byte aligned clip starts (float4 loads)
        out_off += (((long long)MEL_BINS * c.frames) + 31) & ~31LL;
    }
    p.pcm_floats = in_off;
    p.out_floats = out_off;
    p.total_tiles = tile;
    return p;
}

}  // namespace q3
