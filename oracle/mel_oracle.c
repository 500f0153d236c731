/*
 * mel_oracle.c — CPU restatement of the reference's log-mel frontend.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or
 * executed from the product library (libq3asr.so).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker
 * or as the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (Swift + Apple Accelerate) cannot be built or run on
 * Linux and its tests hold no numeric golden vectors for this path.  The
 * restatement is instead cross-checked (tests/test_oracle_mel.py) against
 * transformers.WhisperFeatureExtractor with the three reference quirks switched off, which
 * pins reflect padding, the periodic Hann window, the slaney filterbank and the
 * log/clamp/scale tail; the quirks (Q1-Q3 below) are then switched on for parity fixtures.
 *
 * Follows /root/reference/Sources/Qwen3ASR/AudioPreprocessing.swift:
 *   :39-53    periodic Hann window, 400 taps, Float arithmetic
 *   :61-164   slaney-scale / slaney-norm triangular filterbank on the 257 bins of the
 *             zero-padded 512-point FFT (Q1), Float arithmetic (logf/expf)
 *   :174-192  reflect padding by 200 samples each side
 *   :195      nFrames = (padded - 400)/160 + 1
 *   :209-250  per frame: window, zero-pad to 512, real FFT, power spectrum
 *   :263-268  dense [T,257]x[257,128] product
 *   :275-293  clip 1e-10, log10, global max, clip to max-8, *0.25+1
 *   :296-313  drop last frame, cap at 120000 frames, transpose to [128,T]
 *
 * Third-party semantics assumed (Apple Accelerate, not in the reference tree):
 *   vDSP_fft_zrip forward returns 2x the mathematical DFT (Q2; the reference documents
 *   this itself at Sources/SpeechWakeWord/KaldiFbank.swift:239-246) and the ASR path does
 *   not compensate, so power is 4x.  vDSP_mmul is taken as a plain ascending-k fp32 sum.
 *   vvlog10f is taken as correctly-rounded-ish log10f.
 *
 * Switches (for the HF cross-check only; defaults = reference behaviour):
 *   fft_size 512|400, vdsp_scale2 1|0, max_before_trim 1|0, precise 0|1 (double FFT).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_FFT 400
#define HOP 160
#define N_MELS 128
#define MAX_FRAMES 120000

typedef struct {
    int fft_size;        /* 512 (reference, Q1) or 400 (HF Whisper) */
    int vdsp_scale2;     /* 1: FFT output is 2x DFT (Q2) */
    int max_before_trim; /* 1: max taken over all nFrames incl. the dropped last one (Q3) */
    int precise;         /* 1: FFT in double (accuracy reference), 0: fp32 like the reference */
} q3o_mel_opts;

/* AudioPreprocessing.swift:41-44 */
void q3o_hann(float *w) {
    const float pi = 3.14159265358979323846f;
    for (int i = 0; i < N_FFT; i++)
        w[i] = 0.5f * (1.0f - cosf(2.0f * pi * (float)i / (float)N_FFT));
}

/* AudioPreprocessing.swift:61-164; out is [128, nbins] row-major */
void q3o_mel_filterbank(int fft_size, float *fb) {
    const int nbins = fft_size / 2 + 1;
    const float fmin = 0.0f, fmax = 16000.0f / 2.0f;
    const float min_log_hz = 1000.0f, min_log_mel = 15.0f;
    const float logstep_hz2mel = 27.0f / logf(6.4f);
    const float logstep_mel2hz = logf(6.4f) / 27.0f;
    float pts[N_MELS + 2], hz[N_MELS + 2], diff[N_MELS + 1];
    float mel_min = fmin < min_log_hz ? 3.0f * fmin / 200.0f
                                      : min_log_mel + logf(fmin / min_log_hz) * logstep_hz2mel;
    float mel_max = fmax < min_log_hz ? 3.0f * fmax / 200.0f
                                      : min_log_mel + logf(fmax / min_log_hz) * logstep_hz2mel;
    for (int i = 0; i < N_MELS + 2; i++) {
        pts[i] = mel_min + (float)i * (mel_max - mel_min) / (float)(N_MELS + 1);
        hz[i] = pts[i] < min_log_mel ? 200.0f * pts[i] / 3.0f
                                     : min_log_hz * expf((pts[i] - min_log_mel) * logstep_mel2hz);
    }
    for (int i = 0; i < N_MELS + 1; i++) diff[i] = hz[i + 1] - hz[i];
    for (int b = 0; b < nbins; b++) {
        float f = (float)b * 16000.0f / (float)fft_size;
        for (int m = 0; m < N_MELS; m++) {
            float down = (f - hz[m]) / diff[m];
            float up = (hz[m + 2] - f) / diff[m + 1];
            float v = fminf(down, up);
            if (v < 0.0f) v = 0.0f;
            float enorm = 2.0f / (hz[m + 2] - hz[m]);
            fb[m * nbins + b] = v * enorm;
        }
    }
}

/* in-place radix-2 complex FFT of length n (power of two), fp32, forward (e^{-i..}) */
static void fft_c32(float *re, float *im, int n) {
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            float t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        for (int k = 0; k < len / 2; k++) {
            double ang = -2.0 * M_PI * (double)k / (double)len;
            float wr = (float)cos(ang), wi = (float)sin(ang);
            for (int s = 0; s < n; s += len) {
                int a = s + k, b = a + len / 2;
                float xr = re[b] * wr - im[b] * wi;
                float xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
        }
    }
}

/* power spectrum of one windowed frame y[0..399] on fft_size/2+1 bins */
static void frame_power(const float *y, const q3o_mel_opts *o, float *pw) {
    const int N = o->fft_size, nb = N / 2 + 1;
    const float sc = o->vdsp_scale2 ? 2.0f : 1.0f;
    if (o->precise || N != 512) {
        /* direct DFT in double: accuracy reference and the non-power-of-two HF mode */
        for (int k = 0; k < nb; k++) {
            double sr = 0.0, si = 0.0;
            for (int j = 0; j < N_FFT && j < N; j++) {
                double ang = -2.0 * M_PI * (double)((long)k * j % N) / (double)N;
                sr += (double)y[j] * cos(ang);
                si += (double)y[j] * sin(ang);
            }
            sr *= sc; si *= sc;
            pw[k] = (float)(sr * sr + si * si);
        }
        return;
    }
    /* AudioPreprocessing.swift:221-249 — even/odd packing into a 256-point complex FFT,
     * real-FFT split, 2x scale, DC in realp[0] and Nyquist in imagp[0]. */
    float re[256], im[256];
    for (int i = 0; i < 256; i++) {
        int a = 2 * i, b = 2 * i + 1;
        re[i] = a < N_FFT ? y[a] : 0.0f;
        im[i] = b < N_FFT ? y[b] : 0.0f;
    }
    fft_c32(re, im, 256);
    float dc = (re[0] + im[0]) * sc, ny = (re[0] - im[0]) * sc;
    pw[0] = dc * dc;
    pw[256] = ny * ny;
    for (int k = 1; k < 256; k++) {
        int m = 256 - k;
        float er = 0.5f * (re[k] + re[m]), ei = 0.5f * (im[k] - im[m]);
        float orr = 0.5f * (im[k] + im[m]), oi = -0.5f * (re[k] - re[m]);
        double ang = -2.0 * M_PI * (double)k / 512.0;
        float wr = (float)cos(ang), wi = (float)sin(ang);
        float xr = (er + (orr * wr - oi * wi)) * sc;
        float xi = (ei + (orr * wi + oi * wr)) * sc;
        pw[k] = xr * xr + xi * xi;
    }
}

/* number of output frames for n samples: min(n/160, 120000)  (:195, :296, :304) */
int q3o_mel_frames(long n) {
    long t = n / HOP;
    return (int)(t > MAX_FRAMES ? MAX_FRAMES : t);
}

/*
 * out: [128, q3o_mel_frames(n)] row-major (mel-major), fp32.  Returns frames or <0.
 * opts may be NULL (reference behaviour).
 */
int q3o_mel(const float *x, long n, const q3o_mel_opts *opts, float *out) {
    q3o_mel_opts o = {512, 1, 1, 0};
    if (opts) o = *opts;
    if (n <= 0) return -1;
    const int nb = o.fft_size / 2 + 1;
    const long npad = n + 2 * (N_FFT / 2);
    const long nF = (npad - N_FFT) / HOP + 1;
    float *pad = (float *)malloc(sizeof(float) * (size_t)npad);
    float *fb = (float *)malloc(sizeof(float) * N_MELS * (size_t)nb);
    float *lm = (float *)malloc(sizeof(float) * (size_t)nF * N_MELS);
    float hann[N_FFT], y[N_FFT], pw[257];
    if (!pad || !fb || !lm) { free(pad); free(fb); free(lm); return -2; }
    q3o_hann(hann);
    q3o_mel_filterbank(o.fft_size, fb);
    /* :178-192 */
    for (long i = 0; i < N_FFT / 2; i++) {
        long s = N_FFT / 2 - i; if (s > n - 1) s = n - 1; if (s < 0) s = 0;
        pad[i] = x[s];
    }
    memcpy(pad + N_FFT / 2, x, sizeof(float) * (size_t)n);
    for (long i = 0; i < N_FFT / 2; i++) {
        long s = n - 2 - i; if (s < 0) s = 0;
        pad[N_FFT / 2 + n + i] = x[s];
    }
    /* :209-268 (frame loop + dense product), :275-279 (clip, log10) */
    for (long f = 0; f < nF; f++) {
        const float *src = pad + f * HOP;
        for (int j = 0; j < N_FFT; j++) y[j] = src[j] * hann[j];
        frame_power(y, &o, pw);
        for (int m = 0; m < N_MELS; m++) {
            const float *w = fb + (size_t)m * nb;
            float acc = 0.0f;
            for (int k = 0; k < nb; k++) acc += pw[k] * w[k];
            if (acc < 1e-10f) acc = 1e-10f;
            lm[f * N_MELS + m] = log10f(acc);
        }
    }
    /* :281-293 */
    long nmax = o.max_before_trim ? nF : nF - 1;
    float g = -INFINITY;
    for (long i = 0; i < nmax * N_MELS; i++) if (lm[i] > g) g = lm[i];
    const float lo = g - 8.0f;
    /* :296-316 */
    long T = nF - 1; if (T > MAX_FRAMES) T = MAX_FRAMES;
    for (long f = 0; f < T; f++)
        for (int m = 0; m < N_MELS; m++) {
            float v = lm[f * N_MELS + m];
            if (v < lo) v = lo;
            out[(size_t)m * T + f] = v * 0.25f + 1.0f;
        }
    free(pad); free(fb); free(lm);
    return (int)T;
}

/* Batched helper for the CPU baseline: clips processed one after another (the reference's
 * frame loop is serial, AudioPreprocessing.swift:209); callers may parallelise over clips. */
int q3o_mel_batch(const float *const *x, const long *n, int batch, float *const *out) {
    for (int b = 0; b < batch; b++) {
        int t = q3o_mel(x[b], n[b], NULL, out[b]);
        if (t < 0) return t;
    }
    return 0;
}
