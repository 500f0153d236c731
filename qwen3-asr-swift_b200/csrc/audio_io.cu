// audio_io.cu — the front door of the batched path: WAV PCM16 parser, sample-rate converter,
// long-form windowing.
//
//   AudioFileLoader.loadWAV     /root/reference/Sources/AudioCommon/AudioFileLoader.swift:70-157  -> wav_parse (same checks in
//                               the same order; the reference's SecurityHardeningTests.swift:83-190 are ported in tests/test_audio_io.py)
//   AudioFileLoader.resample    AudioFileLoader.swift:159-213 (AVAudioConverter, an Apple framework: its filter is not in the
//                               reference) -> a polyphase windowed-sinc converter with a stated design, run on the GPU; the output
//                               LENGTH follows the reference (floor(n * out / in), :190-191)
//   Qwen3ASRModel.transcribe    resamples to 16 kHz when sampleRate != 16000 (AudioPreprocessing.swift:323-337) -> batch_upload_sr
//                               converts such clips on the device, straight into the packed sample buffer the mel kernel reads
//   long-form                   BASELINE config 5: fixed windows, each an independent utterance -> longform_plan
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "model.h"

namespace q3 {

// ------------------------------------------------------------------------------------------
// WAV (RIFF / WAVE, PCM 16-bit): first channel only, samples / 32768
// ------------------------------------------------------------------------------------------
namespace {
inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
const char* const kInvalidWav = "Invalid WAV file format";  // AudioLoadError.invalidWAVFile, AudioFileLoader.swift:227
}  // namespace

// Returns the number of frames; writes min(frames, cap) samples when out != nullptr.
size_t wav_parse(const uint8_t* data, size_t size, float* out, size_t cap, int* sample_rate) {
    Q3_CHECK(data != nullptr, Q3ASR_ERR_INVALID, "wav_parse: null data");
    Q3_CHECK(size > 44, Q3ASR_ERR_INVALID, kInvalidWav);                       // :74-76
    Q3_CHECK(memcmp(data, "RIFF", 4) == 0, Q3ASR_ERR_INVALID, kInvalidWav);    // :79-82
    Q3_CHECK(memcmp(data + 8, "WAVE", 4) == 0, Q3ASR_ERR_INVALID, kInvalidWav);  // :85-88
    const uint16_t audio_format = rd16(data + 20), channels = rd16(data + 22), bits = rd16(data + 34);  // :91-94
    const uint32_t rate = rd32(data + 24);
    Q3_CHECK(audio_format == 1, Q3ASR_ERR_INVALID, "Unsupported audio format: Not PCM format");  // :96-98
    Q3_CHECK(channels > 0, Q3ASR_ERR_INVALID, kInvalidWav);                                        // :100-102
    Q3_CHECK(bits == 16, Q3ASR_ERR_INVALID, "Unsupported audio format: Not 16-bit");              // :104-106
    // find the data chunk (:109-127): the scan starts at byte 36 and every chunk advance is validated
    size_t off = 36;
    bool found = false;
    uint32_t chunk_size = 0;
    while (off + 8 < size) {
        const uint32_t sz = rd32(data + off + 4);
        if (memcmp(data + off, "data", 4) == 0) {
            off += 8;
            chunk_size = sz;
            found = true;
            break;
        }
        const size_t next = off + 8 + (size_t)sz;
        Q3_CHECK(next >= off && next <= size, Q3ASR_ERR_INVALID, kInvalidWav);
        off = next;
    }
    Q3_CHECK(found, Q3ASR_ERR_INVALID, kInvalidWav);                                   // :130-132
    Q3_CHECK(off <= size && off + (size_t)chunk_size <= size, Q3ASR_ERR_INVALID, kInvalidWav);  // :134-136
    const size_t frame_bytes = 2 * (size_t)channels;
    const size_t frames = (size_t)chunk_size / frame_bytes;  // :139-143
    if (sample_rate) *sample_rate = (int)rate;
    if (out != nullptr) {
        const uint8_t* s = data + off;
        const size_t n = frames < cap ? frames : cap;
        for (size_t i = 0; i < n; i++) out[i] = (float)(int16_t)rd16(s + i * frame_bytes) / 32768.0f;  // :146-154
    }
    return frames;
}

// ------------------------------------------------------------------------------------------
// Sample-rate conversion.  in -> out with L = out/g, M = in/g (g = gcd):
//   fc   = 0.945 * min(1, L/M)                      cutoff, as a fraction of the INPUT Nyquist rate
//   half = 24 / fc                                  support half-width in input samples (24 zero crossings)
//   h(t) = fc * sinc(fc * t) * I0(10 * sqrt(1 - (t/half)^2)) / I0(10)      |t| < half      (Kaiser, beta = 10)
//   y[j] = sum_{k=-K}^{K+1} tap[p][k] * x[i0 + k],  i0 = floor(j*M/L), p = (j*M) mod L, tap[p][k] = h(p/L - k) / sum_k h(p/L - k)
// (x = 0 outside the clip; K = ceil(half); taps in double, stored as float; the sum runs in ascending k with fp32 FMAs).
// oracle/audio_io.py restates exactly this.
// ------------------------------------------------------------------------------------------
size_t resample_len(size_t n, int in_rate, int out_rate) {
    if (in_rate == out_rate) return n;
    return (size_t)((double)n * ((double)out_rate / (double)in_rate));  // AudioFileLoader.swift:190-191
}

namespace {
double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 200; k++) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}
}  // namespace

void resample_design(int in_rate, int out_rate, int* L_out, int* M_out, int* K_out, std::vector<float>* taps) {
    Q3_CHECK(in_rate > 0 && out_rate > 0 && in_rate <= 768000 && out_rate <= 768000, Q3ASR_ERR_INVALID, "resample: bad sample rate");
    const int g = std::gcd(in_rate, out_rate);
    const int L = out_rate / g, M = in_rate / g;
    Q3_CHECK(L <= 4096, Q3ASR_ERR_INVALID, "resample: the rate pair needs more than 4096 filter phases");
    const double fc = 0.945 * std::min(1.0, (double)L / (double)M);
    const double half = 24.0 / fc;
    const int K = (int)ceil(half);
    const int nt = 2 * K + 2;
    const double beta = 10.0, i0b = bessel_i0(beta);
    taps->assign((size_t)L * nt, 0.0f);
    std::vector<double> h(nt);
    for (int p = 0; p < L; p++) {
        double sum = 0.0;
        for (int k = -K; k <= K + 1; k++) {
            const double t = (double)p / (double)L - (double)k;
            double v = 0.0;
            if (fabs(t) < half) {
                const double a = M_PI * fc * t;
                const double sinc = fabs(a) < 1e-12 ? 1.0 : sin(a) / a;
                const double r = t / half;
                v = fc * sinc * bessel_i0(beta * sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
            }
            h[k + K] = v;
            sum += v;
        }
        for (int i = 0; i < nt; i++) (*taps)[(size_t)p * nt + i] = (float)(h[i] / sum);
    }
    *L_out = L;
    *M_out = M;
    *K_out = K;
}

namespace {
// One output sample per thread.  The input window of neighbouring outputs overlaps almost entirely (L1-resident), the tap
// table is read-only and small: the kernel moves 4*n + 4*n_out bytes of HBM traffic per clip.
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, long long n, const float* __restrict__ taps, int L, int M,
                                                       int K, float* __restrict__ y, long long n_out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const long long jm = j * (long long)M;
    const long long i0 = jm / L;
    const int p = (int)(jm - i0 * L);
    const int nt = 2 * K + 2;
    const float* tp = taps + (size_t)p * nt;
    float acc = 0.f;
    const long long lo = i0 - K;
    if (lo >= 0 && lo + nt <= n) {  // interior: no bounds checks, four taps per step (nt = 2K + 2 is even)
        const float* xp = x + lo;
        int t = 0;
        for (; t + 4 <= nt; t += 4) {
            acc = fmaf(__ldg(tp + t), __ldg(xp + t), acc);
            acc = fmaf(__ldg(tp + t + 1), __ldg(xp + t + 1), acc);
            acc = fmaf(__ldg(tp + t + 2), __ldg(xp + t + 2), acc);
            acc = fmaf(__ldg(tp + t + 3), __ldg(xp + t + 3), acc);
        }
        for (; t < nt; t++) acc = fmaf(__ldg(tp + t), __ldg(xp + t), acc);
    } else {
        for (int t = 0; t < nt; t++) {
            const long long i = lo + t;
            const float v = (i >= 0 && i < n) ? __ldg(x + i) : 0.f;
            acc = fmaf(__ldg(tp + t), v, acc);
        }
    }
    y[j] = acc;
}
}  // namespace

const ResampleTab& resample_tab(Handle* h, int in_rate, int out_rate) {
    const std::pair<int, int> key(in_rate, out_rate);
    auto it = h->resample_tabs.find(key);
    if (it != h->resample_tabs.end()) return it->second;
    ResampleTab t;
    std::vector<float> taps;
    resample_design(in_rate, out_rate, &t.L, &t.M, &t.K, &taps);
    t.taps.total = &h->dev_bytes;
    t.taps.reserve(taps.size() * sizeof(float));
    Q3_CUDA(cudaMemcpyAsync(t.taps.p, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    Q3_CUDA(cudaStreamSynchronize(h->stream));  // `taps` is a pageable local
    return h->resample_tabs.emplace(key, t).first->second;
}

void resample_device(Handle* h, const float* d_in, size_t n, int in_rate, int out_rate, float* d_out, size_t n_out, cudaStream_t st) {
    if (n_out == 0) return;
    const ResampleTab& t = resample_tab(h, in_rate, out_rate);
    resample_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(d_in, (long long)n, t.taps.as<float>(), t.L, t.M, t.K, d_out,
                                                                      (long long)n_out);
    Q3_CUDA(cudaGetLastError());
    h->launches++;
}

// host in, host out (AudioFileLoader.resample as a call of its own)
void resample_host(Handle* h, const float* in, size_t n, int in_rate, int out_rate, float* out, size_t cap, size_t* n_out) {
    Q3_CHECK(in != nullptr && n_out != nullptr && in_rate > 0 && out_rate > 0, Q3ASR_ERR_INVALID, "resample: bad argument");
    const size_t m = resample_len(n, in_rate, out_rate);
    *n_out = m;
    if (out == nullptr) return;
    Q3_CHECK(cap >= m, Q3ASR_ERR_INVALID, "resample: output buffer too small");
    if (in_rate == out_rate || n == 0) {  // AudioFileLoader.swift:160: returned unchanged
        memcpy(out, in, sizeof(float) * n);
        return;
    }
    h->rs_in.reserve(sizeof(float) * n);
    h->rs_out.reserve(sizeof(float) * std::max<size_t>(m, 1));
    h->rs_stage.reserve(sizeof(float) * std::max(n, m));
    memcpy(h->rs_stage.p, in, sizeof(float) * n);
    Q3_CUDA(cudaMemcpyAsync(h->rs_in.p, h->rs_stage.p, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    resample_device(h, h->rs_in.as<float>(), n, in_rate, out_rate, h->rs_out.as<float>(), m, h->stream);
    Q3_CUDA(cudaMemcpyAsync(h->rs_stage.p, h->rs_out.p, sizeof(float) * m, cudaMemcpyDeviceToHost, h->stream));
    Q3_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(out, h->rs_stage.p, sizeof(float) * m);
}

// ------------------------------------------------------------------------------------------
// Long-form windowing: [k*W, min((k+1)*W, n)); a tail shorter than min_tail joins the previous window.
// ------------------------------------------------------------------------------------------
int longform_plan(size_t n, size_t window, size_t min_tail, size_t* starts, size_t* lens, int cap) {
    Q3_CHECK(window > 0, Q3ASR_ERR_INVALID, "longform_plan: window must be positive");
    Q3_CHECK(n / window < ((size_t)1 << 24), Q3ASR_ERR_INVALID, "longform_plan: more than 2^24 windows");  // the count is an int; no endless loop
    int count = 0;
    size_t pos = 0;
    while (pos < n) {
        size_t len = std::min(window, n - pos);
        const size_t rest = n - pos - len;
        if (rest > 0 && rest < min_tail) len += rest;
        if (count < cap && starts && lens) {
            starts[count] = pos;
            lens[count] = len;
        }
        count++;
        pos += len;
    }
    return count;
}

}  // namespace q3

// ---- host-only C entry points (no handle: errors go to a thread-local message) ----
namespace {
thread_local std::string g_io_error;
template <typename F>
int io_guarded(F&& fn) {
    try {
        fn();
        return Q3ASR_OK;
    } catch (const q3::Error& e) {
        g_io_error = e.what();
        return e.code > 0 ? e.code : Q3ASR_ERR_INVALID;
    } catch (const std::exception& e) {
        g_io_error = e.what();
        return Q3ASR_ERR_INVALID;
    }
}
}  // namespace

extern "C" {

const char* q3asr_io_last_error(void) { return g_io_error.c_str(); }

int q3asr_wav_parse(const uint8_t* data, size_t size, float* samples, size_t cap, size_t* n_samples, int* sample_rate) {
    return io_guarded([&]() {
        const size_t n = q3::wav_parse(data, size, samples, cap, sample_rate);
        if (n_samples) *n_samples = n;
    });
}

int q3asr_wav_load(const char* path, float* samples, size_t cap, size_t* n_samples, int* sample_rate) {
    return io_guarded([&]() {
        Q3_CHECK(path != nullptr, Q3ASR_ERR_INVALID, "wav_load: null path");
        std::unique_ptr<FILE, int (*)(FILE*)> f(fopen(path, "rb"), fclose);  // closed on every path out, bad_alloc included
        if (!f) throw q3::Error(Q3ASR_ERR_IO, std::string("wav_load: cannot open ") + path);
        std::vector<uint8_t> buf;
        uint8_t tmp[1 << 16];
        size_t r;
        while ((r = fread(tmp, 1, sizeof(tmp), f.get())) > 0) buf.insert(buf.end(), tmp, tmp + r);
        f.reset();
        const size_t n = q3::wav_parse(buf.data(), buf.size(), samples, cap, sample_rate);
        if (n_samples) *n_samples = n;
    });
}

int q3asr_wav_write(const char* path, const float* samples, size_t n_samples, int sample_rate) {
    return io_guarded([&]() {  // WAVWriter.write (Sources/AudioCommon/WAVWriter.swift:11-47): mono, 16-bit PCM, 44-byte header
        Q3_CHECK(path != nullptr && (samples != nullptr || n_samples == 0), Q3ASR_ERR_INVALID, "wav_write: null argument");
        Q3_CHECK(sample_rate > 0 && n_samples <= 0x7FFFFFEDull / 2, Q3ASR_ERR_INVALID, "wav_write: bad sample rate or too many samples for a RIFF file");
        const uint32_t data_size = (uint32_t)(n_samples * 2);
        std::vector<uint8_t> buf(44 + (size_t)data_size);
        auto put16 = [&](size_t at, uint32_t v) { buf[at] = (uint8_t)v; buf[at + 1] = (uint8_t)(v >> 8); };
        auto put32 = [&](size_t at, uint32_t v) { put16(at, v & 0xFFFF); put16(at + 2, v >> 16); };
        memcpy(&buf[0], "RIFF", 4);
        put32(4, 36 + data_size);
        memcpy(&buf[8], "WAVEfmt ", 8);
        put32(16, 16);
        put16(20, 1);  // PCM
        put16(22, 1);  // mono
        put32(24, (uint32_t)sample_rate);
        put32(28, (uint32_t)sample_rate * 2);
        put16(32, 2);
        put16(34, 16);
        memcpy(&buf[36], "data", 4);
        put32(40, data_size);
        for (size_t i = 0; i < n_samples; i++) {
            float v = samples[i];
            v = !(v >= -1.0f) ? -1.0f : (v > 1.0f ? 1.0f : v);  // max(-1, min(1, x)); NaN (Swift would trap on it) is written as -1
            put16(44 + 2 * i, (uint32_t)(uint16_t)(int16_t)(v * 32767.0f));  // Int16(Float): truncation toward zero (:42)
        }
        std::unique_ptr<FILE, int (*)(FILE*)> f(fopen(path, "wb"), fclose);
        if (!f) throw q3::Error(Q3ASR_ERR_IO, std::string("wav_write: cannot open ") + path);
        if (fwrite(buf.data(), 1, buf.size(), f.get()) != buf.size()) throw q3::Error(Q3ASR_ERR_IO, std::string("wav_write: short write to ") + path);
    });
}

size_t q3asr_resample_len(size_t n_samples, int in_rate, int out_rate) {
    if (in_rate <= 0 || out_rate <= 0) return 0;
    return q3::resample_len(n_samples, in_rate, out_rate);
}

int q3asr_resample_design(int in_rate, int out_rate, int* L, int* M, int* K, float* taps, size_t cap, size_t* n_taps) {
    return io_guarded([&]() {
        Q3_CHECK(L && M && K && n_taps, Q3ASR_ERR_INVALID, "resample_design: null output");
        std::vector<float> t;
        q3::resample_design(in_rate, out_rate, L, M, K, &t);
        *n_taps = t.size();
        if (taps) {
            Q3_CHECK(cap >= t.size(), Q3ASR_ERR_INVALID, "resample_design: buffer too small");
            memcpy(taps, t.data(), t.size() * sizeof(float));
        }
    });
}

int q3asr_longform_plan(size_t n_samples, size_t window, size_t min_tail, size_t* starts, size_t* lens, int cap, int* count) {
    return io_guarded([&]() {
        Q3_CHECK(count != nullptr && cap >= 0, Q3ASR_ERR_INVALID, "longform_plan: bad argument");
        *count = q3::longform_plan(n_samples, window, min_tail, starts, lens, cap);
    });
}

}  // extern "C"
