// transcribe_batch.cpp — the `speech transcribe-batch` front door over the batched B200 path
// (/root/reference/Sources/AudioCLILib/TranscribeBatchCommand.swift:45-139, 217-243): walks a directory, loads every WAV
// (AudioFileLoader.loadWAV), and — unlike the reference's serial per-file loop (:82-125) — hands the files to the
// library in batches, so the utterances of one batch share every kernel launch.  Files at other sample rates are converted
// on the device; recordings longer than --window-seconds are cut into windows (each an independent utterance) and
// their texts joined.  Output lines follow the reference: JSONL {"file","text","time","rtf","duration"} or the
// bracketed progress lines, then the aggregate block.  A file's "time" is its share (by audio duration) of the
// batch it ran in.
//
//   transcribe_batch <inputDir> [--output-dir D] [--model 0.6B|1.7B] [--model-dir DIR] [--language L] [--extensions wav]
//                    [--jsonl] [--batch N] [--max-tokens N] [--window-seconds S] [--device N] [--list]
// Without --model-dir the weights are random-init (this repo has no checkpoint offline); --list only prints the files.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <set>

#include "qwen3_asr.hpp"

using namespace qwen3asr;
namespace fs = std::filesystem;

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static std::string json_escape(const std::string& s) {  // TranscribeBatchCommand.swift:101-103 (+ control characters)
    std::string o;
    for (unsigned char c : s) {
        if (c == '\\') o += "\\\\";
        else if (c == '"') o += "\\\"";
        else if (c == '\n') o += "\\n";
        else if (c == '\r') o += "\\r";
        else if (c == '\t') o += "\\t";
        else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o += (char)c;
    }
    return o;
}

// TranscribeBatchCommand.swift:217-234: regular files with a listed extension, hidden files skipped, sorted by file name
static std::vector<fs::path> find_audio_files(const std::string& dir, const std::string& extensions) {
    std::set<std::string> exts;
    size_t a = 0;
    while (a <= extensions.size()) {
        size_t b = extensions.find(',', a);
        if (b == std::string::npos) b = extensions.size();
        std::string e = extensions.substr(a, b - a);
        std::transform(e.begin(), e.end(), e.begin(), [](unsigned char c) { return (char)std::tolower(c); });
        if (!e.empty()) exts.insert(e);
        a = b + 1;
    }
    std::vector<fs::path> files;
    std::error_code ec;
    for (fs::recursive_directory_iterator it(dir, fs::directory_options::skip_permission_denied, ec), end; !ec && it != end; it.increment(ec)) {
        const std::string name = it->path().filename().string();
        if (!name.empty() && name[0] == '.') {
            if (it->is_directory(ec)) it.disable_recursion_pending();
            continue;
        }
        if (!it->is_regular_file(ec)) continue;
        std::string e = it->path().extension().string();
        if (!e.empty()) e = e.substr(1);
        std::transform(e.begin(), e.end(), e.begin(), [](unsigned char c) { return (char)std::tolower(c); });
        if (exts.count(e)) files.push_back(it->path());
    }
    std::sort(files.begin(), files.end(), [](const fs::path& x, const fs::path& y) { return x.filename().string() < y.filename().string(); });
    return files;
}

int main(int argc, char** argv) {
    std::string inputDir, outputDir, model = "0.6B", modelDir, extensions = "wav,flac,mp3";
    std::optional<std::string> language;
    bool jsonl = false, listOnly = false;
    int batch = 64, maxTokens = 448, device = 0;
    float windowSeconds = 30.f;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--output-dir") outputDir = val();
        else if (a == "--model" || a == "-m") model = val();
        else if (a == "--model-dir") modelDir = val();
        else if (a == "--language") language = val();
        else if (a == "--extensions") extensions = val();
        else if (a == "--jsonl") jsonl = true;
        else if (a == "--list") listOnly = true;
        else if (a == "--batch") batch = std::max(1, atoi(val().c_str()));
        else if (a == "--max-tokens") maxTokens = std::max(1, atoi(val().c_str()));
        else if (a == "--window-seconds") windowSeconds = (float)atof(val().c_str());
        else if (a == "--device") device = atoi(val().c_str());
        else if (!a.empty() && a[0] != '-' && inputDir.empty()) inputDir = a;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (inputDir.empty()) { fprintf(stderr, "usage: transcribe_batch <inputDir> [options]\n"); return 2; }
    const auto files = find_audio_files(inputDir, extensions);
    if (files.empty()) { printf("No audio files found in %s\n", inputDir.c_str()); return 0; }
    printf("Found %zu audio files\n", files.size());
    if (listOnly) {
        for (const auto& f : files) printf("%s\n", f.filename().string().c_str());
        return 0;
    }
    const ASRModelSize size = detectModelSize(model);
    printf("Loading model (%s): %s\n", size == ASRModelSize::large ? "1.7B" : "0.6B", modelDir.empty() ? "random-init weights" : modelDir.c_str());
    const double loadStart = now_s();
    std::unique_ptr<Qwen3ASRModel> asr;
    try {
        asr = modelDir.empty() ? Qwen3ASRModel::randomInit(size, 20260418, device) : Qwen3ASRModel::fromPretrained(model, modelDir, device);
    } catch (const AudioModelError& e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    const double loadTime = now_s() - loadStart;
    printf("  Model loaded in %.2fs\n", loadTime);
    if (!outputDir.empty()) fs::create_directories(outputDir);

    // load every file; a file that fails is reported in its place and skipped (TranscribeBatchCommand.swift:118-124)
    struct Item { std::string name; AudioFileLoader::Wav wav; std::string error; std::vector<size_t> windows; double duration = 0; std::string text; double time = 0; };
    std::vector<Item> items(files.size());
    struct Unit { size_t item, start, len; };
    std::vector<Unit> units;
    for (size_t i = 0; i < files.size(); i++) {
        items[i].name = files[i].stem().string();
        try {
            items[i].wav = AudioFileLoader::loadWAV(files[i].string());
            const auto& w = items[i].wav;
            if (w.sampleRate <= 0 || q3asr_resample_len(w.samples.size(), w.sampleRate, 16000) < 160) throw AudioLoadError(1, "audio shorter than one mel frame");
            items[i].duration = (double)w.samples.size() / w.sampleRate;
            const size_t window = (size_t)((double)windowSeconds * w.sampleRate + 0.5);
            int count = 0;
            q3asr_longform_plan(w.samples.size(), std::max<size_t>(window, 1), std::max<size_t>(160, ((size_t)160 * w.sampleRate + 15999) / 16000), nullptr, nullptr, 0, &count);
            std::vector<size_t> st((size_t)count), ln((size_t)count);
            q3asr_longform_plan(w.samples.size(), std::max<size_t>(window, 1), std::max<size_t>(160, ((size_t)160 * w.sampleRate + 15999) / 16000), st.data(), ln.data(), count, &count);
            for (int k = 0; k < count; k++) units.push_back({i, st[k], ln[k]});
        } catch (const std::exception& e) {
            items[i].error = e.what();
        }
    }
    // warm-up (TranscribeBatchCommand.swift:68-75): one pass over the first unit
    const double warmStart = now_s();
    if (!units.empty()) {
        const Unit& u = units[0];
        std::vector<float> clip(items[u.item].wav.samples.begin() + u.start, items[u.item].wav.samples.begin() + u.start + u.len);
        asr->transcribe(clip, items[u.item].wav.sampleRate, language, std::min(maxTokens, 4));
    }
    const double warmupTime = now_s() - warmStart;
    printf("  Warmup: %.2fs\n", warmupTime);

    double totalInference = 0, totalAudio = 0;
    const double batchStart = now_s();
    for (size_t u0 = 0; u0 < units.size(); u0 += (size_t)batch) {
        const size_t nb = std::min((size_t)batch, units.size() - u0);
        std::vector<std::vector<float>> clips(nb);
        std::vector<const std::vector<float>*> ptrs(nb);
        std::vector<int> rates(nb);
        double audio = 0;
        for (size_t k = 0; k < nb; k++) {
            const Unit& u = units[u0 + k];
            const auto& w = items[u.item].wav;
            clips[k].assign(w.samples.begin() + u.start, w.samples.begin() + u.start + u.len);
            ptrs[k] = &clips[k];
            rates[k] = w.sampleRate;
            audio += (double)u.len / w.sampleRate;
        }
        const double t0 = now_s();
        const auto texts = asr->transcribeBatch(ptrs, language, maxTokens, {}, rates);
        const double elapsed = now_s() - t0;
        totalInference += elapsed;
        for (size_t k = 0; k < nb; k++) {
            const Unit& u = units[u0 + k];
            Item& it = items[u.item];
            if (!it.text.empty()) it.text += " ";
            it.text += texts[k];
            it.time += elapsed * ((double)u.len / it.wav.sampleRate) / std::max(audio, 1e-9);
        }
    }
    const double batchTime = now_s() - batchStart;
    for (size_t i = 0; i < items.size(); i++) {
        const Item& it = items[i];
        if (!it.error.empty()) {
            if (jsonl) printf("{\"file\":\"%s\",\"error\":\"%s\"}\n", json_escape(it.name).c_str(), json_escape(it.error).c_str());
            else printf("  [%zu/%zu] %s: ERROR - %s\n", i + 1, items.size(), it.name.c_str(), it.error.c_str());
            continue;
        }
        totalAudio += it.duration;
        const double rtf = it.time / std::max(it.duration, 0.001);
        if (jsonl)
            printf("{\"file\":\"%s\",\"text\":\"%s\",\"time\":%.3f,\"rtf\":%.4f,\"duration\":%.2f}\n", json_escape(it.name).c_str(),
                   json_escape(it.text).c_str(), it.time, rtf, it.duration);
        else
            printf("  [%zu/%zu] (%.0f%%) %s: %s  (%.2fs, RTF=%.3f)\n", i + 1, items.size(), 100.0 * (i + 1) / items.size(), it.name.c_str(),
                   it.text.c_str(), it.time, rtf);
        if (!outputDir.empty()) {
            std::ofstream f(fs::path(outputDir) / (it.name + ".txt"));
            f << it.text;
        }
    }
    printf("\nBatch complete: %zu files, %.1fs audio\n", files.size(), totalAudio);
    printf("  Total inference: %.2fs, Aggregate RTF: %.4f\n", totalInference, totalInference / std::max(totalAudio, 0.001));
    printf("  Wall time: %.2fs (includes I/O)\n", batchTime);
    printf("  Model load: %.2fs, Warmup: %.2fs\n", loadTime, warmupTime);
    return 0;
}
