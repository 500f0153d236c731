// host_demo.cpp — exercises the C++ host mirror (qwen3_asr.hpp) end to end; used by tests/test_host_cpp.py.
//   host_demo symbols        -> checks option defaults / size detection (no GPU needed)
//   host_demo run <seconds>  -> random-init 0.6B, mel + transcribe of a synthetic clip (needs a B200)
#include <cmath>
#include <cstdio>
#include <cstring>

#include "qwen3_asr.hpp"

using namespace qwen3asr;

int main(int argc, char** argv) {
    const char* mode = argc > 1 ? argv[1] : "symbols";
    if (!strcmp(mode, "symbols")) {
        Qwen3DecodingOptions o;
        if (o.maxTokens != 448 || !o.isGreedyFastPath() || o.language || o.context) return 1;
        if (detectModelSize("aufklarer/Qwen3-ASR-1.7B-MLX-8bit") != ASRModelSize::large) return 2;
        if (detectModelSize("aufklarer/Qwen3-ASR-0.6B-MLX-4bit") != ASRModelSize::small) return 3;
        if (detectModelSize("some-custom/model") != ASRModelSize::small) return 3;
        // testASRModelSizeBitsDetection (Qwen3ASRTests.swift:61-69)
        if (detectModelBits("aufklarer/Qwen3-ASR-0.6B-MLX-8bit") != 8 || detectModelBits("aufklarer/Qwen3-ASR-0.6B-MLX-4bit") != 4 ||
            detectModelBits("aufklarer/Qwen3-ASR-1.7B-MLX-4bit") != 4 || detectModelBits("some-custom/small-model") != 4 ||
            detectModelBits("some/1.7B-model") != 8)
            return 4;
        try {
            auto m = Qwen3ASRModel::randomInit(ASRModelSize::small);
            printf("created on GPU, footprint %zu\n", m->memoryFootprint());
        } catch (const AudioModelError& e) {
            printf("load error (expected without a GPU): %s\n", e.what());
        }
        // tokenizer through the host mirror: a two-token vocabulary written to a temp directory
        if (argc > 2) {
            const std::string dir = argv[2];
            FILE* f = fopen((dir + "/vocab.json").c_str(), "w");
            if (!f) return 7;
            fputs("{\"Hello\": 5, \"\u0120world\": 6, \"<asr_text>\": 7}", f);
            fclose(f);
            Tokenizer t = Tokenizer::fromDirectory(dir);
            if (t.decode({7, 5, 6}) != "<asr_text>Hello world") return 8;
            try {
                Tokenizer::fromDirectory(dir + "/nope");
                return 9;
            } catch (const AudioModelError&) {
            }
            // transcript extraction (Qwen3ASR.swift:283-289): what follows <asr_text>, trimmed of spaces / tabs / Unicode space separators
            if (Qwen3ASRModel::textFromIds(t, {5, 7, 6, 5}) != "worldHello") return 9;  // "Hello<asr_text> worldHello" -> after the marker, trimmed
            if (Qwen3ASRModel::trimSwiftWhitespaces(" \t\xC2\xA0\xE3\x80\x80 a b\n \xE2\x80\x89") != "a b\n") return 9;
            if (Qwen3ASRModel::trimSwiftWhitespaces("  \t ") != "") return 9;
            printf("tokenizer ok\n");
            // preset detection from a checkpoint's own tensor index: header-only safetensors files written here
            auto write_ck = [&](const std::string& sub, const std::string& header, size_t payload) {
                const std::string d = dir + "/" + sub;
                if (system(("mkdir -p '" + d + "'").c_str()) != 0) return false;
                FILE* g = fopen((d + "/model.safetensors").c_str(), "wb");
                if (!g) return false;
                const uint64_t hl = header.size();
                fwrite(&hl, 8, 1, g);
                fwrite(header.data(), 1, header.size(), g);
                const std::string zeros(payload, '\0');
                fwrite(zeros.data(), 1, zeros.size(), g);
                fclose(g);
                return true;
            };
            if (!write_ck("small", "{\"model.norm.weight\":{\"dtype\":\"BF16\",\"shape\":[1024],\"data_offsets\":[0,2048]}}", 2048)) return 10;
            if (!write_ck("large", "{\"model.norm.weight\":{\"dtype\":\"BF16\",\"shape\":[2048],\"data_offsets\":[0,4096]}}", 4096)) return 10;
            if (!write_ck("align", "{\"thinker.lm_head.weight\":{\"dtype\":\"BF16\",\"shape\":[2,4],\"data_offsets\":[0,16]},"
                                   "\"thinker.model.norm.weight\":{\"dtype\":\"BF16\",\"shape\":[1024],\"data_offsets\":[16,2064]}}", 2064)) return 10;
            if (!write_ck("odd", "{\"model.norm.weight\":{\"dtype\":\"BF16\",\"shape\":[128],\"data_offsets\":[0,256]}}", 256)) return 10;
            if (detectPresetFromCheckpoint(dir + "/small") != "0.6B" || detectPresetFromCheckpoint(dir + "/large") != "1.7B" ||
                detectPresetFromCheckpoint(dir + "/align") != "aligner" || detectPresetFromCheckpoint(dir + "/odd") != "" ||
                detectPresetFromCheckpoint(dir + "/nope") != "")
                return 11;
            printf("checkpoint detection ok\n");
            // WAVWriter -> AudioFileLoader round trip on the PCM16 grid (WAVWriterTests.swift:38-52)
            std::vector<float> tone(1000);
            for (size_t i = 0; i < tone.size(); i++) tone[i] = std::sin(0.1f * (float)i);
            WAVWriter::write(tone, 24000, dir + "/tone.wav");
            const AudioFileLoader::Wav back = AudioFileLoader::loadWAV(dir + "/tone.wav");
            if (back.sampleRate != 24000 || back.samples.size() != tone.size()) return 12;
            for (size_t i = 0; i < tone.size(); i++)
                if (back.samples[i] != (float)(int)(tone[i] * 32767.0f) / 32768.0f) return 13;
            printf("wav round trip ok\n");
        }
        printf("%s\n", q3asr_version());
        return 0;
    }
    const int seconds = argc > 2 ? atoi(argv[2]) : 3;
    std::vector<float> x((size_t)seconds * 16000);
    for (size_t i = 0; i < x.size(); i++) x[i] = 0.4f * sinf(2.f * 3.14159265f * 440.f * (float)i / 16000.f);
    auto m = Qwen3ASRModel::randomInit(ASRModelSize::small);
    MelFeatures f = m->featureExtractor.extractFeaturesRaw(x);
    printf("mel %d x %d\n", f.melBins, f.timeFrames);
    Qwen3DecodingOptions opt;
    opt.maxTokens = 8;
    std::string t = m->transcribe(x, 16000, opt);
    printf("text(ids) %s\n", t.c_str());
    auto both = m->transcribeBatch({&x, &x}, {}, 8);
    if (both[0] != both[1] || both[0] != t) return 4;
    // decoder knobs: a no-repeat-bigram mask must leave no bigram twice in the id stream (printed as ids: no tokenizer here)
    Qwen3DecodingOptions knobs;
    knobs.maxTokens = 12;
    knobs.noRepeatNgramSize = 2;
    knobs.repetitionPenalty = 1.2f;
    std::string k = m->transcribe(x, 16000, knobs);
    printf("knobs(ids) %s\n", k.c_str());
    if (k == t || k.empty() || k[0] == '[') return 10;
    // 24 kHz input is converted on the device; long-form windows
    std::vector<float> x24((size_t)seconds * 24000);
    for (size_t i = 0; i < x24.size(); i++) x24[i] = 0.4f * sinf(2.f * 3.14159265f * 440.f * (float)i / 24000.f);
    if (m->transcribe(x24, 24000, {}, 8).empty()) return 11;
    if (AudioFileLoader::resample(*m, x24, 24000, 16000).size() != x.size()) return 12;
    auto segs = m->transcribeLong(x, 16000, 1.0f, 4, 2);
    if ((int)segs.size() != seconds || segs.back().segmentIndex != seconds - 1) return 13;
    {   // forced aligner through the host mirror: tiny configuration, a toy tokenizer (one id per word length)
        auto al = Qwen3ForcedAligner::randomInit(20260418, 0, "tiny-aligner");
        Tokenizer toy;
        toy.encode = [](const std::string& w) { return std::vector<int32_t>{(int32_t)(10 + w.size()), (int32_t)(100 + (unsigned char)w[0])}; };
        al->setTokenizer(toy);
        auto words = al->align(x, "Hello, brave new world -- it's 9 o'clock .");
        if (words.size() != 7 || words[0].text != "Hello," || words[3].text != "world--" || words[6].text != "o'clock.") return 14;
        for (size_t i = 0; i < words.size(); i++) {
            if (words[i].endTime < words[i].startTime || (i && words[i].startTime < words[i - 1].startTime)) return 15;
        }
        if (al->alignLong(x, "one two three").size() != 3) return 16;
        printf("aligner ok: %zu words, last at %.2f s\n", words.size(), words.back().startTime);
    }
    m->unload();
    if (m->isLoaded() || m->transcribe(x).find("not loaded") == std::string::npos) return 5;
    return f.timeFrames == seconds * 100 ? 0 : 6;
}
