// safetensors.cu — reads the checkpoint format the reference loads through MLX.loadArrays
// (/root/reference/Sources/MLXCommon/WeightLoading.swift:9-11; key filtering in
// /root/reference/Sources/Qwen3ASR/WeightLoading.swift:17-126): every *.safetensors file of a directory,
// keys under audio_tower.* and model.*, dtypes F32 / F16 / BF16.  Quantised (U32-packed) tensors of the
// MLX 4-/8-bit repos are rejected with a clear message (bf16 build; dequant-at-load is a later row).
//
// File layout: u64 little-endian header length, JSON header {name: {dtype, shape, data_offsets:[a,b]}},
// then the raw tensor bytes.
#include <dirent.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "model.h"

namespace q3 {

namespace {

struct Entry {
    std::string name, dtype;
    std::vector<int64_t> shape;
    uint64_t begin = 0, end = 0;
};

struct Parser {  // just enough JSON for a safetensors header
    const char* s;
    size_t n, i = 0;
    void ws() { while (i < n && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; }
    bool eat(char c) {
        ws();
        if (i < n && s[i] == c) { i++; return true; }
        return false;
    }
    void expect(char c) {
        if (!eat(c)) throw Error(Q3ASR_ERR_IO, std::string("safetensors header: expected '") + c + "' at byte " + std::to_string(i));
    }
    std::string str() {
        ws();
        expect('"');
        std::string out;
        while (i < n && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < n) i++;
            out.push_back(s[i++]);
        }
        expect('"');
        return out;
    }
    int64_t num() {
        ws();
        int64_t v = 0;
        bool any = false;
        while (i < n && s[i] >= '0' && s[i] <= '9') { v = v * 10 + (s[i++] - '0'); any = true; }
        if (!any) throw Error(Q3ASR_ERR_IO, "safetensors header: expected a number at byte " + std::to_string(i));
        return v;
    }
    void skip_value() {
        ws();
        if (i >= n) return;
        if (s[i] == '"') { str(); return; }
        if (s[i] == '{' || s[i] == '[') {
            const char open = s[i], close = open == '{' ? '}' : ']';
            int depth = 0;
            bool in_str = false;
            for (; i < n; i++) {
                if (in_str) {
                    if (s[i] == '\\') i++;
                    else if (s[i] == '"') in_str = false;
                } else if (s[i] == '"') in_str = true;
                else if (s[i] == open) depth++;
                else if (s[i] == close && --depth == 0) { i++; return; }
            }
            return;
        }
        while (i < n && s[i] != ',' && s[i] != '}' && s[i] != ']') i++;
    }
};

std::vector<Entry> parse_header(const std::string& hdr) {
    Parser p{hdr.data(), hdr.size()};
    std::vector<Entry> out;
    p.expect('{');
    if (p.eat('}')) return out;
    do {
        Entry e;
        e.name = p.str();
        p.expect(':');
        if (e.name == "__metadata__") { p.skip_value(); continue; }
        p.expect('{');
        do {
            const std::string k = p.str();
            p.expect(':');
            if (k == "dtype") e.dtype = p.str();
            else if (k == "shape") {
                p.expect('[');
                if (!p.eat(']')) {
                    do e.shape.push_back(p.num()); while (p.eat(','));
                    p.expect(']');
                }
            } else if (k == "data_offsets") {
                p.expect('[');
                e.begin = (uint64_t)p.num();
                p.expect(',');
                e.end = (uint64_t)p.num();
                p.expect(']');
            } else p.skip_value();
        } while (p.eat(','));
        p.expect('}');
        out.push_back(std::move(e));
    } while (p.eat(','));
    p.expect('}');
    return out;
}

}  // namespace

void model_load_safetensors(Handle* h, const char* dir) {
    Q3_CHECK(dir != nullptr, Q3ASR_ERR_INVALID, "load_safetensors: null directory");
    DIR* d = opendir(dir);
    Q3_CHECK(d != nullptr, Q3ASR_ERR_IO, std::string("load_safetensors: cannot open directory ") + dir);
    std::vector<std::string> files;
    while (dirent* ent = readdir(d)) {
        const std::string f = ent->d_name;
        if (f.size() > 12 && f.compare(f.size() - 12, 12, ".safetensors") == 0) files.push_back(std::string(dir) + "/" + f);
    }
    closedir(d);
    std::sort(files.begin(), files.end());
    // "noWeightsFound" of the reference (MLXCommon/WeightLoading.swift:224-239)
    Q3_CHECK(!files.empty(), Q3ASR_ERR_IO, std::string("load_safetensors: no .safetensors files in ") + dir);
    std::vector<std::pair<std::string, std::vector<int64_t>>> specs;
    model_tensor_specs(h->cfg, &specs);
    size_t loaded = 0;
    std::vector<char> buf;
    for (const std::string& path : files) {
        FILE* fp = fopen(path.c_str(), "rb");
        Q3_CHECK(fp != nullptr, Q3ASR_ERR_IO, "load_safetensors: cannot open " + path);
        try {
            uint64_t hl = 0;
            Q3_CHECK(fread(&hl, 8, 1, fp) == 1 && hl > 1 && hl < (1ull << 30), Q3ASR_ERR_IO, "load_safetensors: bad header length in " + path);
            std::string hdr(hl, 0);
            Q3_CHECK(fread(&hdr[0], 1, hl, fp) == hl, Q3ASR_ERR_IO, "load_safetensors: truncated header in " + path);
            for (const Entry& e : parse_header(hdr)) {
                const bool ours = e.name.compare(0, 12, "audio_tower.") == 0 || e.name.compare(0, 6, "model.") == 0;
                if (!ours) continue;
                bool known = false;
                for (auto& s : specs)
                    if (s.first == e.name) { known = true; break; }
                if (!known) {
                    Q3_CHECK(e.name.find(".scales") == std::string::npos && e.name.find(".biases") == std::string::npos, Q3ASR_ERR_INVALID,
                             "load_safetensors: " + e.name + " belongs to a quantised MLX checkpoint; this build loads fp32/fp16/bf16 weights");
                    continue;
                }
                int dt = e.dtype == "F32" ? 0 : e.dtype == "BF16" ? 1 : e.dtype == "F16" ? 2 : -1;
                Q3_CHECK(dt >= 0, Q3ASR_ERR_INVALID, "load_safetensors: unsupported dtype " + e.dtype + " for " + e.name);
                Q3_CHECK(e.end >= e.begin && !e.shape.empty() && e.shape.size() <= 4, Q3ASR_ERR_IO, "load_safetensors: bad entry " + e.name);
                buf.resize(e.end - e.begin);
                Q3_CHECK(fseek(fp, (long)(8 + hl + e.begin), SEEK_SET) == 0 && fread(buf.data(), 1, buf.size(), fp) == buf.size(), Q3ASR_ERR_IO,
                         "load_safetensors: truncated data for " + e.name);
                size_t numel = 1;
                for (int64_t v : e.shape) numel *= (size_t)v;
                Q3_CHECK(buf.size() == numel * (dt == 0 ? 4 : 2), Q3ASR_ERR_IO, "load_safetensors: size mismatch for " + e.name);
                model_set_tensor(h, e.name.c_str(), buf.data(), dt, e.shape.data(), (int)e.shape.size());
                loaded++;
            }
        } catch (...) {
            fclose(fp);
            throw;
        }
        fclose(fp);
    }
    Q3_CHECK(loaded == specs.size(), Q3ASR_ERR_IO,
             "load_safetensors: found " + std::to_string(loaded) + " of " + std::to_string(specs.size()) + " expected tensors in " + dir);
    model_commit(h);
}

}  // namespace q3
