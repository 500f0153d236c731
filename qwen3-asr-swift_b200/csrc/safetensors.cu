// safetensors.cu — reads the checkpoint format the reference loads through MLX.loadArrays
// (/root/reference/Sources/MLXCommon/WeightLoading.swift:9-11; key filtering in
// /root/reference/Sources/Qwen3ASR/WeightLoading.swift:17-126): every *.safetensors file of a directory,
// keys under audio_tower.* and model.*, dtypes F32 / F16 / BF16, plus the U32-packed tensors of the MLX 4-/8-bit
// repos (weight + scales + biases), which are dequantised to bf16 at load.
//
// File layout: u64 little-endian header length, JSON header {name: {dtype, shape, data_offsets:[a,b]}},
// then the raw tensor bytes.
#include <dirent.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>

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
        int digits = 0;
        while (i < n && s[i] >= '0' && s[i] <= '9') {
            if (++digits > 18) throw Error(Q3ASR_ERR_IO, "safetensors header: number too long at byte " + std::to_string(i));
            v = v * 10 + (s[i++] - '0');
        }
        if (digits == 0) throw Error(Q3ASR_ERR_IO, "safetensors header: expected a number at byte " + std::to_string(i));
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

namespace {

struct Located {
    Entry e;
    int file = -1;
    uint64_t base = 0;  // file offset of the data section
};

// Elements of a tensor, with every dimension and the product bounded so that no later size computation can wrap.
size_t checked_numel(const Entry& e) {
    Q3_CHECK(!e.shape.empty() && e.shape.size() <= 4, Q3ASR_ERR_IO, "load_safetensors: bad rank for " + e.name);
    uint64_t numel = 1;
    for (int64_t v : e.shape) {
        Q3_CHECK(v >= 0 && v <= (int64_t)1 << 31, Q3ASR_ERR_IO, "load_safetensors: bad dimension in " + e.name);
        numel *= (uint64_t)v;
        Q3_CHECK(numel <= (uint64_t)1 << 36, Q3ASR_ERR_IO, "load_safetensors: tensor too large: " + e.name);
    }
    return (size_t)numel;
}

// Every audio_tower.* / model.* / lm_head.* entry of every *.safetensors file of a directory, each checked against the size of the
// file that holds it (a header may claim anything).  Owns the open files.
struct CheckpointIndex {
    std::vector<std::string> files;
    std::vector<FILE*> fps;
    std::map<std::string, Located> index;
    CheckpointIndex() = default;
    CheckpointIndex(const CheckpointIndex&) = delete;
    CheckpointIndex& operator=(const CheckpointIndex&) = delete;
    ~CheckpointIndex() {
        for (FILE* f : fps)
            if (f) fclose(f);
    }

    void open(const char* dir) {
        Q3_CHECK(dir != nullptr, Q3ASR_ERR_INVALID, "load_safetensors: null directory");
        DIR* d = opendir(dir);
        Q3_CHECK(d != nullptr, Q3ASR_ERR_IO, std::string("load_safetensors: cannot open directory ") + dir);
        while (dirent* ent = readdir(d)) {
            const std::string f = ent->d_name;
            if (f.size() > 12 && f.compare(f.size() - 12, 12, ".safetensors") == 0) files.push_back(std::string(dir) + "/" + f);
        }
        closedir(d);
        std::sort(files.begin(), files.end());
        // "noWeightsFound" of the reference (MLXCommon/WeightLoading.swift:224-239)
        Q3_CHECK(!files.empty(), Q3ASR_ERR_IO, std::string("load_safetensors: no .safetensors files in ") + dir);
        fps.assign(files.size(), nullptr);
        for (size_t fi = 0; fi < files.size(); fi++) {
            FILE* fp = fps[fi] = fopen(files[fi].c_str(), "rb");
            Q3_CHECK(fp != nullptr, Q3ASR_ERR_IO, "load_safetensors: cannot open " + files[fi]);
            Q3_CHECK(fseeko(fp, 0, SEEK_END) == 0, Q3ASR_ERR_IO, "load_safetensors: cannot seek in " + files[fi]);
            const off_t fsize = ftello(fp);
            Q3_CHECK(fsize >= 8 && fseeko(fp, 0, SEEK_SET) == 0, Q3ASR_ERR_IO, "load_safetensors: " + files[fi] + " is too small");
            uint64_t hl = 0;
            Q3_CHECK(fread(&hl, 8, 1, fp) == 1 && hl > 1 && hl < (1ull << 30) && hl <= (uint64_t)fsize - 8, Q3ASR_ERR_IO,
                     "load_safetensors: bad header length in " + files[fi]);
            std::string hdr(hl, 0);
            Q3_CHECK(fread(&hdr[0], 1, hl, fp) == hl, Q3ASR_ERR_IO, "load_safetensors: truncated header in " + files[fi]);
            const uint64_t data_bytes = (uint64_t)fsize - 8 - hl;
            for (Entry& e : parse_header(hdr)) {
                // the forced-aligner checkpoints carry a "thinker." prefix and keep the classification head under lm_head.*
                // (WeightLoading.swift:162-179); everything else is ignored
                if (e.name.compare(0, 8, "thinker.") == 0) e.name = e.name.substr(8);
                if (e.name.compare(0, 12, "audio_tower.") != 0 && e.name.compare(0, 6, "model.") != 0 && e.name.compare(0, 8, "lm_head.") != 0)
                    continue;
                Q3_CHECK(e.begin <= e.end && e.end <= data_bytes, Q3ASR_ERR_IO, "load_safetensors: data of " + e.name + " lies outside " + files[fi]);
                checked_numel(e);
                Located l;
                l.e = e;
                l.file = (int)fi;
                l.base = 8 + hl;
                index[e.name] = l;
            }
        }
    }

    void read(const Located& l, std::vector<char>* buf) const {  // offsets were checked against the file size in open()
        buf->resize((size_t)(l.e.end - l.e.begin));
        Q3_CHECK(fseeko(fps[l.file], (off_t)(l.base + l.e.begin), SEEK_SET) == 0 &&
                     fread(buf->data(), 1, buf->size(), fps[l.file]) == buf->size(),
                 Q3ASR_ERR_IO, "load_safetensors: truncated data for " + l.e.name);
    }
};

float half_to_float(uint16_t hbits) {
    const uint32_t sign = (uint32_t)(hbits & 0x8000) << 16;
    uint32_t exp = (hbits >> 10) & 0x1F, man = hbits & 0x3FF, out;
    if (exp == 0) {
        if (man == 0) out = sign;
        else {
            exp = 127 - 15 + 1;
            while (!(man & 0x400)) { man <<= 1; exp--; }
            out = sign | (exp << 23) | ((man & 0x3FF) << 13);
        }
    } else if (exp == 31) out = sign | 0x7F800000u | (man << 13);
    else out = sign | ((exp - 15 + 127) << 23) | (man << 13);
    float f;
    memcpy(&f, &out, 4);
    return f;
}

// element i of a scales / biases tensor as float
float scalar_at(const std::vector<char>& buf, const std::string& dtype, size_t i) {
    if (dtype == "F32") { float f; memcpy(&f, buf.data() + 4 * i, 4); return f; }
    uint16_t u;
    memcpy(&u, buf.data() + 2 * i, 2);
    if (dtype == "BF16") { const uint32_t w = (uint32_t)u << 16; float f; memcpy(&f, &w, 4); return f; }
    return half_to_float(u);
}

}  // namespace

// MLX affine quantisation (mlx-swift `dequantized`, used through PreQuantizedEmbedding / QuantizedLinear,
// /root/reference/Sources/MLXCommon/PreQuantizedEmbedding.swift:12-49, group size 64, Configuration.swift:61-63): a row of
// `in` features is stored as in*bits/32 little-endian uint32 words (value j of a word at bits [j*bits, (j+1)*bits)) plus one
// (scale, bias) pair per group of 64 features; w = scale * q + bias.  Dequantised at load: the B200 path computes in bf16.
void model_load_safetensors(Handle* h, const char* dir) {
    CheckpointIndex ck;
    ck.open(dir);  // pass 1
    const std::map<std::string, Located>& index = ck.index;
    std::vector<std::pair<std::string, std::vector<int64_t>>> specs;
    model_tensor_specs(h->cfg, &specs);
    {
        // pass 2: every tensor the model needs, dequantising MLX-packed ones
        size_t loaded = 0;
        std::vector<char> buf, sbuf, bbuf;
        std::vector<float> deq;
        for (auto& spec : specs) {
            auto it = index.find(spec.first);
            if (it == index.end()) {
                // the reference's applyLinearWeights keeps a Linear's bias as it is when the checkpoint has none
                // (MLXCommon/WeightLoading.swift:113-130): the aligner's classification bias then stays zero
                if (spec.first == "lm_head.bias") {
                    size_t n = 1;
                    for (int64_t d : spec.second) n *= (size_t)d;
                    std::vector<float> zeros(n, 0.f);
                    model_set_tensor(h, spec.first.c_str(), zeros.data(), 0, spec.second.data(), (int)spec.second.size());
                    loaded++;
                }
                continue;
            }
            const Located& l = it->second;
            const Entry& e = l.e;
            ck.read(l, &buf);
            const size_t numel = checked_numel(e);
            if (e.dtype == "U32") {
                Q3_CHECK(e.name.size() > 7, Q3ASR_ERR_INVALID, "load_safetensors: packed tensor with an unexpected name: " + e.name);
                const std::string stem = e.name.substr(0, e.name.size() - 7);  // strip ".weight"
                auto si = index.find(stem + ".scales"), bi = index.find(stem + ".biases");
                Q3_CHECK(e.shape.size() == 2 && si != index.end() && bi != index.end(), Q3ASR_ERR_INVALID,
                         "load_safetensors: " + e.name + " is packed (U32) but its .scales / .biases are missing");
                const Entry& se = si->second.e;
                Q3_CHECK(se.shape.size() == 2 && se.shape[0] == e.shape[0] && bi->second.e.shape == se.shape && bi->second.e.dtype == se.dtype,
                         Q3ASR_ERR_INVALID, "load_safetensors: scales / biases of " + e.name + " do not match it");
                // the group size follows from the columns the model expects (64 in the published checkpoints; MLX also writes 32 / 128)
                Q3_CHECK(spec.second.size() == 2 && se.shape[1] > 0 && spec.second[1] % se.shape[1] == 0, Q3ASR_ERR_INVALID,
                         "load_safetensors: quantisation groups of " + e.name + " do not divide its columns");
                const int64_t rows = e.shape[0], words = e.shape[1], groups = se.shape[1], cols = spec.second[1], gsz = cols / groups;
                Q3_CHECK(gsz == 32 || gsz == 64 || gsz == 128, Q3ASR_ERR_INVALID, "load_safetensors: unsupported quantisation group size for " + e.name);
                Q3_CHECK(buf.size() == numel * 4 && cols > 0 && (words * 32) % cols == 0, Q3ASR_ERR_IO, "load_safetensors: size mismatch for " + e.name);
                const int bits = (int)(words * 32 / cols);
                Q3_CHECK(bits == 2 || bits == 4 || bits == 8, Q3ASR_ERR_INVALID, "load_safetensors: unsupported quantisation width for " + e.name);
                ck.read(si->second, &sbuf);
                ck.read(bi->second, &bbuf);
                const size_t sb = se.dtype == "F32" ? 4 : 2;
                Q3_CHECK(se.dtype == "F32" || se.dtype == "BF16" || se.dtype == "F16", Q3ASR_ERR_INVALID, "load_safetensors: unsupported scales dtype for " + e.name);
                Q3_CHECK(sbuf.size() == (size_t)(rows * groups) * sb && bbuf.size() == sbuf.size(), Q3ASR_ERR_IO, "load_safetensors: scales size mismatch for " + e.name);
                deq.resize((size_t)rows * cols);
                const uint32_t mask = (1u << bits) - 1;
                const int per = 32 / bits;
                const uint32_t* w32 = reinterpret_cast<const uint32_t*>(buf.data());
                for (int64_t r = 0; r < rows; r++)
                    for (int64_t g = 0; g < groups; g++) {
                        const float sc = scalar_at(sbuf, se.dtype, (size_t)(r * groups + g)), bs = scalar_at(bbuf, se.dtype, (size_t)(r * groups + g));
                        for (int64_t c = 0; c < gsz; c++) {
                            const int64_t col = g * gsz + c;
                            const uint32_t q = (w32[r * words + col / per] >> ((col % per) * bits)) & mask;
                            deq[(size_t)(r * cols + col)] = (float)((double)sc * (double)q + (double)bs);  // exact product, one rounding
                        }
                    }
                const int64_t shape[2] = {rows, cols};
                model_set_tensor(h, e.name.c_str(), deq.data(), 0, shape, 2);
            } else {
                const int dt = e.dtype == "F32" ? 0 : e.dtype == "BF16" ? 1 : e.dtype == "F16" ? 2 : -1;
                Q3_CHECK(dt >= 0, Q3ASR_ERR_INVALID, "load_safetensors: unsupported dtype " + e.dtype + " for " + e.name);
                Q3_CHECK(buf.size() == numel * (dt == 0 ? 4 : 2), Q3ASR_ERR_IO, "load_safetensors: size mismatch for " + e.name);
                model_set_tensor(h, e.name.c_str(), buf.data(), dt, e.shape.data(), (int)e.shape.size());
            }
            loaded++;
        }
        Q3_CHECK(loaded == specs.size(), Q3ASR_ERR_IO,
                 "load_safetensors: found " + std::to_string(loaded) + " of " + std::to_string(specs.size()) + " expected tensors in " + dir);
    }
    model_commit(h);
}

// q3asr_checkpoint_list: the validated index as text, one tensor per line (host only)
std::string checkpoint_list(const char* dir) {
    CheckpointIndex ck;
    ck.open(dir);
    std::string out;
    for (const auto& kv : ck.index) {
        const Entry& e = kv.second.e;
        auto printable = [](std::string v) {  // the names come from the file: keep the line format intact whatever they hold
            for (char& c : v)
                if ((unsigned char)c < 0x20) c = '?';
            return v;
        };
        out += printable(e.name) + "\t" + printable(e.dtype) + "\t";
        for (size_t i = 0; i < e.shape.size(); i++) out += (i ? "x" : "") + std::to_string(e.shape[i]);
        out += "\t" + std::to_string(e.end - e.begin) + "\n";
    }
    return out;
}

}  // namespace q3
