// tokenizer.cu — host-only: the Qwen3 byte-level BPE tokenizer behind q3asr_tokenizer_* (no device code).
//
// Follows /root/reference/Sources/AudioCommon/Tokenizer.swift:
//   load(from:) :38-62           vocab.json {token: id}; tokenizer_config.json "added_tokens_decoder" overrides (:65-91);
//                                merges.txt, '#' lines and malformed lines skipped, rank = line index (:94-109)
//   decode(tokens:) :111-142     unknown ids skipped; "<|...|>" specials dropped; "<...>" markers without '|' kept verbatim;
//                                every other character goes through the GPT-2 unicode->byte table into ONE byte buffer, which
//                                is decoded as UTF-8 at the end (so characters split across tokens survive), invalid
//                                sequences become U+FFFD; the result is trimmed of leading / trailing whitespace
//   byteToUnicode :146-173       bytes 33-126, 161-172, 174-255 map to themselves, the rest to U+0100 + n in byte order
//   encode(_:) :195-278          pre-tokenise on ' ', '\n', '\t' (the separator starts the next word), byte-level map, lowest-rank
//                                merge first (all occurrences of the pair in one sweep), pieces missing from the vocabulary are
//                                dropped; without merges: per-character vocabulary lookup
// The reference's own unit tests for this code (Tests/Qwen3ASRTests/Qwen3ASRTests.swift:275-451) are ported in
// tests/test_tokenizer.py and pin the behaviour.
#include <stdio.h>
#include <string.h>

#include <climits>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/q3asr.h"

namespace {

// ---- UTF-8 helpers ----
void append_utf8(std::string& s, uint32_t cp) {
    if (cp < 0x80) {
        s += (char)cp;
    } else if (cp < 0x800) {
        s += (char)(0xC0 | (cp >> 6));
        s += (char)(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
        s += (char)(0xE0 | (cp >> 12));
        s += (char)(0x80 | ((cp >> 6) & 0x3F));
        s += (char)(0x80 | (cp & 0x3F));
    } else {
        s += (char)(0xF0 | (cp >> 18));
        s += (char)(0x80 | ((cp >> 12) & 0x3F));
        s += (char)(0x80 | ((cp >> 6) & 0x3F));
        s += (char)(0x80 | (cp & 0x3F));
    }
}

// Decodes one scalar starting at s[i]; returns its length in bytes, or 0 if the bytes there are not valid UTF-8
// (overlongs, surrogates and values above U+10FFFF are invalid).  *bad_len = bytes of the maximal invalid prefix.
int decode_utf8(const std::string& s, size_t i, uint32_t* cp, int* bad_len) {
    const unsigned char b0 = (unsigned char)s[i];
    *bad_len = 1;
    if (b0 < 0x80) {
        *cp = b0;
        return 1;
    }
    int need;
    uint32_t v;
    unsigned char lo = 0x80, hi = 0xBF;
    if (b0 >= 0xC2 && b0 <= 0xDF) { need = 1; v = b0 & 0x1F; }
    else if (b0 >= 0xE0 && b0 <= 0xEF) { need = 2; v = b0 & 0x0F; if (b0 == 0xE0) lo = 0xA0; if (b0 == 0xED) hi = 0x9F; }
    else if (b0 >= 0xF0 && b0 <= 0xF4) { need = 3; v = b0 & 0x07; if (b0 == 0xF0) lo = 0x90; if (b0 == 0xF4) hi = 0x8F; }
    else return 0;
    for (int k = 1; k <= need; k++) {
        if (i + k >= s.size()) { *bad_len = k; return 0; }
        const unsigned char b = (unsigned char)s[i + k];
        const unsigned char l = k == 1 ? lo : 0x80, h = k == 1 ? hi : 0xBF;
        if (b < l || b > h) { *bad_len = k; return 0; }
        v = (v << 6) | (b & 0x3F);
    }
    *cp = v;
    return need + 1;
}

std::vector<uint32_t> scalars_of(const std::string& s) {
    std::vector<uint32_t> out;
    for (size_t i = 0; i < s.size();) {
        uint32_t cp;
        int bad;
        const int n = decode_utf8(s, i, &cp, &bad);
        if (n == 0) { out.push_back(0xFFFD); i += bad; } else { out.push_back(cp); i += n; }
    }
    return out;
}

// String(decoding:as: UTF8.self): every maximal invalid subsequence becomes one U+FFFD
std::string repair_utf8(const std::string& s) {
    std::string out;
    for (size_t i = 0; i < s.size();) {
        uint32_t cp;
        int bad;
        const int n = decode_utf8(s, i, &cp, &bad);
        if (n == 0) { append_utf8(out, 0xFFFD); i += bad; } else { out.append(s, i, n); i += n; }
    }
    return out;
}

bool is_ws_scalar(uint32_t c) {  // CharacterSet.whitespaces: Unicode Zs + TAB
    return c == 0x20 || c == 0x09 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x202F || c == 0x205F || c == 0x3000;
}

std::string trim_ws(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b) {
        uint32_t cp;
        int bad;
        const int n = decode_utf8(s, a, &cp, &bad);
        if (n == 0 || !is_ws_scalar(cp)) break;
        a += n;
    }
    while (b > a) {
        size_t st = b - 1;
        while (st > a && ((unsigned char)s[st] & 0xC0) == 0x80) st--;
        uint32_t cp;
        int bad;
        const int n = decode_utf8(s, st, &cp, &bad);
        if (n == 0 || st + n != b || !is_ws_scalar(cp)) break;
        b = st;
    }
    return s.substr(a, b - a);
}

// ---- minimal JSON (objects, arrays, strings, numbers, literals) ----
struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json* get(const std::string& k) const {
        for (auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct JsonParser {
    const std::string& s;
    size_t i = 0;
    std::string err;
    explicit JsonParser(const std::string& src) : s(src) {}
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; }
    bool fail(const char* m) { if (err.empty()) err = std::string(m) + " at byte " + std::to_string(i); return false; }
    bool hex4(uint32_t* v) {
        if (i + 4 > s.size()) return fail("short \\u escape");
        uint32_t x = 0;
        for (int k = 0; k < 4; k++) {
            const char c = s[i + k];
            x <<= 4;
            if (c >= '0' && c <= '9') x |= c - '0';
            else if (c >= 'a' && c <= 'f') x |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') x |= c - 'A' + 10;
            else return fail("bad \\u escape");
        }
        i += 4;
        *v = x;
        return true;
    }
    bool string(std::string* out) {
        if (s[i] != '"') return fail("expected string");
        i++;
        out->clear();
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\') {
                if (++i >= s.size()) return fail("dangling escape");
                const char c = s[i++];
                switch (c) {
                    case 'n': *out += '\n'; break;
                    case 't': *out += '\t'; break;
                    case 'r': *out += '\r'; break;
                    case 'b': *out += '\b'; break;
                    case 'f': *out += '\f'; break;
                    case 'u': {
                        uint32_t cp = 0;
                        if (!hex4(&cp)) return false;
                        if (cp >= 0xD800 && cp <= 0xDBFF && i + 1 < s.size() && s[i] == '\\' && s[i + 1] == 'u') {
                            i += 2;
                            uint32_t lo = 0;
                            if (!hex4(&lo)) return false;
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        append_utf8(*out, cp);
                        break;
                    }
                    default: *out += c;  // \" \\ \/
                }
            } else {
                *out += s[i++];
            }
        }
        if (i >= s.size()) return fail("unterminated string");
        i++;
        return true;
    }
    int depth = 0;
    struct Nest {  // a file of a million '[' must not turn into a million stack frames
        int& d;
        explicit Nest(int& x) : d(x) { d++; }
        ~Nest() { d--; }
    };
    bool value(Json* v) {
        Nest nest(depth);
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (i >= s.size()) return fail("unexpected end");
        const char c = s[i];
        if (c == '{') {
            v->kind = Json::Obj;
            i++;
            ws();
            if (i < s.size() && s[i] == '}') { i++; return true; }
            for (;;) {
                ws();
                std::string k;
                if (i >= s.size() || !string(&k)) return fail("expected key");
                ws();
                if (i >= s.size() || s[i] != ':') return fail("expected ':'");
                i++;
                v->obj.emplace_back(k, Json());
                if (!value(&v->obj.back().second)) return false;
                ws();
                if (i < s.size() && s[i] == ',') { i++; continue; }
                if (i < s.size() && s[i] == '}') { i++; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            v->kind = Json::Arr;
            i++;
            ws();
            if (i < s.size() && s[i] == ']') { i++; return true; }
            for (;;) {
                v->arr.emplace_back();
                if (!value(&v->arr.back())) return false;
                ws();
                if (i < s.size() && s[i] == ',') { i++; continue; }
                if (i < s.size() && s[i] == ']') { i++; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') { v->kind = Json::Str; return string(&v->str); }
        if (!s.compare(i, 4, "true")) { v->kind = Json::Bool; v->b = true; i += 4; return true; }
        if (!s.compare(i, 5, "false")) { v->kind = Json::Bool; i += 5; return true; }
        if (!s.compare(i, 4, "null")) { i += 4; return true; }
        char* end = nullptr;
        v->num = strtod(s.c_str() + i, &end);
        if (end == s.c_str() + i) return fail("unexpected character");
        v->kind = Json::Num;
        i = end - s.c_str();
        return true;
    }
};

bool read_file(const std::string& path, std::string* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

}  // namespace

struct q3asr_tokenizer {
    std::unordered_map<int, std::string> id_to_token;
    std::unordered_map<std::string, int> token_to_id;
    std::unordered_map<std::string, int> merge_rank;  // "a b" -> rank
    size_t n_merges = 0;
    uint32_t byte_to_uni[256];
    std::unordered_map<uint32_t, unsigned char> uni_to_byte;
    std::string last_error;

    q3asr_tokenizer() {
        // Tokenizer.swift:146-173
        bool direct[256] = {false};
        for (int b = 33; b <= 126; b++) direct[b] = true;
        for (int b = 161; b <= 172; b++) direct[b] = true;
        for (int b = 174; b <= 255; b++) direct[b] = true;
        int n = 0;
        for (int b = 0; b < 256; b++) {
            byte_to_uni[b] = direct[b] ? (uint32_t)b : (uint32_t)(0x100 + n++);
            uni_to_byte[byte_to_uni[b]] = (unsigned char)b;
        }
    }

    void add(int id, const std::string& tok) {
        id_to_token[id] = tok;
        token_to_id[tok] = id;
    }

    std::string decode(const int32_t* ids, int n) const {
        std::string buf;
        for (int k = 0; k < n; k++) {
            auto it = id_to_token.find(ids[k]);
            if (it == id_to_token.end()) continue;
            const std::string& t = it->second;
            const bool lt = !t.empty() && t.front() == '<' && t.back() == '>';
            if (t.size() >= 2 && !t.compare(0, 2, "<|") && !t.compare(t.size() - 2, 2, "|>")) continue;
            if (lt && t.find('|') == std::string::npos) { buf += t; continue; }
            for (uint32_t cp : scalars_of(t)) {
                auto b = uni_to_byte.find(cp);
                if (b != uni_to_byte.end()) buf += (char)b->second;
                else append_utf8(buf, cp);
            }
        }
        return trim_ws(repair_utf8(buf));
    }

    std::string byte_level(const std::string& raw) const {
        std::string out;
        for (unsigned char c : raw) append_utf8(out, byte_to_uni[c]);
        return out;
    }

    std::vector<std::string> bpe(const std::string& word) const {
        std::vector<std::string> pieces;
        for (uint32_t cp : scalars_of(word)) {
            pieces.emplace_back();
            append_utf8(pieces.back(), cp);
        }
        while (pieces.size() > 1) {
            int best = INT_MAX;
            size_t at = 0;
            for (size_t i = 0; i + 1 < pieces.size(); i++) {
                auto it = merge_rank.find(pieces[i] + " " + pieces[i + 1]);
                if (it != merge_rank.end() && it->second < best) { best = it->second; at = i; }
            }
            if (best == INT_MAX) break;
            const std::string first = pieces[at], second = pieces[at + 1];
            std::vector<std::string> next;
            for (size_t i = 0; i < pieces.size();) {
                if (i + 1 < pieces.size() && pieces[i] == first && pieces[i + 1] == second) { next.push_back(first + second); i += 2; }
                else next.push_back(pieces[i++]);
            }
            pieces.swap(next);
        }
        return pieces;
    }

    std::vector<int32_t> encode(const std::string& text) const {
        std::vector<int32_t> out;
        if (n_merges == 0) {  // characterEncode, Tokenizer.swift:268-278
            for (uint32_t cp : scalars_of(text)) {
                std::string c;
                append_utf8(c, cp);
                auto it = token_to_id.find(c);
                if (it != token_to_id.end()) out.push_back(it->second);
            }
            return out;
        }
        std::vector<std::string> words;  // preTokenize, Tokenizer.swift:218-238
        std::string cur;
        // The reference walks Swift Characters (grapheme clusters): "\r\n" is ONE Character that is not equal to "\n", so a CRLF
        // does not split; every other splitting character here is a single-byte cluster on its own.
        for (size_t i = 0; i < text.size(); i++) {
            const char ch = text[i];
            if (ch == '\r' && i + 1 < text.size() && text[i + 1] == '\n') {
                cur += "\r\n";
                i++;
            } else if (ch == ' ' || ch == '\n' || ch == '\t') {
                if (!cur.empty()) words.push_back(byte_level(cur));
                cur.assign(1, ch);
            } else {
                cur += ch;
            }
        }
        if (!cur.empty()) words.push_back(byte_level(cur));
        for (const std::string& w : words)
            for (const std::string& piece : bpe(w)) {
                auto it = token_to_id.find(piece);
                if (it != token_to_id.end()) out.push_back(it->second);
            }
        return out;
    }

    bool load_dir(const std::string& path) {
        std::string vocab_path = path, dir = path;
        if (path.size() > 5 && !path.compare(path.size() - 5, 5, ".json")) {
            const size_t sl = path.find_last_of('/');
            dir = sl == std::string::npos ? "." : path.substr(0, sl);
        } else {
            vocab_path = path + "/vocab.json";
        }
        std::string src;
        if (!read_file(vocab_path, &src)) { last_error = "cannot read " + vocab_path; return false; }
        {
            JsonParser jp(src);
            Json root;
            if (!jp.value(&root) || root.kind != Json::Obj) { last_error = "Invalid tokenizer format: Expected {token: id} dictionary (" + jp.err + ")"; return false; }
            for (auto& kv : root.obj) {
                // ids are Ints in the reference; a value outside the int range (or NaN) cannot be one, and casting it would be undefined
                if (kv.second.kind != Json::Num || !(kv.second.num >= -2147483648.0 && kv.second.num <= 2147483647.0)) {
                    last_error = "Invalid tokenizer format: Expected {token: id} dictionary";
                    return false;
                }
                add((int)kv.second.num, kv.first);
            }
        }
        if (read_file(dir + "/tokenizer_config.json", &src)) {
            JsonParser jp(src);
            Json root;
            if (jp.value(&root) && root.kind == Json::Obj)
                if (const Json* added = root.get("added_tokens_decoder"))
                    for (auto& kv : added->obj) {
                        char* end = nullptr;
                        const long id = strtol(kv.first.c_str(), &end, 10);
                        const Json* content = kv.second.get("content");
                        if (end == kv.first.c_str() || *end || id < INT_MIN || id > INT_MAX || !content || content->kind != Json::Str) continue;
                        add((int)id, content->str);
                    }
        }
        if (read_file(dir + "/merges.txt", &src)) {
            // components(separatedBy: .newlines): every scalar of CharacterSet.newlines (U+000A-U+000D, U+0085, U+2028, U+2029) starts
            // a new line (so "\r\n" yields an empty line, skipped); the line INDEX is the merge rank, so all of them must count
            auto newline_at = [&](size_t i) -> size_t {  // byte length of the line terminator at i, or 0
                const unsigned char c0 = (unsigned char)src[i];
                if (c0 >= 0x0A && c0 <= 0x0D) return 1;
                if (c0 == 0xC2 && i + 1 < src.size() && (unsigned char)src[i + 1] == 0x85) return 2;
                if (c0 == 0xE2 && i + 2 < src.size() && (unsigned char)src[i + 1] == 0x80 &&
                    ((unsigned char)src[i + 2] == 0xA8 || (unsigned char)src[i + 2] == 0xA9))
                    return 3;
                return 0;
            };
            size_t pos = 0;
            int index = 0;
            while (pos <= src.size()) {
                size_t e = pos, nl = 0;
                while (e < src.size() && (nl = newline_at(e)) == 0) e++;
                if (e >= src.size()) nl = 1;
                const std::string line = src.substr(pos, e - pos);
                if (!line.empty() && line[0] != '#') {
                    const size_t sp = line.find(' ');
                    if (sp != std::string::npos && line.find(' ', sp + 1) == std::string::npos) {
                        merge_rank[line] = index;
                        n_merges++;
                    }
                }
                index++;
                pos = e + nl;
            }
        }
        return true;
    }
};

namespace {
// Nothing may leave through the C boundary: host allocation failures come back as Q3ASR_ERR_NOMEM.
template <typename F>
int tok_guarded(F&& fn) {
    try {
        return fn();
    } catch (const std::bad_alloc&) {
        return Q3ASR_ERR_NOMEM;
    } catch (const std::exception&) {
        return Q3ASR_ERR_INVALID;
    }
}
}  // namespace

extern "C" {

int q3asr_tokenizer_load(const char* path, q3asr_tokenizer** out) {
    if (path == nullptr || out == nullptr) return Q3ASR_ERR_INVALID;
    *out = nullptr;
    return tok_guarded([&]() {
        std::unique_ptr<q3asr_tokenizer> t(new q3asr_tokenizer());
        const bool ok = t->load_dir(path);
        *out = t.release();  // returned even on failure so the caller can read the message, then destroy it
        return ok ? Q3ASR_OK : Q3ASR_ERR_IO;
    });
}

int q3asr_tokenizer_from_pairs(const int32_t* ids, const char* const* tokens, int n, q3asr_tokenizer** out) {
    if ((n > 0 && (ids == nullptr || tokens == nullptr)) || n < 0 || out == nullptr) return Q3ASR_ERR_INVALID;
    *out = nullptr;
    for (int i = 0; i < n; i++)
        if (tokens[i] == nullptr) return Q3ASR_ERR_INVALID;
    return tok_guarded([&]() {
        std::unique_ptr<q3asr_tokenizer> t(new q3asr_tokenizer());
        for (int i = 0; i < n; i++) t->add(ids[i], tokens[i]);
        *out = t.release();
        return Q3ASR_OK;
    });
}

int q3asr_tokenizer_add_merge(q3asr_tokenizer* t, const char* first, const char* second) {
    if (t == nullptr || first == nullptr || second == nullptr) return Q3ASR_ERR_INVALID;
    return tok_guarded([&]() {
        t->merge_rank[std::string(first) + " " + second] = (int)t->n_merges++;
        return Q3ASR_OK;
    });
}

void q3asr_tokenizer_destroy(q3asr_tokenizer* t) { delete t; }

const char* q3asr_tokenizer_last_error(const q3asr_tokenizer* t) { return t ? t->last_error.c_str() : ""; }

int q3asr_tokenizer_size(const q3asr_tokenizer* t, int* n_tokens, int* n_merges) {
    if (t == nullptr) return Q3ASR_ERR_INVALID;
    if (n_tokens) *n_tokens = (int)t->id_to_token.size();
    if (n_merges) *n_merges = (int)t->n_merges;
    return Q3ASR_OK;
}

int q3asr_tokenizer_decode(const q3asr_tokenizer* t, const int32_t* ids, int n, char* out, size_t cap, size_t* needed) {
    if (t == nullptr || (n > 0 && ids == nullptr) || n < 0) return Q3ASR_ERR_INVALID;
    return tok_guarded([&]() {
        const std::string s = t->decode(ids, n);
        if (needed) *needed = s.size() + 1;
        if (out == nullptr || cap < s.size() + 1) return out == nullptr && needed ? Q3ASR_OK : Q3ASR_ERR_NOMEM;
        memcpy(out, s.c_str(), s.size() + 1);
        return Q3ASR_OK;
    });
}

int q3asr_tokenizer_encode(const q3asr_tokenizer* t, const char* text, int32_t* ids, int cap, int* n) {
    if (t == nullptr || text == nullptr || n == nullptr) return Q3ASR_ERR_INVALID;
    return tok_guarded([&]() {
        const std::vector<int32_t> v = t->encode(text);
        *n = (int)v.size();
        if (ids == nullptr) return Q3ASR_OK;
        if (cap < (int)v.size()) return Q3ASR_ERR_NOMEM;
        if (!v.empty()) memcpy(ids, v.data(), sizeof(int32_t) * v.size());
        return Q3ASR_OK;
    });
}

int q3asr_tokenizer_token_id(const q3asr_tokenizer* t, const char* token) {
    if (t == nullptr || token == nullptr) return -1;
    try {
        auto it = t->token_to_id.find(token);
        return it == t->token_to_id.end() ? -1 : it->second;
    } catch (const std::exception&) {
        return -1;
    }
}

}  // extern "C"
