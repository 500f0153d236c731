// text.cu — the forced aligner's word splitter on the host (no GPU work in this file).
//
// TextPreprocessor.splitIntoWordPairs, default path (/root/reference/Sources/Qwen3ASR/TextPreprocessing.swift:97-115, 163-243,
// 262-306): split on Unicode white space; inside a segment every Han ideograph is a word of its own; runs of other scalars are a
// word when they hold a letter, number or combining mark, otherwise (pure punctuation) they ride on a neighbour's surface form.
// `surface` keeps the punctuation, `cleaned` is what the tokenizer sees.  The Japanese / Korean / Thai / Lao / Khmer / Burmese /
// Tibetan paths of the reference call Apple's NLTokenizer (:101-160) and stay on the Swift side: those languages are refused here.
//
// Input is UTF-8.  A byte sequence that is not valid UTF-8 is carried through on the surface form byte by byte and never kept in
// the cleaned form (Swift strings cannot hold one, so the reference has no behaviour to match there).
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "model.h"

namespace q3 {

namespace {

struct Range {
    uint32_t lo, hi;
};
const Range kKept[] = {
#include "unicode_kept.inc"
};

bool is_kept(uint32_t c) {  // :273-290 (+ the ASCII apostrophe)
    if (c == '\'') return true;
    size_t lo = 0, hi = sizeof(kKept) / sizeof(kKept[0]);
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (c > kKept[mid].hi) lo = mid + 1;
        else if (c < kKept[mid].lo) hi = mid;
        else return true;
    }
    return false;
}

bool is_han(uint32_t v) {  // :296-306
    return (v >= 0x4E00 && v <= 0x9FFF) || (v >= 0x3400 && v <= 0x4DBF) || (v >= 0x20000 && v <= 0x2A6DF) || (v >= 0x2A700 && v <= 0x2B73F) ||
           (v >= 0x2B740 && v <= 0x2B81F) || (v >= 0x2B820 && v <= 0x2CEAF) || (v >= 0xF900 && v <= 0xFAFF);
}

bool is_space(uint32_t c) {  // Unicode White_Space, what Character.isWhitespace tests (:166)
    return (c >= 9 && c <= 13) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 ||
           c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}

constexpr uint32_t kInvalid = 0xFFFFFFFFu;

// One scalar of UTF-8 text: its value (kInvalid for a byte that starts no valid sequence) and its byte length.
uint32_t next_scalar(const std::string& s, size_t i, size_t* len) {
    const unsigned char b0 = (unsigned char)s[i];
    *len = 1;
    if (b0 < 0x80) return b0;
    int need;
    uint32_t v, min;
    if (b0 >= 0xC2 && b0 <= 0xDF) { need = 1; v = b0 & 0x1F; min = 0x80; }
    else if (b0 >= 0xE0 && b0 <= 0xEF) { need = 2; v = b0 & 0x0F; min = 0x800; }
    else if (b0 >= 0xF0 && b0 <= 0xF4) { need = 3; v = b0 & 0x07; min = 0x10000; }
    else return kInvalid;
    if (i + (size_t)need >= s.size()) return kInvalid;  // truncated sequence
    for (int k = 1; k <= need; k++) {
        const unsigned char b = (unsigned char)s[i + (size_t)k];
        if ((b & 0xC0) != 0x80) return kInvalid;
        v = (v << 6) | (b & 0x3F);
    }
    if (v < min || v > 0x10FFFF || (v >= 0xD800 && v <= 0xDFFF)) return kInvalid;
    *len = (size_t)need + 1;
    return v;
}

struct Scalar {
    uint32_t value;
    size_t at, len;
};

std::vector<Scalar> scalars_of(const std::string& s) {
    std::vector<Scalar> out;
    for (size_t i = 0; i < s.size();) {
        size_t len;
        const uint32_t v = next_scalar(s, i, &len);
        out.push_back(Scalar{v, i, len});
        i += len;
    }
    return out;
}

std::string clean_token(const std::string& s) {  // :262-271
    std::string out;
    for (const Scalar& c : scalars_of(s))
        if (c.value != kInvalid && is_kept(c.value)) out.append(s, c.at, c.len);
    return out;
}

void pairs_for_segment(const std::string& seg, std::vector<WordPair>* pairs_out) {  // :191-243
    const std::vector<Scalar> sc = scalars_of(seg);
    bool has_han = false;
    for (const Scalar& c : sc) has_han = has_han || (c.value != kInvalid && is_han(c.value));
    if (!has_han) {
        std::string cleaned = clean_token(seg);
        if (!cleaned.empty()) pairs_out->push_back(WordPair{seg, std::move(cleaned)});
        return;
    }
    std::vector<WordPair> pairs;
    std::string buf;
    auto flush = [&](bool before_han) {
        if (buf.empty()) return;
        std::string cleaned = clean_token(buf);
        if (cleaned.empty()) {
            if (!pairs.empty()) {  // pure punctuation rides on the previous pair's surface
                pairs.back().surface += buf;
                buf.clear();
            } else if (!before_han) {  // trailing punctuation with no anchor at all: dropped
                buf.clear();
            }
            return;  // leading punctuation waits for the upcoming Han
        }
        pairs.push_back(WordPair{buf, std::move(cleaned)});
        buf.clear();
    };
    for (const Scalar& c : sc) {
        if (c.value != kInvalid && is_han(c.value)) {
            flush(true);
            const std::string han = seg.substr(c.at, c.len);
            pairs.push_back(WordPair{buf + han, han});
            buf.clear();
        } else {
            buf.append(seg, c.at, c.len);
        }
    }
    flush(false);
    for (WordPair& p : pairs) pairs_out->push_back(std::move(p));
}

std::string lowercase_ascii(std::string s) {
    for (char& c : s)
        if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
    return s;
}

}  // namespace

std::vector<WordPair> split_into_word_pairs(const std::string& text, const std::string& language) {  // :97-115
    const std::string lang = lowercase_ascii(language);
    static const struct { const char* name; const char* code; } kNLOnly[] = {
        {"japanese", "ja"}, {"korean", "ko"}, {"thai", "th"}, {"lao", "lo"}, {"khmer", "km"}, {"burmese", "my"}, {"myanmar", nullptr},
        {"tibetan", "bo"}};
    for (const auto& nl : kNLOnly)
        Q3_CHECK(lang.find(nl.name) == std::string::npos && !(nl.code && lang == nl.code), Q3ASR_ERR_INVALID,
                 "text: " + language + " is segmented with Apple's NLTokenizer in the reference (TextPreprocessing.swift:101-160); "
                 "split it on the Swift side and pass the slotted ids");
    std::vector<WordPair> pairs;
    const std::vector<Scalar> sc = scalars_of(text);
    size_t k = 0;
    while (k < sc.size()) {  // :163-184
        while (k < sc.size() && sc[k].value != kInvalid && is_space(sc[k].value)) k++;
        if (k == sc.size()) break;
        const size_t start = sc[k].at;
        while (k < sc.size() && !(sc[k].value != kInvalid && is_space(sc[k].value))) k++;
        const size_t end = k < sc.size() ? sc[k].at : text.size();
        const std::string segment = text.substr(start, end - start);
        const size_t before = pairs.size();
        pairs_for_segment(segment, &pairs);
        if (pairs.size() == before && !pairs.empty()) pairs.back().surface += segment;  // stray punctuation joins the previous word
    }
    return pairs;
}

}  // namespace q3

namespace {
thread_local std::string g_text_error;
}

extern "C" {

const char* q3asr_text_last_error(void) { return g_text_error.c_str(); }

int q3asr_text_word_pairs(const char* text, const char* language, char* buf, size_t cap, size_t* needed, int* n_pairs) {
    if (text == nullptr || needed == nullptr || n_pairs == nullptr) {
        g_text_error = "text_word_pairs: null argument";
        return Q3ASR_ERR_INVALID;
    }
    try {
        const std::vector<q3::WordPair> pairs = q3::split_into_word_pairs(text, language ? language : "English");
        std::string flat;
        for (const q3::WordPair& p : pairs) {
            flat.append(p.surface).push_back('\0');
            flat.append(p.cleaned).push_back('\0');
        }
        *needed = flat.size();
        *n_pairs = (int)pairs.size();
        if (buf == nullptr) return Q3ASR_OK;
        if (cap < flat.size()) {
            g_text_error = "text_word_pairs: buffer too small";
            return Q3ASR_ERR_NOMEM;
        }
        if (!flat.empty()) memcpy(buf, flat.data(), flat.size());
        return Q3ASR_OK;
    } catch (const q3::Error& e) {
        g_text_error = e.what();
        return e.code > 0 ? e.code : Q3ASR_ERR_INVALID;
    } catch (const std::exception& e) {
        g_text_error = e.what();
        return Q3ASR_ERR_NOMEM;
    }
}

int q3asr_text_prepare_for_alignment(const q3asr_tokenizer* tok, const char* text, const char* language, int32_t timestamp_id, int32_t* ids,
                                     int ids_cap, int* n_ids, int* positions, int pos_cap, int* n_positions, char* words, size_t words_cap,
                                     size_t* words_needed) {
    if (tok == nullptr || text == nullptr || n_ids == nullptr || n_positions == nullptr || words_needed == nullptr || ids_cap < 0 ||
        pos_cap < 0) {
        g_text_error = "text_prepare_for_alignment: null argument";
        return Q3ASR_ERR_INVALID;
    }
    try {  // TextPreprocessing.swift:48-87
        const std::vector<q3::WordPair> pairs = q3::split_into_word_pairs(text, language ? language : "English");
        std::vector<int32_t> out_ids, word_ids;
        std::vector<int> out_pos;
        std::vector<std::string> out_words;
        for (const q3::WordPair& p : pairs) {
            int n = 0;
            int rc = q3asr_tokenizer_encode(tok, p.cleaned.c_str(), nullptr, 0, &n);
            word_ids.assign((size_t)(n > 0 ? n : 0), 0);
            if (rc == Q3ASR_OK && n > 0) rc = q3asr_tokenizer_encode(tok, p.cleaned.c_str(), word_ids.data(), n, &n);
            if (rc != Q3ASR_OK) throw q3::Error(rc, "text_prepare_for_alignment: tokenizer failed on a word");
            if (n <= 0) {  // unencodable: its surface joins the previous word (:63-70)
                if (!out_words.empty()) out_words.back() += p.surface;
                continue;
            }
            out_pos.push_back((int)out_ids.size());
            out_ids.push_back(timestamp_id);
            out_ids.insert(out_ids.end(), word_ids.begin(), word_ids.begin() + n);
            out_pos.push_back((int)out_ids.size());
            out_ids.push_back(timestamp_id);
            out_words.push_back(p.surface);
        }
        std::string flat;
        for (const std::string& w : out_words) flat.append(w).push_back('\0');
        *n_ids = (int)out_ids.size();
        *n_positions = (int)out_pos.size();
        *words_needed = flat.size();
        if (ids == nullptr && positions == nullptr && words == nullptr) return Q3ASR_OK;  // sizing call
        if (ids == nullptr || positions == nullptr || words == nullptr || ids_cap < *n_ids || pos_cap < *n_positions || words_cap < flat.size()) {
            g_text_error = "text_prepare_for_alignment: buffer missing or too small";
            return Q3ASR_ERR_NOMEM;
        }
        std::copy(out_ids.begin(), out_ids.end(), ids);
        std::copy(out_pos.begin(), out_pos.end(), positions);
        if (!flat.empty()) memcpy(words, flat.data(), flat.size());
        return Q3ASR_OK;
    } catch (const q3::Error& e) {
        g_text_error = e.what();
        return e.code > 0 ? e.code : Q3ASR_ERR_INVALID;
    } catch (const std::exception& e) {
        g_text_error = e.what();
        return Q3ASR_ERR_NOMEM;
    }
}

}  // extern "C"
