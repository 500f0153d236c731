"""TextPreprocessor of the forced aligner (/root/reference/Sources/Qwen3ASR/TextPreprocessing.swift:44-308), host side.

The default path — whitespace split, one token per Han ideograph, punctuation kept on the surface form only — is reproduced here
(pure Python, `unicodedata` for the general categories).  The Japanese / Korean / Thai / Lao / Khmer / Burmese / Tibetan paths of the
reference call Apple's NLTokenizer (TextPreprocessing.swift:2, 101-160) and are not available: those languages raise.
The reference's unit tests for this path (Tests/Qwen3ASRTests/ForcedAlignerTests.swift:14-48, 140-211) are ported in
tests/test_text_preprocessing.py."""
import unicodedata
from collections import namedtuple

WordPair = namedtuple("WordPair", ["surface", "cleaned"])
SlottedText = namedtuple("SlottedText", ["token_ids", "timestamp_positions", "words"])

TIMESTAMP_TOKEN_ID = 151705  # <|timestamp|>, Qwen3ASR.swift:62

_KEPT = {"Lu", "Ll", "Lt", "Lm", "Lo", "Nd", "Nl", "No", "Mn", "Mc", "Me"}  # TextPreprocessing.swift:273-290
_NL_ONLY = (("japanese", "ja"), ("korean", "ko"), ("thai", "th"), ("lao", "lo"), ("khmer", "km"), ("burmese", "my"), ("myanmar", None),
            ("tibetan", "bo"))


def is_kept_scalar(ch):
    return ch == "'" or unicodedata.category(ch) in _KEPT


def clean_token(token):
    """Letters, numbers, combining marks and the ASCII apostrophe (TextPreprocessing.swift:262-271)."""
    return "".join(ch for ch in token if is_kept_scalar(ch))


def is_han_ideograph(ch):
    """TextPreprocessing.swift:296-306: CJK Unified + Extensions A-E + Compatibility; kana and Hangul excluded."""
    v = ord(ch)
    return (0x4E00 <= v <= 0x9FFF or 0x3400 <= v <= 0x4DBF or 0x20000 <= v <= 0x2A6DF or 0x2A700 <= v <= 0x2B73F
            or 0x2B740 <= v <= 0x2B81F or 0x2B820 <= v <= 0x2CEAF or 0xF900 <= v <= 0xFAFF)


_WHITE_SPACE = frozenset([*range(9, 14), 0x20, 0x85, 0xA0, 0x1680, *range(0x2000, 0x200B), 0x2028, 0x2029, 0x202F, 0x205F, 0x3000])


def _segments(text):
    """text.split(whereSeparator: \\.isWhitespace) (TextPreprocessing.swift:166): the Unicode White_Space property (str.split() would
    also break on U+001C..U+001F)."""
    out, cur = [], ""
    for ch in text:
        if ord(ch) in _WHITE_SPACE:
            if cur:
                out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur:
        out.append(cur)
    return out


def _pairs_for_segment(seg):
    """TextPreprocessing.swift:191-243."""
    if not any(is_han_ideograph(ch) for ch in seg):
        cleaned = clean_token(seg)
        return [WordPair(seg, cleaned)] if cleaned else []
    pairs, buf = [], ""

    def flush(before_han):
        nonlocal buf
        if not buf:
            return
        cleaned = clean_token(buf)
        if not cleaned:
            if pairs:                       # pure punctuation rides on the previous pair's surface
                pairs[-1] = WordPair(pairs[-1].surface + buf, pairs[-1].cleaned)
                buf = ""
            elif not before_han:            # trailing punctuation with no anchor at all: dropped
                buf = ""
            return                          # leading punctuation waits for the upcoming Han
        pairs.append(WordPair(buf, cleaned))
        buf = ""

    for ch in seg:
        if is_han_ideograph(ch):
            flush(True)
            if buf:
                pairs.append(WordPair(buf + ch, ch))
                buf = ""
            else:
                pairs.append(WordPair(ch, ch))
        else:
            buf += ch
    flush(False)
    return pairs


def tokenize_space_lang_pairs(text):
    """TextPreprocessing.swift:163-184."""
    pairs = []
    for segment in _segments(text):
        seg_pairs = _pairs_for_segment(segment)
        if not seg_pairs:
            if pairs:
                pairs[-1] = WordPair(pairs[-1].surface + segment, pairs[-1].cleaned)
            continue
        pairs.extend(seg_pairs)
    return pairs


def split_into_word_pairs(text, language="English"):
    """TextPreprocessing.swift:97-115."""
    lang = language.lower()
    for name, code in _NL_ONLY:
        if name in lang or (code is not None and lang == code):
            raise NotImplementedError(f"{language}: the reference segments this language with Apple's NLTokenizer "
                                      "(TextPreprocessing.swift:101-160), which is not available here")
    return tokenize_space_lang_pairs(text)


def split_into_words(text, language="English"):
    return [p.cleaned for p in split_into_word_pairs(text, language)]


def prepare_for_alignment(text, tokenizer, language="English", timestamp_token_id=TIMESTAMP_TOKEN_ID):
    """TextPreprocessor.prepareForAlignment (TextPreprocessing.swift:48-87): <timestamp> word-tokens <timestamp> per word; a word
    the tokenizer cannot encode hands its surface to the previous word."""
    token_ids, positions, words = [], [], []
    for pair in split_into_word_pairs(text, language):
        toks = [int(t) for t in tokenizer.encode(pair.cleaned)]
        if not toks:
            if words:
                words[-1] += pair.surface
            continue
        positions.append(len(token_ids))
        token_ids.append(timestamp_token_id)
        token_ids.extend(toks)
        positions.append(len(token_ids))
        token_ids.append(timestamp_token_id)
        words.append(pair.surface)
    return SlottedText(token_ids, positions, words)
