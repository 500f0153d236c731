"""TextPreprocessor of the forced aligner, default (whitespace + per-Han) path: the reference's own unit tests
(Tests/Qwen3ASRTests/ForcedAlignerTests.swift:14-48, 140-211) ported one to one.  Host logic, no GPU."""
import pytest

from q3asr import text as tp


def test_english():                                   # testTextPreprocessingEnglish
    assert tp.split_into_words("Hello world test", "English") == ["Hello", "world", "test"]


def test_chinese():                                   # testTextPreprocessingChinese
    assert tp.split_into_words("你好世界", "Chinese") == ["你", "好", "世", "界"]


def test_mixed_han_latin():                           # testTextPreprocessingMixedHanLatin
    assert tp.split_into_words("Hello你好world", "Chinese") == ["Hello", "你", "好", "world"]


def test_punctuation_stripped():                      # testTextPreprocessingPunctuationStripped
    assert tp.split_into_words("Hello, world!", "English") == ["Hello", "world"]


def test_apostrophe_kept():                           # testTextPreprocessingApostropheKept
    assert tp.split_into_words("don't stop", "English") == ["don't", "stop"]


def test_hindi_marks_preserved():                     # testTextPreprocessingHindiMarksPreserved
    words = tp.split_into_words("नमस्ते दोस्त", "hindi")
    assert words == ["नमस्ते", "दोस्त"]


def test_bengali_marks_preserved():                   # testTextPreprocessingBengaliMarksPreserved
    words = tp.split_into_words("নমস্কার বন্ধু", "bengali")
    assert len(words) == 2 and words[0] == "নমস্কার"


def test_german():                                    # testTextPreprocessingGerman
    assert tp.split_into_words("Guten Morgen, Donaudampfschifffahrtsgesellschaft!", "german") == \
        ["Guten", "Morgen", "Donaudampfschifffahrtsgesellschaft"]


def _surf(pairs):
    return [p.surface for p in pairs], [p.cleaned for p in pairs]


def test_surface_english_punctuation():               # testSurfacePreservesEnglishPunctuation
    assert _surf(tp.split_into_word_pairs("Hello, world! How are you?", "English")) == \
        (["Hello,", "world!", "How", "are", "you?"], ["Hello", "world", "How", "are", "you"])


def test_surface_apostrophe_and_trailing_period():    # testSurfacePreservesApostropheAndTrailingPeriod
    assert _surf(tp.split_into_word_pairs("you're great.", "English")) == (["you're", "great."], ["you're", "great"])


def test_surface_leading_punctuation():               # testSurfacePreservesLeadingPunctuation
    assert _surf(tp.split_into_word_pairs('"Hello" she said.', "English")) == (['"Hello"', "she", "said."], ["Hello", "she", "said"])


def test_surface_cjk_punctuation():                   # testSurfacePreservesCJKPunctuation
    assert _surf(tp.split_into_word_pairs("你好，世界。", "Chinese")) == (["你", "好，", "世", "界。"], ["你", "好", "世", "界"])


def test_surface_mixed_han_latin_punctuation():       # testSurfacePreservesMixedHanLatinPunctuation
    assert _surf(tp.split_into_word_pairs("Hello, 你好world.", "Chinese")) == \
        (["Hello,", "你", "好", "world."], ["Hello", "你", "好", "world"])


def test_more_edge_cases():
    # a stray punctuation-only segment rides on the previous word; leading punctuation waits for a Han anchor; kana are not split
    assert _surf(tp.split_into_word_pairs("wait — what", "English")) == (["wait—", "what"], ["wait", "what"])
    assert _surf(tp.split_into_word_pairs("「你好」", "Chinese")) == (["「你", "好」"], ["你", "好"])
    assert tp.split_into_words("カタカナ test", "English") == ["カタカナ", "test"]
    assert tp.split_into_words("  ", "English") == [] and tp.split_into_words("!!!", "English") == []
    assert tp.is_han_ideograph("你") and not tp.is_han_ideograph("か") and not tp.is_han_ideograph("한")
    for lang in ("Japanese", "ja", "korean", "Thai", "lo", "khmer", "myanmar", "bo"):
        with pytest.raises(NotImplementedError):
            tp.split_into_word_pairs("x", lang)


def test_prepare_for_alignment_slots():
    class Tok:                                        # a toy tokenizer: one id per character, nothing for digits
        def encode(self, w):
            return [] if w.isdigit() else [100 + (ord(c) % 50) for c in w]
    st = tp.prepare_for_alignment("Hi, 42 you!", Tok(), "English", timestamp_token_id=7)
    assert st.words == ["Hi,42", "you!"]              # "42" is unencodable: its surface joins the previous word (TextPreprocessing.swift:63-70)
    assert st.token_ids[0] == 7 and st.token_ids.count(7) == 4
    assert st.timestamp_positions == [0, 3, 4, 8]
    assert all(st.token_ids[p] == 7 for p in st.timestamp_positions)


# ---- the library's splitter (csrc/text.cu, q3asr_text_word_pairs): same reference cases, then agreement with the Python one ----
REFERENCE_CASES = [  # (text, language, surfaces, cleaned) — ForcedAlignerTests.swift:14-48, 140-211
    ("Hello world test", "English", ["Hello", "world", "test"], ["Hello", "world", "test"]),
    ("你好世界", "Chinese", ["你", "好", "世", "界"], ["你", "好", "世", "界"]),
    ("Hello你好world", "Chinese", ["Hello", "你", "好", "world"], ["Hello", "你", "好", "world"]),
    ("Hello, world!", "English", ["Hello,", "world!"], ["Hello", "world"]),
    ("don't stop", "English", ["don't", "stop"], ["don't", "stop"]),
    ("नमस्ते दोस्त", "hindi", ["नमस्ते", "दोस्त"], ["नमस्ते", "दोस्त"]),
    ("Guten Morgen, Donaudampfschifffahrtsgesellschaft!", "german", ["Guten", "Morgen,", "Donaudampfschifffahrtsgesellschaft!"],
     ["Guten", "Morgen", "Donaudampfschifffahrtsgesellschaft"]),
    ("Hello, world! How are you?", "English", ["Hello,", "world!", "How", "are", "you?"], ["Hello", "world", "How", "are", "you"]),
    ("you're great.", "English", ["you're", "great."], ["you're", "great"]),
    ('"Hello" she said.', "English", ['"Hello"', "she", "said."], ["Hello", "she", "said"]),
    ("你好，世界。", "Chinese", ["你", "好，", "世", "界。"], ["你", "好", "世", "界"]),
    ("Hello, 你好world.", "Chinese", ["Hello,", "你", "好", "world."], ["Hello", "你", "好", "world"]),
    ("wait — what", "English", ["wait—", "what"], ["wait", "what"]),
    ("「你好」", "Chinese", ["「你", "好」"], ["你", "好"]),
    ("  ", "English", [], []),
    ("!!!", "English", [], []),
]


@pytest.fixture(scope="module")
def q3(built_lib):
    return built_lib


@pytest.mark.parametrize("text,lang,surfaces,cleaned", REFERENCE_CASES)
def test_c_splitter_reference_cases(q3, text, lang, surfaces, cleaned):
    got = q3.text_word_pairs(text, lang)
    assert [s for s, _ in got] == surfaces and [c for _, c in got] == cleaned


def test_c_splitter_refuses_nltokenizer_languages(q3):
    for lang in ("Japanese", "ja", "korean", "Thai", "lo", "khmer", "myanmar", "bo"):
        with pytest.raises(q3.Q3Error) as e:
            q3.text_word_pairs("x", lang)
        assert e.value.code == 1 and "NLTokenizer" in str(e.value)
    assert q3.text_word_pairs("x y", None) == [("x", "x"), ("y", "y")]     # NULL language = English


def test_c_splitter_matches_python_on_mixed_scripts(q3):
    """Random text over many scripts, punctuation, combining marks, astral code points and every kind of white space."""
    import random
    rnd = random.Random(11)
    alphabet = (list("abcXYZ019'’\".,!?-—()[]«»…·") + list("äöüßéñçøåğİıșț") + list("привет") + list("Ελληνικά") + list("שלוםمرحبا")
                + list("नमस्तेবন্ধু") + list("你好世界漢字㐀𠀀𪜀") + list("かなカナ한글") + list("ᠮᠣᠩ") + list("́̈⃝ा")
                + list("①Ⅷ½٣") + list("😀𝒜€©™°") + list(" \t\n\r\x0b\x0c\x85       　") + list("\x1c\x1f​﻿"))
    for _ in range(600):
        text = "".join(rnd.choice(alphabet) for _ in range(rnd.randrange(0, 40)))
        want = [(p.surface, p.cleaned) for p in tp.split_into_word_pairs(text, "English")]
        assert q3.text_word_pairs(text, "English") == want, repr(text)


def test_c_splitter_every_code_point_classified_like_python(q3):
    """The generated category table (csrc/unicode_kept.inc) against unicodedata, one probe per code point in bulk: the cleaned form of a
    string holding every code point of a block keeps exactly the letters, numbers and marks."""
    import unicodedata
    for base in range(0, 0x110000, 0x1000):
        cps = [cp for cp in range(max(base, 1), base + 0x1000) if not 0xD800 <= cp <= 0xDFFF and cp not in tp._WHITE_SPACE
               and not tp.is_han_ideograph(chr(cp))]
        if not cps:
            continue
        s = "".join(map(chr, cps))
        got = q3.text_word_pairs(s, "English")
        want = "".join(c for c in s if c == "'" or unicodedata.category(c) in tp._KEPT)
        assert (got[0][1] if got else "") == want, hex(base)


def test_c_splitter_invalid_utf8_is_carried_on_the_surface_only(q3):
    got = q3.text_word_pairs(b"ab\xff\xfecd \xe4\xbd x\xc0\xaf \xed\xa0\x80y \xf4\x90\x80\x80z", "English")
    # stray 0xFF/0xFE, a truncated 3-byte sequence (rides on the previous word like punctuation), an overlong "/", a surrogate, > U+10FFFF
    assert got == [(b"ab\xff\xfecd\xe4\xbd", b"abcd"), (b"x\xc0\xaf", b"x"), (b"\xed\xa0\x80y", b"y"), (b"\xf4\x90\x80\x80z", b"z")]
    # a truncated sequence at the very end of the text must not be read past
    assert q3.text_word_pairs(b"ok \xe4\xbd", "English") == [(b"ok\xe4\xbd", b"ok")]
    assert q3.text_word_pairs(b"\xf0\x9f", "English") == []


def test_c_prepare_for_alignment_matches_python(q3):
    """q3asr_text_prepare_for_alignment (native splitter + native tokenizer) against q3asr.text.prepare_for_alignment driving the same
    tokenizer: slotted ids, <timestamp> positions and surface words."""
    import random
    # a character-level vocabulary without digits: "42" is unencodable and must ride on the previous word (TextPreprocessing.swift:63-70)
    # (Han and accented letters are missing too: "你" is a word of its own that cannot be encoded)
    vocab = {i + 10: c for i, c in enumerate("abcdefghijklmnopqrstuvwxyzHW'")}
    tok = q3.Qwen3Tokenizer(id_to_token=vocab)
    try:
        ids, pos, words = q3.text_prepare_for_alignment(tok, "Hi, 42 you!", "English", timestamp_token_id=7)
        assert words == ["Hi,42", "you!"] and pos == [0, 3, 4, 8] and ids.count(7) == 4 and all(ids[p] == 7 for p in pos)
        assert q3.text_prepare_for_alignment(tok, "  ...  ", "English", 7) == ([], [], [])
        assert q3.text_prepare_for_alignment(tok, "12 34", "English", 7) == ([], [], [])          # nothing encodable, no previous word
        rnd = random.Random(3)
        alphabet = list("abcxyzHW'") + list(" ,.!?-—\t\n") + list("0123") + ["你", "好", "。", "，", "ä", "😀"]
        for _ in range(300):
            text = "".join(rnd.choice(alphabet) for _ in range(rnd.randrange(0, 30)))
            want = tp.prepare_for_alignment(text, tok, "English", timestamp_token_id=7)
            got = q3.text_prepare_for_alignment(tok, text, "English", timestamp_token_id=7)
            assert got == (want.token_ids, want.timestamp_positions, want.words), repr(text)
        with pytest.raises(q3.Q3Error):
            q3.text_prepare_for_alignment(tok, "x", "Japanese", 7)
    finally:
        tok.close()
