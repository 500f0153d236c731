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
