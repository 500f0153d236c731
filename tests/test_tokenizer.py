"""The tokenizer behind q3asr_tokenizer_* against the reference's own unit tests, ported one for one from
/root/reference/Tests/Qwen3ASRTests/Qwen3ASRTests.swift:270-451 (byte-level decode across token boundaries, CJK, special
tokens, truncated UTF-8), plus encode / file-loading checks built from Tokenizer.swift's documented behaviour.  No GPU needed."""
import json
import os

import pytest


def bpe_char(byte):
    """GPT-2 byte-to-unicode (Qwen3ASRTests.swift:332-343)."""
    direct = lambda b: 33 <= b <= 126 or 161 <= b <= 172 or 174 <= b <= 255
    if direct(byte):
        return chr(byte)
    n = 0
    for b in range(256):
        if direct(b):
            continue
        if b == byte:
            return chr(0x100 + n)
        n += 1
    raise AssertionError


def bpe_token(bs):
    return "".join(bpe_char(b) for b in bs)


def token_map(bytes_=(), literals=()):
    m = {i: bpe_token(b) for i, b in bytes_}
    m.update(dict(literals))
    return m


UTF8_MAP = {  # makeUTF8TokenMap, Qwen3ASRTests.swift:274-294
    100: "Hello", 101: "Ġworld", 102: "ä", 103: "¾", 104: "Ĩ", 105: "å", 106: "¥", 107: "½",
    108: "Ġä¾Ĩ", 200: "<asr_text>", 201: "<|im_start|>",
}


@pytest.fixture()
def tok(built_lib):
    made = []

    def make(id_to_token, merges=()):
        t = built_lib.Qwen3Tokenizer(id_to_token=id_to_token, merges=merges)
        made.append(t)
        return t
    yield make
    for t in made:
        t.close()


@pytest.mark.parametrize("ids,want", [
    ([102, 103, 104], "來"),                                   # testDecodeCJKSplitAcrossThreeTokens
    ([100, 102, 103, 104], "Hello來"),                         # testDecodeMixedASCIIAndCJK
    ([102, 103, 104, 105, 106, 107], "來好"),                  # testDecodeConsecutiveMultiByteCharacters
    ([100, 108], "Hello 來"),                                  # testDecodeGPrefixBeforeMultiByteCharacter
    ([102, 103, 104, 200, 105, 106, 107], "來<asr_text>好"),   # testDecodeASRTextMarkerBetweenMultiByteSequences
    ([100, 101], "Hello world"),
    ([201, 100, 201], "Hello"),                                # <|...|> specials are dropped
    ([101], "world"),                                          # result is trimmed
])
def test_reference_decode_vectors(tok, ids, want):
    assert tok(UTF8_MAP).decode(ids) == want


def test_truncated_utf8_falls_back_to_replacement_char(tok):
    t = tok(token_map([(1, [0xE4]), (2, [0xBE])]))
    assert "�" in t.decode([1, 2])


def test_special_token_skipped_between_cjk_bytes(tok):
    t = tok(token_map([(1, [0xE4]), (2, [0xBE]), (3, [0x86])], [(900, "<|im_end|>")]))
    assert t.decode([1, 2, 900, 3]) == "來"


def test_4byte_utf8_cjk_extension_b(tok):
    t = tok(token_map([(1, [0xF0]), (2, [0xA0]), (3, [0x80]), (4, [0x80])]))
    assert t.decode([1, 2, 3, 4]) == "\U00020000"


def test_bpe_token_spanning_utf8_boundary(tok):
    t = tok(token_map([(1, [0xE4]), (2, [0xBE]), (3, [0x86, 0xE5]), (4, [0xA5]), (5, [0xBD])]))
    assert t.decode([1, 2, 3, 4, 5]) == "來好"


def test_korean_hangul(tok):
    t = tok(token_map([(1, [0xED]), (2, [0x95]), (3, [0x9C]), (4, [0xEA]), (5, [0xB5]), (6, [0xAD])]))
    assert t.decode([1, 2, 3, 4, 5, 6]) == "한국"


def test_japanese_mixed_hiragana_kanji(tok):
    t = tok(token_map([(1, [0xE3]), (2, [0x81]), (3, [0x93]), (4, [0x82]), (5, [0xAB]), (6, [0xA1]), (7, [0xAF]), (8, [0xE6]),
                       (9, [0x97]), (10, [0xA5]), (11, [0x9C]), (12, [0xAC]), (13, [0xE8]), (14, [0xAA]), (15, [0x9E])]))
    ids = [1, 2, 3, 1, 4, 3, 1, 2, 5, 1, 2, 6, 1, 2, 7, 8, 9, 10, 8, 11, 12, 13, 14, 15]
    assert t.decode(ids) == "こんにちは日本語"


def test_unknown_token_id_between_cjk_bytes(tok):
    t = tok(token_map([(1, [0xE4]), (2, [0xBE]), (3, [0x86])]))
    assert t.decode([1, 2, 999, 3]) == "來"


def test_every_byte_round_trips_through_the_byte_table(tok):
    t = tok({b: bpe_char(b) for b in range(256)})
    for s in ("naïve café", "日本語", "tab\there", "😀 ok"):
        raw = s.encode("utf-8")
        assert t.decode(list(raw)) == s.strip(" \t")


def test_encode_with_merges_and_character_fallback(tok):
    vocab = {0: "l", 1: "o", 2: "w", 3: "e", 4: "r", 5: "lo", 6: "low", 7: "er", 8: "Ġ", 9: "Ġlow", 10: "Ġl", 11: "n"}
    merges = [("l", "o"), ("lo", "w"), ("e", "r"), ("Ġ", "l"), ("Ġl", "ow")]
    t = tok(vocab, merges)
    assert t.encode("low") == [6]
    assert t.encode("lower") == [6, 7]
    # the separator starts the next word; "Ġ l" merges (rank 3) after "l o" / "lo w" (ranks 0, 1) have fired
    assert t.encode("low low") == [6, 8, 6]
    assert t.encode("lowx") == [6]           # pieces missing from the vocabulary are dropped (Tokenizer.swift:208-212)
    assert t.decode(t.encode("low lower")) == "low lower"
    plain = tok(vocab)                          # no merges: per-character lookup (Tokenizer.swift:268-278)
    assert plain.encode("lone") == [0, 1, 11, 3]


def test_crlf_is_one_character_and_every_newline_scalar_counts_a_merge_line(tmp_path, built_lib):
    """Tokenizer.swift:218-238 walks Swift Characters: "\r\n" is ONE grapheme that is not equal to "\n", so a CRLF inside a
    context string does not start a new word.  Tokenizer.swift:94: merges.txt is split on CharacterSet.newlines, which also holds
    VT, FF, U+0085, U+2028 and U+2029 — the line index is the merge rank, so lines ended by those must be counted."""
    vocab = {"a": 0, "b": 1, "c": 2, "ab": 3, "bc": 4, "ĊĊ": 5, "Ċ": 6, "č": 7, "čĊ": 8, "ačĊb": 9, "Ċb": 10}
    (tmp_path / "vocab.json").write_text(json.dumps(vocab), encoding="utf-8")
    # ranks: "b c" sits on line index 1 only if the U+2028 after the first line is a line break; then "bc" wins over "ab" for "abc"
    # exactly when its rank is lower than that of "a b", which comes two (U+0085, FF) line breaks later
    (tmp_path / "merges.txt").write_bytes("#v\u2028b c\u0085\x0ca b\nč Ċ\na čĊ\načĊ b\n".encode("utf-8"))
    t = built_lib.Qwen3Tokenizer(path=str(tmp_path))
    try:
        assert t.size()[1] == 5
        assert t.encode("abc") == [0, 4]           # "b c" (line 1) outranks "a b" (line 3): a + bc
        assert t.encode("a\r\nb") == [9]            # one word: a, CR, LF, b merge all the way ("\r\n" did not split)
        assert t.encode("a\nb") == [0, 6, 1]        # a bare LF does split: "a", then the word "\nb" (no merge for it: two pieces)
    finally:
        t.close()


def test_load_from_directory(tmp_path, built_lib):
    vocab = {"Hello": 100, "Ġworld": 101, "!": 0, "Ċ": 1, "H": 2, "e": 3, "l": 4, "o": 5, "He": 6, "ll": 7, "Hell": 8}
    (tmp_path / "vocab.json").write_text(json.dumps(vocab), encoding="utf-8")
    (tmp_path / "tokenizer_config.json").write_text(json.dumps({
        "added_tokens_decoder": {"151644": {"content": "<|im_start|>", "special": True}, "151704": {"content": "<asr_text>"},
                                 "bogus": {"content": "x"}}, "model_max_length": 1000}), encoding="utf-8")
    (tmp_path / "merges.txt").write_text("#version: 0.2\nH e\nl l\nHe ll\n\nHell o\nbad line here\n", encoding="utf-8")
    t = built_lib.Qwen3Tokenizer(path=str(tmp_path))
    try:
        n_tok, n_merges = t.size()
        assert (n_tok, n_merges) == (len(vocab) + 2, 4)
        assert t.token_id("<asr_text>") == 151704 and t.token_id("nope") is None
        assert t.decode([151644, 100, 101, 0, 151704]) == "Hello world!<asr_text>"
        assert t.encode("Hello") == [100]        # H e -> He, l l -> ll, He ll -> Hell, Hell o -> Hello
        assert t.decode([100, 1]) == "Hello\n"   # newlines are not trimmed (CharacterSet.whitespaces)
        t2 = built_lib.Qwen3Tokenizer(path=str(tmp_path / "vocab.json"))
        assert t2.size() == (n_tok, n_merges)
        t2.close()
    finally:
        t.close()
    with pytest.raises(built_lib.Q3Error):
        built_lib.Qwen3Tokenizer(path=str(tmp_path / "missing"))
    (tmp_path / "bad").mkdir()
    (tmp_path / "bad" / "vocab.json").write_text("[1, 2]")
    with pytest.raises(built_lib.Q3Error) as e:
        built_lib.Qwen3Tokenizer(path=str(tmp_path / "bad"))
    assert "Invalid tokenizer format" in str(e.value)


def test_hostile_tokenizer_files_are_refused_not_crashed_on(tmp_path, built_lib):
    """A vocab.json of a million '[' must not become a million stack frames; truncated escapes and lone surrogates must not read past the end."""
    for name, body in [("deep", '{"a":' + "[" * 1_000_000), ("deep_obj", '{"a":' * 200_000), ("esc", '{"a\\'), ("uni", '{"\\u12'),
                       ("sur", '{"\\ud83d\\u": 1}'), ("num", '{"a": 1e99999, "b": -}'), ("inf", '{"a": 1e999}'), ("big", '{"a": 4294967296}'), ("empty", "")]:
        d = tmp_path / name
        d.mkdir()
        (d / "vocab.json").write_text(body)
        with pytest.raises(built_lib.Q3Error) as e:
            built_lib.Qwen3Tokenizer(path=str(d))
        assert "Invalid tokenizer format" in str(e.value) or "vocab" in str(e.value), (name, str(e.value))
    # a lone high surrogate followed by a normal escape is tolerated (garbage in the token, nothing else)
    d = tmp_path / "lone"
    d.mkdir()
    (d / "vocab.json").write_text('{"\\ud83dx": 1, "ok": 2}')
    t = built_lib.Qwen3Tokenizer(path=str(d))
    assert t.token_id("ok") == 2
    t.close()


def test_from_pairs_argument_checks(built_lib):
    import ctypes
    import numpy as np
    L = built_lib.lib()
    ids = np.array([1, 2], dtype=np.int32)
    toks = (ctypes.c_char_p * 2)(b"a", None)                      # a NULL token string
    out = ctypes.c_void_p(0xdead)
    assert L.q3asr_tokenizer_from_pairs(ids.ctypes.data, ctypes.cast(toks, ctypes.c_void_p), 2, ctypes.byref(out)) == 1 and not out.value
    assert L.q3asr_tokenizer_from_pairs(None, None, 2, ctypes.byref(out)) == 1
    assert L.q3asr_tokenizer_from_pairs(None, None, -1, ctypes.byref(out)) == 1
    assert L.q3asr_tokenizer_from_pairs(None, None, 0, ctypes.byref(out)) == 0 and out.value    # an empty tokenizer is fine
    n = ctypes.c_int(-1)
    assert L.q3asr_tokenizer_encode(out, b"abc", None, 0, ctypes.byref(n)) == 0 and n.value == 0
    assert L.q3asr_tokenizer_add_merge(out, None, b"b") == 1 and L.q3asr_tokenizer_token_id(out, None) == -1
    L.q3asr_tokenizer_destroy(out)
    L.q3asr_tokenizer_destroy(None)


def test_transcript_extraction_from_generated_ids(built_lib):
    """Qwen3ASRModel.transcribe's tail (Qwen3ASR.swift:283-293): decode, keep what follows "<asr_text>" trimmed of CharacterSet.whitespaces
    (space separators and TAB, not line breaks); without the marker the decoded text as it is; without a tokenizer the ids joined by
    spaces.  The layout is the one testTokenizerDecodeWithASRMarker builds: "language English<asr_text>Hello"."""
    m = built_lib.Qwen3ASRModel.__new__(built_lib.Qwen3ASRModel)      # no GPU handle needed for this host logic
    m._h = None
    m.tokenizer = None
    assert m._text_of([11528, 6364, 151704, 9707]) == "11528 6364 151704 9707"
    tok = built_lib.Qwen3Tokenizer(id_to_token={1: "language", 2: "ĠEnglish", 3: "<asr_text>", 4: "Hello", 5: "Ġworld", 6: "ĉ", 7: "Ċ",
                                                8: "Âł", 9: "<|im_end|>"})
    m.tokenizer = tok
    try:
        assert tok.decode([1, 2, 3, 4]) == "language English<asr_text>Hello"
        assert m._text_of([1, 2, 3, 4, 5, 9]) == "Hello world"
        assert m._text_of([1, 2, 3, 5, 4]) == "worldHello"               # the space of "Ġworld" right after the marker is trimmed
        assert m._text_of([3, 6, 8, 4, 7]) == "Hello\n"                   # TAB and NBSP go, the trailing line break stays
        assert m._text_of([4, 5]) == "Hello world"                        # no marker: the decoded text
        assert m._text_of([3]) == "" and m._text_of([]) == ""
    finally:
        m.tokenizer = None
        tok.close()
