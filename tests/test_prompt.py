"""Prompt layout (host logic of the C ABI, no GPU): q3asr_prompt_ids against the reference's own ContextInjectionTests
(Tests/Qwen3ASRTests/Qwen3ASRTests.swift:484-552, ported one to one), its special-token ids (:219-235), the template of
Qwen3ASR.swift:196-233 / ForcedAligner.swift:338-378 written out literally, and the oracle's prompt builder."""
import ctypes

import numpy as np
import pytest

from oracle import model as omodel
from oracle import weights

IM_START, IM_END, SYSTEM, USER, ASSISTANT, NEWLINE = 151644, 151645, 8948, 872, 77091, 198
AUDIO_START, AUDIO_END, AUDIO_PAD, ASR_TEXT = 151669, 151670, 151676, 151704


@pytest.fixture(scope="module")
def q3(built_lib):
    import q3asr
    return q3asr


def test_default_transcribe_has_no_context(q3):
    """testDefaultTranscribeHasNoContext: the system turn is <|im_start|>system\\n<|im_end|>\\n"""
    ids, at = q3.prompt_ids(q3.preset("0.6B"), 3)
    assert ids[:5].tolist() == [IM_START, SYSTEM, NEWLINE, IM_END, NEWLINE]


def test_context_inserts_tokens(q3):
    """testContextInsertsTokens: context ids sit between system\\n and <|im_end|>"""
    ids, at = q3.prompt_ids(q3.preset("0.6B"), 3, context=[100, 200, 300])
    assert ids[:8].tolist() == [IM_START, SYSTEM, NEWLINE, 100, 200, 300, IM_END, NEWLINE]


def test_empty_context_same_as_nil(q3):
    """testEmptyContextSameAsNil"""
    cfg = q3.preset("0.6B")
    a, at_a = q3.prompt_ids(cfg, 7)
    b, at_b = q3.prompt_ids(cfg, 7, context=[])
    assert a.tolist() == b.tolist() and at_a == at_b


@pytest.mark.parametrize("name", ["0.6B", "1.7B"])
def test_full_template_literal(q3, name):
    """Qwen3ASR.swift:196-233 written out; the ids are the ones the reference's tokenizer tests assert (:219-235)."""
    ids, at = q3.prompt_ids(q3.preset(name), 4, context=[11, 12], language=[21, 22, 23])
    assert ids.tolist() == [IM_START, SYSTEM, NEWLINE, 11, 12, IM_END, NEWLINE,
                            IM_START, USER, NEWLINE, AUDIO_START, AUDIO_PAD, AUDIO_PAD, AUDIO_PAD, AUDIO_PAD, AUDIO_END, IM_END, NEWLINE,
                            IM_START, ASSISTANT, NEWLINE, 21, 22, 23, ASR_TEXT]
    assert at == 11 and ids[at - 1] == AUDIO_START and ids[at + 4] == AUDIO_END


def test_aligner_template_has_no_asr_text(q3):
    """ForcedAligner.swift:338-378: empty system turn, audio, assistant turn, then the slotted text verbatim."""
    slotted = [501, 151705, 151705, 502, 151705, 151705]
    ids, at = q3.prompt_ids(q3.preset("aligner"), 2, language=slotted, raw_suffix=True)
    assert ids.tolist() == [IM_START, SYSTEM, NEWLINE, IM_END, NEWLINE, IM_START, USER, NEWLINE, AUDIO_START, AUDIO_PAD, AUDIO_PAD,
                            AUDIO_END, IM_END, NEWLINE, IM_START, ASSISTANT, NEWLINE] + slotted
    assert at == 9


@pytest.mark.parametrize("n_audio", [0, 1, 13, 390, 1500])
@pytest.mark.parametrize("ctx,lang", [(None, None), ([5, 6, 7], None), (None, [9]), ([1] * 40, [2, 3])])
def test_matches_oracle(q3, n_audio, ctx, lang):
    cfg_d = weights.preset("tiny")
    m = omodel.Oracle.__new__(omodel.Oracle)  # only the template is needed, no weights
    m.cfg = cfg_d
    want, want_at = m.prompt_ids(n_audio, ctx, lang)
    got, at = q3.prompt_ids(q3.preset("tiny"), n_audio, context=ctx, language=lang)
    assert got.tolist() == want and at == want_at
    assert int((got == cfg_d["tok_audio_pad"]).sum()) == n_audio


def test_sizing_protocol_and_errors(q3):
    L = q3.lib()
    cfg = q3.preset("0.6B")
    n, at = ctypes.c_int(0), ctypes.c_int(0)
    # cap 0 sizes the buffer
    assert L.q3asr_prompt_ids(ctypes.byref(cfg), 10, None, None, 0, ctypes.byref(n), ctypes.byref(at)) == 1
    assert n.value == 9 + 10 + 6 + 1 and at.value == 9
    buf = np.full(n.value + 1, -1, dtype=np.int32)
    # one id short: refused, nothing written
    assert L.q3asr_prompt_ids(ctypes.byref(cfg), 10, None, buf.ctypes.data, n.value - 1, ctypes.byref(n), None) == 1
    assert (buf == -1).all()
    assert L.q3asr_prompt_ids(ctypes.byref(cfg), 10, None, buf.ctypes.data, n.value, ctypes.byref(n), None) == 0
    assert buf[-1] == -1 and buf[n.value - 1] == ASR_TEXT
    # bad arguments
    assert L.q3asr_prompt_ids(None, 10, None, buf.ctypes.data, 64, ctypes.byref(n), None) == 1
    assert L.q3asr_prompt_ids(ctypes.byref(cfg), -1, None, buf.ctypes.data, 64, ctypes.byref(n), None) == 1
    assert L.q3asr_prompt_ids(ctypes.byref(cfg), 2 ** 31 - 1, None, buf.ctypes.data, 64, ctypes.byref(n), None) == 1   # no 8 GB vector
    assert L.q3asr_prompt_ids(ctypes.byref(cfg), 10, None, buf.ctypes.data, 64, None, None) == 1
    bad = q3.Prompt()
    bad.n_context = 3  # count without a pointer
    assert L.q3asr_prompt_ids(ctypes.byref(cfg), 10, ctypes.byref(bad), buf.ctypes.data, 64, ctypes.byref(n), None) == 1
