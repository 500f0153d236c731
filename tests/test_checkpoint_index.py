"""The checkpoint index behind q3asr_load_safetensors, on the host (q3asr_checkpoint_list): which keys it keeps
(Sources/Qwen3ASR/WeightLoading.swift:17-126: audio_tower.* / model.*; the aligner's "thinker." prefix and lm_head.*, :162-179),
and that a header cannot make the loader read outside its file — the hardening the reference applies to the other file format on
this path (SecurityHardeningTests.swift, WAV) applied to this one.  The dequantisation itself is a GPU test (tests/test_gpu_weights.py)."""
import json
import os
import random
import struct

import numpy as np
import pytest


def _write(path, header, payload=b"", pad=True):
    h = header if isinstance(header, bytes) else json.dumps(header).encode()
    if pad:
        h += b" " * ((8 - len(h) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)) + h + payload)


def _entry(dtype, shape, a, b):
    return {"dtype": dtype, "shape": shape, "data_offsets": [a, b]}


def test_lists_the_keys_the_loader_keeps(built_lib, tmp_path):
    hdr = {
        "__metadata__": {"format": "mlx", "nested": {"a": "}\"{"}},
        "model.embed_tokens.weight": _entry("U32", [8, 16], 0, 512),
        "model.embed_tokens.scales": _entry("BF16", [8, 2], 512, 544),
        "model.embed_tokens.biases": _entry("BF16", [8, 2], 544, 576),
        "audio_tower.ln_post.weight": _entry("F32", [4], 576, 592),
        "thinker.lm_head.weight": _entry("F16", [2, 3], 592, 604),
        "optimizer.state": _entry("F32", [1], 604, 608),        # ignored: not a model key
    }
    _write(tmp_path / "model.safetensors", hdr, bytes(608))
    got = built_lib.checkpoint_list(tmp_path)
    assert got == [("audio_tower.ln_post.weight", "F32", (4,), 16), ("lm_head.weight", "F16", (2, 3), 12),
                   ("model.embed_tokens.biases", "BF16", (8, 2), 32), ("model.embed_tokens.scales", "BF16", (8, 2), 32),
                   ("model.embed_tokens.weight", "U32", (8, 16), 512)]


def test_shards_are_merged(built_lib, tmp_path):
    _write(tmp_path / "model-00001-of-00002.safetensors", {"model.norm.weight": _entry("BF16", [4], 0, 8)}, bytes(8))
    _write(tmp_path / "model-00002-of-00002.safetensors", {"model.layers.0.mlp.up_proj.weight": _entry("BF16", [2, 4], 0, 16)}, bytes(16))
    (tmp_path / "config.json").write_text("{}")
    assert [t[0] for t in built_lib.checkpoint_list(tmp_path)] == ["model.layers.0.mlp.up_proj.weight", "model.norm.weight"]


@pytest.mark.parametrize("case,needle", [
    ("missing_dir", "cannot open directory"),
    ("no_files", "no .safetensors files"),
    ("tiny_file", "too small"),
    ("header_longer_than_file", "bad header length"),
    ("header_len_zero", "bad header length"),
    ("data_past_eof", "lies outside"),
    ("end_before_begin", "lies outside"),
    ("huge_offsets", "number too long"),
    ("negative_dim", "expected a number"),
    ("huge_dim", "bad dimension"),
    ("product_overflow", "too large"),
    ("rank_5", "bad rank"),
    ("rank_0", "bad rank"),
    ("not_json", "expected"),
    ("truncated_json", "expected"),
])
def test_malformed_checkpoints_are_refused(built_lib, tmp_path, case, needle):
    d = tmp_path / "ck"
    if case != "missing_dir":
        d.mkdir()
    f = d / "model.safetensors"
    ok = _entry("F32", [4], 0, 16)
    if case == "no_files":
        (d / "weights.bin").write_bytes(b"x")
    elif case == "tiny_file":
        f.write_bytes(b"\x01\x02\x03")
    elif case == "header_longer_than_file":
        f.write_bytes(struct.pack("<Q", 1 << 20) + b"{}")
    elif case == "header_len_zero":
        f.write_bytes(struct.pack("<Q", 0) + b"{}")
    elif case == "data_past_eof":
        _write(f, {"model.norm.weight": ok}, bytes(15))
    elif case == "end_before_begin":
        _write(f, {"model.norm.weight": _entry("F32", [4], 16, 0)}, bytes(16))
    elif case == "huge_offsets":
        _write(f, b'{"model.norm.weight":{"dtype":"F32","shape":[4],"data_offsets":[0,99999999999999999999999]}}', bytes(16))
    elif case == "negative_dim":
        _write(f, {"model.norm.weight": _entry("F32", [-4], 0, 16)}, bytes(16))
    elif case == "huge_dim":
        _write(f, {"model.norm.weight": _entry("F32", [1 << 40], 0, 16)}, bytes(16))
    elif case == "product_overflow":   # 2^31 * 2^31 * 4 would wrap a 64-bit byte count to 0 == an empty payload
        _write(f, {"model.norm.weight": _entry("F32", [1 << 31, 1 << 31, 4], 0, 0)}, b"")
    elif case == "rank_5":
        _write(f, {"model.norm.weight": _entry("F32", [1, 1, 1, 1, 4], 0, 16)}, bytes(16))
    elif case == "rank_0":
        _write(f, {"model.norm.weight": _entry("F32", [], 0, 4)}, bytes(4))
    elif case == "not_json":
        _write(f, b"\x00\x01garbage\xff", bytes(16))
    elif case == "truncated_json":
        _write(f, b'{"model.norm.weight":{"dtype":"F32","shape":[4],"data_offs', bytes(16), pad=False)
    with pytest.raises(built_lib.Q3Error) as e:
        built_lib.checkpoint_list(d)
    assert e.value.code == 5 and needle in str(e.value), str(e.value)


def test_random_corruption_never_gets_past_the_checks(built_lib, tmp_path):
    """Mutated headers either fail with a message or list only tensors whose bytes lie inside the file."""
    rnd = random.Random(7)
    base = json.dumps({"model.norm.weight": _entry("BF16", [4], 0, 8), "model.embed_tokens.weight": _entry("U32", [8, 16], 8, 520),
                       "__metadata__": {"format": "mlx"}}).encode()
    f = tmp_path / "model.safetensors"
    listed = refused = 0
    for _ in range(400):
        h = bytearray(base)
        for _ in range(rnd.randrange(1, 5)):
            k = rnd.randrange(4)
            i = rnd.randrange(len(h))
            if k == 0:
                h[i] = rnd.randrange(256)
            elif k == 1:
                del h[i:i + rnd.randrange(1, 6)]
            elif k == 2:
                h[i:i] = rnd.choice([b"9" * 12, b"{", b"}", b"[", b"]", b'"', b"\\", b",", b":", b"-"])
            else:
                h = h[:i]
            if not h:
                h = bytearray(b"{")
        payload = bytes(rnd.choice([0, 8, 519, 520, 4096]))
        _write(f, bytes(h), payload, pad=False)
        try:
            got = built_lib.checkpoint_list(tmp_path)
        except built_lib.Q3Error as e:
            assert e.code == 5 and str(e)
            refused += 1
            continue
        listed += 1
        for name, dtype, shape, nbytes in got:
            assert 0 <= nbytes <= len(payload) and 1 <= len(shape) <= 4 and all(0 <= v <= 1 << 31 for v in shape)
    assert listed > 0 and refused > 0


def test_sizing_protocol(built_lib, tmp_path):
    import ctypes
    L = built_lib.lib()
    _write(tmp_path / "m.safetensors", {"model.norm.weight": _entry("BF16", [4], 0, 8)}, bytes(8))
    need = ctypes.c_size_t()
    d = os.fsencode(tmp_path)
    assert L.q3asr_checkpoint_list(d, None, 0, ctypes.byref(need)) == 0 and need.value == len("model.norm.weight\tBF16\t4\t8\n") + 1
    small = ctypes.create_string_buffer(4)
    assert L.q3asr_checkpoint_list(d, small, 4, ctypes.byref(need)) == 4          # Q3ASR_ERR_NOMEM, nothing written past the buffer
    assert L.q3asr_checkpoint_list(d, None, 0, None) == 1
    assert L.q3asr_checkpoint_list(None, small, 4, ctypes.byref(need)) == 5 and small.value == b"loa"   # message, truncated to fit


def test_preset_detection_from_the_index(built_lib, tmp_path):
    """detect_preset_from_checkpoint: the decoder width / the classification head decide, not the directory name."""
    for sub, hdr, want in [
        ("a", {"model.norm.weight": _entry("BF16", [1024], 0, 2048)}, "0.6B"),
        ("Qwen3-ASR-0.6B-named-but-large", {"model.norm.weight": _entry("BF16", [2048], 0, 4096)}, "1.7B"),
        ("c", {"thinker.model.norm.weight": _entry("BF16", [1024], 0, 2048), "thinker.lm_head.weight": _entry("BF16", [2, 4], 2048, 2064)}, "aligner"),
        ("d", {"model.norm.weight": _entry("BF16", [128], 0, 256)}, None),
        ("e", {"audio_tower.ln_post.weight": _entry("F32", [4], 0, 16)}, None),
    ]:
        d = tmp_path / sub
        d.mkdir()
        _write(d / "model.safetensors", hdr, bytes(max(e["data_offsets"][1] for e in hdr.values())))
        assert built_lib.detect_preset_from_checkpoint(d) == want, sub
