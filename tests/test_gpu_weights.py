"""Checkpoint loading through q3asr_load_safetensors: fp32 / fp16 / bf16 tensors under the reference's key names and the
U32-packed tensors of the MLX 4-/8-bit repos (weight + scales + biases, group size 64), dequantised to bf16 at load
(/root/reference/Sources/MLXCommon/PreQuantizedEmbedding.swift:12-49, Sources/Qwen3ASR/WeightLoading.swift:17-126)."""
import json
import os
import struct

import numpy as np
import pytest

from oracle import synth
from oracle.weights import bf16_round

pytestmark = pytest.mark.gpu


def _bf16_bits(x):
    return (bf16_round(x.astype(np.float32)).view(np.uint32) >> 16).astype(np.uint16)


def write_safetensors(path, tensors):
    """tensors: name -> (dtype string, numpy array already in its storage dtype)"""
    header, blobs, off = {}, [], 0
    for name, (dt, arr) in tensors.items():
        raw = np.ascontiguousarray(arr).tobytes()
        header[name] = {"dtype": dt, "shape": list(arr.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    header["__metadata__"] = {"format": "pt"}
    h = json.dumps(header).encode()
    h += b" " * ((8 - len(h) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for b in blobs:
            f.write(b)


def mlx_quantize(w, bits):
    """Affine group quantisation (group 64): returns (uint32 packed, bf16 scales bits, bf16 biases bits, dequantised fp32)."""
    rows, cols = w.shape
    g = w.reshape(rows, cols // 64, 64).astype(np.float64)
    lo, hi = g.min(axis=2), g.max(axis=2)
    qmax = (1 << bits) - 1
    scale = bf16_round(np.maximum((hi - lo) / qmax, 1e-8).astype(np.float32))
    bias = bf16_round(lo.astype(np.float32))
    q = np.clip(np.rint((g - bias[..., None]) / scale[..., None]), 0, qmax).astype(np.uint32).reshape(rows, cols)
    per = 32 // bits
    packed = np.zeros((rows, cols // per), dtype=np.uint32)
    for j in range(per):
        packed |= q[:, j::per] << np.uint32(j * bits)
    deq = (scale[..., None].astype(np.float64) * q.reshape(rows, cols // 64, 64) + bias[..., None].astype(np.float64)).reshape(rows, cols)
    return packed, _bf16_bits(scale), _bf16_bits(bias), bf16_round(deq.astype(np.float32))


@pytest.mark.parametrize("bits", [4, 8])
def test_load_quantised_and_float_checkpoint(built_lib, tmp_path, bits):
    src = built_lib.Qwen3ASRModel.random_init("tiny", seed=99)
    try:
        names = src.tensor_names()
        sd = {n: src.get_tensor(n, shp) for n, shp in names}
    finally:
        src.close()
    rng = np.random.default_rng(bits)
    enc, dec, expect = {}, {}, {}
    for i, (n, shp) in enumerate(names):
        w = sd[n]
        quantisable = n.startswith("model.") and w.ndim == 2 and w.shape[1] % 64 == 0 and n.endswith(".weight")
        if quantisable:
            packed, sc, bs, deq = mlx_quantize(w + 0.01 * rng.standard_normal(w.shape).astype(np.float32), bits)
            stem = n[:-len(".weight")]
            dec[n] = ("U32", packed)
            dec[stem + ".scales"] = ("BF16", sc)
            dec[stem + ".biases"] = ("BF16", bs)
            expect[n] = deq
        elif i % 3 == 0:
            (enc if n.startswith("audio_tower.") else dec)[n] = ("F32", w.astype(np.float32))
            expect[n] = bf16_round(w)
        elif i % 3 == 1:
            (enc if n.startswith("audio_tower.") else dec)[n] = ("BF16", _bf16_bits(w))
            expect[n] = bf16_round(w)
        else:
            (enc if n.startswith("audio_tower.") else dec)[n] = ("F16", w.astype(np.float16))
            expect[n] = bf16_round(w.astype(np.float16).astype(np.float32))
    dec["lm_head.unrelated"] = ("F32", np.zeros((2, 2), np.float32))  # keys outside audio_tower.* / model.* are ignored
    write_safetensors(tmp_path / "model-00001-of-00002.safetensors", enc)
    write_safetensors(tmp_path / "model-00002-of-00002.safetensors", dec)
    m = built_lib.Qwen3ASRModel.from_pretrained(str(tmp_path), size="tiny")
    ref = built_lib.Qwen3ASRModel.random_init("tiny", seed=1)
    try:
        assert m.is_loaded and m.tokenizer is None
        for n, shp in names:
            got = m.get_tensor(n, shp)
            assert np.array_equal(got, expect[n].reshape(shp)), n
            ref.set_tensor(n, expect[n].reshape(shp))
        ref.commit_weights()
        x = synth.clip(4, 40000)
        a = m.transcribe_ids([x], max_tokens=12, stop_on_eos=False)[0]
        b = ref.transcribe_ids([x], max_tokens=12, stop_on_eos=False)[0]
        assert a.tolist() == b.tolist()
    finally:
        m.close()
        ref.close()


def test_load_errors(built_lib, tmp_path):
    m = built_lib.Qwen3ASRModel("tiny")
    try:
        with pytest.raises(built_lib.Q3Error) as e:
            m._ck(built_lib.lib().q3asr_load_safetensors(m._h, str(tmp_path).encode()))
        assert "no .safetensors" in str(e.value)
        write_safetensors(tmp_path / "a.safetensors", {"model.norm.weight": ("F32", np.ones(128, np.float32))})
        with pytest.raises(built_lib.Q3Error) as e:
            m._ck(built_lib.lib().q3asr_load_safetensors(m._h, str(tmp_path).encode()))
        assert "found 1 of" in str(e.value)
        assert not m.is_loaded
    finally:
        m.close()


def test_load_forced_aligner_checkpoint(built_lib, tmp_path):
    """The aligner checkpoints prefix every key with "thinker." and keep the classification head under lm_head.*
    (WeightLoading.swift:162-179, 229)."""
    src = built_lib.Qwen3ASRModel.random_init("tiny-aligner", seed=5)
    try:
        names = src.tensor_names()
        assert ("lm_head.weight", (70, 128)) in names and ("lm_head.bias", (70,)) in names
        sd = {n: src.get_tensor(n, shp) for n, shp in names}
        x = synth.clip(2, 30000)
        ids, pos = [2007, 11, 12, 2007, 2007, 13, 2007], [0, 3, 4, 6]
        want = src.align_indices([x], [ids], [pos])[0]
    finally:
        src.close()
    write_safetensors(tmp_path / "model.safetensors", {"thinker." + n: ("BF16", _bf16_bits(w)) for n, w in sd.items()})
    m = built_lib.Qwen3ASRModel.from_pretrained(str(tmp_path), size="tiny-aligner")
    try:
        assert m.is_loaded
        assert m.align_indices([x], [ids], [pos])[0].tolist() == want.tolist()
    finally:
        m.close()
