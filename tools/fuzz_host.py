"""Mutation fuzzer for the host-side parsers of libq3asr.so: tokenizer files (vocab.json / merges.txt / tokenizer_config.json), the
WAV parser and the word splitter.  No GPU needed.  Meant to run against a sanitizer build (tools/sanitize.sh), where any
out-of-bounds access or undefined behaviour aborts the process; against the normal build it only catches crashes.
Run: python tools/fuzz_host.py SEED SCRATCH_DIR ITERATIONS   (Q3LIB=/path/to/libq3asr.so picks the library)"""
import os, sys, json, struct, random, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qwen3-asr-swift_b200"))
import numpy as np
import q3asr
q3asr.LIB_PATH = os.environ.get("Q3LIB", q3asr.LIB_PATH)
L = q3asr.lib()
rnd = random.Random(int(sys.argv[1]))
d = sys.argv[2]
os.makedirs(d, exist_ok=True)
vocab = {"a": 0, "b": 1, "ab": 2, "Ġc": 3, "ä½": 4, "\\u": 5, "\"q": 6}
cfg = {"added_tokens_decoder": {"10": {"content": "<|im_start|>", "special": True}, "11": {"content": "<asr_text>"}}}
base = {"vocab.json": json.dumps(vocab).encode(), "merges.txt": b"#version: 0.2\na b\n\xc4\xa0 c\n", "tokenizer_config.json": json.dumps(cfg).encode()}
def mutate(b):
    b = bytearray(b)
    k = rnd.randrange(6)
    if k == 0 and b: b = b[:rnd.randrange(len(b))]
    elif k == 1:
        for _ in range(rnd.randrange(1, 6)):
            if b: b[rnd.randrange(len(b))] = rnd.randrange(256)
    elif k == 2:
        for _ in range(rnd.randrange(1, 4)):
            i = rnd.randrange(len(b) + 1); b[i:i] = bytes(rnd.choice([b'{', b'}', b'"', b'\\', b'\\u', b'\\ud8', b':', b',', b'[', b'9'*30, b'-', b'\x00']))
    elif k == 3: b = b * rnd.randrange(2, 4)
    elif k == 4: b = bytearray(rnd.randbytes(rnd.randrange(0, 64)))
    elif k == 5 and b:
        i = rnd.randrange(len(b)); del b[i:i + rnd.randrange(1, 8)]
    return bytes(b)
n_ok = 0
for it in range(int(sys.argv[3])):
    for name, data in base.items():
        with open(os.path.join(d, name), "wb") as f:
            f.write(mutate(data) if rnd.random() < 0.6 else data)
    t = ctypes.c_void_p()
    rc = L.q3asr_tokenizer_load(d.encode(), ctypes.byref(t))
    if rc == 0:
        n_ok += 1
        ids = np.array([rnd.randrange(-5, 20) for _ in range(12)], dtype=np.int32)
        need = ctypes.c_size_t()
        buf = ctypes.create_string_buffer(4096)
        L.q3asr_tokenizer_decode(t, ids.ctypes.data, ids.size, buf, 4096, ctypes.byref(need))
        out = np.zeros(64, dtype=np.int32); n = ctypes.c_int()
        L.q3asr_tokenizer_encode(t, "ab c 你 q".encode(), out.ctypes.data, 64, ctypes.byref(n))
    L.q3asr_tokenizer_destroy(t)
print("tokenizer loads ok:", n_ok)
# wav
pcm = np.arange(100, dtype="<i2").tobytes()
wav = struct.pack("<4sI4s4sIHHIIHH", b"RIFF", 36 + len(pcm), b"WAVE", b"fmt ", 16, 1, 1, 16000, 32000, 2, 16) + b"LIST" + struct.pack("<I", 4) + b"abcd" + b"data" + struct.pack("<I", len(pcm)) + pcm
ok = 0
for it in range(int(sys.argv[3]) * 4):
    b = mutate(wav)
    arr = np.frombuffer(b, dtype=np.uint8) if b else np.zeros(0, np.uint8)
    n = ctypes.c_size_t(); sr = ctypes.c_int()
    out = np.zeros(4096, dtype=np.float32)
    rc = L.q3asr_wav_parse(arr.ctypes.data if arr.size else None, arr.size, out.ctypes.data, rnd.choice([0, 10, 4096]), ctypes.byref(n), ctypes.byref(sr))
    ok += rc == 0
print("wav ok:", ok)

# word splitter: random byte strings (valid and invalid UTF-8)
pairs = 0
for it in range(int(sys.argv[3]) * 10):
    k = rnd.randrange(0, 24)
    b = bytes(rnd.choice([rnd.randrange(1, 256), 0x20, 0xe4, 0xbd, 0xa0, 0xf0, 0x9f, 0x98, 0x80, 0x41, 0x2c]) for _ in range(k))
    try:
        pairs += len(q3asr.text_word_pairs(b, rnd.choice(["English", "Chinese", None, "xx"])))
    except q3asr.Q3Error:
        pass
print("word pairs:", pairs)
