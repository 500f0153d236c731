"""The C-ABI library loads and exports every symbol include/q3asr.h declares; host-side logic that needs no GPU."""
import os
import re

import numpy as np
import pytest

from conftest import HAS_GPU, ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "q3asr.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(q3asr_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built_lib):
    L = built_lib.lib()
    names = _declared()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(built_lib.EXPORTS) == names  # the binding covers the whole header
    assert "sm_100a" in built_lib.version()


def test_presets_match_reference_configs(built_lib):
    # Tests/Qwen3ASRTests/Qwen3ASRTests.swift:11-118 pins these integers
    s = built_lib.preset("0.6B")
    assert (s.enc_d_model, s.enc_layers, s.enc_heads, s.enc_ffn, s.enc_out_dim) == (896, 18, 14, 3584, 1024)
    assert (s.dec_hidden, s.dec_layers, s.dec_heads, s.dec_kv_heads, s.dec_inter, s.dec_head_dim) == (1024, 28, 16, 8, 3072, 128)
    l = built_lib.preset("1.7B")
    assert (l.enc_d_model, l.enc_layers, l.enc_heads, l.enc_ffn, l.enc_out_dim) == (1024, 24, 16, 4096, 2048)
    assert (l.dec_hidden, l.dec_inter, l.dec_vocab) == (2048, 6144, 151936)
    assert (s.enc_conv_ch, s.enc_n_window, s.enc_n_window_infer, s.enc_conv_ch * 16) == (480, 50, 800, 7680)
    assert (s.tok_im_start, s.tok_im_end, s.tok_audio_start, s.tok_audio_end, s.tok_audio_pad, s.tok_asr_text) == (
        151644, 151645, 151669, 151670, 151676, 151704)
    assert (s.tok_system, s.tok_user, s.tok_assistant, s.tok_newline, s.tok_eos) == (8948, 872, 77091, 198, 151645)
    with pytest.raises(built_lib.Q3Error):
        built_lib.preset("7B")
    # the NumPy twin of the presets used by the oracle agrees
    from oracle import weights
    for name in ("0.6B", "1.7B", "tiny", "aligner", "tiny-aligner"):
        c, o = built_lib.preset(name).as_dict(), weights.preset(name)
        for k, v in o.items():
            assert abs(c[k] - v) <= 1e-6 * max(1.0, abs(v)), (name, k)


def test_frame_and_token_counts(built_lib):
    from oracle import model as omodel
    from oracle import mel as omel
    for n in (0, 159, 160, 1600, 47999, 480000, 16000 * 1300):
        assert built_lib.mel_frames(n) == omel.mel_frames(n)
    # AudioEncoder.swift:287-303
    for t, want in ((3000, 390), (1500, 195), (100, 13), (101, 14), (60, 8), (1, 1), (1730, 221 + 0)):
        assert built_lib.encoder_tokens(t) == omodel.output_length(t)
    assert built_lib.encoder_tokens(3000) == 390 and built_lib.encoder_tokens(1500) == 195


def test_scheduler_assignment(built_lib):
    n = np.array([480000] * 64, dtype=np.uint64)
    for g in (1, 2, 4, 8):
        a = built_lib.schedule(n, g)
        assert np.bincount(a, minlength=g).tolist() == [64 // g] * g
    # longest-first: a mixed load is balanced within one short clip
    rng = np.random.default_rng(0)
    n = rng.integers(16000, 480000, size=200).astype(np.uint64)
    a = built_lib.schedule(n, 8)
    load = np.array([n[a == g].sum() for g in range(8)], dtype=np.float64)
    assert load.max() - load.min() <= 480000
    assert built_lib.schedule(np.zeros(0, np.uint64), 4).size == 0


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(built_lib):
    # no CPU fallback: the product path must refuse to run without a Blackwell GPU
    with pytest.raises(built_lib.Q3Error) as e:
        built_lib.Qwen3ASRModel("tiny")
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)
    with pytest.raises(built_lib.Q3Error) as e:
        built_lib.Pool("tiny", devices=(0,))
    assert "device 0" in str(e.value) and "no CPU fallback" in str(e.value)  # the reason survives although no pool came back


def test_pool_and_job_argument_checks(built_lib):
    """Bad arguments come back as Q3ASR_ERR_INVALID (1) before any GPU work, and no call leaves a dangling out-pointer."""
    import ctypes
    L = built_lib.lib()
    cfg = built_lib.preset("tiny")
    out = ctypes.c_void_p(0xdead)
    assert L.q3asr_pool_create(ctypes.byref(cfg), None, 0, 1, None, ctypes.byref(out)) == 1 and not out.value
    assert b"no devices" in L.q3asr_pool_last_error(None)
    assert L.q3asr_pool_transcribe_ids_opts(None, None, None, None, 1, None, None, 8, 1, 0, None, None) == 1
    job = ctypes.c_void_p(0xdead)
    assert L.q3asr_pool_submit(None, None, None, None, 1, None, None, 8, 1, 0, ctypes.byref(job)) == 1 and not job.value
    assert L.q3asr_job_done(None) == 0 and L.q3asr_job_wait(None, None, None) == 1 and L.q3asr_job_last_error(None) == b""
    L.q3asr_job_free(None)
    L.q3asr_pool_destroy(None)


def test_model_size_and_bits_detection(built_lib):
    """testASRModelSizeDetection / testASRModelSizeBitsDetection (Qwen3ASRTests.swift:61-69, 93-103)"""
    import q3asr
    assert q3asr.detect_model_size("aufklarer/Qwen3-ASR-0.6B-MLX-4bit") == "0.6B"
    assert q3asr.detect_model_size("aufklarer/Qwen3-ASR-1.7B-MLX-8bit") == "1.7B"
    assert q3asr.detect_model_size("some-custom/model") == "0.6B"
    assert q3asr.detect_bits("aufklarer/Qwen3-ASR-0.6B-MLX-8bit") == 8
    assert q3asr.detect_bits("aufklarer/Qwen3-ASR-0.6B-MLX-4bit") == 4
    assert q3asr.detect_bits("aufklarer/Qwen3-ASR-1.7B-MLX-4bit") == 4
    assert q3asr.detect_bits("some-custom/small-model") == 4
    assert q3asr.detect_bits("some/1.7B-model") == 8


def test_missing_library_fails_loudly(built_lib, tmp_path):
    """No CPU fallback anywhere above the C ABI either: without libq3asr.so the mirror raises on first use, naming the build command."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import q3asr; q3asr.LIB_PATH = %r\n"
            "try:\n    q3asr.mel_frames(16000)\nexcept q3asr.Q3Error as e:\n    print('RAISED', e)\n"
            % (os.path.dirname(os.path.dirname(built_lib.__file__)), str(tmp_path / "libq3asr.so")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "RAISED" in r.stdout and "no CPU fallback" in r.stdout and "g.build()" in r.stdout, r.stdout + r.stderr


def test_header_is_plain_c99_and_links_from_c(built_lib, tmp_path):
    """The boundary is a C ABI: include/q3asr.h must compile as strict C99 (cgo / Swift's Clang importer / ctypes read it as C, not C++)
    and a C program must link against the library and call a host-only entry point."""
    import subprocess
    src = tmp_path / "use.c"
    src.write_text('#include "q3asr.h"\n#include <stdio.h>\n#include <string.h>\n'
                   "int main(void) {\n"
                   "    q3asr_config c; int ids[64]; int n = 0, at = 0;\n"
                   '    if (q3asr_config_preset("0.6B", &c) != Q3ASR_OK) return 1;\n'
                   "    if (q3asr_prompt_ids(&c, 3, NULL, ids, 64, &n, &at) != Q3ASR_OK || n != 19 || at != 9) return 2;\n"
                   "    if (q3asr_encoder_tokens(3000) != 390 || q3asr_mel_frames(480000) != 3000) return 3;\n"
                   '    printf("%s\\n", q3asr_version());\n'
                   "    return 0;\n}\n")
    exe = tmp_path / "use"
    lib_dir = os.path.join(ROOT, "qwen3-asr-swift_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-L", lib_dir,
                    "-lq3asr", "-Wl,-rpath," + lib_dir, "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "sm_100a" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_host_entry_points_are_reentrant(built_lib, tmp_path):
    """The host-only entry points keep no shared mutable state (their error strings are per thread): eight threads calling them at once
    (ctypes drops the GIL during a call) get what a serial run gets, and a failure on one thread does not leak its message into another."""
    import threading
    q = built_lib
    wav = tmp_path / "t.wav"
    q.AudioFileLoader.write_wav(wav, np.sin(np.arange(4000) * 0.05), 16000)
    data = wav.read_bytes()
    tok = q.Qwen3Tokenizer(id_to_token={i + 10: c for i, c in enumerate("abcdefghijklmnopqrstuvwxyz")})
    text = "the quick brown fox, 你好 jumps over the lazy dog! " * 20

    def work(k):
        out = []
        for it in range(40):
            out.append(q.text_word_pairs(text, "English"))
            out.append(q.AudioFileLoader.parse_wav(data)[0].tobytes())
            out.append(tok.encode("quick brown fox"))
            out.append(q.enforce_monotonicity([1, 3, 2, 7, 9, 11, 4, 12] * 5))
            out.append(q.prompt_ids(q.preset("0.6B"), 50 + k % 2)[0].tobytes())
            if k % 2:                                   # odd threads also fail on purpose, with a thread-specific message
                try:
                    q.text_word_pairs("x", "Japanese")
                except q.Q3Error as e:
                    out.append("NLTokenizer" in str(e))
                try:
                    q.AudioFileLoader.parse_wav(b"RIFF" + bytes(60))
                except q.AudioLoadError as e:
                    out.append("Invalid WAV" in str(e))
        return out

    want = {k: work(k) for k in (0, 1)}
    got = {}
    threads = [threading.Thread(target=lambda k=k: got.__setitem__(k, work(k))) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    tok.close()
    assert all(got[k] == want[k % 2] for k in range(8))
    assert all(v is True for v in want[1] if isinstance(v, bool))


def test_integration_appendix_lists_every_entry_point():
    """INTEGRATION.md's appendix and include/q3asr.h name the same symbols."""
    header = open(os.path.join(ROOT, "include", "q3asr.h")).read()
    declared = set(re.findall(r"\b(q3asr_[a-z0-9_]+)\s*\(", header))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    appendix = doc[doc.index("## Appendix: every entry point"):]
    listed = set(re.findall(r"`(q3asr_[a-z0-9_]+)`", appendix))
    assert declared - listed == set(), sorted(declared - listed)
    assert listed - declared == set(), sorted(listed - declared)


def test_every_binding_declares_its_argument_types(built_lib):
    """A ctypes call without argtypes passes Python ints as 32-bit C ints — a truncated pointer on the first 64-bit address.  Only the
    entry points that take no argument may go without."""
    L = built_lib.lib()
    bare = [n for n in built_lib.EXPORTS if getattr(L, n).argtypes is None]
    assert sorted(bare) == ["q3asr_io_last_error", "q3asr_text_last_error", "q3asr_version"]


def test_binding_arity_matches_the_header(built_lib):
    """Number of parameters of every declaration in include/q3asr.h against the length of the ctypes argtypes of its binding."""
    header = open(os.path.join(ROOT, "include", "q3asr.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    L = built_lib.lib()
    checked = 0
    for m in re.finditer(r"\b(q3asr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        arity = 0 if params in ("", "void") else params.count(",") + 1
        fn = getattr(L, name)
        got = 0 if fn.argtypes is None else len(fn.argtypes)
        assert got == arity, (name, got, arity)
        checked += 1
    assert checked == len(built_lib.EXPORTS)


def test_struct_layouts_match_the_header(built_lib, tmp_path):
    """sizeof / offsetof of the three structs of the ABI as a C compiler sees them against the ctypes mirrors."""
    import ctypes
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text('#include "q3asr.h"\n#include <stddef.h>\n#include <stdio.h>\nint main(void) {\n'
                   '    printf("%zu %zu %zu %zu\\n", sizeof(q3asr_config), offsetof(q3asr_config, dec_rope_theta), offsetof(q3asr_config, tok_eos),'
                   " offsetof(q3asr_config, tok_timestamp));\n"
                   '    printf("%zu %zu %zu\\n", sizeof(q3asr_prompt), offsetof(q3asr_prompt, language_ids), offsetof(q3asr_prompt, raw_suffix));\n'
                   '    printf("%zu %zu %zu\\n", sizeof(q3asr_sampling), offsetof(q3asr_sampling, seed), offsetof(q3asr_sampling, force_device_sampler));\n'
                   "    return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    C, P, S = built_lib.Config, built_lib.Prompt, built_lib.Sampling
    want = [ctypes.sizeof(C), C.dec_rope_theta.offset, C.tok_eos.offset, C.tok_timestamp.offset,
            ctypes.sizeof(P), P.language_ids.offset, P.raw_suffix.offset,
            ctypes.sizeof(S), S.seed.offset, S.force_device_sampler.offset]
    assert [int(v) for v in out] == want


def test_swift_shim_calls_only_declared_entry_points():
    """The Swift shim cannot be compiled here (no toolchain); at least every q3asr_* it calls must exist in the header, with as many
    arguments as it passes (top-level commas of the call)."""
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "q3asr.h")).read(), flags=re.S)
    arity = {}
    for m in re.finditer(r"\b(q3asr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        params = m.group(2).strip()
        arity[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    structs = set(re.findall(r"typedef struct (q3asr_\w+)", header))
    swift = open(os.path.join(ROOT, "qwen3-asr-swift_b200", "swift", "Qwen3ASRB200.swift")).read()
    swift = re.sub(r"//.*", "", swift)
    calls = 0
    for m in re.finditer(r"\b(q3asr_[a-z0-9_]+)\s*\(", swift):
        name = m.group(1)
        if name in structs:          # q3asr_config(), q3asr_prompt(...): struct initialisers
            continue
        assert name in arity, name
        depth, i, commas, empty = 1, m.end(), 0, True
        while depth:
            ch = swift[i]
            depth += ch in "([{"
            depth -= ch in ")]}"
            if depth == 1 and ch == ",":
                commas += 1
            if depth >= 1 and not ch.isspace():
                empty = False
            i += 1
        assert (0 if empty else commas + 1) == arity[name], (name, commas + 1, arity[name])
        calls += 1
    assert calls >= 15
