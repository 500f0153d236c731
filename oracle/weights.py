"""NumPy twin of the library's deterministic random initialisation (csrc/weights.cu, csrc/ops.cu).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Weights are not available offline, so parity runs on
random-init weights: value = bf16(scale * irwin_hall4) from a splitmix64 stream keyed by
(seed, tensor name); norm scales are constant.  scale = 0.02 for the audio tower and the aligner's head; for
the text decoder 0.08 (o_proj 0.04), the tied embedding 0.15 with "loud" rows (row r is multiplied by
2^min(ctz(splitmix(rowseed + r)) // 4, 4)), q_norm / k_norm scales 2 (other norms 1).  Why: with 0.02 everywhere
greedy ids collapse to one repeated id (the hidden state is an average over nearly identical audio rows that
ignores the last token) and id parity is vacuous; with these values the token path carries weight, the
attention is peaked enough to pick out individual earlier tokens (history dependence, no short cycles) and the
heavy-tailed rows keep top-1/top-2 margins of the bf16 logits above rounding noise (measured and asserted by
tests/golden/make_golden.py).
The generator is integer-only up to one fp32 multiply and an exact power-of-two scale, so this
twin is bit-identical to the CUDA kernel (csrc/ops.cu random_init_kernel); tests/test_gpu_weights.py
checks that against q3asr_get_tensor.

Tensor names/shapes follow the reference's safetensors keys
(/root/reference/Sources/Qwen3ASR/WeightLoading.swift:17-126, 235-323).
"""
import numpy as np

M64 = (1 << 64) - 1


def fnv1a(name):
    h = 0xcbf29ce484222325
    for c in name.encode():
        h ^= c
        h = (h * 0x100000001b3) & M64
    return h


def splitmix_scalar(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def _splitmix_vec(x):
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def bf16_round(x):
    """float32 -> nearest-even bf16, returned as float32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)) << np.uint64(16)
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(x))


EMBED = "model.embed_tokens.weight"


def scale_for(name):
    if name == EMBED:
        return 0.15
    if name.startswith("model."):
        return 0.04 if "self_attn.o_proj." in name else 0.08
    return 0.02


def norm_fill(name):
    return 2.0 if ("self_attn.q_norm." in name or "self_attn.k_norm." in name) else 1.0


def loud_levels(seed, name, rows):
    """min(ctz(splitmix(rowseed + r)) // 4, 4) per row."""
    s = splitmix_scalar(seed ^ fnv1a(name + "#loud"))
    with np.errstate(over="ignore"):
        z = _splitmix_vec(np.uint64(s) + np.arange(rows, dtype=np.uint64))
        low = z & (~z + np.uint64(1))  # lowest set bit
        tz = np.where(z == 0, 64, np.bitwise_count(low - np.uint64(1))).astype(np.int64)
    return np.minimum(tz // 4, 4)


def random_tensor(seed, name, shape, scale=None):
    if scale is None:
        scale = scale_for(name)
    n = int(np.prod(shape))
    s = splitmix_scalar(seed ^ fnv1a(name))
    with np.errstate(over="ignore"):
        z = _splitmix_vec(np.uint64(s) + np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    m = np.uint64(0xFFFF)
    tot = ((z & m) + ((z >> np.uint64(16)) & m) + ((z >> np.uint64(32)) & m) + (z >> np.uint64(48))).astype(np.int64) - 131070
    mult = np.float32(float(np.float32(scale)) / 37837.22668596909)  # the C side divides (double)(float)scale
    out = bf16_round(tot.astype(np.float32) * mult).reshape(shape)
    if name == EMBED:
        out = out * np.exp2(loud_levels(seed, name, shape[0])).astype(np.float32)[:, None]
    return out


def is_norm_weight(n):
    return (n.endswith("layer_norm.weight") or n.endswith("ln_post.weight") or n.endswith("layernorm.weight")
            or n.endswith("_norm.weight") or n == "model.norm.weight")


def tensor_specs(cfg):
    """cfg: dict with the q3asr_config fields."""
    C, d, f = cfg["enc_conv_ch"], cfg["enc_d_model"], cfg["enc_ffn"]
    out = []
    a = "audio_tower."
    out += [(a + "conv2d1.weight", (C, 3, 3, 1)), (a + "conv2d1.bias", (C,)),
            (a + "conv2d2.weight", (C, 3, 3, C)), (a + "conv2d2.bias", (C,)),
            (a + "conv2d3.weight", (C, 3, 3, C)), (a + "conv2d3.bias", (C,)),
            (a + "conv_out.weight", (d, C * 16))]
    for l in range(cfg["enc_layers"]):
        p = f"{a}layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            out += [(p + f"self_attn.{nm}.weight", (d, d)), (p + f"self_attn.{nm}.bias", (d,))]
        out += [(p + "self_attn_layer_norm.weight", (d,)), (p + "self_attn_layer_norm.bias", (d,)),
                (p + "fc1.weight", (f, d)), (p + "fc1.bias", (f,)), (p + "fc2.weight", (d, f)), (p + "fc2.bias", (d,)),
                (p + "final_layer_norm.weight", (d,)), (p + "final_layer_norm.bias", (d,))]
    o = cfg["enc_out_dim"]
    out += [(a + "ln_post.weight", (d,)), (a + "ln_post.bias", (d,)), (a + "proj1.weight", (d, d)), (a + "proj1.bias", (d,)),
            (a + "proj2.weight", (o, d)), (a + "proj2.bias", (o,))]
    h, hd, I = cfg["dec_hidden"], cfg["dec_head_dim"], cfg["dec_inter"]
    out.append(("model.embed_tokens.weight", (cfg["dec_vocab"], h)))
    for l in range(cfg["dec_layers"]):
        p = f"model.layers.{l}."
        out += [(p + "self_attn.q_proj.weight", (cfg["dec_heads"] * hd, h)), (p + "self_attn.k_proj.weight", (cfg["dec_kv_heads"] * hd, h)),
                (p + "self_attn.v_proj.weight", (cfg["dec_kv_heads"] * hd, h)), (p + "self_attn.o_proj.weight", (h, cfg["dec_heads"] * hd)),
                (p + "self_attn.q_norm.weight", (hd,)), (p + "self_attn.k_norm.weight", (hd,)),
                (p + "input_layernorm.weight", (h,)), (p + "post_attention_layernorm.weight", (h,)),
                (p + "mlp.gate_proj.weight", (I, h)), (p + "mlp.up_proj.weight", (I, h)), (p + "mlp.down_proj.weight", (h, I))]
    out.append(("model.norm.weight", (h,)))
    if cfg.get("classify_num", 0) > 0:  # the forced aligner's classification head keeps the lm_head.* keys (WeightLoading.swift:177-179)
        out += [("lm_head.weight", (cfg["classify_num"], h)), ("lm_head.bias", (cfg["classify_num"],))]
    return out


def random_state_dict(cfg, seed=20260418, only=None):
    sd = {}
    for name, shape in tensor_specs(cfg):
        if only is not None and not name.startswith(only):
            continue
        if is_norm_weight(name):
            sd[name] = np.full(shape, norm_fill(name), dtype=np.float32)
        else:
            sd[name] = random_tensor(seed, name, shape)
    return sd


# the presets of q3asr_config_preset (csrc/api.cu); the reference's values are in
# Sources/Qwen3ASR/AudioEncoder.swift:28-68 and Sources/Qwen3ASR/Configuration.swift:47-100
def preset(name):
    c = dict(enc_conv_ch=480, enc_n_window=50, enc_n_window_infer=800, enc_ln_eps=1e-5, dec_vocab=151936, dec_layers=28,
             dec_heads=16, dec_kv_heads=8, dec_head_dim=128, dec_rope_theta=1e6, dec_rms_eps=1e-6,
             tok_im_start=151644, tok_im_end=151645, tok_audio_start=151669, tok_audio_end=151670, tok_audio_pad=151676,
             tok_asr_text=151704, tok_newline=198, tok_system=8948, tok_user=872, tok_assistant=77091, tok_eos=151645,
             classify_num=0, tok_timestamp=151705)
    if name == "0.6B":
        c.update(enc_d_model=896, enc_heads=14, enc_ffn=3584, enc_layers=18, enc_out_dim=1024, dec_hidden=1024, dec_inter=3072)
    elif name == "1.7B":
        c.update(enc_d_model=1024, enc_heads=16, enc_ffn=4096, enc_layers=24, enc_out_dim=2048, dec_hidden=2048, dec_inter=6144)
    elif name == "aligner":  # AudioEncoder.swift:71-88 + TextDecoderConfig.small + 5000 classes (Configuration.swift:132)
        c.update(enc_d_model=1024, enc_heads=16, enc_ffn=4096, enc_layers=24, enc_out_dim=1024, dec_hidden=1024, dec_inter=3072,
                 classify_num=5000)
    elif name in ("tiny", "tiny-aligner"):
        c.update(enc_d_model=128, enc_heads=2, enc_ffn=256, enc_layers=2, enc_out_dim=128, enc_conv_ch=32, dec_vocab=2048,
                 dec_hidden=128, dec_layers=2, dec_heads=4, dec_kv_heads=2, dec_inter=256, tok_im_start=2001, tok_im_end=2002,
                 tok_audio_start=2003, tok_audio_end=2004, tok_audio_pad=2005, tok_asr_text=2006, tok_newline=198,
                 tok_system=1948, tok_user=872, tok_assistant=1091, tok_eos=2002, tok_timestamp=2007,
                 classify_num=70 if name == "tiny-aligner" else 0)
    else:
        raise ValueError(name)
    return c
