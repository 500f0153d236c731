"""CPU restatement of the reference's audio encoder, Qwen3 text decoder and greedy loop.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for tests/, smoke() and bench.py's CPU
baseline — never imported by the product (qwen3-asr-swift_b200/).

PARITY UNPINNED: the reference (Swift + MLX/Metal) cannot run here and its tests hold no numeric vectors
for this path.  The restatement is cross-checked instead against independent
implementations of the same architecture that do exist in this image (tests/test_oracle_model.py):
transformers' Qwen3OmniMoeAudioEncoder (same conv stack / c*16+f flatten / per-chunk positions / window
rule) and Qwen3Model (same decoder block), sharing weights.

Follows (paths relative to /root/reference/Sources):
  Qwen3ASR/AudioEncoder.swift:362-511   chunk -> pad to the longest chunk -> 3x(conv3x3 s2 p1 + gelu) -> [C,13,c*16+f]
                                        -> conv_out -> + sinusoid(pos 0..12 per chunk, :171-199) -> keep valid tokens
                                        -> windows of maxlen*8 tokens -> pre-LN layers (:130-164, :93-126) -> ln_post -> proj1/gelu/proj2
  MLXCommon/SDPA.swift:18-101           softmax(scale * q k^T + mask) v, kv head j/(nq/nkv) for query head j
  Qwen3ASR/FloatTextDecoder.swift:35-226 RMSNorm -> q,k,v -> per-head RMSNorm(q),(k) -> split-half RoPE(theta 1e6, offset = cache len)
                                        -> cache append -> causal GQA -> o -> +res -> RMSNorm -> down(silu(gate) * up) -> +res; final RMSNorm
  MLXCommon/PreQuantizedEmbedding.swift:45-49  tied LM head: logits = h E^T
  Qwen3ASR/Qwen3ASR.swift:181-256, 317-390     prompt layout, audio splice, prefill, greedy loop (append, then stop on EOS)

Third-party semantics assumed (mlx-swift >= 0.30, not in the tree): Linear y = x W^T + b; Conv2d NHWC with
[O,kH,kW,I] weights and zero padding; gelu = exact erf form; LayerNorm biased variance; RMSNorm
x * rsqrt(mean(x^2) + eps) * w; argMax = lowest index on ties.

`emulate_bf16=True` rounds to bf16 at the points where the B200 kernels store bf16 (every GEMM/conv/norm/
attention output; the residual sum is rounded again), with fp32 accumulation in between — that is the
contract token parity is judged on.  `emulate_bf16=False` is the plain fp32 model (what the reference's
encoder computes on an fp32 mel), used to state the bf16 tolerance of the encoder.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _rbf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def conv_len(x):
    return (x - 1) // 2 + 1


def conv_len3(x):
    return conv_len(conv_len(conv_len(x)))


def output_length(frames, chunk=100):
    """AudioEncoder.swift:287-303"""
    rem = frames % chunk
    full = (frames // chunk) * conv_len3(chunk)
    return full + (max(conv_len3(rem), 1) if rem > 0 else 0)


def sinusoid_positions(n, d):
    """AudioEncoder.swift:171-199, Float arithmetic."""
    half = d // 2
    inc = np.float32(np.log(np.float32(10000.0))) / np.float32(half - 1)
    inv = np.exp(np.arange(half, dtype=np.float32) * (-inc)).astype(np.float32)
    t = np.arange(n, dtype=np.float32)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


class Oracle:
    def __init__(self, cfg, weights, emulate_bf16=True, threads=None, decoder_fp64=False):
        """decoder_fp64: the decoder's products and attention accumulate in float64 (rounded to fp32, then to bf16 at the same
        points) — a second accumulation order, used by tests/golden/make_golden.py to keep only fixtures whose greedy ids do
        not depend on it."""
        self.cfg = dict(cfg)
        self.emu = emulate_bf16
        self.fp64 = decoder_fp64
        if threads:
            torch.set_num_threads(threads)
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in weights.items()}
        self.r = _rbf if emulate_bf16 else (lambda x: x)

    # ------------------------------------------------------------------------------------------
    # encoder
    # ------------------------------------------------------------------------------------------
    def _conv(self, x, name):
        w = self.w[f"audio_tower.{name}.weight"].permute(0, 3, 1, 2)  # [O,kH,kW,I] -> [O,I,kH,kW]
        b = self.w[f"audio_tower.{name}.bias"]
        return self.r(F.gelu(F.conv2d(x, w, b, stride=2, padding=1)))

    def _mm(self, a, w):
        """a @ w.T for the decoder (fp32, or float64 accumulation when decoder_fp64)."""
        if self.fp64:
            return (a.double() @ w.double().t()).float()
        return a @ w.t()

    def _attention(self, q, k, v, heads, kv_heads, scale, causal, fp64=False):
        """q [T, heads*hd], k/v [S, kv_heads*hd]; one segment."""
        if fp64:
            return self._attention64(q, k, v, heads, kv_heads, scale, causal)
        T, S = q.shape[0], k.shape[0]
        hd = q.shape[1] // heads
        qh = q.view(T, heads, hd).transpose(0, 1)
        kh = k.view(S, kv_heads, hd).transpose(0, 1).repeat_interleave(heads // kv_heads, dim=0)
        vh = v.view(S, kv_heads, hd).transpose(0, 1).repeat_interleave(heads // kv_heads, dim=0)
        s = torch.matmul(qh, kh.transpose(1, 2)) * scale
        if causal:
            i = torch.arange(T)[:, None] + (S - T)
            j = torch.arange(S)[None, :]
            s = s.masked_fill(j > i, float("-inf"))
        m = s.max(dim=-1, keepdim=True).values
        p = torch.exp(s - m)
        l = p.sum(dim=-1, keepdim=True)
        o = torch.matmul(self.r(p), vh) / l
        return self.r(o.transpose(0, 1).reshape(T, heads * hd))

    def _attention64(self, q, k, v, heads, kv_heads, scale, causal):
        T, S = q.shape[0], k.shape[0]
        hd = q.shape[1] // heads
        qh = q.double().view(T, heads, hd).transpose(0, 1)
        kh = k.double().view(S, kv_heads, hd).transpose(0, 1).repeat_interleave(heads // kv_heads, dim=0)
        vh = v.double().view(S, kv_heads, hd).transpose(0, 1).repeat_interleave(heads // kv_heads, dim=0)
        s = (torch.matmul(qh, kh.transpose(1, 2)) * scale).float()
        if causal:
            i = torch.arange(T)[:, None] + (S - T)
            j = torch.arange(S)[None, :]
            s = s.masked_fill(j > i, float("-inf"))
        m = s.max(dim=-1, keepdim=True).values
        p = torch.exp(s - m)
        l = p.double().sum(dim=-1, keepdim=True)
        o = (torch.matmul(self.r(p).double(), vh) / l).float()
        return self.r(o.transpose(0, 1).reshape(T, heads * hd))

    def encode(self, mel, return_stages=False):
        """mel: float32 [128, T] -> [tokens, enc_out_dim] float32."""
        c = self.cfg
        r = self.r
        mel = torch.from_numpy(np.ascontiguousarray(mel, dtype=np.float32))
        T = mel.shape[1]
        chunk = 2 * c["enc_n_window"]
        n_chunks = (T + chunk - 1) // chunk
        lens = [chunk] * n_chunks
        if T % chunk:
            lens[-1] = T % chunk
        L = max(lens)
        x = torch.zeros(n_chunks, 1, 128, L)
        pos = 0
        for i, cl in enumerate(lens):
            x[i, 0, :, :cl] = mel[:, pos:pos + cl]
            pos += cl
        stages = {}
        x = self._conv(x, "conv2d1")
        x = self._conv(x, "conv2d2")
        x = self._conv(x, "conv2d3")  # [n, C, 16, t]
        n, C, Fq, Tt = x.shape
        stages["conv3"] = x
        x = x.permute(0, 3, 1, 2).reshape(n, Tt, C * Fq)  # feature index c*16 + f
        d = c["enc_d_model"]
        x = torch.matmul(x, self.w["audio_tower.conv_out.weight"].t())
        x = r(x + torch.from_numpy(sinusoid_positions(Tt, d))[None])
        valid = [conv_len3(cl) for cl in lens]
        h = torch.cat([x[i, :valid[i]] for i in range(n)], dim=0)
        stages["embed"] = h
        total = h.shape[0]
        win = max(valid) * (c["enc_n_window_infer"] // chunk)
        bounds = list(range(0, total, win)) + [total]
        heads = c["enc_heads"]
        scale = 1.0 / math.sqrt(d // heads)
        eps = c["enc_ln_eps"]
        for l in range(c["enc_layers"]):
            p = f"audio_tower.layers.{l}."
            W = lambda s: self.w[p + s]
            xn = r(F.layer_norm(h, (d,), W("self_attn_layer_norm.weight"), W("self_attn_layer_norm.bias"), eps))
            q = r(xn @ W("self_attn.q_proj.weight").t() + W("self_attn.q_proj.bias"))
            k = r(xn @ W("self_attn.k_proj.weight").t() + W("self_attn.k_proj.bias"))
            v = r(xn @ W("self_attn.v_proj.weight").t() + W("self_attn.v_proj.bias"))
            att = torch.cat([self._attention(q[a:b], k[a:b], v[a:b], heads, heads, scale, False)
                             for a, b in zip(bounds[:-1], bounds[1:])], dim=0)
            h = r(h + r(att @ W("self_attn.out_proj.weight").t() + W("self_attn.out_proj.bias")))
            xn = r(F.layer_norm(h, (d,), W("final_layer_norm.weight"), W("final_layer_norm.bias"), eps))
            f = r(F.gelu(xn @ W("fc1.weight").t() + W("fc1.bias")))
            h = r(h + r(f @ W("fc2.weight").t() + W("fc2.bias")))
        a = "audio_tower."
        stages["layers"] = h
        xn = r(F.layer_norm(h, (d,), self.w[a + "ln_post.weight"], self.w[a + "ln_post.bias"], eps))
        p1 = r(F.gelu(xn @ self.w[a + "proj1.weight"].t() + self.w[a + "proj1.bias"]))
        out = r(p1 @ self.w[a + "proj2.weight"].t() + self.w[a + "proj2.bias"])
        if return_stages:
            return out.numpy(), {k: v.numpy() for k, v in stages.items()}
        return out.numpy()

    # ------------------------------------------------------------------------------------------
    # decoder
    # ------------------------------------------------------------------------------------------
    def prompt_ids(self, n_audio, context=None, language=None):
        """Qwen3ASR.swift:196-233"""
        c = self.cfg
        ids = [c["tok_im_start"], c["tok_system"], c["tok_newline"]]
        ids += list(context or [])
        ids += [c["tok_im_end"], c["tok_newline"], c["tok_im_start"], c["tok_user"], c["tok_newline"], c["tok_audio_start"]]
        audio_at = len(ids)
        ids += [c["tok_audio_pad"]] * n_audio
        ids += [c["tok_audio_end"], c["tok_im_end"], c["tok_newline"], c["tok_im_start"], c["tok_assistant"], c["tok_newline"]]
        ids += list(language or [])
        ids.append(c["tok_asr_text"])
        return ids, audio_at

    def _rmsnorm(self, x, w, eps):
        return self.r(x * torch.rsqrt((x * x).mean(dim=-1, keepdim=True) + eps) * w)

    def _rope(self, x, pos, heads):
        """x [T, heads*hd], split-half rotation, positions pos [T] (MLXNN.RoPE traditional:false)."""
        c = self.cfg
        hd = c["dec_head_dim"]
        half = hd // 2
        inv = np.power(float(c["dec_rope_theta"]), -np.arange(half, dtype=np.float64) * 2.0 / hd).astype(np.float32)
        ang = (np.asarray(pos, dtype=np.float32)[:, None] * inv[None, :]).astype(np.float32)  # fp32 product
        cs = torch.from_numpy(np.cos(ang.astype(np.float64)).astype(np.float32))[:, None, :]
        sn = torch.from_numpy(np.sin(ang.astype(np.float64)).astype(np.float32))[:, None, :]
        xh = x.view(x.shape[0], heads, hd)
        a, b = xh[..., :half], xh[..., half:]
        out = torch.cat([a * cs - b * sn, b * cs + a * sn], dim=-1)
        return self.r(out.reshape(x.shape[0], heads * hd))

    def _decoder_forward(self, x, cache, rows=None):
        """x [T, h] new token embeddings; cache: list of (K, V) per layer (or None).  Returns final-norm hidden
        of the LAST position and the new cache."""
        c = self.cfg
        r = self.r
        T = x.shape[0]
        nh, nkv, hd = c["dec_heads"], c["dec_kv_heads"], c["dec_head_dim"]
        eps = c["dec_rms_eps"]
        off = 0 if cache is None else cache[0][0].shape[0]
        pos = np.arange(off, off + T)
        scale = 1.0 / math.sqrt(hd)
        new_cache = []
        for l in range(c["dec_layers"]):
            p = f"model.layers.{l}."
            W = lambda s: self.w[p + s]
            xn = self._rmsnorm(x, W("input_layernorm.weight"), eps)
            q = r(self._mm(xn, W("self_attn.q_proj.weight")))
            k = r(self._mm(xn, W("self_attn.k_proj.weight")))
            v = r(self._mm(xn, W("self_attn.v_proj.weight")))
            q = self._rmsnorm(q.view(T, nh, hd), W("self_attn.q_norm.weight"), eps).reshape(T, nh * hd)
            k = self._rmsnorm(k.view(T, nkv, hd), W("self_attn.k_norm.weight"), eps).reshape(T, nkv * hd)
            q = self._rope(q, pos, nh)
            k = self._rope(k, pos, nkv)
            if cache is not None:
                k = torch.cat([cache[l][0], k], dim=0)
                v = torch.cat([cache[l][1], v], dim=0)
            new_cache.append((k, v))
            att = self._attention(q, k, v, nh, nkv, scale, causal=T > 1, fp64=self.fp64)
            x = r(x + r(self._mm(att, W("self_attn.o_proj.weight"))))
            xn = self._rmsnorm(x, W("post_attention_layernorm.weight"), eps)
            g = r(self._mm(xn, W("mlp.gate_proj.weight")))
            u = r(self._mm(xn, W("mlp.up_proj.weight")))
            act = r(r(F.silu(g)) * u)
            x = r(x + r(self._mm(act, W("mlp.down_proj.weight"))))
        if rows is not None:  # final norm of the requested rows (the forced aligner classifies several positions of one pass)
            return self._rmsnorm(x[torch.tensor(list(rows), dtype=torch.long)], self.w["model.norm.weight"], eps), new_cache
        last = self._rmsnorm(x[-1:], self.w["model.norm.weight"], eps)
        return last, new_cache

    def _logits(self, last):
        return self.r(self._mm(last, self.w["model.embed_tokens.weight"]))[0]

    def prefill(self, audio_embeds, context=None, language=None):
        ids, at = self.prompt_ids(audio_embeds.shape[0], context, language)
        E = self.w["model.embed_tokens.weight"]
        x = E[torch.tensor(ids, dtype=torch.long)].clone()
        x[at:at + audio_embeds.shape[0]] = self.r(torch.from_numpy(np.ascontiguousarray(audio_embeds, dtype=np.float32)))
        last, cache = self._decoder_forward(x, None)
        return self._logits(last), cache, len(ids)

    def greedy(self, audio_embeds, max_tokens, stop_on_eos=True, forced=None, context=None, language=None):
        """Qwen3ASR.swift:317-390.  forced: optional token stream fed instead of the argmax (teacher forcing).
        Returns (ids, top1 logit per step, top1-top2 margin per step); self.topk_ids / self.topk_vals hold the best bf16
        logits of every step (tests accept a deviation only towards a candidate within the stated noise bound of the best)."""
        E = self.w["model.embed_tokens.weight"]
        logits, cache, _ = self.prefill(audio_embeds, context, language)
        ids, tops, margins = [], [], []
        self.topk_ids, self.topk_vals = [], []
        steps = max_tokens if forced is None else len(forced) + 1
        for step in range(steps):
            top2 = torch.topk(logits, 2)
            best = int(torch.argmax(logits))  # first maximal index
            mx = float(logits[best])
            cand = torch.nonzero(logits == mx)
            best = int(cand.min())
            ids.append(best)
            tops.append(mx)
            margins.append(float(top2.values[0] - top2.values[1]))
            top4 = torch.topk(logits, 8)
            self.topk_ids.append(top4.indices.numpy().astype(np.int32))
            self.topk_vals.append(top4.values.numpy().astype(np.float32))
            if forced is None and stop_on_eos and best == self.cfg["tok_eos"]:
                break
            if step + 1 == steps:
                break
            nxt = best if forced is None else int(forced[step])
            last, cache = self._decoder_forward(E[nxt:nxt + 1].clone(), cache)
            logits = self._logits(last)
        return np.array(ids, dtype=np.int32), np.array(tops, dtype=np.float32), np.array(margins, dtype=np.float32)

    def align_indices(self, audio_embeds, slotted_ids, positions):
        """Qwen3ForcedAligner.align steps 3-7 (ForcedAligner.swift:257-299): the aligner template (:338-378: empty system turn,
        audio, assistant turn, then the slotted text and no <asr_text>), one causal pass, classification head (Linear with bias,
        WeightLoading.swift:229) at the given positions of the slotted text, first-maximum class.
        Returns (raw indices, per-position margin between the two best logits, best logit)."""
        c = self.cfg
        n_audio = audio_embeds.shape[0]
        ids = [c["tok_im_start"], c["tok_system"], c["tok_newline"], c["tok_im_end"], c["tok_newline"],
               c["tok_im_start"], c["tok_user"], c["tok_newline"], c["tok_audio_start"]]
        at = len(ids)
        ids += [c["tok_audio_pad"]] * n_audio
        ids += [c["tok_audio_end"], c["tok_im_end"], c["tok_newline"], c["tok_im_start"], c["tok_assistant"], c["tok_newline"]]
        start = len(ids)
        ids += [int(t) for t in slotted_ids]
        E = self.w["model.embed_tokens.weight"]
        x = E[torch.tensor(ids, dtype=torch.long)].clone()
        x[at:at + n_audio] = self.r(torch.from_numpy(np.ascontiguousarray(audio_embeds, dtype=np.float32)))
        hsel, _ = self._decoder_forward(x, None, rows=[start + int(p) for p in positions])
        logits = self.r(hsel @ self.w["lm_head.weight"].t() + self.w["lm_head.bias"])
        top2 = torch.topk(logits, 2, dim=-1).values
        raw = []
        for r in range(logits.shape[0]):
            mx = logits[r].max()
            raw.append(int(torch.nonzero(logits[r] == mx).min()))
        return np.array(raw, dtype=np.int32), (top2[:, 0] - top2[:, 1]).numpy(), top2[:, 0].numpy()

    def generate_slow(self, audio_embeds, max_tokens, repetition_penalty=1.0, no_repeat_ngram_size=0, stop_on_eos=True):
        """Qwen3ASR.swift:396-447 (generateSlow) with pickNextToken (oracle/sampler.py) on every step's logits; temperature 0
        (the noise stream is the library's own, see oracle/sampler.py).  Returns (ids, per-step margin between the two best
        adjusted scores, per-step best adjusted score)."""
        from . import sampler
        E = self.w["model.embed_tokens.weight"]
        logits, cache, _ = self.prefill(audio_embeds)
        ids, margins, tops = [], [], []
        for step in range(max_tokens):
            sc = sampler.adjusted_scores(logits.numpy(), ids, repetition_penalty, no_repeat_ngram_size)
            tok = sampler.pick_next_token(logits.numpy(), ids, repetition_penalty, no_repeat_ngram_size)
            top = np.sort(sc[np.isfinite(sc)])[-2:]
            margins.append(float(top[-1] - top[0]) if top.size == 2 else np.inf)
            tops.append(float(top[-1]) if top.size else 0.0)
            ids.append(tok)
            if (stop_on_eos and tok == self.cfg["tok_eos"]) or step + 1 == max_tokens:
                break
            last, cache = self._decoder_forward(E[tok:tok + 1].clone(), cache)
            logits = self._logits(last)
        return np.array(ids, dtype=np.int32), np.array(margins, dtype=np.float32), np.array(tops, dtype=np.float32)

    def transcribe_ids(self, pcm, max_tokens=448, stop_on_eos=True):
        from . import mel as mel_mod
        feats = mel_mod.mel(pcm)
        emb = self.encode(feats)
        return self.greedy(emb, max_tokens, stop_on_eos)[0]
