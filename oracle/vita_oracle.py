"""CPU oracle: a plain fp32 restatement of the reference's omni-modal forward path.

*** TEST INFRASTRUCTURE -- NOT PRODUCT CODE ***
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this module, and
only as the checker / reported CPU baseline.  The product (vita_b200/) never imports it and has no CPU path.

Parity status: PINNED.  Every function here is checked against the reference's own Python classes
(vita.model.* executed on CPU through oracle/ref_shim.py with identical weights) by oracle/make_golden.py; the
resulting input/output vectors are committed under tests/golden/ and re-checked by tests/test_oracle_golden.py.
The reference itself ships no tests or golden vectors.

The Mixtral decoder arithmetic is not in the reference tree: it lives in the third-party `transformers` package
(pinned 4.41.1 in requirements.txt:24; 5.5.0 installed here).  Those functions restate its published algorithm and
cite modeling_mixtral.py of the installed version; they are anchored on the reference's call sites
(vita/model/language_model/vita_mixtral.py:158-173).

All math is fp32 on CPU tensors.  `state` maps the reference's parameter names (vita_b200.weights) to tensors;
weights are up-cast at use so a bf16 state costs 2 bytes/parameter of host RAM.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200
AUDIO_TOKEN_INDEX = -500

PV = "model.vision_tower.vision_tower."
PA = "model.audio_encoder."
PP = "model.mm_projector."


def _f(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32)


def linear(x, w, b=None):
    return F.linear(x, _f(w), None if b is None else _f(b))


# ================================================================================================ InternViT
def vit_embeddings(state, vcfg, pixel_values):
    """InternVisionEmbeddings.forward, internvit/modeling_intern_vit.py:107-122 (448 px: pos-embed resample is identity)."""
    x = F.conv2d(_f(pixel_values), _f(state[PV + "embeddings.patch_embedding.weight"]),
                 _f(state[PV + "embeddings.patch_embedding.bias"]), stride=vcfg.patch_size)      # :109
    b = x.shape[0]
    x = x.flatten(2).transpose(1, 2)                                                               # :111
    cls = _f(state[PV + "embeddings.class_embedding"]).expand(b, 1, -1)                            # :112
    x = torch.cat([cls, x], dim=1)                                                                 # :113
    return x + _f(state[PV + "embeddings.position_embedding"])                                     # :114-121


def vit_layer(state, vcfg, l, h):
    """InternVisionEncoderLayer.forward :245-253 with InternAttention._naive_attn :158-177 and InternMLP :213-217."""
    p = PV + f"encoder.layers.{l}."
    H, nh = vcfg.hidden_size, vcfg.num_attention_heads
    B, N, _ = h.shape
    x = F.layer_norm(h, (H,), _f(state[p + "norm1.weight"]), _f(state[p + "norm1.bias"]), vcfg.layer_norm_eps)
    qkv = linear(x, state[p + "attn.qkv.weight"], state[p + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, nh, H // nh).permute(2, 0, 3, 1, 4)                                # :160-162
    q, k, v = qkv.unbind(0)
    attn = (q * (H // nh) ** -0.5) @ k.transpose(-2, -1)                                           # :170
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, H)                                                # :174
    x = linear(x, state[p + "attn.proj.weight"], state[p + "attn.proj.bias"])
    h = h + x * _f(state[p + "ls1"])                                                               # :245-247
    x = F.layer_norm(h, (H,), _f(state[p + "norm2.weight"]), _f(state[p + "norm2.bias"]), vcfg.layer_norm_eps)
    x = linear(x, state[p + "mlp.fc1.weight"], state[p + "mlp.fc1.bias"])
    x = F.gelu(x)                                                                                  # ACT2FN["gelu"]
    x = linear(x, state[p + "mlp.fc2.weight"], state[p + "mlp.fc2.bias"])
    return h + x * _f(state[p + "ls2"])                                                            # :249-251


def pixel_shuffle(x, scale_factor=0.5):
    """InternViTVisionTower.pixel_shuffle, internvit/internvit_encoder.py:42-53."""
    n, w, h, c = x.size()
    x = x.view(n, w, int(h * scale_factor), int(c / scale_factor))
    x = x.permute(0, 2, 1, 3).contiguous()
    x = x.view(n, int(h * scale_factor), int(w * scale_factor), int(c / (scale_factor * scale_factor)))
    x = x.permute(0, 2, 1, 3).contiguous()
    return x


def vision_tower(state, vcfg, images, return_hidden=False):
    """InternViTVisionTower.forward, internvit_encoder.py:55-79 (select_layer=-1, drop CLS, x0.5, pixel shuffle)."""
    h = vit_embeddings(state, vcfg, images)
    for l in range(vcfg.num_hidden_layers):
        h = vit_layer(state, vcfg, l, h)
    feats = h[:, 1:]                                                                               # :38
    g = int(feats.shape[1] ** 0.5)
    assert feats.shape[1] == g * g                                                                 # :72
    feats = feats.reshape(feats.shape[0], g, g, -1)
    feats = pixel_shuffle(feats * vcfg.scale_pix_shuffle)                                          # :74
    feats = feats.reshape(feats.shape[0], -1, feats.shape[-1])
    return (feats, h) if return_hidden else feats


def mm_projector(state, x):
    """mlp2x_gelu, multimodal_projector/builder.py:160-168."""
    x = linear(x, state[PP + "0.weight"], state[PP + "0.bias"])
    x = F.gelu(x)
    return linear(x, state[PP + "2.weight"], state[PP + "2.bias"])


def encode_images(state, cfg, images):
    """VITAMetaForCausalLM.encode_images, vita/model/vita_arch.py:131-134."""
    return mm_projector(state, vision_tower(state, cfg.vision, images))


# ================================================================================================ Whale audio
def whale_position_table(max_len, d_model):
    """PositionalEncoding.__init__, whale/module/layer/attention.py:24-36."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def whale_attention(state, acfg, q_prefix, x, mask, pos_emb):
    """MultiHeadedAttention.forward (rel-enc branch), whale/module/layer/attention.py:358-419."""
    nh, dk = acfg.num_attention_heads, acfg.head_dim
    B = x.shape[0]
    q = linear(x, state[q_prefix + "linear_q.weight"], state[q_prefix + "linear_q.bias"]).view(B, -1, nh, dk)
    k = linear(x, state[q_prefix + "linear_k.weight"], state[q_prefix + "linear_k.bias"]).view(B, -1, nh, dk)
    v = linear(x, state[q_prefix + "linear_v.weight"], state[q_prefix + "linear_v.bias"]).view(B, -1, nh, dk)
    k = k.transpose(1, 2)
    v = v.transpose(1, 2)
    p = linear(pos_emb, state[q_prefix + "linear_pos.weight"]).view(pos_emb.shape[0], -1, nh, dk).transpose(1, 2)  # :381
    q_u = (q + _f(state[q_prefix + "pos_bias_u"])).transpose(1, 2)                                 # :384
    q_v = (q + _f(state[q_prefix + "pos_bias_v"])).transpose(1, 2)                                 # :386
    matrix_ac = torch.matmul(q_u, k.transpose(-2, -1))                                             # :391
    matrix_bd = torch.matmul(q_v, p.transpose(-2, -1))                                             # :394 (no rel_shift)
    scores = (matrix_ac + matrix_bd) / math.sqrt(dk)                                               # :398
    m = mask.unsqueeze(1).eq(0)                                                                    # :404
    scores = scores.masked_fill(m, float(torch.finfo(torch.float16).min))                          # :295,405
    attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)                                       # :406-408
    x = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, -1, nh * dk)                    # :415-418
    return linear(x, state[q_prefix + "linear_out.weight"], state[q_prefix + "linear_out.bias"])


def whale_encoder(state, acfg, feats, lengths):
    """whaleEncoder.forward (encoder.py:121-147): pad mask -> GlobalCMVN -> Conv2dSubsampling4 -> Transformer."""
    C = acfg.hidden_size
    T = feats.shape[1]
    lengths = lengths.to(torch.int64)
    masks = ~(torch.arange(T)[None, :] >= lengths[:, None])                                        # utils.py:78-85
    masks = masks.unsqueeze(1)                                                                     # (B, 1, T)
    x = _f(feats)
    if (PA + "encoder.global_cmvn.mean") in state:
        x = (x - _f(state[PA + "encoder.global_cmvn.mean"])) * _f(state[PA + "encoder.global_cmvn.istd"])  # cmvn.py:30-32
    # Conv2dSubsampling4.forward, component/subsampling.py:37-43
    x = x.unsqueeze(1)
    x = F.relu(F.conv2d(x, _f(state[PA + "encoder.enc.0.core.conv.0.weight"]),
                        _f(state[PA + "encoder.enc.0.core.conv.0.bias"]), stride=2))
    x = F.relu(F.conv2d(x, _f(state[PA + "encoder.enc.0.core.conv.2.weight"]),
                        _f(state[PA + "encoder.enc.0.core.conv.2.bias"]), stride=2))
    b, c, t, f = x.size()
    x = linear(x.transpose(1, 2).contiguous().view(b, t, c * f), state[PA + "encoder.enc.0.core.out.0.weight"],
               state[PA + "encoder.enc.0.core.out.0.bias"])
    masks = masks[:, :, 2::2][:, :, 2::2]
    # Transformer.forward, component/transformer.py:374-394 (dynamic chunks off -> plain padding mask, utils.py:141-146)
    x = linear(x, state[PA + "encoder.enc.1.embed.0.weight"], state[PA + "encoder.enc.1.embed.0.bias"])
    x = F.layer_norm(x, (C,), _f(state[PA + "encoder.enc.1.embed.1.weight"]), _f(state[PA + "encoder.enc.1.embed.1.bias"]))
    x = F.relu(x)                                                                                  # :313-318
    x = x * math.sqrt(C)                                                                           # attention.py:109
    pos_emb = whale_position_table(acfg.max_len, C)[None, : x.shape[1]]                            # attention.py:110
    for l in range(acfg.num_blocks):
        p = PA + f"encoder.enc.1.encoders.{l}."
        r = x                                                                                      # transformer.py:106-113
        y = F.layer_norm(x, (C,), _f(state[p + "norm1.weight"]), _f(state[p + "norm1.bias"]))
        x = r + whale_attention(state, acfg, p + "self_attn.", y, masks, pos_emb)
        r = x                                                                                      # :117-120
        y = F.layer_norm(x, (C,), _f(state[p + "norm2.weight"]), _f(state[p + "norm2.bias"]))
        y = linear(F.relu(linear(y, state[p + "feed_forward.w_1.weight"], state[p + "feed_forward.w_1.bias"])),
                   state[p + "feed_forward.w_2.weight"], state[p + "feed_forward.w_2.bias"])       # attention.py:145-147
        x = r + y
    x = F.layer_norm(x, (C,), _f(state[PA + "encoder.enc.1.after_norm.weight"]),
                     _f(state[PA + "encoder.enc.1.after_norm.bias"]))                              # :391-392
    return x, masks


def whale_adapter(state, acfg, x, mask_pad):
    """CNNSubsampling.forward (cnn_num == 1 branch), whale/adapter.py:107-136."""
    k = acfg.adapter_kernel
    x = x.transpose(1, 2)
    if mask_pad.size(2) > 0:
        x = x.masked_fill(~mask_pad, 0.0)                                                          # :115-116
    x = F.pad(x, (0, k - 1))                                                                       # :93,124
    x = F.conv1d(x, _f(state[PA + "adpter.conv1d2.weight"]), _f(state[PA + "adpter.conv1d2.bias"]), stride=2)
    x = x.transpose(1, 2)
    x = F.layer_norm(x, (x.shape[-1],), _f(state[PA + "adpter.bn2.weight"]), _f(state[PA + "adpter.bn2.bias"]),
                     acfg.adapter_ln_eps)                                                          # :98,126-130
    x = F.gelu(x)                                                                                  # :100,131
    x = linear(x, state[PA + "adpter.project.weight"], state[PA + "adpter.project.bias"])          # :104,134
    return x, mask_pad[:, :, 0::2]


def encode_audios(state, cfg, feats, lengths):
    """audioEncoder.forward, whale/init_model.py:114-139."""
    enc, mask = whale_encoder(state, cfg.audio, feats, lengths)
    emb, mask = whale_adapter(state, cfg.audio, enc, mask)
    return {"inputs_embeds": emb, "attention_mask": mask.squeeze(1)}


# ================================================================================================ Mixtral decoder
def rmsnorm(x, w, eps):
    """MixtralRMSNorm.forward, transformers modeling_mixtral.py:148-153."""
    x = _f(x)
    var = x.pow(2).mean(-1, keepdim=True)
    return _f(w) * (x * torch.rsqrt(var + eps))


def rope_cos_sin(positions, head_dim, theta):
    """MixtralRotaryEmbedding.forward, modeling_mixtral.py:210-221 (default rope, attention_scaling = 1)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = positions.to(torch.float32)[..., None] * inv_freq
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    """modeling_mixtral.py:224-228."""
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q, k, cos, sin):
    """apply_rotary_pos_emb, modeling_mixtral.py:232-254; q,k [B, heads, S, D], cos/sin [B, S, D]."""
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


def router_topk(xn, gate_w, top_k=2):
    """MixtralTopKRouter.forward, modeling_mixtral.py:109-116."""
    logits = linear(xn, gate_w)
    probs = F.softmax(logits.float(), dim=-1)
    top_v, top_i = torch.topk(probs, top_k, dim=-1)
    top_v = top_v / top_v.sum(dim=-1, keepdim=True)
    return probs, top_v, top_i


def expert_weights(state, lcfg, l, e):
    p = f"model.layers.{l}.block_sparse_moe.experts.{e}."
    return state[p + "w1.weight"], state[p + "w3.weight"], state[p + "w2.weight"]


def sparse_moe(state, lcfg, l, xn, trace=None, route=None):
    """MixtralSparseMoeBlock / MixtralExperts.forward, modeling_mixtral.py:74-98,127-136; xn [T, H].

    `route` (checker aid, default None = the reference's own routing): {"ids": [T, 2] expert ids}.  The two experts of
    every token are then the given ones, their weights the oracle's OWN soft-max probabilities of those two experts,
    renormalised (:113-115 applied to the given pair); the dict receives "logits" [T, E] (the oracle's router logits)
    and "own_ids" [T, 2] (the pair the oracle itself would have picked) so the caller can judge every difference."""
    gate_w = state[f"model.layers.{l}.block_sparse_moe.gate.weight"]
    probs, top_v, top_i = router_topk(xn, gate_w, lcfg.num_experts_per_tok)
    if trace is not None:
        trace["router_probs"] = probs
    if route is not None:
        route["logits"], route["own_ids"] = linear(xn, gate_w).float(), top_i
        top_i = route["ids"].to(top_i.device).long()
        top_v = probs.gather(1, top_i)
        top_v = top_v / top_v.sum(dim=-1, keepdim=True)
    out = torch.zeros_like(xn)
    for e in range(lcfg.num_local_experts):
        tok, kpos = torch.where(top_i == e)
        if tok.numel() == 0:
            continue
        w1, w3, w2 = expert_weights(state, lcfg, l, e)
        cur = xn[tok]
        hcur = F.silu(linear(cur, w1)) * linear(cur, w3)                                            # :92-93
        hcur = linear(hcur, w2) * top_v[tok, kpos, None]                                            # :94-95
        out.index_add_(0, tok, hcur)                                                                # :96
    return out, top_i, top_v


def decoder_layer(state, lcfg, l, h, positions, past_kv=None, trace=None, route=None):
    """MixtralDecoderLayer.forward :365-390 + MixtralAttention.forward :312-351 (eager/sdpa causal GQA).

    h [B, S, H]; positions [B, S]; past_kv = (k, v) each [B, n_kv, P, D] or None.  Returns (h, (k, v))."""
    p = f"model.layers.{l}."
    B, S, H = h.shape
    nq, nkv, D = lcfg.num_attention_heads, lcfg.num_key_value_heads, lcfg.head_dim
    x = rmsnorm(h, state[p + "input_layernorm.weight"], lcfg.rms_norm_eps)
    q = linear(x, state[p + "self_attn.q_proj.weight"]).view(B, S, nq, D).transpose(1, 2)
    k = linear(x, state[p + "self_attn.k_proj.weight"]).view(B, S, nkv, D).transpose(1, 2)
    v = linear(x, state[p + "self_attn.v_proj.weight"]).view(B, S, nkv, D).transpose(1, 2)
    cos, sin = rope_cos_sin(positions, D, lcfg.rope_theta)
    q, k = apply_rope(q, k, cos, sin)
    if past_kv is not None:
        k = torch.cat([past_kv[0], k], dim=2)
        v = torch.cat([past_kv[1], v], dim=2)
    P = k.shape[2]
    kk = k.repeat_interleave(nq // nkv, dim=1)                                                      # repeat_kv :257-266
    vv = v.repeat_interleave(nq // nkv, dim=1)
    scores = torch.matmul(q, kk.transpose(2, 3)) * (D ** -0.5)                                      # :281
    qpos = torch.arange(P - S, P)[:, None]
    causal = torch.arange(P)[None, :] > qpos
    scores = scores.masked_fill(causal[None, None], float("-inf"))
    attn = F.softmax(scores, dim=-1, dtype=torch.float32)                                           # :287
    o = torch.matmul(attn, vv).transpose(1, 2).reshape(B, S, nq * D)
    h = h + linear(o, state[p + "self_attn.o_proj.weight"])                                         # :378-380
    xn = rmsnorm(h, state[p + "post_attention_layernorm.weight"], lcfg.rms_norm_eps)
    if trace is not None:
        trace["h_mid"] = h
    y, _, _ = sparse_moe(state, lcfg, l, xn.reshape(-1, H), trace, route)
    return h + y.reshape(B, S, H), (k, v)                                                           # :386-389


def mixtral_forward(state, lcfg, inputs_embeds, positions=None, past=None, last_only=False, trace=None):
    """custom_forward (vita/model/language_model/vita_mixtral.py:101-215): self.model(...) then lm_head on all rows.

    `trace` (a list) collects per-layer {h_in, h_mid, router_probs, h_out} for the layer-wise parity tests."""
    B, S, _ = inputs_embeds.shape
    past_len = 0 if past is None else past[0][0].shape[2]
    if positions is None:
        positions = torch.arange(past_len, past_len + S)[None].expand(B, S)
    h = _f(inputs_embeds)
    new_past = []
    for l in range(lcfg.num_hidden_layers):
        t = None
        if trace is not None:
            t = {"h_in": h}
            trace.append(t)
        h, kv = decoder_layer(state, lcfg, l, h, positions, None if past is None else past[l], t)
        if t is not None:
            t["h_out"] = h
        new_past.append(kv)
    h = rmsnorm(h, state["model.norm.weight"], lcfg.rms_norm_eps)
    if last_only:
        h = h[:, -1:]
    logits = linear(h, state["lm_head.weight"])                                                     # vita_mixtral.py:172
    return logits, new_past, h


# ================================================================================================ multimodal splice
def prepare_inputs_embeds(state, cfg, input_ids: torch.Tensor, images: Optional[torch.Tensor],
                          audios: Optional[dict], image_features=None, audio_features=None):
    """prepare_inputs_labels_for_multimodal, vita/model/vita_arch.py:151-407 (inference subset: no labels,
    attention_mask=None, right padding).  Returns (inputs_embeds [B, S, H], lengths list).

    `image_features` / `audio_features` may be supplied (e.g. produced by the CUDA encoders) to check the splice
    arithmetic in isolation."""
    embed = _f(state["model.embed_tokens.weight"])
    if image_features is None:
        image_features = encode_images(state, cfg, images)                                          # :177-184
    if audio_features is None:
        audio_features = encode_audios(state, cfg, audios["audios"], audios["lengths"])["inputs_embeds"]  # :186-189
    ids_list = [row for row in input_ids]
    n_img_ph = sum(int((r == IMAGE_TOKEN_INDEX).sum()) for r in ids_list)
    n_aud_ph = sum(int((r == AUDIO_TOKEN_INDEX).sum()) for r in ids_list)
    assert n_img_ph + sum(int(IMAGE_TOKEN_INDEX not in r) for r in ids_list) == image_features.shape[0]   # :227-231
    assert n_aud_ph + sum(int(AUDIO_TOKEN_INDEX not in r) for r in ids_list) == audio_features.shape[0]   # :232-236
    new_embeds = []
    ii = ai = 0
    for cur in ids_list:
        n_i = int((cur == IMAGE_TOKEN_INDEX).sum())
        n_a = int((cur == AUDIO_TOKEN_INDEX).sum())
        if n_i == 0 and n_a == 0:                                                                   # :240-252
            new_embeds.append(embed[cur])
            ii += 1
            ai += 1
            continue
        idx = [-1] + torch.where((cur == IMAGE_TOKEN_INDEX) | (cur == AUDIO_TOKEN_INDEX))[0].tolist() + [cur.shape[0]]
        parts = []
        for i in range(len(idx) - 1):
            parts.append(embed[cur[idx[i] + 1: idx[i + 1]]])                                        # :263-276
            if i < n_i + n_a:
                tok = int(cur[idx[i + 1]])
                if tok == IMAGE_TOKEN_INDEX:
                    parts.append(image_features[ii]); ii += 1                                       # :281-292
                elif tok == AUDIO_TOKEN_INDEX:
                    parts.append(audio_features[ai]); ai += 1                                       # :293-304
                else:
                    raise ValueError                                                                # :305-306
        if n_i != 0 and n_a == 0:
            ai += 1                                                                                 # :309-312
        elif n_i == 0 and n_a != 0:
            ii += 1                                                                                 # :313-316
        new_embeds.append(torch.cat(parts))
    assert ii == image_features.shape[0] and ai == audio_features.shape[0]                          # :323-324
    max_model = cfg.llm.tokenizer_model_max_length
    if max_model is not None:
        new_embeds = [x[:max_model] for x in new_embeds]                                            # :326-329
    lens = [x.shape[0] for x in new_embeds]
    S = max(lens)
    out = torch.zeros(len(new_embeds), S, embed.shape[1])
    for i, x in enumerate(new_embeds):
        out[i, : x.shape[0]] = x                                                                    # :372-392 (right pad)
    return out, lens


def forward(state, cfg, input_ids, images=None, audios=None, past=None, last_only=False):
    """VITAMixtralForCausalLM.forward, vita/model/language_model/vita_mixtral.py:249-289."""
    if images is None or input_ids.shape[1] == 1:                                                   # vita_arch.py:155-175
        emb = _f(state["model.embed_tokens.weight"])[input_ids]
    else:
        emb, _ = prepare_inputs_embeds(state, cfg, input_ids, images, audios)
    return mixtral_forward(state, cfg.llm, emb, past=past, last_only=last_only)


def greedy_generate(state, cfg, input_ids, images=None, audios=None, max_new_tokens=8,
                    teacher: Optional[List[int]] = None):
    """Greedy decode through `forward` (the manual loop of SURVEY.md Appendix C; HF generate() at
    video_audio_demo.py:257-270 is its reference).  Returns (tokens, per-step last-row logits).
    With `teacher`, the given tokens are fed instead of the arg-max (teacher forcing)."""
    logits, past, _ = forward(state, cfg, input_ids, images, audios, last_only=True)
    toks, all_logits = [], []
    for step in range(max_new_tokens):
        row = logits[0, -1]
        all_logits.append(row)
        nxt = int(row.argmax())
        toks.append(nxt)
        feed = nxt if teacher is None else teacher[step]
        if step + 1 < max_new_tokens:
            logits, past, _ = forward(state, cfg, torch.tensor([[feed]]), past=past, last_only=True)
    return toks, torch.stack(all_logits)
