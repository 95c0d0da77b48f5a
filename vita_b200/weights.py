"""Weights: reference-named state dicts <-> the kernel-native packed layout.

* ``synthetic_state(cfg, seed)`` -- deterministic, seeded, bf16-representable tensors under the *reference's* parameter
  names (the names `VITAMixtralForCausalLM.state_dict()` produces with transformers 4.41, i.e. the shipped
  checkpoint; see SURVEY.md section 5 "Weight-name map").  Each tensor has its own generator seeded from
  (seed, crc32(name)), so any subset / depth reproduces the same values on any machine.
* ``pack(state, cfg, device)`` -- the layouts the CUDA kernels consume: fused q|k|v, stacked expert tensors
  [E, 2I, H] (gate rows then up rows) and [E, H, I], conv weights re-ordered for the im2col column order, Whale
  position table.  Works for synthetic and real checkpoints alike (name map of
  web_demo/vllm_tools/vllm_file/mixtral.py:1190-1229 and vita/model/builder.py:245-257).
* ``random_packed(cfg, device, seed)`` -- packed weights drawn directly on the GPU (bench only: the full 46.9 B
  parameter model does not fit host RAM).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable

import torch

from .config import VitaConfig

BF16 = torch.bfloat16
PREFIX_VISION = "model.vision_tower.vision_tower."
PREFIX_AUDIO = "model.audio_encoder."
PREFIX_PROJ = "model.mm_projector."


# ------------------------------------------------------------------------------------------------ names / shapes
def llm_param_shapes(cfg: VitaConfig) -> Dict[str, tuple]:
    c = cfg.llm
    H, I, D = c.hidden_size, c.intermediate_size, c.head_dim
    s = {"model.embed_tokens.weight": (c.vocab_size, H), "model.norm.weight": (H,), "lm_head.weight": (c.vocab_size, H)}
    for l in range(c.num_hidden_layers):
        p = f"model.layers.{l}."
        s[p + "input_layernorm.weight"] = (H,)
        s[p + "post_attention_layernorm.weight"] = (H,)
        s[p + "self_attn.q_proj.weight"] = (c.num_attention_heads * D, H)
        s[p + "self_attn.k_proj.weight"] = (c.num_key_value_heads * D, H)
        s[p + "self_attn.v_proj.weight"] = (c.num_key_value_heads * D, H)
        s[p + "self_attn.o_proj.weight"] = (H, c.num_attention_heads * D)
        s[p + "block_sparse_moe.gate.weight"] = (c.num_local_experts, H)
        for e in range(c.num_local_experts):
            q = p + f"block_sparse_moe.experts.{e}."
            s[q + "w1.weight"] = (I, H)   # gate
            s[q + "w3.weight"] = (I, H)   # up
            s[q + "w2.weight"] = (H, I)   # down
    return s


def vision_param_shapes(cfg: VitaConfig) -> Dict[str, tuple]:
    v = cfg.vision
    H, M = v.hidden_size, v.intermediate_size
    p = PREFIX_VISION
    s = {p + "embeddings.class_embedding": (1, 1, H),
         p + "embeddings.patch_embedding.weight": (H, v.num_channels, v.patch_size, v.patch_size),
         p + "embeddings.patch_embedding.bias": (H,),
         p + "embeddings.position_embedding": (1, v.num_patches + 1, H)}
    for l in range(v.num_hidden_layers):
        q = p + f"encoder.layers.{l}."
        s[q + "attn.qkv.weight"] = (3 * H, H); s[q + "attn.qkv.bias"] = (3 * H,)
        s[q + "attn.proj.weight"] = (H, H); s[q + "attn.proj.bias"] = (H,)
        s[q + "mlp.fc1.weight"] = (M, H); s[q + "mlp.fc1.bias"] = (M,)
        s[q + "mlp.fc2.weight"] = (H, M); s[q + "mlp.fc2.bias"] = (H,)
        s[q + "norm1.weight"] = (H,); s[q + "norm1.bias"] = (H,)
        s[q + "norm2.weight"] = (H,); s[q + "norm2.bias"] = (H,)
        s[q + "ls1"] = (H,); s[q + "ls2"] = (H,)
    return s


def projector_param_shapes(cfg: VitaConfig) -> Dict[str, tuple]:
    H = cfg.llm.hidden_size
    p = PREFIX_PROJ
    return {p + "0.weight": (H, cfg.vision.out_dim), p + "0.bias": (H,), p + "2.weight": (H, H), p + "2.bias": (H,)}


def audio_param_shapes(cfg: VitaConfig) -> Dict[str, tuple]:
    a = cfg.audio
    C, F2, nh, dk = a.hidden_size, a.freq_bins, a.num_attention_heads, a.head_dim
    p = PREFIX_AUDIO
    s = {p + "encoder.global_cmvn.mean": (a.input_dim,), p + "encoder.global_cmvn.istd": (a.input_dim,),
         p + "encoder.enc.0.core.conv.0.weight": (C, 1, 3, 3), p + "encoder.enc.0.core.conv.0.bias": (C,),
         p + "encoder.enc.0.core.conv.2.weight": (C, C, 3, 3), p + "encoder.enc.0.core.conv.2.bias": (C,),
         p + "encoder.enc.0.core.out.0.weight": (C, C * F2), p + "encoder.enc.0.core.out.0.bias": (C,),
         p + "encoder.enc.1.embed.0.weight": (C, C), p + "encoder.enc.1.embed.0.bias": (C,),
         p + "encoder.enc.1.embed.1.weight": (C,), p + "encoder.enc.1.embed.1.bias": (C,),
         p + "encoder.enc.1.after_norm.weight": (C,), p + "encoder.enc.1.after_norm.bias": (C,),
         p + "adpter.conv1d2.weight": (2 * C, C, a.adapter_kernel), p + "adpter.conv1d2.bias": (2 * C,),
         p + "adpter.bn2.weight": (2 * C,), p + "adpter.bn2.bias": (2 * C,),
         p + "adpter.project.weight": (cfg.llm.hidden_size, 2 * C), p + "adpter.project.bias": (cfg.llm.hidden_size,)}
    for l in range(a.num_blocks):
        q = p + f"encoder.enc.1.encoders.{l}."
        s[q + "self_attn.pos_bias_u"] = (nh, dk); s[q + "self_attn.pos_bias_v"] = (nh, dk)
        for n in ("q", "k", "v", "out"):
            s[q + f"self_attn.linear_{n}.weight"] = (C, C); s[q + f"self_attn.linear_{n}.bias"] = (C,)
        s[q + "self_attn.linear_pos.weight"] = (C, C)
        s[q + "feed_forward.w_1.weight"] = (a.linear_units, C); s[q + "feed_forward.w_1.bias"] = (a.linear_units,)
        s[q + "feed_forward.w_2.weight"] = (C, a.linear_units); s[q + "feed_forward.w_2.bias"] = (C,)
        for n in ("norm1", "norm2"):
            s[q + n + ".weight"] = (C,); s[q + n + ".bias"] = (C,)
    return s


def all_param_shapes(cfg: VitaConfig, parts: Iterable[str] = ("llm", "vision", "projector", "audio")):
    out: Dict[str, tuple] = {}
    for part in parts:
        out.update({"llm": llm_param_shapes, "vision": vision_param_shapes, "projector": projector_param_shapes,
                    "audio": audio_param_shapes}[part](cfg))
    return out


# ------------------------------------------------------------------------------------------------ synthetic values
def _kind(name: str) -> str:
    last = name.rsplit(".", 1)[-1]
    if "global_cmvn.mean" in name:
        return "cmvn_mean"
    if "global_cmvn.istd" in name:
        return "cmvn_istd"
    if "layernorm" in name or name.endswith("model.norm.weight") or ".norm1." in name or ".norm2." in name \
            or "after_norm" in name or ".bn2." in name or ".embed.1." in name:
        return "gain" if last == "weight" else "bias"
    if last in ("ls1", "ls2"):
        return "layerscale"
    if last == "bias" or "pos_bias" in name:
        return "bias"
    if "class_embedding" in name or "position_embedding" in name:
        return "embedding"
    return "matrix"


def synthetic_tensor(name: str, shape: tuple, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    kind = _kind(name)
    n = torch.randn(shape, generator=g, dtype=torch.float32)
    if kind == "matrix":
        t = n * 0.02
    elif kind == "gain":
        t = 1.0 + 0.1 * n
    elif kind == "layerscale":
        t = 0.5 + 0.05 * n
    elif kind == "bias":
        t = 0.02 * n
    elif kind == "embedding":
        t = 0.02 * n
    elif kind == "cmvn_mean":
        t = 0.5 * n
    elif kind == "cmvn_istd":
        t = 1.0 + 0.1 * n.abs()
    else:
        raise AssertionError(kind)
    if kind.startswith("cmvn"):
        return t  # fp32 buffers in the reference
    return t.to(BF16)


def synthetic_state(cfg: VitaConfig, seed: int = 0, parts=("llm", "vision", "projector", "audio")):
    """name -> tensor (bf16, except the fp32 CMVN statistics), reference naming."""
    return {name: synthetic_tensor(name, shape, seed) for name, shape in all_param_shapes(cfg, parts).items()}


# ------------------------------------------------------------------------------------------------ packing
def whale_position_table(max_len: int, d_model: int) -> torch.Tensor:
    """PositionalEncoding.__init__ (whale/module/layer/attention.py:24-36), fp32 [max_len, d_model]."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def rope_table(max_pos: int, head_dim: int, theta: float) -> torch.Tensor:
    """MixtralRotaryEmbedding (modeling_mixtral.py:159-221): fp32 [max_pos, 2, head_dim/2] = (cos, sin)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv_freq[None, :]
    return torch.stack([freqs.cos(), freqs.sin()], dim=1).contiguous()


def default_rope_len(c) -> int:
    """Positions covered by the packed cos/sin table: the prompt limit plus room for the reply; the decoder grows the
    table when it is built with a longer max_seq_len (MixtralDecoder.__init__)."""
    return min(c.max_position_embeddings, max(c.tokenizer_model_max_length + 1024, 2048))


def pack_llm(state, cfg: VitaConfig, device, ep=None) -> dict:
    c = cfg.llm
    e_lo, e_hi = (0, c.num_local_experts) if ep is None else expert_range(c.num_local_experts, ep[0], ep[1])
    dev = lambda t: t.to(device=device, dtype=BF16).contiguous()
    out = {"embed": dev(state["model.embed_tokens.weight"]), "norm": dev(state["model.norm.weight"]),
           "lm_head": dev(state["lm_head.weight"]), "layers": []}
    for l in range(c.num_hidden_layers):
        p = f"model.layers.{l}."
        moe = "block_sparse_moe." if (p + "block_sparse_moe.gate.weight") in state else "mlp."
        if (p + moe + "experts.gate_up_proj") in state:           # transformers >= 5 naming
            w13 = state[p + moe + "experts.gate_up_proj"][e_lo:e_hi]
            w2 = state[p + moe + "experts.down_proj"][e_lo:e_hi]
        else:
            ex = [p + moe + f"experts.{e}." for e in range(e_lo, e_hi)]
            w13 = torch.stack([torch.cat([state[q + "w1.weight"], state[q + "w3.weight"]], dim=0) for q in ex])
            w2 = torch.stack([state[q + "w2.weight"] for q in ex])
        out["layers"].append({
            "ln1": dev(state[p + "input_layernorm.weight"]),
            "wqkv": dev(torch.cat([state[p + "self_attn.q_proj.weight"], state[p + "self_attn.k_proj.weight"],
                                   state[p + "self_attn.v_proj.weight"]], dim=0)),
            "wo": dev(state[p + "self_attn.o_proj.weight"]),
            "ln2": dev(state[p + "post_attention_layernorm.weight"]),
            "gate": dev(state[p + moe + "gate.weight"]),
            "w13": dev(w13), "w2": dev(w2)})
    out["rope"] = rope_table(default_rope_len(c),
                             c.head_dim, c.rope_theta).to(device)
    out["ep"] = (0, 1) if ep is None else tuple(ep)
    return out


def pack_vision(state, cfg: VitaConfig, device) -> dict:
    v = cfg.vision
    p = PREFIX_VISION
    dev = lambda t: t.to(device=device, dtype=BF16).contiguous()
    pw = state[p + "embeddings.patch_embedding.weight"].reshape(v.hidden_size, v.patch_k)
    pw_pad = torch.zeros(v.hidden_size, v.patch_k_pad, dtype=pw.dtype)
    pw_pad[:, :v.patch_k] = pw
    out = {"patch_w": dev(pw_pad), "patch_b": dev(state[p + "embeddings.patch_embedding.bias"]),
           "cls": dev(state[p + "embeddings.class_embedding"].reshape(-1)),
           "pos": dev(state[p + "embeddings.position_embedding"].reshape(v.num_patches + 1, v.hidden_size)),
           "layers": []}
    for l in range(v.num_hidden_layers):
        q = p + f"encoder.layers.{l}."
        out["layers"].append({k: dev(state[q + n]) for k, n in (
            ("ln1_w", "norm1.weight"), ("ln1_b", "norm1.bias"), ("qkv_w", "attn.qkv.weight"),
            ("qkv_b", "attn.qkv.bias"), ("proj_w", "attn.proj.weight"), ("proj_b", "attn.proj.bias"), ("ls1", "ls1"),
            ("ln2_w", "norm2.weight"), ("ln2_b", "norm2.bias"), ("fc1_w", "mlp.fc1.weight"), ("fc1_b", "mlp.fc1.bias"),
            ("fc2_w", "mlp.fc2.weight"), ("fc2_b", "mlp.fc2.bias"), ("ls2", "ls2"))})
    return out


def pack_projector(state, cfg: VitaConfig, device) -> dict:
    dev = lambda t: t.to(device=device, dtype=BF16).contiguous()
    p = PREFIX_PROJ
    return {"w0": dev(state[p + "0.weight"]), "b0": dev(state[p + "0.bias"]), "w2": dev(state[p + "2.weight"]),
            "b2": dev(state[p + "2.bias"])}


def pack_audio(state, cfg: VitaConfig, device) -> dict:
    a = cfg.audio
    C, F2 = a.hidden_size, a.freq_bins
    p = PREFIX_AUDIO
    dev = lambda t: t.to(device=device, dtype=BF16).contiguous()
    f32 = lambda t: t.to(device=device, dtype=torch.float32).contiguous()
    conv2 = state[p + "encoder.enc.0.core.conv.2.weight"]                      # [C_out, C_in, kt, kf]
    lin = state[p + "encoder.enc.0.core.out.0.weight"]                         # [C, C_in * F2], column = c * F2 + f
    c1d = state[p + "adpter.conv1d2.weight"]                                   # [2C, C, k]
    out = {
        "cmvn_mean": f32(state[p + "encoder.global_cmvn.mean"]) if (p + "encoder.global_cmvn.mean") in state else None,
        "cmvn_istd": f32(state[p + "encoder.global_cmvn.istd"]) if (p + "encoder.global_cmvn.istd") in state else None,
        "conv1_w": dev(state[p + "encoder.enc.0.core.conv.0.weight"].reshape(C, 9)),
        "conv1_b": dev(state[p + "encoder.enc.0.core.conv.0.bias"]),
        # im2col column order is (kt, kf, c_in)
        "conv2_w": dev(conv2.permute(0, 2, 3, 1).reshape(C, 9 * C)),
        "conv2_b": dev(state[p + "encoder.enc.0.core.conv.2.bias"]),
        # conv2 GEMM output is [b, t, f, c]; the reference flattens (c, f) -> re-order the Linear's columns to (f, c)
        "sub_out_w": dev(lin.reshape(C, C, F2).permute(0, 2, 1).reshape(C, F2 * C)),
        "sub_out_b": dev(state[p + "encoder.enc.0.core.out.0.bias"]),
        "embed_w": dev(state[p + "encoder.enc.1.embed.0.weight"]), "embed_b": dev(state[p + "encoder.enc.1.embed.0.bias"]),
        "embed_ln_w": dev(state[p + "encoder.enc.1.embed.1.weight"]),
        "embed_ln_b": dev(state[p + "encoder.enc.1.embed.1.bias"]),
        "after_w": dev(state[p + "encoder.enc.1.after_norm.weight"]),
        "after_b": dev(state[p + "encoder.enc.1.after_norm.bias"]),
        "ad_conv_w": dev(c1d.permute(0, 2, 1).reshape(2 * C, a.adapter_kernel * C)),   # column = kk * C + c
        "ad_conv_b": dev(state[p + "adpter.conv1d2.bias"]),
        "ad_ln_w": dev(state[p + "adpter.bn2.weight"]), "ad_ln_b": dev(state[p + "adpter.bn2.bias"]),
        "ad_proj_w": dev(state[p + "adpter.project.weight"]), "ad_proj_b": dev(state[p + "adpter.project.bias"]),
        "pos_table": dev(whale_position_table(a.max_len, C)),
        "layers": []}
    for l in range(a.num_blocks):
        q = p + f"encoder.enc.1.encoders.{l}.self_attn."
        r = p + f"encoder.enc.1.encoders.{l}."
        out["layers"].append({
            "ln1_w": dev(state[r + "norm1.weight"]), "ln1_b": dev(state[r + "norm1.bias"]),
            "qkv_w": dev(torch.cat([state[q + "linear_q.weight"], state[q + "linear_k.weight"],
                                    state[q + "linear_v.weight"]], dim=0)),
            "qkv_b": dev(torch.cat([state[q + "linear_q.bias"], state[q + "linear_k.bias"],
                                    state[q + "linear_v.bias"]], dim=0)),
            "pos_w": dev(state[q + "linear_pos.weight"]),
            "bias_u": dev(state[q + "pos_bias_u"].reshape(-1)), "bias_v": dev(state[q + "pos_bias_v"].reshape(-1)),
            "out_w": dev(state[q + "linear_out.weight"]), "out_b": dev(state[q + "linear_out.bias"]),
            "ln2_w": dev(state[r + "norm2.weight"]), "ln2_b": dev(state[r + "norm2.bias"]),
            "w1": dev(state[r + "feed_forward.w_1.weight"]), "b1": dev(state[r + "feed_forward.w_1.bias"]),
            "w2": dev(state[r + "feed_forward.w_2.weight"]), "b2": dev(state[r + "feed_forward.w_2.bias"])})
    return out


def pack(state, cfg: VitaConfig, device) -> dict:
    out = {"llm": pack_llm(state, cfg, device)}
    if any(k.startswith(PREFIX_VISION) for k in state):
        out["vision"] = pack_vision(state, cfg, device)
        out["projector"] = pack_projector(state, cfg, device)
    if any(k.startswith(PREFIX_AUDIO) for k in state):
        out["audio"] = pack_audio(state, cfg, device)
    return out


# ------------------------------------------------------------------------------------------------ on-device random
def expert_range(num_experts: int, rank: int, world: int):
    """Contiguous expert ownership for expert parallelism: rank r owns experts [lo, hi)."""
    assert num_experts % world == 0, "the number of experts must be divisible by the EP world size"
    per = num_experts // world
    return rank * per, (rank + 1) * per


def random_packed(cfg: VitaConfig, device, seed: int = 0, parts=("llm", "vision", "projector", "audio"), ep=None) -> dict:
    """Packed weights with the same shapes/statistics as pack(synthetic_state()), drawn on the GPU (bench only).

    Every tensor (and every expert of the stacked expert tensors) has its own generator seed derived from
    (seed, name), so a rank that holds only experts [lo, hi) (`ep=(rank, world)`) gets exactly the values the
    single-GPU model has for them, and all ranks agree on the replicated tensors."""
    g = torch.Generator(device=device)

    def draw(name, shape, kind):
        g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
        if kind == "gain":
            return (1.0 + 0.1 * torch.randn(shape, device=device, generator=g)).to(BF16)
        if kind == "ls":
            return (0.5 + 0.05 * torch.randn(shape, device=device, generator=g)).to(BF16)
        return torch.empty(shape, device=device, dtype=BF16).normal_(0.0, 0.02, generator=g)

    def fill(tree, prefix):
        for k, v in (tree.items() if isinstance(tree, dict) else enumerate(tree)):
            name = f"{prefix}.{k}"
            if isinstance(v, (dict, list)):
                fill(v, name)
            elif isinstance(v, tuple):
                shape, kind = v
                if kind == "experts":                       # stacked [E_local, rows, cols], one seed per expert
                    lo, hi, rows, cols = shape
                    tree[k] = torch.stack([draw(f"{name}.{e}", (rows, cols), "m") for e in range(lo, hi)])
                else:
                    tree[k] = draw(name, shape, kind)

    c, v, a = cfg.llm, cfg.vision, cfg.audio
    H, I, E = c.hidden_size, c.intermediate_size, c.num_local_experts
    lo, hi = (0, E) if ep is None else expert_range(E, ep[0], ep[1])
    out = {}
    if "llm" in parts:
        out["llm"] = {"embed": ((c.vocab_size, H), "m"), "norm": ((H,), "gain"), "lm_head": ((c.vocab_size, H), "m"),
                      "layers": [{"ln1": ((H,), "gain"), "wqkv": ((c.qkv_rows, H), "m"),
                                  "wo": ((H, c.num_attention_heads * c.head_dim), "m"), "ln2": ((H,), "gain"),
                                  "gate": ((E, H), "m"), "w13": ((lo, hi, 2 * I, H), "experts"),
                                  "w2": ((lo, hi, H, I), "experts")}
                                 for _ in range(c.num_hidden_layers)]}
    if "vision" in parts:
        VH, VM = v.hidden_size, v.intermediate_size
        out["vision"] = {"patch_w": ((VH, v.patch_k_pad), "m"), "patch_b": ((VH,), "m"), "cls": ((VH,), "m"),
                         "pos": ((v.num_patches + 1, VH), "m"),
                         "layers": [{"ln1_w": ((VH,), "gain"), "ln1_b": ((VH,), "m"), "qkv_w": ((3 * VH, VH), "m"),
                                     "qkv_b": ((3 * VH,), "m"), "proj_w": ((VH, VH), "m"), "proj_b": ((VH,), "m"),
                                     "ls1": ((VH,), "ls"), "ln2_w": ((VH,), "gain"), "ln2_b": ((VH,), "m"),
                                     "fc1_w": ((VM, VH), "m"), "fc1_b": ((VM,), "m"), "fc2_w": ((VH, VM), "m"),
                                     "fc2_b": ((VH,), "m"), "ls2": ((VH,), "ls")} for _ in range(v.num_hidden_layers)]}
    if "projector" in parts:
        out["projector"] = {"w0": ((H, v.out_dim), "m"), "b0": ((H,), "m"), "w2": ((H, H), "m"), "b2": ((H,), "m")}
    if "audio" in parts:
        C, F2, k = a.hidden_size, a.freq_bins, a.adapter_kernel
        out["audio"] = {"conv1_w": ((C, 9), "m"), "conv1_b": ((C,), "m"), "conv2_w": ((C, 9 * C), "m"),
                        "conv2_b": ((C,), "m"), "sub_out_w": ((C, F2 * C), "m"), "sub_out_b": ((C,), "m"),
                        "embed_w": ((C, C), "m"), "embed_b": ((C,), "m"), "embed_ln_w": ((C,), "gain"),
                        "embed_ln_b": ((C,), "m"), "after_w": ((C,), "gain"), "after_b": ((C,), "m"),
                        "ad_conv_w": ((2 * C, k * C), "m"), "ad_conv_b": ((2 * C,), "m"), "ad_ln_w": ((2 * C,), "gain"),
                        "ad_ln_b": ((2 * C,), "m"), "ad_proj_w": ((H, 2 * C), "m"), "ad_proj_b": ((H,), "m"),
                        "layers": [{"ln1_w": ((C,), "gain"), "ln1_b": ((C,), "m"), "qkv_w": ((3 * C, C), "m"),
                                    "qkv_b": ((3 * C,), "m"), "pos_w": ((C, C), "m"), "bias_u": ((C,), "m"),
                                    "bias_v": ((C,), "m"), "out_w": ((C, C), "m"), "out_b": ((C,), "m"),
                                    "ln2_w": ((C,), "gain"), "ln2_b": ((C,), "m"), "w1": ((a.linear_units, C), "m"),
                                    "b1": ((a.linear_units,), "m"), "w2": ((C, a.linear_units), "m"),
                                    "b2": ((C,), "m")} for _ in range(a.num_blocks)]}
    fill(out, "w")
    if "llm" in out:
        out["llm"]["rope"] = rope_table(default_rope_len(c),
                                        c.head_dim, c.rope_theta).to(device)
        out["llm"]["ep"] = (0, 1) if ep is None else tuple(ep)
    if "audio" in out:
        out["audio"]["cmvn_mean"] = torch.zeros(a.input_dim, device=device)
        out["audio"]["cmvn_istd"] = torch.ones(a.input_dim, device=device)
        out["audio"]["pos_table"] = whale_position_table(a.max_len, a.hidden_size).to(device=device, dtype=BF16)
    return out
