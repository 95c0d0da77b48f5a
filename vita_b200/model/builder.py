"""`load_pretrained_model` with the reference's signature and return value (vita/model/builder.py:14-24,306).

Checkpoint ingestion streams the HF safetensors shards of `model_path` and packs them into the kernel-native layout
(vita_b200/weights.py).  The tokenizer and the CPU image / audio pre-processors are the reference's own (out of
scope for the kernel tier) and are returned when their files are present next to the checkpoint.

`model_path="synthetic:tiny"` / `"synthetic:full[:L]"` builds seeded random weights instead (tests, demos, bench).
"""
from __future__ import annotations

import json
import os
from pathlib import Path

import torch

from .. import weights as W
from ..config import VitaConfig, LLMConfig, VisionConfig, AudioConfig
from .vita_mixtral import VITAMixtralForCausalLM


def config_from_hf(cfg_json: dict) -> VitaConfig:
    """config.json of the shipped checkpoint (web_demo/vllm_tools/model_weight_file/config.json) -> VitaConfig."""
    t = cfg_json.get("text_config", cfg_json)
    llm = LLMConfig(vocab_size=t.get("vocab_size", 51760), hidden_size=t.get("hidden_size", 4096),
                    intermediate_size=t.get("intermediate_size", 14336),
                    num_hidden_layers=t.get("num_hidden_layers", 32),
                    num_attention_heads=t.get("num_attention_heads", 32),
                    num_key_value_heads=t.get("num_key_value_heads", 8),
                    num_local_experts=t.get("num_local_experts", 8),
                    num_experts_per_tok=t.get("num_experts_per_tok", 2), rms_norm_eps=t.get("rms_norm_eps", 1e-5),
                    rope_theta=t.get("rope_theta", 1e6),
                    max_position_embeddings=t.get("max_position_embeddings", 32768),
                    tokenizer_model_max_length=cfg_json.get("tokenizer_model_max_length", 4600))
    v = cfg_json.get("vision_config", {})
    vision = VisionConfig(hidden_size=v.get("hidden_size", 1024), intermediate_size=v.get("intermediate_size", 4096),
                          num_hidden_layers=v.get("num_hidden_layers", 24),
                          num_attention_heads=v.get("num_attention_heads", 16), image_size=v.get("image_size", 448),
                          patch_size=v.get("patch_size", 14), layer_norm_eps=v.get("layer_norm_eps", 1e-6))
    a = cfg_json.get("audio_config", {})
    audio = AudioConfig(input_dim=a.get("num_mel_bins", 80), hidden_size=a.get("hidden_size", 1024),
                        num_attention_heads=a.get("num_attention_heads", 16),
                        linear_units=a.get("intermediate_size", 4096), num_blocks=a.get("num_hidden_layers", 24))
    return VitaConfig(llm=llm, vision=vision, audio=audio)


class LazySafetensors:
    """Read-only mapping over the `*.safetensors` shards of a checkpoint directory that reads a tensor only when it is
    asked for.  `weights.pack` walks the model layer by layer and moves every tensor to the GPU as soon as it has it,
    so the host never holds more than one layer of the 93.7 GB checkpoint instead of the
    whole state dict.  Uses `model.safetensors.index.json` when present, else the shards' own key lists."""

    def __init__(self, path: Path, prefix: str = ""):
        from safetensors import safe_open
        self._open = safe_open
        self._prefix = prefix
        self._where = {}
        index = path / "model.safetensors.index.json"
        if index.exists():
            for k, shard in json.loads(index.read_text())["weight_map"].items():
                self._where[prefix + k] = path / shard
        else:
            for shard in sorted(path.glob("*.safetensors")):
                with safe_open(str(shard), framework="pt", device="cpu") as f:
                    for k in f.keys():
                        self._where[prefix + k] = shard
        if not self._where:
            raise ValueError(f"no *.safetensors shards under {path}")
        self._handles = {}

    def __contains__(self, key) -> bool:
        return key in self._where

    def __iter__(self):
        return iter(self._where)

    def __len__(self) -> int:
        return len(self._where)

    def keys(self):
        return self._where.keys()

    def __getitem__(self, key):
        shard = self._where[key]                       # KeyError for unknown names, like a dict
        f = self._handles.get(shard)
        if f is None:
            f = self._handles[shard] = self._open(str(shard), framework="pt", device="cpu").__enter__()
        return f.get_tensor(key[len(self._prefix):])

    def close(self):
        for f in self._handles.values():
            f.__exit__(None, None, None)
        self._handles.clear()


class _Overlay:
    """`primary` wins over `base` for the keys it has (the reference's separate vision-tower checkpoint override,
    vita/model/builder.py:245-257)."""

    def __init__(self, primary, base):
        self.primary, self.base = primary, base

    def __contains__(self, key):
        return key in self.primary or key in self.base

    def __getitem__(self, key):
        return self.primary[key] if key in self.primary else self.base[key]

    def __iter__(self):
        yield from self.primary
        yield from (k for k in self.base if k not in self.primary)


def _load_safetensors_dir(path: Path) -> LazySafetensors:
    return LazySafetensors(path)


def load_pretrained_model(model_path, model_base=None, model_name=None, model_type="mixtral-8x7b", load_8bit=False,
                          load_4bit=False, device_map="auto", device="cuda", **kwargs):
    """-> (tokenizer, model, image_processor, context_len), as vita/model/builder.py:306."""
    if model_type not in {"mixtral-8x7b"}:
        raise ValueError(f"Unknown Model Type {model_type}")                      # builder.py:25-26
    if load_8bit or load_4bit:
        raise ValueError("vita_b200 runs bf16 only (bitsandbytes paths are out of scope)")
    dev = torch.device(device if device != "cuda" else "cuda:0")
    model_kwargs = {k: kwargs[k] for k in ("max_batch", "max_seq_len", "max_new_tokens") if k in kwargs}
    if str(model_path).startswith("synthetic:"):
        parts = str(model_path).split(":")
        cfg = VitaConfig.tiny() if parts[1] == "tiny" else VitaConfig.full(int(parts[2]) if len(parts) > 2 else 32)
        seed = int(kwargs.get("seed", 0))
        packed = W.pack(W.synthetic_state(cfg, seed), cfg, dev) if parts[1] == "tiny" else W.random_packed(cfg, dev, seed)
        return None, VITAMixtralForCausalLM(cfg, packed, dev, **model_kwargs), None, cfg.llm.tokenizer_model_max_length
    path = Path(model_path)
    cfg = config_from_hf(json.loads((path / "config.json").read_text()))
    state = _load_safetensors_dir(path)
    vt = kwargs.get("vision_tower_path")
    if vt:                                                                        # builder.py:245-257 override
        state = _Overlay(LazySafetensors(Path(vt), prefix=W.PREFIX_VISION), state)
    model = VITAMixtralForCausalLM(cfg, W.pack(state, cfg, dev), dev, **model_kwargs)
    for s_ in (state.primary, state.base) if isinstance(state, _Overlay) else (state,):
        s_.close()
    tokenizer = image_processor = None
    try:
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(str(path), use_fast=True)
    except Exception:
        pass
    try:
        from transformers import CLIPImageProcessor
        image_processor = CLIPImageProcessor.from_pretrained(str(path))
        model.get_vision_tower().image_processor = image_processor
    except Exception:
        pass
    context_len = getattr(cfg.llm, "max_position_embeddings", 2048)               # builder.py:295-304
    return tokenizer, model, image_processor, context_len
