"""Checkpoint ingestion on the host: the lazy safetensors mapping (vita_b200/model/builder.py) must give `weights.pack`
exactly what an in-memory state dict gives it -- shard layout, index file and the separate vision-tower override
included -- while reading tensors only on demand."""
import json

import pytest
import torch

from vita_b200 import weights as W
from vita_b200.config import VitaConfig


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert a.keys() == b.keys(), path
        for k in a:
            _same(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    elif torch.is_tensor(a):
        assert torch.equal(a, b), path
    else:
        assert a == b, path


@pytest.fixture(scope="module")
def tiny_state():
    cfg = VitaConfig.tiny()
    return cfg, W.synthetic_state(cfg, 0)


@pytest.mark.parametrize("with_index", [False, True])
def test_lazy_shards_pack_like_a_dict(tiny_state, tmp_path, with_index):
    from safetensors.torch import save_file
    from vita_b200.model.builder import LazySafetensors
    cfg, state = tiny_state
    names = sorted(state)
    shards = {"model-00001-of-00003.safetensors": names[0::3], "model-00002-of-00003.safetensors": names[1::3],
              "model-00003-of-00003.safetensors": names[2::3]}
    for fn, keys in shards.items():
        save_file({k: state[k].contiguous() for k in keys}, str(tmp_path / fn))
    if with_index:
        (tmp_path / "model.safetensors.index.json").write_text(
            json.dumps({"weight_map": {k: fn for fn, keys in shards.items() for k in keys}}))
    lazy = LazySafetensors(tmp_path)
    assert len(lazy) == len(state) and set(lazy) == set(state)
    assert "model.norm.weight" in lazy and "no.such.tensor" not in lazy
    with pytest.raises(KeyError):
        lazy["no.such.tensor"]
    assert not lazy._handles, "nothing is opened before a tensor is asked for" if with_index else True
    _same(W.pack(lazy, cfg, "cpu"), W.pack(state, cfg, "cpu"))
    lazy.close()


def test_vision_tower_override_overlay(tiny_state, tmp_path):
    """vita/model/builder.py:245-257: a separate InternViT checkpoint replaces the vision weights of the main one."""
    from safetensors.torch import save_file
    from vita_b200.model.builder import LazySafetensors, _Overlay
    cfg, state = tiny_state
    main, vit = tmp_path / "main", tmp_path / "vit"
    main.mkdir(); vit.mkdir()
    save_file({k: v.contiguous() for k, v in state.items()}, str(main / "model.safetensors"))
    g = torch.Generator().manual_seed(7)
    other = {k[len(W.PREFIX_VISION):]: torch.randn(v.shape, generator=g).to(v.dtype)
             for k, v in state.items() if k.startswith(W.PREFIX_VISION)}
    save_file(other, str(vit / "model.safetensors"))
    merged = _Overlay(LazySafetensors(vit, prefix=W.PREFIX_VISION), LazySafetensors(main))
    want = dict(state)
    want.update({W.PREFIX_VISION + k: v for k, v in other.items()})
    assert set(merged) == set(want)
    _same(W.pack(merged, cfg, "cpu"), W.pack(want, cfg, "cpu"))


def test_vllm_loader_name_map():
    """vita_b200/vllm_adapter.py: names handed over by vLLM's loader -> checkpoint names (mixtral.py:1190-1229)."""
    from vita_b200.vllm_adapter import MixtralForConditionalGeneration as M
    assert M.checkpoint_name("language_model.model.layers.3.self_attn.q_proj.weight") == \
        "model.layers.3.self_attn.q_proj.weight"
    assert M.checkpoint_name("language_model.lm_head.weight") == "lm_head.weight"
    assert M.checkpoint_name("model.layers.0.self_attn.rotary_emb.inv_freq") is None
    assert M.checkpoint_name("model.vision_tower.vision_tower.embeddings.class_embedding") == \
        "model.vision_tower.vision_tower.embeddings.class_embedding"


def test_lora_adapter_is_merged_while_streaming(tiny_state, tmp_path):
    """vita/model/builder.py:51-145 (`lora` in the model name + model_base): base weights + non_lora_trainables.bin +
    the PEFT adapter merged as W + alpha / r * B @ A, tensor by tensor."""
    from safetensors.torch import save_file
    from vita_b200.model.builder import LazySafetensors, LoraMerged, _Overlay, _load_adapter, _non_lora_trainables
    cfg, state = tiny_state
    base, lora = tmp_path / "base", tmp_path / "vita-lora"
    base.mkdir(); lora.mkdir()
    save_file({k: v.contiguous() for k, v in state.items()}, str(base / "model.safetensors"))
    g = torch.Generator().manual_seed(3)
    r, alpha = 4, 8.0
    targets = ["model.layers.0.self_attn.q_proj", "model.layers.1.block_sparse_moe.experts.2.w1", "lm_head"]
    adapter = {}
    for tname in targets:
        out_f, in_f = state[tname + ".weight"].shape
        adapter[f"base_model.model.{tname}.lora_A.weight"] = torch.randn(r, in_f, generator=g).bfloat16()
        adapter[f"base_model.model.{tname}.lora_B.weight"] = torch.randn(out_f, r, generator=g).bfloat16() * 0.1
    save_file(adapter, str(lora / "adapter_model.safetensors"))
    (lora / "adapter_config.json").write_text(json.dumps({"r": r, "lora_alpha": alpha, "target_modules": ["q_proj", "w1"]}))
    extra_name = "model.mm_projector.0.bias"
    extra = {"base_model.model." + extra_name: torch.randn(state[extra_name].shape, generator=g).bfloat16()}
    torch.save(extra, str(lora / "non_lora_trainables.bin"))

    tensors, a, rr, rs = _load_adapter(lora)
    merged = LoraMerged(_Overlay(_non_lora_trainables(lora), LazySafetensors(base)), tensors, a, rr, rs)
    want = dict(state)
    want[extra_name] = extra["base_model.model." + extra_name]
    for tname in targets:
        A, B = adapter[f"base_model.model.{tname}.lora_A.weight"], adapter[f"base_model.model.{tname}.lora_B.weight"]
        want[tname + ".weight"] = (state[tname + ".weight"].float() + alpha / r * (B.float() @ A.float())).bfloat16()
    assert set(merged) == set(want)
    assert not torch.equal(merged["lm_head.weight"], state["lm_head.weight"])
    _same(W.pack(merged, cfg, "cpu"), W.pack(want, cfg, "cpu"))
