"""vLLM-facing model class on the B200 kernels (reference path B).

Mirrors `MixtralForConditionalGeneration` of the reference's vLLM fork
(web_demo/vllm_tools/vllm_file/mixtral.py:897-1330): `forward` over FLATTENED tokens with the engine's paged KV cache
(:1130-1172), `merge_multimodal_embeddings` (:1084-1128, placeholder ids -> feature rows, in place), `compute_logits`
(:1174-1178), greedy `sample` (:1180-1186) and `load_weights` with the checkpoint -> kernel name map (:1189-1330).

What the engine hands the model per step is captured by `TokenBatch` (the fields of vLLM's flash-attention metadata:
slot mapping, block tables, sequence / query lengths); the engine's block tables and KV tensors are used as they are --
`kv_caches[l]` is vLLM's flash layout [2, num_blocks, block_size, n_kv_heads, head_dim], which is exactly the layout
`vita_gemm_qkv_rope` / `vita_decode_attention` address (page = block).  A step may mix fresh prompts (attention over the
prompt itself, FlashAttention kernel) and single-token decodes (paged decode kernel); chunked prefill / prefix caching
(a prompt continuing cached pages) is not supported and raises.

The reference class was written against vLLM 0.5.5, whose model API no longer exists in the installed 0.22 (probed,
SURVEY.md section 8b).  `register()` registers this class under the reference's architecture name with the installed
`ModelRegistry`; the 0.22 engine drives models through `vllm_config` / forward-context objects that cannot be
exercised without a live engine on a GPU, so the tested contract is the call sequence above
(tests/test_vllm_adapter_gpu.py), not an engine run.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

import torch

from . import ops, weights as W
from .config import VitaConfig
from .model.internvit import InternViTVisionTower, VisionProjector
from .model.mixtral import MixtralDecoder
from .model.whale import AudioEncoder

BF16 = torch.bfloat16


@dataclass
class TokenBatch:
    """One engine step over T flattened tokens of B sequences (vLLM FlashAttentionMetadata, reduced to what is used)."""
    slot_mapping: torch.Tensor          # int32 [T] device: KV slot (block * block_size + offset) of every token
    block_tables: torch.Tensor          # int32 [B, max_blocks] device
    query_start_loc: Sequence[int]      # host, B + 1 entries: tokens of sequence i are rows [qsl[i], qsl[i + 1])
    seq_lens: Sequence[int]             # host, B entries: context length of sequence i including this step's tokens

    @staticmethod
    def from_vllm(md) -> "TokenBatch":
        """Best-effort view of a vLLM attention-metadata object (field names of the flash-attention backend)."""
        qsl = md.query_start_loc.tolist() if torch.is_tensor(md.query_start_loc) else list(md.query_start_loc)
        sl = md.seq_lens.tolist() if torch.is_tensor(md.seq_lens) else list(md.seq_lens)
        bt = getattr(md, "block_table", None)
        if bt is None:
            bt = md.block_tables
        return TokenBatch(md.slot_mapping.to(torch.int32), bt.to(torch.int32), qsl, sl)


class MixtralForConditionalGeneration:
    """Reference path-B surface; weights arrive through `load_weights` (or pre-packed for tests)."""

    def __init__(self, config: VitaConfig, multimodal_config=None, cache_config=None, lora_config=None,
                 quant_config=None, device="cuda", image_token_index: int = 51000, audio_token_index: int = 51001,
                 packed: Optional[dict] = None):
        if lora_config is not None or quant_config is not None:
            raise ValueError("vita_b200 serves bf16 weights without LoRA / quantisation")
        self.config = config
        self.device = torch.device(device)
        self.image_token_index, self.audio_token_index = image_token_index, audio_token_index
        self.block_size = getattr(cache_config, "block_size", 16) if cache_config is not None else 16
        self.packed = None
        self._dec_ws = {}
        if packed is not None:
            self._bind(packed)

    # ------------------------------------------------------------------------------------------ weights
    @staticmethod
    def checkpoint_name(name: str) -> Optional[str]:
        """Name found in a checkpoint / handed over by vLLM's loader -> the shipped checkpoint's own name (what
        vita_b200.weights.pack consumes).  Returns None for tensors the kernels do not use (mixtral.py:1247-1258)."""
        if "rotary_emb.inv_freq" in name:
            return None
        for a, b in (("language_model.lm_head.", "lm_head."), ("language_model.model.", "model.")):
            if name.startswith(a):
                name = b + name[len(a):]
        return name

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]):
        """mixtral.py:1189-1330: consumes (name, tensor) pairs, stacks q|k|v and the experts' w1|w3 / w2 into the
        kernel-native layout.  Returns the set of consumed names."""
        state, used = {}, set()
        for name, t in weights:
            key = self.checkpoint_name(name)
            if key is None:
                continue
            state[key] = t
            used.add(name)
        if not any("global_cmvn" in k for k in state):
            # path B applies CMVN in the feature extractor (processor_whale.py:385-393) and skips these buffers
            # (mixtral.py:1257-1258): identity statistics
            a = self.config.audio
            state[W.PREFIX_AUDIO + "encoder.global_cmvn.mean"] = torch.zeros(a.input_dim)
            state[W.PREFIX_AUDIO + "encoder.global_cmvn.istd"] = torch.ones(a.input_dim)
        self._bind(W.pack(state, self.config, self.device))
        return used

    def _bind(self, packed: dict):
        cfg = self.config
        self.packed = packed
        # the decoder object supplies kernels + workspaces; its own (1-slot) KV cache is unused: the engine owns the KV
        self.llm = MixtralDecoder(cfg.llm, packed["llm"], self.device, max_batch=1, max_seq_len=64, max_new_tokens=8)
        self.vision_tower = InternViTVisionTower(cfg.vision, packed["vision"], self.device) if "vision" in packed else None
        self.vision_projector = VisionProjector(packed["projector"]) if "projector" in packed else None
        self.audio_tower = AudioEncoder(cfg.audio, cfg.llm.hidden_size, packed["audio"], self.device) \
            if "audio" in packed else None

    # ------------------------------------------------------------------------------------------ multimodal
    def _validate_pixel_values(self, data: torch.Tensor) -> torch.Tensor:                       # mixtral.py:964-981
        h = self.config.vision.image_size
        for d in data:
            if tuple(d.shape) != (3, h, h):
                raise ValueError("The expected shape of pixel values per image per batch  per patch is "
                                 f"{(3, h, h)}. You supplied {tuple(d.shape)}.")
        return data

    @torch.no_grad()
    def merge_multimodal_embeddings(self, input_ids: torch.Tensor, input_embeds: torch.Tensor, embeddings,
                                    masks: Optional[torch.Tensor], token_id: int) -> torch.Tensor:
        """mixtral.py:1084-1128: overwrite the rows of `input_embeds` at the positions of `token_id` with the feature
        rows (only the valid ones when a mask is given).  In place, one device row-copy."""
        H = input_embeds.shape[-1]
        dst = (input_ids.reshape(-1) == token_id).nonzero().reshape(-1).to(torch.int32)
        feats = embeddings.reshape(-1, H) if torch.is_tensor(embeddings) else torch.cat(list(embeddings)).reshape(-1, H)
        src = None
        if masks is not None:
            src = masks.reshape(-1).bool().nonzero().reshape(-1).to(torch.int32)
        n = feats.shape[0] if src is None else src.numel()
        if dst.numel() != n:
            raise ValueError(f"Attempted to assign {n} multimodal tokens to {dst.numel()} placeholders")
        if n:
            ops.row_copy(feats.contiguous(), src, dst.to(self.device), input_embeds.view(-1, H), n)
        return input_embeds

    # ------------------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, positions: torch.Tensor, kv_caches: List[torch.Tensor], attn_metadata,
                intermediate_tensors=None, **kwargs) -> torch.Tensor:
        """-> hidden states [T, H] after the final norm (what `compute_logits` consumes), mixtral.py:1130-1172."""
        meta = attn_metadata if isinstance(attn_metadata, TokenBatch) else TokenBatch.from_vllm(attn_metadata)
        dev, H = self.device, self.config.llm.hidden_size
        ids = input_ids.reshape(-1).to(dev)
        T = ids.numel()
        emb = torch.empty(T, H, dtype=BF16, device=dev)
        ops.row_copy(self.packed["llm"]["embed"], ids.to(torch.int32), None, emb, T)       # embed_tokens, :1139
        pixel_values = kwargs.pop("pixel_values", None)
        if pixel_values is not None:                                                        # :1141-1150
            if not isinstance(pixel_values, torch.Tensor):
                raise ValueError(f"Incorrect type of pixel values. Got type: {type(pixel_values)}")
            px = self._validate_pixel_values(pixel_values.reshape(-1, *pixel_values.shape[-3:]))
            feats = self.vision_projector(self.vision_tower(px))
            self.merge_multimodal_embeddings(ids, emb, feats, None, self.image_token_index)
        audio_input = kwargs.pop("audio_input", None)
        if audio_input is not None:                                                         # :1152-1161
            audio_mask = kwargs.pop("audio_mask", None)
            lengths = audio_mask.sum(-1) if audio_mask is not None else \
                torch.full((audio_input.shape[0],), audio_input.shape[1])
            out = self.audio_tower(audio_input, lengths)
            self.merge_multimodal_embeddings(ids, emb, out["inputs_embeds"], out["attention_mask"],
                                             self.audio_token_index)
        return self._language_model(emb, positions.reshape(-1).to(device=dev, dtype=torch.int32), kv_caches, meta)

    __call__ = forward

    def _language_model(self, h: torch.Tensor, positions: torch.Tensor, kv_caches, meta: TokenBatch) -> torch.Tensor:
        llm, c, w = self.llm, self.config.llm, self.packed["llm"]
        T, H = h.shape
        nq, nkv, D, E = c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.num_local_experts
        W_ = c.qkv_rows
        qsl, seq_lens = list(meta.query_start_loc), list(meta.seq_lens)
        assert qsl[-1] == T and len(seq_lens) == len(qsl) - 1
        prefills, decodes = [], []
        for i, sl in enumerate(seq_lens):
            n = qsl[i + 1] - qsl[i]
            if n == sl:
                prefills.append((qsl[i], n))
            elif n == 1:
                decodes.append(i)
            else:
                raise NotImplementedError("chunked prefill / prefix caching: a multi-token query over cached pages")
        if decodes:
            d0 = qsl[decodes[0]]
            assert [qsl[i] for i in decodes] == list(range(d0, d0 + len(decodes))), "decode tokens must be contiguous"
            nd = len(decodes)
            bt_dec = meta.block_tables[decodes].contiguous()
            cur_dec = torch.tensor([seq_lens[i] - 1 for i in decodes], dtype=torch.int32).to(self.device)
            if nd not in self._dec_ws:
                self._dec_ws[nd] = ops.decode_attention_workspace(nd, nkv, llm.decode_splits, self.device)
        ws = llm._ws(max(T, 16))
        xn, qkv, attn, xn2 = ws["xn"][:T], ws["qkv"][:T], ws["attn"][:T], ws["xn2"][:T]
        ids, tw = ws["ids"][:T], ws["tw"][:T]
        perm, rtok, rw = ws["perm"][:2 * T], ws["rtok"][:2 * T], ws["rw"][:2 * T]
        xp, act, yp = ws["xp"][:2 * T], ws["act"][:2 * T], ws["yp"][:2 * T]
        slots = meta.slot_mapping.to(device=self.device, dtype=torch.int32)
        layers = w["layers"]
        ops.rmsnorm(h, layers[0]["ln1"], c.rms_norm_eps, out=xn)
        for li, lw in enumerate(layers):
            kc = kv_caches[li][0].reshape(-1, nkv, D)       # engine-owned pages, vLLM flash layout
            vc = kv_caches[li][1].reshape(-1, nkv, D)
            ops.linear_qkv_rope(xn, lw["wqkv"], qkv, positions, slots, w["rope"], kc, vc, nq, nkv, D)
            for r0, n in prefills:
                q = qkv[r0:r0 + n]
                ops.attention(q, q[:, nq * D:], q[:, (nq + nkv) * D:], attn[r0:r0 + n], (0, W_, D), (0, W_, D),
                              (0, W_, D), (0, nq * D, D), 1, nq, nkv, n, n, D, D, None, True, D ** -0.5)
            if decodes:
                ops.decode_attention(qkv[d0:d0 + nd], kc, vc, bt_dec, cur_dec, attn[d0:d0 + nd], self._dec_ws[nd], nq,
                                     nkv, D, self.block_size, llm.decode_splits, D ** -0.5, q_stride=W_)
            ops.linear(attn, lw["wo"], residual=h, out=h)
            ops.moe_router(h, lw["ln2"], lw["gate"], xn2, ids, tw, c.rms_norm_eps)
            ops.moe_align(ids, tw, ws["offs"], perm, rtok, rw, T, E)
            ops.row_copy(xn2, rtok, None, xp, 2 * T)
            ops.moe_gate_up(xp, lw["w13"], act, ws["offs"], 2 * T)
            ops.moe_down(act, lw["w2"], yp, ws["offs"], rw, 2 * T)
            nxt = layers[li + 1]["ln1"] if li + 1 < len(layers) else w["norm"]
            ops.moe_combine(h, yp, perm, nxt, xn, c.rms_norm_eps)
        return xn.clone()                                   # final RMSNorm(h)

    # ------------------------------------------------------------------------------------------ logits / sampling
    @torch.no_grad()
    def compute_logits(self, hidden_states: torch.Tensor, sampling_metadata=None) -> torch.Tensor:
        """mixtral.py:1174-1178.  `sampling_metadata.selected_token_indices` (if given) picks the rows to score."""
        sel = getattr(sampling_metadata, "selected_token_indices", None)
        if sel is not None:
            hidden_states = hidden_states[sel.to(hidden_states.device)]
        return ops.linear(hidden_states.contiguous(), self.packed["llm"]["lm_head"])

    @torch.no_grad()
    def sample(self, logits: torch.Tensor, sampling_metadata=None) -> torch.Tensor:
        """Greedy next tokens (the reference demos sample with temperature 0.001, web_ability_demo.py:353)."""
        best = torch.zeros(logits.shape[0], dtype=torch.int64, device=logits.device)
        ops.argmax_rows(logits.contiguous(), best)
        return (0xFFFFFFFF - (best & 0xFFFFFFFF)).to(torch.int64)


def register() -> bool:
    """Register the class with the installed vLLM under the reference's architecture name
    (web_demo/vllm_tools/vllm_file/__init__.py:83-84).  Returns False when vLLM is not importable."""
    try:
        from vllm import ModelRegistry
    except Exception:
        return False
    ModelRegistry.register_model("MixtralForConditionalGeneration", "vita_b200.vllm_adapter:MixtralForConditionalGeneration")
    return True
