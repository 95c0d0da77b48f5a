"""Full-depth parity checker: the 32-layer Mixtral of the benchmarked configuration against the fp32 oracle, layer
streamed on the same GPU (plain torch fp32 on cuda, TF32 off; the bf16 weights of one layer at a time are widened to
fp32 in the reference's parameter names and handed to `oracle.vita_oracle.decoder_layer`).

TEST INFRASTRUCTURE: used by tests/test_full_depth_gpu.py and by bench.py's parity record (outside every timed region).
"""
from __future__ import annotations

import torch

from oracle import vita_oracle as O


def layer_state(packed_llm: dict, lcfg, l: int) -> dict:
    """Kernel-native packed layer l -> reference-named fp32 tensors (inverse of vita_b200.weights.pack_llm)."""
    lw = packed_llm["layers"][l]
    nq, nkv, D, I = lcfg.num_attention_heads, lcfg.num_key_value_heads, lcfg.head_dim, lcfg.intermediate_size
    p = f"model.layers.{l}."
    wqkv = lw["wqkv"].float()
    st = {p + "input_layernorm.weight": lw["ln1"].float(), p + "post_attention_layernorm.weight": lw["ln2"].float(),
          p + "self_attn.q_proj.weight": wqkv[: nq * D], p + "self_attn.k_proj.weight": wqkv[nq * D: (nq + nkv) * D],
          p + "self_attn.v_proj.weight": wqkv[(nq + nkv) * D:], p + "self_attn.o_proj.weight": lw["wo"].float(),
          p + "block_sparse_moe.gate.weight": lw["gate"].float()}
    for e in range(lw["w13"].shape[0]):
        q = p + f"block_sparse_moe.experts.{e}."
        w13 = lw["w13"][e].float()
        st[q + "w1.weight"], st[q + "w3.weight"], st[q + "w2.weight"] = w13[:I], w13[I:], lw["w2"][e].float()
    return st


class StreamedOracle:
    """fp32 Mixtral forward over packed bf16 weights, one layer resident at a time, with an fp32 KV cache."""

    def __init__(self, packed_llm: dict, lcfg, device):
        self.w, self.c, self.dev = packed_llm, lcfg, torch.device(device)
        self.past = None
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False

    @torch.no_grad()
    def forward(self, emb: torch.Tensor) -> torch.Tensor:
        """emb [1, S, H] (any float dtype) -> last-row logits [V] fp32; appends to the KV cache."""
        with torch.device(self.dev):
            c = self.c
            past_len = 0 if self.past is None else self.past[0][0].shape[2]
            S = emb.shape[1]
            pos = torch.arange(past_len, past_len + S)[None]
            h = emb.to(self.dev).float()
            new_past = []
            for l in range(c.num_hidden_layers):
                st = layer_state(self.w, c, l)
                h, kv = O.decoder_layer(st, c, l, h, pos, None if self.past is None else self.past[l])
                new_past.append(kv)
                del st
            self.past = new_past
            hn = O.rmsnorm(h[:, -1:], self.w["norm"].float(), c.rms_norm_eps)
            return O.linear(hn, self.w["lm_head"].float())[0, -1]

    def embed(self, tok: int) -> torch.Tensor:
        return self.w["embed"][tok].float()[None, None]


@torch.no_grad()
def check_mixtral(model, emb: torch.Tensor, n_tokens: int = 8, sharpen_router: float = 128.0) -> dict:
    """`model`: vita_b200 VITAMixtralForCausalLM; `emb` [S, H] bf16 spliced prompt embeddings (consumed by value).
    Runs prefill + n_tokens greedy steps on both sides; the oracle is teacher-forced with ITS OWN tokens, the CUDA path
    runs free; reports the first-row logit error, the token agreement and the oracle's margins.

    `sharpen_router`: random-init routers put ~4 % of all (token, layer) decisions within bf16 noise of a rank-2 / rank-3
    tie; a flipped expert replaces that layer's whole MoE output, and over 32 layers x S tokens the bf16 and fp32
    trajectories decorrelate completely (observed: relative logit error 1.1 with every kernel correct).  For the duration
    of the check the router weights of BOTH sides are multiplied by this power of two (exact in bf16, exactly undone
    afterwards): the routing soft-max becomes sharp, the second expert of a near-tied pair carries a negligible weight,
    and the comparison measures the arithmetic instead of the chaos.  1.0 = leave the weights alone."""
    llm = model.llm
    cfg = model.config.llm
    if sharpen_router != 1.0:
        for lw in model.packed["llm"]["layers"]:
            lw["gate"].mul_(sharpen_router)
    try:
        return _check_mixtral(model, emb, n_tokens, sharpen_router)
    finally:
        if sharpen_router != 1.0:
            for lw in model.packed["llm"]["layers"]:
                lw["gate"].div_(sharpen_router)


def _check_mixtral(model, emb, n_tokens, sharpen_router):
    llm = model.llm
    cfg = model.config.llm
    orc = StreamedOracle(model.packed["llm"], cfg, emb.device)
    row = orc.forward(emb[None])
    ref_rows, ref_toks = [row], [int(row.argmax())]
    for _ in range(n_tokens - 1):
        row = orc.forward(orc.embed(ref_toks[-1]))
        ref_rows.append(row)
        ref_toks.append(int(row.argmax()))
    ref_rows = torch.stack(ref_rows)
    # CUDA path: free-running greedy with the logits of every step
    llm.check_capacity(emb.shape[0], n_tokens)
    llm.reset()
    log = llm.enable_score_log()
    first = llm.prefill(emb.clone().contiguous(), slot=0, want_last_logits=True)
    log[0].copy_(first[0])
    for _ in range(n_tokens):
        llm.decode_step(1, use_graph=False, want_logits=True)
    got_toks = llm.generated_tokens(0)[:n_tokens]
    got_rows = log[:n_tokens].float()
    top = ref_rows.topk(2, dim=-1).values
    rel_gap = ((top[:, 0] - top[:, 1]) / top[:, 0].abs().clamp_min(1e-9)).tolist()
    n_same = 0
    while n_same < n_tokens and got_toks[n_same] == ref_toks[n_same]:
        n_same += 1
    # rows are comparable while both sides have consumed the same tokens: row i depends on tokens < i
    cmp = min(n_same + 1, n_tokens)
    err = (got_rows[:cmp] - ref_rows[:cmp]).abs().amax(-1) / ref_rows[:cmp].abs().amax(-1)
    return {"tokens": n_tokens, "router_weights_scaled_by": sharpen_router, "ids_equal_prefix": n_same, "oracle_ids": ref_toks, "cuda_ids": got_toks,
            "first_row_rel_err": float(err[0]), "max_row_rel_err_on_common_prefix": float(err.max()),
            "oracle_top2_rel_gap": rel_gap,
            "first_mismatch_gap": None if n_same == n_tokens else rel_gap[n_same]}
