"""Mixtral-8x7B sparse-MoE decoder on the B200 kernels: paged KV cache, prefill, CUDA-graphed greedy decode.

Replaces the third-party arithmetic the reference reaches through `self.model(...)` / `lm_head`
(vita/model/language_model/vita_mixtral.py:158-173 -> transformers MixtralModel; vLLM twin
web_demo/vllm_tools/vllm_file/mixtral.py:375-628).  Host code only sequences kernels; no arithmetic in torch.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from .. import ops
from ..config import LLMConfig

BF16 = torch.bfloat16


class PagedKVCache:
    """k/v per layer: [num_pages * page_size, n_kv_heads, head_dim] bf16; block_table [max_batch, pages_per_seq]."""

    def __init__(self, cfg: LLMConfig, max_batch: int, max_seq_len: int, device, page_size: int = 16,
                 shuffle_pages: bool = False, storage=None):
        """`storage(layer, "k" | "v", shape) -> tensor` lets the owner place the pages (e.g. in NVLink symmetric memory
        for the sequence-sharded expert-parallel prefill); default: ordinary device allocations."""
        self.page_size = page_size
        self.pages_per_seq = (max_seq_len + page_size - 1) // page_size
        self.max_seq_len = self.pages_per_seq * page_size
        self.max_batch = max_batch
        n_pages = max_batch * self.pages_per_seq
        slots = n_pages * page_size
        shape = (slots, cfg.num_key_value_heads, cfg.head_dim)
        if storage is None:
            storage = lambda layer, which, shp: torch.zeros(shp, dtype=BF16, device=device)
        self.k = [storage(l, "k", shape) for l in range(cfg.num_hidden_layers)]
        self.v = [storage(l, "v", shape) for l in range(cfg.num_hidden_layers)]
        order = torch.randperm(n_pages, generator=torch.Generator().manual_seed(0)) if shuffle_pages \
            else torch.arange(n_pages)
        table = order.view(max_batch, self.pages_per_seq).to(torch.int32)
        self.block_table = table.to(device)
        pos = torch.arange(self.max_seq_len)
        self.slot_map = (table[:, pos // page_size].to(torch.int64) * page_size + (pos % page_size)[None, :]) \
            .to(torch.int32).to(device)                                      # [max_batch, max_seq_len]
        self.cache_len = torch.zeros(max_batch, dtype=torch.int32, device=device)
        self.cur_pos = torch.zeros(max_batch, dtype=torch.int32, device=device)

    def reset(self):
        self.cache_len.zero_()
        self.cur_pos.zero_()


class MixtralDecoder:
    def __init__(self, cfg: LLMConfig, weights: dict, device, max_batch: int = 1, max_seq_len: Optional[int] = None,
                 max_new_tokens: int = 1024, page_size: int = 16, decode_splits: int = 16, shuffle_pages: bool = False):
        self.cfg = cfg
        self.w = weights
        self.device = torch.device(device)
        self.max_batch = max_batch
        max_seq_len = max_seq_len or (cfg.tokenizer_model_max_length + max_new_tokens)
        if max_seq_len > cfg.max_position_embeddings:
            raise ValueError(f"max_seq_len {max_seq_len} exceeds max_position_embeddings {cfg.max_position_embeddings}")
        page_rounded = (max_seq_len + page_size - 1) // page_size * page_size
        if page_rounded > weights["rope"].shape[0]:      # cos/sin table sized from the KV capacity, not a constant
            from ..weights import rope_table
            weights["rope"] = rope_table(page_rounded, cfg.head_dim, cfg.rope_theta).to(self.device)
        # expert parallelism: this rank holds experts [e_lo, e_hi) of every layer.  Modes (VITA_B200_EP):
        #   seq  (default) sequence-sharded residual stream: every rank owns S/N tokens for the dense part (qkv, causal
        #        attention over the all-gathered K/V, o-proj, router); K/V rows and routed activations are all-gathered
        #        and the expert outputs pushed to the token owners by P2P stores over NVLink symmetric memory
        #   p2p  replicated dense part, fused P2P combine (round-1 design, kept as the A/B baseline)
        #   nccl replicated dense part, partial sums + NCCL all-reduce (library baseline)
        self.ep_rank, self.ep_world = weights.get("ep", (0, 1))
        self.ep_mode = os.environ.get("VITA_B200_EP", "seq") if self.ep_world > 1 else None
        assert self.ep_mode in (None, "seq", "p2p", "nccl"), "VITA_B200_EP must be seq, p2p or nccl"
        self.ep_p2p = None
        storage = None
        if self.ep_mode == "seq":
            assert max_batch == 1 and not shuffle_pages, "sequence-sharded EP serves one sequence with in-order pages"
            storage = self._init_ep_seq(page_rounded)
        self.cache = PagedKVCache(cfg, max_batch, max_seq_len, device, page_size, shuffle_pages, storage)
        self.max_new_tokens = max_new_tokens
        self.decode_splits = int(os.environ.get("VITA_B200_ATTN_SPLITS", decode_splits))
        H, I = cfg.hidden_size, cfg.intermediate_size
        B = max_batch
        dev = self.device
        # decode-step state (all device resident: a step needs no host input, so it can be replayed as a CUDA graph)
        self.best = torch.zeros(B, dtype=torch.int64, device=dev)         # packed (logit, ~index) arg-max
        self.token_log = torch.zeros(B, max_new_tokens + 8, dtype=torch.int32, device=dev)
        self.gen_count = torch.zeros(B, dtype=torch.int32, device=dev)
        self.d_h = torch.zeros(B, H, dtype=BF16, device=dev)
        self.d_q = torch.zeros(B, cfg.num_attention_heads * cfg.head_dim, dtype=BF16, device=dev)
        self.d_attn = torch.zeros_like(self.d_q)
        self.d_xn = torch.zeros(B, H, dtype=BF16, device=dev)
        self.d_ids = torch.zeros(B, 2, dtype=torch.int32, device=dev)
        self.d_w = torch.zeros(B, 2, dtype=torch.float32, device=dev)
        self.d_route = torch.zeros(B, dtype=torch.int64, device=dev)      # early routing hand-over gate|up -> down
        self.d_act = torch.zeros(B, 2, I, dtype=BF16, device=dev)
        self.d_logits = torch.zeros(B, cfg.vocab_size, dtype=BF16, device=dev)
        self.attn_ws = ops.decode_attention_workspace(B, cfg.num_key_value_heads, self.decode_splits, dev)
        # decode linears: tcgen05 swap-AB GEMVs (decode_tc.cu); their stream-K partial sums meet in one workspace
        if H % 64 or I % 128:
            raise ValueError("the decode GEMVs need hidden_size % 64 == 0 and intermediate_size % 128 == 0")
        max_rb = max((cfg.vocab_size + 127) // 128, 2 * (I // 128), cfg.num_attention_heads
                     + 2 * cfg.num_key_value_heads, (H + 127) // 128)
        self.tc_ws = ops.TcWorkspace(B, max_rb, dev)
        # completion counters of the bs = 1 decode chain: [serial, one counter per chain kernel (5 per layer + LM head)]
        self.chain_mem = torch.zeros(2 + 5 * cfg.num_hidden_layers, dtype=torch.int64, device=dev)
        self._graphs = {}             # single-sequence decode step: captured CUDA graph per (B, want_logits)
        self.scores_buf = None        # [max_new_tokens + 1, V] logits log of slot 0 (generate(output_scores=True))
        self._bgraphs = {}            # batched decode step: captured CUDA graph per batch size
        self.d_slots = torch.zeros(B, dtype=torch.int32, device=dev)
        self._prefill_ws = {}
        # router decisions of every layer (the reference returns them as `router_logits` when asked,
        # vita_mixtral.py:108,185-190): set to a list to have prefill() and the eager decode step append one
        # (top-2 ids [T, 2] int32, renormalised weights [T, 2] fp32) pair per layer
        self.route_trace: Optional[list] = None
        # decode: the gate|up kernel can publish the expert pair as soon as its router is done so that the down projection
        # streams its weight rows while gate|up is still running.  Bit-identical, but two same-box A/Bs measured it 0.7 %
        # SLOWER (5.12 vs 5.08 and 5.27 vs 5.23 ms/token, profiles/r02_decode_ab.txt): off unless VITA_B200_EARLY_ROUTE=1
        self.early_route = os.environ.get("VITA_B200_EARLY_ROUTE", "0") == "1"
        # prefill / batched decode: fused top-2 router + token permute over per-expert slot ranges (0 = align + gather)
        self.moe_fused = os.environ.get("VITA_B200_MOE_FUSED", "1") == "1"
        per = cfg.num_local_experts // self.ep_world
        self.e_lo, self.e_hi = self.ep_rank * per, (self.ep_rank + 1) * per
        assert weights["layers"][0]["w13"].shape[0] == per, "expert tensors do not match the EP layout"
        if self.ep_mode == "p2p":
            self._init_ep_p2p()

    def _init_ep_seq(self, S_max: int):
        """One NVLink symmetric allocation for the sequence-sharded expert-parallel prefill: receive slots of the
        combine, residual stream / normed rows (own tokens), the all-gathered post-attention activations with their
        routing records, epoch flags, and the K/V pages of every layer.  Returns the page-storage callback."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        c, N, dev = self.cfg, self.ep_world, self.device
        H, L = c.hidden_size, c.num_hidden_layers
        kv_row = c.num_key_value_heads * c.head_dim * 2                       # bytes of one K (or V) row
        chunk_max = self.ep_chunk(S_max)
        al = lambda n: (n + 255) // 256 * 256
        names = ["rs", "h", "xn", "xn2", "ids", "tw", "flags", "kv"]
        sizes = [al(chunk_max * 2 * H * 2), al(S_max * H * 2), al(S_max * H * 2), al(S_max * H * 2), al(S_max * 8),
                 al(S_max * 8), al(4 * N * 4), 2 * L * al(S_max * kv_row)]
        offs = [0]
        for z in sizes:
            offs.append(offs[-1] + z)
        sym = symm_mem.empty(offs[-1], dtype=torch.uint8, device=dev)
        sym.zero_()
        torch.cuda.synchronize()
        hdl = symm_mem.rendezvous(sym, dist.group.WORLD)
        delta = sym.data_ptr() - int(hdl.buffer_ptrs[hdl.rank])
        bases = [int(p) + delta for p in hdl.buffer_ptrs]
        assert hdl.world_size == N and hdl.rank == self.ep_rank
        off = dict(zip(names, offs))
        view = lambda n, dt: sym[off[n]:off[n] + sizes[names.index(n)]].view(dt)
        ptrs = lambda n: torch.tensor([b + off[n] for b in bases], dtype=torch.int64, device=dev)
        kv_layer = al(S_max * kv_row)
        self.ep_p2p = dict(
            sym=sym, hdl=hdl, chunk_max=chunk_max, off=off, kv_layer=kv_layer, kv_row=kv_row,
            base_ptrs=torch.tensor(bases, dtype=torch.int64, device=dev),
            rs=view("rs", BF16), h=view("h", BF16)[:S_max * H].view(S_max, H),
            xn=view("xn", BF16)[:S_max * H].view(S_max, H), xn2=view("xn2", BF16)[:S_max * H].view(S_max, H),
            ids=view("ids", torch.int32)[:S_max * 2].view(S_max, 2),      # (the regions are padded to 256 bytes)
            tw=view("tw", torch.float32)[:S_max * 2].view(S_max, 2), flags=view("flags", torch.int32),
            rs_ptrs=ptrs("rs"), h_ptrs=ptrs("h"), xn_ptrs=ptrs("xn"), flag_ptrs=ptrs("flags"), epoch=0)
        dist.barrier()

        def storage(layer, which, shape):
            assert shape[0] == S_max, "K/V pages of the sequence-sharded prefill: one sequence, in-order pages"
            o = off["kv"] + (2 * layer + (which == "v")) * kv_layer
            return sym[o:o + S_max * kv_row].view(BF16).view(shape)
        return storage

    def ep_chunk(self, S: int) -> int:
        """Tokens per rank of the sequence-sharded stream (multiple of 8: 16-byte aligned routing records)."""
        from ..parallel import sequence_chunk
        return sequence_chunk(S, self.ep_world)

    def _init_ep_p2p(self):
        """Symmetric (peer-mapped) buffers for the fused expert-parallel combine: every rank can store into every
        other rank's receive slots, residual stream and normed activations over NVLink."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        c, N, dev = self.cfg, self.ep_world, self.device
        H, S_max = c.hidden_size, self.cache.max_seq_len
        chunk_max = (S_max + N - 1) // N
        al = lambda n: (n + 255) // 256 * 256
        sizes = [al(chunk_max * 2 * H * 2), al(S_max * H * 2), al(S_max * H * 2), al(2 * N * 4)]
        offs = [0]
        for z in sizes:
            offs.append(offs[-1] + z)
        sym = symm_mem.empty(offs[-1], dtype=torch.uint8, device=dev)
        sym.zero_()
        torch.cuda.synchronize()
        hdl = symm_mem.rendezvous(sym, dist.group.WORLD)
        delta = sym.data_ptr() - int(hdl.buffer_ptrs[hdl.rank])      # offset of this tensor inside the allocation
        bases = [int(p) + delta for p in hdl.buffer_ptrs]
        assert hdl.world_size == N and hdl.rank == self.ep_rank
        view = lambda i, dt: sym[offs[i]:offs[i] + sizes[i]].view(dt)
        ptrs = lambda i: torch.tensor([b + offs[i] for b in bases], dtype=torch.int64, device=dev)
        self.ep_p2p = dict(
            sym=sym, hdl=hdl, chunk_max=chunk_max,
            rs=view(0, BF16), h=view(1, BF16)[:S_max * H].view(S_max, H), xn=view(2, BF16)[:S_max * H].view(S_max, H),
            flags=view(3, torch.int32), rs_ptrs=ptrs(0), h_ptrs=ptrs(1), xn_ptrs=ptrs(2), flag_ptrs=ptrs(3), epoch=0)
        dist.barrier()

    # ------------------------------------------------------------------------------------------ prefill
    def _ws(self, S: int):
        """Prefill workspaces, grown geometrically and reused."""
        cap = self._prefill_ws.get("cap", 0)
        if S > cap:
            c = self.cfg
            cap = max(S, 2 * cap, 128)
            self._bgraphs = {}    # the batched-step graphs reference the old workspaces
            H, I, E, dev = c.hidden_size, c.intermediate_size, c.num_local_experts, self.device
            with torch.inference_mode(False):      # reused across calls: keep them ordinary tensors
                self._prefill_ws = self._alloc_ws(cap, H, I, E, dev)
        return self._prefill_ws

    def _alloc_ws(self, cap, H, I, E, dev):
        c = self.cfg
        e = lambda *shape, dt=BF16: torch.empty(*shape, dtype=dt, device=dev)
        # MoE activations in the "slot" layout of the fused router + permute: expert e owns rows [e * cap, (e + 1) * cap)
        # (any expert may receive every token).  The compact buffers of the align / gather flow (expert parallelism, the
        # vLLM-shaped adapter) are the first 2 * cap rows of the same storage.
        xs, acts, ys, rws = e(E * cap, H), e(E * cap, I), e(E * cap, H), e(E * cap, dt=torch.float32)
        return dict(
            cap=cap, xn=e(cap, H), qkv=e(cap, c.qkv_rows), attn=e(cap, c.num_attention_heads * c.head_dim),
            xn2=e(cap, H), ids=e(cap, 2, dt=torch.int32), tw=e(cap, 2, dt=torch.float32),
            offs=e(E + 1, dt=torch.int32), perm=e(cap * 2, dt=torch.int32), rtok=e(cap * 2, dt=torch.int32),
            xs=xs, acts=acts, ys=ys, rws=rws,
            cnt=torch.zeros(c.num_hidden_layers, E, dtype=torch.int32, device=dev),
            rw=rws[:cap * 2], xp=xs[:cap * 2], act=acts[:cap * 2], yp=ys[:cap * 2],
            ybuf=e(cap, H) if self.ep_world > 1 else None, rassign=e(cap * 2, dt=torch.int32),
            pos=torch.arange(cap, dtype=torch.int32, device=dev))

    @torch.no_grad()
    def prefill(self, inputs_embeds: torch.Tensor, slot: int = 0, all_logits: bool = False,
                want_last_logits: bool = False):
        """One sequence: inputs_embeds [S, H] bf16 (consumed as the residual stream, modified in place).

        Appends S tokens to the KV cache of batch slot `slot` (which must be empty), leaves the arg-max of the last
        position in self.best[slot] (the first generated token) and returns logits [S, V] if `all_logits`."""
        c, w = self.cfg, self.w
        S, H = inputs_embeds.shape
        assert inputs_embeds.dtype == BF16 and inputs_embeds.is_cuda and inputs_embeds.is_contiguous()
        if S > self.cache.max_seq_len:
            raise ValueError(f"prompt of {S} tokens exceeds the KV capacity of {self.cache.max_seq_len} per sequence")
        if self.ep_mode == "seq":
            return self._prefill_ep_seq(inputs_embeds, slot, all_logits, want_last_logits)
        ws = self._ws(S)
        h = inputs_embeds
        p2p = self.ep_p2p
        if p2p is not None:          # the residual stream and the normed activations live in symmetric memory
            h = p2p["h"][:S]
            h.copy_(inputs_embeds)
        nq, nkv, D, E = c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.num_local_experts
        xn, qkv, attn, xn2 = ws["xn"][:S], ws["qkv"][:S], ws["attn"][:S], ws["xn2"][:S]
        if p2p is not None:
            xn = p2p["xn"][:S]
        ids, tw = ws["ids"][:S], ws["tw"][:S]
        perm, rtok, rw = ws["perm"][:2 * S], ws["rtok"][:2 * S], ws["rw"][:2 * S]
        xp, act, yp = ws["xp"][:2 * S], ws["act"][:2 * S], ws["yp"][:2 * S]
        pos = ws["pos"][:S]
        slots = self.cache.slot_map[slot, :S]
        W = c.qkv_rows
        layers = w["layers"]
        fused = self.ep_world == 1 and self.moe_fused
        if fused:
            ws["cnt"].zero_()          # per-layer expert counters of the fused router + permute
        ops.rmsnorm(h, layers[0]["ln1"], c.rms_norm_eps, out=xn)
        for li, lw in enumerate(layers):
            # qkv projection, RoPE and the KV append in one kernel (GEMM epilogue)
            ops.linear_qkv_rope(xn, lw["wqkv"], qkv, pos, slots, w["rope"], self.cache.k[li], self.cache.v[li], nq, nkv, D)
            ops.attention(qkv, qkv[:, nq * D:], qkv[:, (nq + nkv) * D:], attn, (0, W, D), (0, W, D), (0, W, D),
                          (0, nq * D, D), 1, nq, nkv, S, S, D, D, None, True, D ** -0.5)
            ops.linear(attn, lw["wo"], residual=h, out=h)
            nxt = layers[li + 1]["ln1"] if li + 1 < len(layers) else (w["norm"] if all_logits else None)
            if fused:
                # router + permute in one kernel (slot layout), grouped GEMMs over the slots, gather-combine
                cnt = ws["cnt"][li]
                ops.moe_route_scatter(h, lw["ln2"], lw["gate"], ws["xs"], cnt, perm, ws["rws"], c.rms_norm_eps, ids, tw)
                if self.route_trace is not None:
                    self.route_trace.append((ids.clone(), tw.clone()))
                ops.moe_gate_up_slots(ws["xs"], lw["w13"], ws["acts"], cnt, 2 * S)
                ops.moe_down_slots(ws["acts"], lw["w2"], ws["ys"], cnt, ws["rws"], 2 * S)
                ops.moe_combine(h, ws["ys"], perm, nxt, xn if nxt is not None else None, c.rms_norm_eps)
                continue
            ops.moe_router(h, lw["ln2"], lw["gate"], xn2, ids, tw, c.rms_norm_eps)
            if self.route_trace is not None:
                self.route_trace.append((ids.clone(), tw.clone()))
            ops.moe_align(ids, tw, ws["offs"], perm, rtok, rw, S, E, row_assign=ws["rassign"][:2 * S])
            ops.row_copy(xn2, rtok, None, xp, 2 * S)
            if p2p is not None:
                # fused expert-parallel combine over NVLink peer memory: the down-projection epilogue pushes every
                # (token, k) row to the token's owner; owners reduce + norm + all-gather by P2P stores
                N, r = self.ep_world, self.ep_rank
                chunk = (S + N - 1) // N
                offs_local = ws["offs"][self.e_lo:]
                p2p["epoch"] += 1
                ep = p2p["epoch"]
                ops.moe_gate_up(xp, lw["w13"], act, offs_local, 2 * S)
                ops.moe_down_ep(act, lw["w2"], offs_local, rw, ws["rassign"][:2 * S], p2p["rs_ptrs"], 2 * S, chunk)
                ops.ep_signal(p2p["flag_ptrs"], 0, N, r, ep)
                n_owned = max(0, min(chunk, S - r * chunk))
                ops.ep_reduce_norm_gather(p2p["rs"], p2p["flags"], p2p["h_ptrs"], p2p["xn_ptrs"], nxt, r * chunk,
                                          n_owned, N, r, ep, H, c.rms_norm_eps)
                ops.ep_signal(p2p["flag_ptrs"], 1, N, r, ep)
                ops.ep_wait(p2p["flags"], 1, N, ep)
            elif self.ep_world == 1:
                ops.moe_gate_up(xp, lw["w13"], act, ws["offs"], 2 * S)
                ops.moe_down(act, lw["w2"], yp, ws["offs"], rw, 2 * S)
                ops.moe_combine(h, yp, perm, nxt, xn if nxt is not None else None, c.rms_norm_eps)
            else:
                # local experts only: the grouped GEMMs walk offs[e_lo : e_hi + 1]; rows of remote experts stay zero
                import torch.distributed as dist
                offs_local = ws["offs"][self.e_lo:]
                yp.zero_()
                ops.moe_gate_up(xp, lw["w13"], act, offs_local, 2 * S)
                ops.moe_down(act, lw["w2"], yp, offs_local, rw, 2 * S)
                ybuf = ws["ybuf"][:S]
                ybuf.zero_()
                ops.moe_combine(ybuf, yp, perm, None, None, c.rms_norm_eps)       # partial sum of the local experts
                dist.all_reduce(ybuf)                                              # NCCL over NVLink (sum, bf16)
                ops.add_rmsnorm(h, ybuf, nxt, xn if nxt is not None else None, c.rms_norm_eps)
        self.cache.cache_len[slot:slot + 1] += S
        # first generated token: final norm + lm_head + arg-max on the last row only
        self.best[slot:slot + 1].zero_()
        last_logits = self.d_logits[slot:slot + 1] if (want_last_logits or all_logits) else None
        ops.tc_lm_head_argmax(h[S - 1:], H, w["norm"], w["lm_head"], last_logits, self.best[slot:slot + 1], 1,
                              self.tc_ws, c.rms_norm_eps)
        if all_logits:
            return ops.linear(xn, w["lm_head"])      # [S, V]; xn = final RMSNorm(h) written by the last combine
        return last_logits

    @torch.no_grad()
    def _prefill_ep_seq(self, inputs_embeds, slot, all_logits, want_last_logits):
        """Sequence-sharded expert-parallel prefill (BASELINE configs[3]).  Rank r owns tokens
        [t0, t1) for everything that is per-token (norms, qkv, o-proj, router, residual stream) and for their causal
        attention over keys [0, t1); per layer three exchanges over NVLink peer memory, all by direct stores:
          1. K/V rows of the own tokens -> every rank's pages (all-gather; a rank waits only for the ranks before it)
          2. post-attention normed rows + routing records -> every rank (all-gather), so every rank can run the
             grouped GEMMs of its local experts over the rows routed to them
          3. the down-projection epilogue pushes each (token, k) output row to the token's owner (fused combine);
             the owner sums its two slots into the residual stream and applies the next RMSNorm.
        No partial sums anywhere: logits are bit-identical to the single-GPU model.  Returns the logits of the OWN rows
        when `all_logits`, the last-row logits on the rank that owns the last token otherwise (None elsewhere)."""
        assert slot == 0
        c, w, sy = self.cfg, self.w, self.ep_p2p
        S, H = inputs_embeds.shape
        N, r = self.ep_world, self.ep_rank
        from ..parallel import token_range
        chunk = self.ep_chunk(S)
        t0, t1 = token_range(S, r, N)
        n = t1 - t0
        ws = self._ws(S)
        nq, nkv, D, E = c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.num_local_experts
        W = c.qkv_rows
        h_own, xn_own = sy["h"][t0:t1], sy["xn"][t0:t1]
        xn2, ids, tw = sy["xn2"][:S], sy["ids"][:S], sy["tw"][:S]
        qkv, attn = ws["qkv"][:n], ws["attn"][:n]
        perm, rtok, rw = ws["perm"][:2 * S], ws["rtok"][:2 * S], ws["rw"][:2 * S]
        xp, act = ws["xp"][:2 * S], ws["act"][:2 * S]
        pos, slots = ws["pos"][t0:t1], self.cache.slot_map[0, t0:t1]
        off, kvl, kvr = sy["off"], sy["kv_layer"], sy["kv_row"]
        offs_local = ws["offs"][self.e_lo:]
        layers = w["layers"]
        if n:
            h_own.copy_(inputs_embeds[t0:t1])
            ops.rmsnorm(h_own, layers[0]["ln1"], c.rms_norm_eps, out=xn_own)
        for li, lw in enumerate(layers):
            sy["epoch"] += 1
            ep = sy["epoch"]
            if n:
                ops.linear_qkv_rope(xn_own, lw["wqkv"], qkv, pos, slots, w["rope"], self.cache.k[li],
                                    self.cache.v[li], nq, nkv, D)
                ko = off["kv"] + 2 * li * kvl + t0 * kvr
                ops.ep_push(sy["base_ptrs"], [(ko, n * kvr), (ko + kvl, n * kvr)], N, r)
            ops.ep_signal(sy["flag_ptrs"], 2, N, r, ep)
            ops.ep_wait(sy["flags"], 2, N, ep, n_wait=r)          # K/V rows of the ranks before this one
            if n:
                ops.attention(qkv, self.cache.k[li], self.cache.v[li], attn, (0, W, D), (0, nkv * D, D),
                              (0, nkv * D, D), (0, nq * D, D), 1, nq, nkv, n, t1, D, D, None, True, D ** -0.5,
                              q_pos0=t0)
                ops.linear(attn, lw["wo"], residual=h_own, out=h_own)
                ops.moe_router(h_own, lw["ln2"], lw["gate"], xn2[t0:t1], ids[t0:t1], tw[t0:t1], c.rms_norm_eps)
                ops.ep_push(sy["base_ptrs"], [(off["xn2"] + t0 * H * 2, n * H * 2), (off["ids"] + t0 * 8, n * 8),
                                              (off["tw"] + t0 * 8, n * 8)], N, r)
            ops.ep_signal(sy["flag_ptrs"], 3, N, r, ep)
            ops.ep_wait(sy["flags"], 3, N, ep)
            ops.moe_align(ids, tw, ws["offs"], perm, rtok, rw, S, E, row_assign=ws["rassign"][:2 * S])
            ops.row_copy(xn2, rtok, None, xp, 2 * S)
            ops.moe_gate_up(xp, lw["w13"], act, offs_local, 2 * S)
            ops.moe_down_ep(act, lw["w2"], offs_local, rw, ws["rassign"][:2 * S], sy["rs_ptrs"], 2 * S, chunk)
            ops.ep_signal(sy["flag_ptrs"], 0, N, r, ep)
            nxt = layers[li + 1]["ln1"] if li + 1 < len(layers) else (w["norm"] if all_logits else None)
            # (waits for flags[0] of every rank: all expert rows of the own tokens have landed)
            ops.ep_reduce_norm_gather(sy["rs"], sy["flags"], sy["h_ptrs"], sy["xn_ptrs"], nxt, t0, n, N, r, ep, H,
                                      c.rms_norm_eps, gather=False)
            if n == 0:
                ops.ep_wait(sy["flags"], 0, N, ep)   # keep the epochs of a token-less rank in step with the others
        self.cache.cache_len[slot:slot + 1] += S
        self.best[slot:slot + 1].zero_()
        last_logits = None
        if t0 <= S - 1 < t1:     # this rank owns the last token: first generated token
            last_logits = self.d_logits[slot:slot + 1] if (want_last_logits or all_logits) else None
            ops.tc_lm_head_argmax(sy["h"][S - 1:S], H, w["norm"], w["lm_head"], last_logits,
                                  self.best[slot:slot + 1], 1, self.tc_ws, c.rms_norm_eps)
        if all_logits:
            return ops.linear(xn_own, w["lm_head"]) if n else torch.empty(0, c.vocab_size, dtype=BF16, device=self.device)
        return last_logits

    # ------------------------------------------------------------------------------------------ decode
    def _decode_step_kernels(self, B: int, want_logits: bool):
        """One greedy token for slots [0, B): embed + per layer (qkv+RoPE+KV append -> paged attention -> o-proj ->
        router+gate/up -> down) + LM head with arg-max, 2 + 5 L launches chained by programmatic dependent launch.
        For B = 1 the kernels are additionally linked by completion counters (include/vita_b200.h, vita_chain_begin)."""
        c, w, cache = self.cfg, self.w, self.cache
        nq, nkv, D = c.num_attention_heads, c.num_key_value_heads, c.head_dim
        h = self.d_h[:B]
        ws = self.tc_ws
        chained = B == 1
        ops.decode_embed(self.best[:B], self.token_log[:B], self.gen_count[:B], cache.cache_len[:B], cache.cur_pos[:B],
                         w["embed"], h, cache.max_seq_len, self.chain_mem if chained else None)
        if chained:
            ops.chain_begin(self.chain_mem)
        try:
            for li, lw in enumerate(w["layers"]):
                ops.decode_tc_qkv_rope(h, lw["ln1"], lw["wqkv"], w["rope"], cache.cur_pos[:B], cache.block_table[:B],
                                       self.d_q[:B], cache.k[li], cache.v[li], ws, nq, nkv, D, cache.page_size,
                                       c.rms_norm_eps)
                ops.decode_attention(self.d_q[:B], cache.k[li], cache.v[li], cache.block_table[:B], cache.cur_pos[:B],
                                     self.d_attn[:B], self.attn_ws, nq, nkv, D, cache.page_size, self.decode_splits,
                                     D ** -0.5)
                ops.decode_tc_oproj(self.d_attn[:B], lw["wo"], h, ws)
                tracing = self.route_trace is not None and not torch.cuda.is_current_stream_capturing()
                route = None if (tracing or not self.early_route) else self.d_route[:B]
                ops.decode_tc_moe_gate_up(h, lw["ln2"], lw["gate"], lw["w13"], self.d_ids[:B], self.d_w[:B],
                                          self.d_act[:B], ws, c.rms_norm_eps, route, li + 1)
                if tracing:
                    self.route_trace.append((self.d_ids[:B].clone(), self.d_w[:B].clone()))
                ops.decode_tc_moe_down(self.d_act[:B], lw["w2"], self.d_ids[:B], self.d_w[:B], h, ws, route, li + 1)
            lg = self.d_logits[:B] if want_logits else None
            ops.tc_lm_head_argmax(h, c.hidden_size, w["norm"], w["lm_head"], lg, self.best[:B], B, ws, c.rms_norm_eps)
        finally:
            if chained:
                ops.chain_end()
        self._log_scores(B, want_logits)

    def enable_score_log(self):
        """Allocate the per-step logits log (HF `output_scores=True`): row i holds the logits token i was chosen from.
        The decode step appends to it on the device (row index = gen_count), so it works under CUDA-graph replay."""
        if self.scores_buf is None:
            with torch.inference_mode(False):   # persistent state must not become an inference tensor when the first
                #                                 call happens under torch.inference_mode() (video_audio_demo.py:256)
                self.scores_buf = torch.zeros(self.max_new_tokens + 1, self.cfg.vocab_size, dtype=BF16,
                                              device=self.device)
            self._graphs = {k: g for k, g in self._graphs.items() if not k[1]}   # re-capture the logits variants
        return self.scores_buf

    def _log_scores(self, B: int, want_logits: bool):
        if want_logits and B == 1 and self.scores_buf is not None:
            ops.row_copy(self.d_logits[:1], None, self.gen_count[:1], self.scores_buf, 1)

    @property
    def launches_per_decode_step(self) -> int:
        return 2 + 5 * self.cfg.num_hidden_layers

    def launches_per_decode_step_with_scores(self) -> int:
        return self.launches_per_decode_step + 1

    @torch.no_grad()
    def decode_step(self, B: int = 1, use_graph: bool = True, want_logits: bool = False):
        assert self.ep_world == 1, "expert-parallel mode covers the prefill (BASELINE configs[3]); decode is replicated"
        self._decode_step(B, use_graph, want_logits)

    def _decode_step(self, B: int = 1, use_graph: bool = True, want_logits: bool = False):
        """Generate one token for batch slots [0, B): consumes self.best, appends to token_log, leaves the next
        arg-max in self.best.  With use_graph the whole step (2 + 5 * layers kernels) replays as one CUDA graph."""
        if not use_graph:
            self._decode_step_kernels(B, want_logits)
            return
        key = (B, want_logits)
        if key not in self._graphs:
            # capture (the kernels were warmed up by an eager step so every cudaFuncSetAttribute already ran)
            self._snapshot = self._save_state()
            self._decode_step_kernels(B, want_logits)
            torch.cuda.synchronize()
            self._restore_state(self._snapshot)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    self._decode_step_kernels(B, want_logits)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._restore_state(self._snapshot)   # capture does not execute, but keep the invariant explicit
            self._graphs[key] = g
        self._graphs[key].replay()

    # ------------------------------------------------------------------------------------------ batched decode
    def _batched_step_kernels(self, B: int, want_logits: bool):
        """One token for each of B sequences through the GEMM path (weights of every touched expert are streamed once
        for the whole batch; BASELINE configs[4]).  Same kernels as the prefill, with M = B rows at B different
        positions, + the paged decode attention."""
        c, w, cache = self.cfg, self.w, self.cache
        nq, nkv, D, E, H = c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.num_local_experts, c.hidden_size
        ws = self._ws(max(B, 16))
        h = self.d_h[:B]
        ops.decode_embed(self.best[:B], self.token_log[:B], self.gen_count[:B], cache.cache_len[:B], cache.cur_pos[:B],
                         w["embed"], h, cache.max_seq_len)
        slots = self.d_slots[:B]
        ops.decode_slots(cache.cur_pos[:B], cache.block_table[:B], slots, cache.page_size)
        xn, qkv, attn, xn2 = ws["xn"][:B], ws["qkv"][:B], ws["attn"][:B], ws["xn2"][:B]
        ids, tw = ws["ids"][:B], ws["tw"][:B]
        perm, rtok, rw = ws["perm"][:2 * B], ws["rtok"][:2 * B], ws["rw"][:2 * B]
        xp, act, yp = ws["xp"][:2 * B], ws["act"][:2 * B], ws["yp"][:2 * B]
        layers = w["layers"]
        if self.moe_fused:
            ws["cnt"].zero_()
        ops.rmsnorm(h, layers[0]["ln1"], c.rms_norm_eps, out=xn)
        for li, lw in enumerate(layers):
            ops.linear_qkv_rope(xn, lw["wqkv"], qkv, cache.cur_pos[:B], slots, w["rope"], cache.k[li], cache.v[li],
                                nq, nkv, D)
            ops.decode_attention(qkv, cache.k[li], cache.v[li], cache.block_table[:B], cache.cur_pos[:B], attn,
                                 self.attn_ws, nq, nkv, D, cache.page_size, self.decode_splits, D ** -0.5,
                                 q_stride=c.qkv_rows)
            ops.linear(attn, lw["wo"], residual=h, out=h)
            nxt = layers[li + 1]["ln1"] if li + 1 < len(layers) else w["norm"]
            if self.moe_fused:
                cnt = ws["cnt"][li]
                ops.moe_route_scatter(h, lw["ln2"], lw["gate"], ws["xs"], cnt, perm, ws["rws"], c.rms_norm_eps)
                ops.moe_gate_up_slots(ws["xs"], lw["w13"], ws["acts"], cnt, 2 * B)
                ops.moe_down_slots(ws["acts"], lw["w2"], ws["ys"], cnt, ws["rws"], 2 * B)
                ops.moe_combine(h, ws["ys"], perm, nxt, xn, c.rms_norm_eps)
                continue
            ops.moe_router(h, lw["ln2"], lw["gate"], xn2, ids, tw, c.rms_norm_eps)
            ops.moe_align(ids, tw, ws["offs"], perm, rtok, rw, B, E)
            ops.row_copy(xn2, rtok, None, xp, 2 * B)
            ops.moe_gate_up(xp, lw["w13"], act, ws["offs"], 2 * B)
            ops.moe_down(act, lw["w2"], yp, ws["offs"], rw, 2 * B)
            ops.moe_combine(h, yp, perm, nxt, xn, c.rms_norm_eps)
        logits = ops.linear(xn, w["lm_head"], out=self.d_logits[:B])
        ops.argmax_rows(logits, self.best[:B])

    @property
    def launches_per_batched_step(self) -> int:
        return (6 + 8 * self.cfg.num_hidden_layers) if self.moe_fused else (5 + 10 * self.cfg.num_hidden_layers)

    @torch.no_grad()
    def decode_step_batched(self, B: int, use_graph: bool = True):
        assert self.ep_world == 1 and 1 <= B <= self.max_batch
        if not use_graph:
            self._batched_step_kernels(B, False)
            return
        if B not in self._bgraphs:             # one captured graph per batch size (continuous batching changes B)
            snap = self._save_state()
            self._batched_step_kernels(B, False)          # warm-up (cudaFuncSetAttribute etc.)
            torch.cuda.synchronize()
            self._restore_state(snap)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    self._batched_step_kernels(B, False)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._restore_state(snap)
            self._bgraphs[B] = g
        self._bgraphs[B].replay()

    def _save_state(self):
        c = self.cache
        return (self.best.clone(), self.gen_count.clone(), c.cache_len.clone(), c.cur_pos.clone(), self.d_h.clone(),
                self.token_log.clone())

    def _restore_state(self, s):
        c = self.cache
        self.best.copy_(s[0]); self.gen_count.copy_(s[1]); c.cache_len.copy_(s[2]); c.cur_pos.copy_(s[3])
        self.d_h.copy_(s[4]); self.token_log.copy_(s[5])
        # the KV slot written by the warm-up step is rewritten by the real step (same position): nothing to undo

    def check_capacity(self, prompt_len: int, new_tokens: int):
        """Every decode entry point calls this: prompt + reply must fit the paged KV cache of one sequence (the
        kernels index block_table[pos / page_size] and the rope table without a bound of their own)."""
        if prompt_len + new_tokens > self.cache.max_seq_len:
            raise ValueError(f"prompt ({prompt_len}) + max_new_tokens ({new_tokens}) exceeds the KV capacity "
                             f"({self.cache.max_seq_len} positions per sequence); build the model with a larger "
                             f"max_seq_len")
        if new_tokens > self.max_new_tokens:
            raise ValueError(f"max_new_tokens {new_tokens} exceeds the token log ({self.max_new_tokens}); build the "
                             f"model with a larger max_new_tokens")

    def reset(self):
        self.cache.reset()
        self.chain_mem.zero_()     # completion counters restart with the request (also heals an aborted step)
        self.best.zero_()
        self.gen_count.zero_()
        self.token_log.zero_()

    def generated_tokens(self, slot: int = 0) -> List[int]:
        n = int(self.gen_count[slot])
        return self.token_log[slot, :n].tolist()
