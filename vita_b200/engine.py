"""Continuous batching for concurrent requests (BASELINE configs[4]: 16 audio queries arriving at different times, one
token per request per step through the batched paged-KV decode step).

The reference serves this shape through vLLM's engine loop (web_demo/web_interactive_demo.py:270-378,
`llm.generate` inside an asyncio loop; scheduler = vLLM's).  Here the scheduling is host-side index bookkeeping over
the decoder's slot tensors:

  * active requests always occupy the slot prefix [0, n_active) -- the batched decode step (and its CUDA graph per
    batch size) works on a prefix;
  * an arriving request is prefilled into slot n_active;
  * a finished request is retired by swapping the last active slot into its place: only the *rows* of the per-slot
    tensors move (block-table row = the request's KV pages, positions, pending arg-max, token log); no KV data is
    copied.

Duplex semantics of the interactive demo as scheduler events (web_interactive_demo.py:251-253,286-293,340-370):
the reference runs TWO full engine replicas that hand the microphone stream to each other; here one weight copy serves
every query of a session and the hand-off is bookkeeping:

  * negative-audio abort -- the first generated token of a query is its state token; `<2>` marks noise / a query not
    addressed to the assistant (`judge_negative`, :251-253): the request is retired right after its prefill, before
    it ever occupies a decode slot (the reference breaks out of its generation loop, :368-370);
  * interrupt hand-off -- a query whose first token is NOT negative interrupts the running answer of the same session
    (`other_stop_event.set()`, :340-353): the older request is retired at that step with the tokens it has produced.

`ContinuousBatcher` only talks to an `engine` object (prefill / step / swap / read_tokens / reset_slot / first_token), so
the policy is unit-tested on the CPU with a deterministic stand-in; `DecoderEngine` binds it to `VITAMixtralForCausalLM`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import torch


@dataclass
class Request:
    rid: int
    payload: dict                       # {"input_ids", "images", "audios"} for DecoderEngine; opaque to the batcher
    max_new_tokens: int
    eos_token_id: Optional[int] = None
    arrival_step: int = 0               # engine step at which the request becomes visible
    session: Optional[int] = None       # duplex: queries of one conversation; a real query interrupts the running answer
    negative_token_id: Optional[int] = None   # duplex: state token `<2>`; a query answering with it is dropped at once
    # filled in by the batcher
    outcome: str = ""                   # "finished" | "negative" | "interrupted"
    steps: int = 0
    admitted_step: int = -1
    finished_step: int = -1
    tokens: List[int] = field(default_factory=list)


def swap_rows(tensors: Sequence[torch.Tensor], i: int, j: int) -> None:
    """Exchange rows i and j of every per-slot tensor (first dimension = slot)."""
    if i == j:
        return
    idx, rev = [i, j], [j, i]
    for t in tensors:
        t[idx] = t[rev].clone()


class ContinuousBatcher:
    def __init__(self, engine, max_batch: int, sync_every: int = 1):
        assert max_batch >= 1 and sync_every >= 1
        self.engine, self.max_batch, self.sync_every = engine, max_batch, sync_every

    def run(self, requests: Sequence[Request], on_step: Optional[Callable[[int, int], None]] = None) -> Dict[int, List[int]]:
        """Runs every request to completion; returns {rid: generated tokens (cut after EOS / max_new_tokens)}.
        `on_step(step_index, n_active)` is called after each decode-step launch (latency harnesses record events)."""
        pending = sorted(requests, key=lambda r: (r.arrival_step, r.rid))
        active: List[Request] = []          # active[s] lives in slot s
        done: Dict[int, List[int]] = {}
        step = 0
        since_sync = 0
        while pending or active:
            while pending and pending[0].arrival_step <= step and len(active) < self.max_batch:
                r = pending.pop(0)
                slot = len(active)
                self.engine.reset_slot(slot)
                self.engine.prefill(slot, r)
                r.admitted_step, r.steps = step, 0
                if r.negative_token_id is not None or r.session is not None:
                    first = self.engine.first_token(slot)            # the state token: known right after the prefill
                    if r.negative_token_id is not None and first == r.negative_token_id:
                        # negative audio: never enters the decode batch (the slot is simply reused by the next arrival)
                        r.tokens, r.finished_step, r.outcome = [first], step, "negative"
                        done[r.rid] = r.tokens
                        continue
                    if r.session is not None:                        # a real query: interrupt the running answer
                        for old_slot in sorted((i for i, o in enumerate(active) if o.session == r.session),
                                               reverse=True):
                            old = active[old_slot]
                            old.tokens = self.engine.read_tokens(old_slot)[: min(old.steps, old.max_new_tokens)]
                            old.finished_step, old.outcome = step, "interrupted"
                            done[old.rid] = old.tokens
                            last = len(active) - 1
                            # the new request sits in slot len(active) (not yet appended): move it down with the rest
                            if old_slot != last:
                                self.engine.swap(old_slot, last)
                                active[old_slot] = active[last]
                            active.pop()
                            self.engine.swap(last, last + 1)          # keep the new request right behind the prefix
                active.append(r)
            if not active:
                if not pending:
                    break                           # the last arrivals were dropped as negative queries
                step = max(step, pending[0].arrival_step)   # idle until the next arrival
                continue
            self.engine.step(len(active))
            for r in active:
                r.steps += 1
            if on_step is not None:
                on_step(step, len(active))
            step += 1
            since_sync += 1
            budget_hit = any(r.steps >= r.max_new_tokens for r in active)
            if since_sync >= self.sync_every or budget_hit:
                since_sync = 0
                self._retire(active, done, step)
        return done

    def _retire(self, active: List[Request], done: Dict[int, List[int]], step: int) -> None:
        finished = []
        for slot, r in enumerate(active):
            toks = self.engine.read_tokens(slot)[: r.steps]
            cut = None
            if r.eos_token_id is not None and r.eos_token_id in toks:
                cut = toks.index(r.eos_token_id) + 1
            if cut is None and r.steps >= r.max_new_tokens:
                cut = r.max_new_tokens
            if cut is not None:
                r.tokens, r.finished_step, r.outcome = toks[: min(cut, r.max_new_tokens)], step, "finished"
                finished.append(slot)
        for slot in sorted(finished, reverse=True):     # highest slot first: the swap partner is never a finished one
            last = len(active) - 1
            done[active[slot].rid] = active[slot].tokens
            if slot != last:
                self.engine.swap(slot, last)
                active[slot] = active[last]
            active.pop()


class DecoderEngine:
    """Binds the batcher to `VITAMixtralForCausalLM`: prefill into a slot, batched decode step over the slot prefix."""

    def __init__(self, model, use_graph: bool = True, overrun: int = 8, lone_fast_path: bool = False):
        """`overrun`: decode steps a finished request may still take before the batcher looks at its tokens
        (its `sync_every` - 1); reserved in the KV capacity check at admission.
        `lone_fast_path`: a step with a single running request takes the bs = 1 weight-streaming GEMV chain instead of
        the batched GEMM step (faster; the two paths round differently, so tokens are no longer batch-invariant)."""
        self.model, self.llm, self.use_graph, self.overrun = model, model.llm, use_graph, overrun
        self.lone_fast_path = lone_fast_path
        c = self.llm.cache
        self._slot_tensors = [c.block_table, c.slot_map, c.cache_len, c.cur_pos, self.llm.best, self.llm.token_log,
                              self.llm.gen_count]

    def reset_slot(self, slot: int) -> None:
        c = self.llm.cache
        for t in (c.cache_len, c.cur_pos, self.llm.best, self.llm.gen_count):
            t[slot:slot + 1].zero_()
        self.llm.token_log[slot].zero_()

    def prefill(self, slot: int, request: Request) -> None:
        p = request.payload
        emb, lens = self.model._embeds_for(p["input_ids"], p.get("images"), p.get("audios"))
        need = int(lens[0]) + request.max_new_tokens + self.overrun
        if need > self.llm.cache.max_seq_len or request.max_new_tokens > self.llm.max_new_tokens:
            raise ValueError(f"request {request.rid}: {need} KV positions / {request.max_new_tokens} new tokens exceed "
                             f"the engine's capacity ({self.llm.cache.max_seq_len} / {self.llm.max_new_tokens})")
        self.llm.prefill(emb[0, : lens[0]].contiguous(), slot=slot)

    def step(self, n_active: int) -> None:
        if n_active == 1 and self.lone_fast_path:      # a lone request takes the weight-streaming GEMV chain (2 experts read, not a GEMM tile)
            self.llm.decode_step(1, use_graph=self.use_graph)
        else:
            self.llm.decode_step_batched(n_active, use_graph=self.use_graph)

    def swap(self, i: int, j: int) -> None:
        swap_rows(self._slot_tensors, i, j)

    def read_tokens(self, slot: int) -> List[int]:
        return self.llm.generated_tokens(slot)

    def first_token(self, slot: int) -> int:
        """arg-max of the prompt's last position (packed as (logit, ~index) by the LM-head kernel); one host sync"""
        return int(0xFFFFFFFF - (int(self.llm.best[slot]) & 0xFFFFFFFF))
