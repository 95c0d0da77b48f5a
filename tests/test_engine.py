"""Continuous batching policy (vita_b200/engine.py) on the CPU: a deterministic stand-in engine produces, for every
request, a token stream that depends only on the request -- so whatever the arrival pattern, slot swaps and retirement
order, each request must come back with exactly its own stream, cut at EOS / max_new_tokens."""
import random

import pytest
import torch

from vita_b200.engine import ContinuousBatcher, Request, swap_rows


def stream(seed: int, n: int):
    return [(seed * 7919 + i * 104729 + (i * i) % 13) % 1000 for i in range(n)]


class FakeEngine:
    """Per-slot state lives in tensors whose rows are swapped like the real decoder's; `step` appends the next token of
    whatever request currently sits in each active slot."""

    def __init__(self, max_batch: int, cap: int):
        self.seed = torch.zeros(max_batch, dtype=torch.int64)
        self.count = torch.zeros(max_batch, dtype=torch.int64)
        self.log = torch.full((max_batch, cap), -1, dtype=torch.int64)
        self.max_active = 0
        self.steps = 0
        self.swaps = 0

    def reset_slot(self, slot):
        self.seed[slot] = 0; self.count[slot] = 0; self.log[slot] = -1

    def prefill(self, slot, request):
        assert int(self.count[slot]) == 0, "slot must be empty"
        self.seed[slot] = request.payload["seed"]

    def step(self, n_active):
        self.max_active = max(self.max_active, n_active)
        self.steps += 1
        for s in range(n_active):
            i = int(self.count[s])
            self.log[s, i] = stream(int(self.seed[s]), i + 1)[i]
            self.count[s] += 1

    def swap(self, i, j):
        self.swaps += 1
        swap_rows([self.seed, self.count, self.log], i, j)

    def read_tokens(self, slot):
        return self.log[slot, : int(self.count[slot])].tolist()

    def first_token(self, slot):
        return stream(int(self.seed[slot]), 1)[0]


def expected(r: Request):
    toks = stream(r.payload["seed"], r.max_new_tokens)
    if r.eos_token_id is not None and r.eos_token_id in toks:
        toks = toks[: toks.index(r.eos_token_id) + 1]
    return toks


def test_swap_rows():
    a = torch.arange(12).view(4, 3).clone()
    b = torch.arange(4).clone()
    swap_rows([a, b], 0, 3)
    assert a.tolist() == [[9, 10, 11], [3, 4, 5], [6, 7, 8], [0, 1, 2]] and b.tolist() == [3, 1, 2, 0]
    swap_rows([a, b], 2, 2)
    assert b.tolist() == [3, 1, 2, 0]


@pytest.mark.parametrize("sync_every", [1, 3])
@pytest.mark.parametrize("trial", range(6))
def test_every_request_gets_its_own_stream(trial, sync_every):
    rng = random.Random(trial)
    max_batch = rng.choice([1, 2, 4, 16])
    reqs = []
    for rid in range(rng.randint(1, 40)):
        n = rng.randint(1, 24)
        seed = rng.randint(1, 10 ** 6)
        eos = None
        if rng.random() < 0.5:                       # stop at a token that really occurs somewhere in the stream
            eos = stream(seed, n)[rng.randrange(n)]
        reqs.append(Request(rid, {"seed": seed}, n, eos, arrival_step=rng.randint(0, 60)))
    eng = FakeEngine(max_batch, cap=24 + sync_every + 1)
    out = ContinuousBatcher(eng, max_batch, sync_every).run(reqs)
    assert set(out) == {r.rid for r in reqs}
    for r in reqs:
        assert out[r.rid] == expected(r), (r.rid, r.max_new_tokens, r.eos_token_id)
        assert r.admitted_step >= r.arrival_step and r.finished_step > r.admitted_step
    assert eng.max_active <= max_batch


def test_arrivals_fill_freed_slots_and_idle_gaps_are_skipped():
    reqs = [Request(0, {"seed": 1}, 5, None, 0), Request(1, {"seed": 2}, 2, None, 0), Request(2, {"seed": 3}, 3, None, 1),
            Request(3, {"seed": 4}, 2, None, 100)]
    eng = FakeEngine(2, cap=16)
    out = ContinuousBatcher(eng, 2).run(reqs)
    assert [len(out[i]) for i in range(4)] == [5, 2, 3, 2]
    # request 2 has to wait for request 1's slot; request 3 arrives long after everything finished
    assert reqs[2].admitted_step == 2 and reqs[3].admitted_step == 100
    assert eng.steps == 5 + 2        # steps 0..4 for the first three, then 2 for the late arrival (no idle steps)


def test_duplex_negative_query_is_dropped_after_prefill():
    """web_interactive_demo.py:251-253,368-370: a reply that starts with the `<2>` state token is abandoned at once."""
    neg = stream(77, 1)[0]
    reqs = [Request(0, {"seed": 5}, 6, None, 0, session=1, negative_token_id=neg),
            Request(1, {"seed": 77}, 6, None, 2, session=1, negative_token_id=neg),     # noise while request 0 talks
            Request(2, {"seed": 9}, 4, None, 3, session=2, negative_token_id=neg)]
    eng = FakeEngine(4, cap=16)
    out = ContinuousBatcher(eng, 4).run(reqs)
    assert out[1] == [neg] and reqs[1].outcome == "negative" and reqs[1].finished_step == reqs[1].admitted_step
    # the noise query neither interrupted request 0 nor took a decode slot
    assert out[0] == stream(5, 6) and reqs[0].outcome == "finished"
    assert out[2] == stream(9, 4) and eng.max_active == 2


def test_duplex_real_query_interrupts_the_running_answer_of_its_session():
    """web_interactive_demo.py:340-353: the first non-negative token of a new query stops the other answer."""
    reqs = [Request(0, {"seed": 5}, 20, None, 0, session=1, negative_token_id=-1),
            Request(1, {"seed": 6}, 20, None, 0, session=2, negative_token_id=-1),       # another conversation
            Request(2, {"seed": 7}, 5, None, 4, session=1, negative_token_id=-1)]       # barge-in at step 4
    eng = FakeEngine(4, cap=32)
    out = ContinuousBatcher(eng, 4).run(reqs)
    assert reqs[0].outcome == "interrupted" and out[0] == stream(5, 4)      # 4 tokens were produced before the barge-in
    assert reqs[0].finished_step == 4 == reqs[2].admitted_step
    assert out[1] == stream(6, 20) and reqs[1].outcome == "finished"        # the other session is untouched
    assert out[2] == stream(7, 5) and reqs[2].outcome == "finished"
    assert eng.max_active == 2


@pytest.mark.parametrize("trial", range(4))
def test_duplex_random_sessions_keep_streams_intact(trial):
    rng = random.Random(100 + trial)
    reqs, t = [], 0
    for rid in range(30):
        t += rng.randint(0, 6)
        seed = rng.randint(1, 10 ** 6)
        reqs.append(Request(rid, {"seed": seed}, rng.randint(1, 12), None, t, session=rng.randint(0, 3),
                            negative_token_id=stream(seed, 1)[0] if rng.random() < 0.3 else -1))
    eng = FakeEngine(4, cap=40)
    out = ContinuousBatcher(eng, 4, sync_every=rng.choice([1, 2])).run(reqs)
    assert set(out) == {r.rid for r in reqs}
    for r in reqs:
        full = stream(r.payload["seed"], r.max_new_tokens)
        if r.outcome == "negative":
            assert out[r.rid] == full[:1]
        elif r.outcome == "interrupted":
            assert out[r.rid] == full[: len(out[r.rid])] and len(out[r.rid]) <= r.max_new_tokens
        else:
            assert out[r.rid] == full
