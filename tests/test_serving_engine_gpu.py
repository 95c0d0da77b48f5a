"""Continuous batching on the real decoder (vita_b200/engine.py::DecoderEngine): requests that arrive at different
steps, share the batch with changing neighbours, get their slots swapped when others retire -- and must still produce
exactly the tokens they produce when served alone through the same batched decode step."""
import pytest
import torch

from vita_b200 import weights as W
from vita_b200.config import VitaConfig
from vita_b200.engine import ContinuousBatcher, DecoderEngine, Request

pytestmark = pytest.mark.gpu


def _requests(cfg, arrivals):
    g = torch.Generator().manual_seed(11)
    reqs = []
    for rid, (arr, plen, n_new) in enumerate(arrivals):
        ids = torch.randint(0, cfg.llm.vocab_size, (1, plen), generator=g)
        reqs.append(Request(rid, {"input_ids": ids}, n_new, None, arrival_step=arr))
    return reqs


def test_staggered_requests_match_solo_runs():
    from vita_b200.model.vita_mixtral import VITAMixtralForCausalLM
    cfg = VitaConfig.tiny()
    packed = {"llm": W.pack_llm(W.synthetic_state(cfg, 0, parts=("llm",)), cfg, "cuda")}
    model = VITAMixtralForCausalLM(cfg, packed, "cuda", max_batch=4, max_seq_len=128, max_new_tokens=24)
    spec = [(0, 9, 10), (0, 17, 4), (1, 5, 12), (2, 30, 6), (3, 12, 9), (9, 7, 5), (40, 21, 7)]
    eng = DecoderEngine(model)
    solo = {}
    for r in _requests(cfg, spec):
        r.arrival_step = 0
        solo.update(ContinuousBatcher(eng, 1).run([r]))
    seen = []
    out = ContinuousBatcher(eng, 3, sync_every=2).run(_requests(cfg, spec), on_step=lambda s, n: seen.append(n))
    assert set(out) == set(solo)
    for rid in solo:
        assert len(out[rid]) == spec[rid][2]
        assert out[rid] == solo[rid], f"request {rid} changed under continuous batching"
    assert max(seen) == 3 and min(seen) >= 1
    # capacity is checked at admission
    big = Request(99, {"input_ids": torch.zeros(1, 120, dtype=torch.long)}, 20)
    with pytest.raises(ValueError):
        ContinuousBatcher(eng, 1).run([big])
