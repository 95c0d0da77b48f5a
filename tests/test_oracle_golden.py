"""The oracle restatement vs the golden vectors minted from the reference's own classes (oracle/make_golden.py)."""
import torch

from oracle import vita_oracle as O
from vita_b200 import weights as W
from vita_b200.config import VitaConfig

TOL = 2e-5


def _t(a):
    return torch.from_numpy(a)


def _close(got, want, what):
    err = (got - want).abs().max().item()
    assert err <= TOL * max(1.0, want.abs().max().item()), f"{what}: {err}"


def _state():
    cfg = VitaConfig.tiny()
    return cfg, W.synthetic_state(cfg, 0)


def test_vision_and_audio_match_reference(golden):
    inp, out = golden
    cfg, state = _state()
    images = _t(inp["images"])
    _close(O.vision_tower(state, cfg.vision, images), _t(out["vision_tower"]), "vision tower")
    _close(O.encode_images(state, cfg, images), _t(out["image_features"]), "encode_images")
    a = O.encode_audios(state, cfg, _t(inp["feats"]), _t(inp["lengths"]))
    _close(a["inputs_embeds"], _t(out["audio_embeds"]), "audio embeds")
    assert torch.equal(a["attention_mask"].to(torch.int32), _t(out["audio_mask"]))


def test_text_prefill_and_greedy_match_reference(golden):
    inp, out = golden
    cfg, state = _state()
    ids = _t(inp["text_ids"])
    logits, _, _ = O.forward(state, cfg, ids)
    _close(logits, _t(out["text_prefill_logits"]), "prefill logits")
    toks, rows = O.greedy_generate(state, cfg, ids, max_new_tokens=8)
    assert toks == out["text_greedy_tokens"].tolist()
    _close(rows, _t(out["text_decode_logits"]), "decode logits")


def test_omni_splice_and_greedy_match_reference(golden):
    inp, out = golden
    cfg, state = _state()
    images, feats, lengths = _t(inp["images"]), _t(inp["feats"]), _t(inp["lengths"])
    audios = {"audios": feats[:1], "lengths": lengths[:1]}
    emb, _ = O.prepare_inputs_embeds(state, cfg, _t(inp["omni_ids"]), images[:1], audios)
    _close(emb, _t(out["omni_inputs_embeds"]), "omni inputs_embeds")
    toks, rows = O.greedy_generate(state, cfg, _t(inp["omni_ids"]), images[:1], audios, max_new_tokens=6)
    assert toks == out["omni_greedy_tokens"].tolist()
    _close(rows, _t(out["omni_decode_logits"]), "omni decode logits")


def test_batched_splice_matches_reference(golden):
    inp, out = golden
    cfg, state = _state()
    images, feats = _t(inp["images"]), _t(inp["feats"])
    b_images = torch.cat([images[:1], images[1:2], images[:1]])
    b_audios = {"audios": torch.cat([feats[:1], feats[1:2], feats[:1], feats[1:2]]),
                "lengths": torch.tensor([100, 77, 100, 77])}
    emb, lens = O.prepare_inputs_embeds(state, cfg, _t(inp["batch_ids"]), b_images, b_audios)
    _close(emb, _t(out["batch_inputs_embeds"]), "batched inputs_embeds")
    assert lens == out["batch_lens"].tolist()


def test_synthetic_weights_are_reproducible_and_bf16_exact():
    cfg = VitaConfig.tiny()
    a = W.synthetic_state(cfg, 0, parts=("projector",))
    b = W.synthetic_state(cfg, 0, parts=("projector",))
    c = W.synthetic_state(cfg, 1, parts=("projector",))
    for k in a:
        assert torch.equal(a[k], b[k]) and a[k].dtype == torch.bfloat16
    assert any(not torch.equal(a[k], c[k]) for k in a)
    assert set(W.all_param_shapes(cfg)) == set(W.synthetic_state(cfg, 0))


def test_oracle_route_override_is_identity_on_its_own_decisions_and_reports_the_margins():
    """`sparse_moe(route=...)` (the checker aid of tests/full_depth.py): given the oracle's own expert pairs it reproduces
    the default path bit for bit; given another pair it evaluates those experts with the oracle's renormalised weights."""
    import torch
    from oracle import vita_oracle as O
    from vita_b200.config import VitaConfig
    from vita_b200 import weights as W
    cfg = VitaConfig.tiny() if hasattr(VitaConfig, "tiny") else None
    if cfg is None:
        import pytest
        pytest.skip("no tiny config")
    state = W.synthetic_state(cfg, 3, parts=("llm",))
    lc = cfg.llm
    xn = torch.randn(11, lc.hidden_size, generator=torch.Generator().manual_seed(0))
    y0, ids0, w0 = O.sparse_moe(state, lc, 0, xn)
    route = {"ids": ids0.clone()}
    y1, ids1, w1 = O.sparse_moe(state, lc, 0, xn, None, route)
    assert torch.equal(y0, y1) and torch.equal(ids0, ids1) and torch.equal(w0, w1)
    assert torch.equal(route["own_ids"], ids0) and route["logits"].shape == (11, lc.num_local_experts)
    # swap the second expert of token 0 for another one: only that token's output changes, weights renormalise over the pair
    other = next(e for e in range(lc.num_local_experts) if e not in ids0[0].tolist())
    forced = ids0.clone()
    forced[0, 1] = other
    route = {"ids": forced}
    y2, ids2, w2 = O.sparse_moe(state, lc, 0, xn, None, route)
    assert torch.allclose(y2[1:], y0[1:], rtol=1e-5, atol=1e-6) and not torch.allclose(y2[0], y0[0], rtol=1e-3)
    p = torch.softmax(route["logits"][0], -1)[forced[0]]
    assert torch.allclose(w2[0], p / p.sum(), atol=1e-6)
