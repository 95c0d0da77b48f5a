"""End-to-end parity of the CUDA path (through the reference-shaped Python surface) against
 (a) the golden vectors minted from the reference itself (tiny geometry, tests/golden/), and
 (b) the oracle at full Mixtral / InternViT / Whale layer width (reduced depth where host RAM requires it).
Tolerances are for bf16 storage with fp32 accumulation against an fp32 oracle and are written next to each check."""
import pytest
import torch

from tests.util import assert_close, bf16_round
from oracle import vita_oracle as O
from vita_b200 import weights as W
from vita_b200.config import VitaConfig, IMAGE_TOKEN_INDEX, AUDIO_TOKEN_INDEX

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(a)


@pytest.fixture(scope="module")
def tiny():
    from vita_b200.model.vita_mixtral import VITAMixtralForCausalLM
    cfg = VitaConfig.tiny()
    state = W.synthetic_state(cfg, 0)
    model = VITAMixtralForCausalLM(cfg, W.pack(state, cfg, "cuda"), "cuda", max_batch=3, max_new_tokens=32,
                                   shuffle_pages=True)
    return cfg, state, model


def _margin_ok(rows, tol):
    top2 = rows.topk(2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]) > tol


def test_tiny_encoders_match_reference_golden(tiny, golden):
    cfg, state, model = tiny
    inp, out = golden
    images = _t(inp["images"])
    assert_close(model.get_vision_tower()(images), _t(out["vision_tower"]), rel=3e-2, what="vision tower")
    assert_close(model.encode_images(images), _t(out["image_features"]), rel=3e-2, what="encode_images")
    a = model.encode_audios(_t(inp["feats"]), _t(inp["lengths"]))
    assert_close(a["inputs_embeds"], _t(out["audio_embeds"]), rel=3e-2, what="audio inputs_embeds")
    assert torch.equal(a["attention_mask"].cpu().to(torch.int32), _t(out["audio_mask"]))


def test_tiny_text_prefill_logits_and_greedy_match_reference_golden(tiny, golden):
    cfg, state, model = tiny
    inp, out = golden
    ids = _t(inp["text_ids"])
    logits = model(input_ids=ids).logits
    ref = _t(out["text_prefill_logits"])
    assert_close(logits, ref, rel=3e-2, what="text prefill logits")
    gen = model.generate(ids, max_new_tokens=8, output_scores=True, use_graph=False)
    new = gen.sequences[0, ids.shape[1]:].tolist()
    ref_rows = _t(out["text_decode_logits"])
    got_rows = torch.cat([s.float().cpu() for s in gen.scores])
    # teacher-free greedy: exact ids as long as every reference step had a clear margin (2% of the logit range)
    ok = _margin_ok(ref_rows, 0.03 * ref_rows.abs().max())
    want = out["text_greedy_tokens"].tolist()
    n = 0
    while n < len(want) and ok[n]:
        n += 1
    assert n >= 1
    assert new[:n] == want[:n], (new, want, ok.tolist())
    assert_close(got_rows[:n], ref_rows[:n], rel=3e-2, what="decode logits")
    # CUDA-graph replay produces the same tokens as eager launches
    gen2 = model.generate(ids, max_new_tokens=8, use_graph=True, sync_every=3)
    assert gen2.sequences.tolist() == gen.sequences.tolist()


def test_tiny_omni_splice_prefill_greedy_match_reference_golden(tiny, golden):
    cfg, state, model = tiny
    inp, out = golden
    ids = _t(inp["omni_ids"])
    images = _t(inp["images"])[:1]
    audios = {"audios": _t(inp["feats"])[:1], "lengths": _t(inp["lengths"])[:1]}
    emb = model.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, images, audios)[4]
    assert_close(emb, _t(out["omni_inputs_embeds"]), rel=3e-2, what="spliced inputs_embeds")
    logits = model(input_ids=ids, images=images, audios=audios).logits
    assert_close(logits[:, -1], _t(out["omni_last_logits"]), rel=4e-2, what="omni last-row logits")
    gen = model.generate(ids, images=images, audios=audios, max_new_tokens=6, output_scores=True)
    assert gen.sequences[0, : ids.shape[1]].tolist() == ids[0].tolist()          # prompt echoed incl. placeholders
    ref_rows = _t(out["omni_decode_logits"])
    ok = _margin_ok(ref_rows, 0.03 * ref_rows.abs().max())
    want = out["omni_greedy_tokens"].tolist()
    new = gen.sequences[0, ids.shape[1]:].tolist()
    n = 0
    while n < len(want) and ok[n]:
        n += 1
    assert new[:n] == want[:n], (new, want, ok.tolist())


def test_tiny_batched_splice_matches_reference_golden(tiny, golden):
    cfg, state, model = tiny
    inp, out = golden
    images, feats = _t(inp["images"]), _t(inp["feats"])
    b_images = torch.cat([images[:1], images[1:2], images[:1]])
    b_audios = {"audios": torch.cat([feats[:1], feats[1:2], feats[:1], feats[1:2]]),
                "lengths": torch.tensor([100, 77, 100, 77])}
    emb = model.prepare_inputs_labels_for_multimodal(_t(inp["batch_ids"]), None, None, None, None, b_images, b_audios)[4]
    assert_close(emb, _t(out["batch_inputs_embeds"]), rel=3e-2, what="batched spliced inputs_embeds")


def test_teacher_forced_decode_matches_oracle(tiny, golden):
    """Feed the oracle's tokens step by step (forward() with past_key_values): every step's logits must agree and the
    arg-max must be identical wherever the oracle's margin is above the bf16 noise floor."""
    cfg, state, model = tiny
    inp, _ = golden
    ids = _t(inp["text_ids"])
    toks, rows = O.greedy_generate(state, cfg, ids, max_new_tokens=12)
    out = model(input_ids=ids)
    got = [out.logits[0, -1].float().cpu()]
    pkv = out.past_key_values
    for t in toks[:-1]:
        out = model(input_ids=torch.tensor([[t]]), past_key_values=pkv)
        got.append(out.logits[0, -1].float().cpu())
    got = torch.stack(got)
    assert_close(got, rows, rel=3e-2, what="teacher-forced logits")
    ok = _margin_ok(rows, 0.03 * rows.abs().max())
    assert ok.float().mean() > 0.5
    assert torch.equal(got.argmax(-1)[ok], rows.argmax(-1)[ok])


def test_full_width_vit_one_tile_matches_oracle():
    """InternViT-300M at full size (24 layers, 1025 tokens) + projector on one 448 px tile."""
    from vita_b200.model.internvit import InternViTVisionTower, VisionProjector
    cfg = VitaConfig.full(num_hidden_layers=1)
    state = W.synthetic_state(cfg, 0, parts=("vision", "projector"))
    img = bf16_round(torch.randn(1, 3, 448, 448, generator=torch.Generator().manual_seed(1)))
    tower = InternViTVisionTower(cfg.vision, W.pack_vision(state, cfg, "cuda"), "cuda")
    proj = VisionProjector(W.pack_projector(state, cfg, "cuda"))
    feats = tower(img)
    assert tuple(feats.shape) == (1, 256, 4096)
    ref = O.vision_tower(state, cfg.vision, img)
    assert_close(feats, ref, rel=5e-2, what="full ViT tower (24 layers of bf16 residual stream)")
    assert_close(proj(feats), O.mm_projector(state, ref), rel=5e-2, what="projector")


def test_full_width_whale_10s_matches_oracle():
    """Whale encoder + adapter at full size on 998 fbank frames (10 s) with one padded batch entry."""
    from vita_b200.model.whale import AudioEncoder
    cfg = VitaConfig.full(num_hidden_layers=1)
    state = W.synthetic_state(cfg, 0, parts=("audio",))
    feats = bf16_round(torch.randn(2, 998, 80, generator=torch.Generator().manual_seed(2)) * 2)
    lengths = torch.tensor([998, 640])
    enc = AudioEncoder(cfg.audio, cfg.llm.hidden_size, W.pack_audio(state, cfg, "cuda"), "cuda")
    got = enc(feats, lengths)
    ref = O.encode_audios(state, cfg, feats, lengths)
    assert tuple(got["inputs_embeds"].shape) == (2, 124, 4096)
    assert torch.equal(got["attention_mask"].cpu(), ref["attention_mask"])
    m = ref["attention_mask"]
    assert_close(got["inputs_embeds"][0], ref["inputs_embeds"][0], rel=5e-2, what="whale 10 s")
    assert_close(got["inputs_embeds"][1][m[1]], ref["inputs_embeds"][1][m[1]], rel=5e-2, what="whale padded entry")


def _routing_stable(trace, min_log_ratio=0.06):
    """Rows whose top-2 expert *set* is decided by a clear margin in every traced layer: log(p2 / p3) above the
    bf16 noise floor of the router logits.  (Swapping 1st and 2nd changes nothing: the pair is renormalised.)"""
    ok = None
    for t in trace:
        srt = t["router_probs"].sort(dim=-1, descending=True).values
        m = (srt[:, 1] / srt[:, 2]).log() > min_log_ratio
        ok = m if ok is None else (ok & m)
    return ok


def test_full_width_mixtral_two_layers_prefill_and_decode_match_oracle():
    """BASELINE configs[0] shape: text-only, 128-token prompt, bs=1, full layer width (H=4096, I=14336, 8 experts,
    V=51760) at depth 2, greedy decode.

    Random-init routers produce near-ties that bf16 arithmetic (the reference's own bf16 mode included) cannot order
    like the fp32 oracle; a token whose top-2 expert set flips gets a different -- equally valid -- MoE output.  The
    test therefore (1) checks every layer in isolation from the oracle's own layer input, tightly, on all tokens
    whose routing margin is clear (the vast majority), and (2) bounds the end-to-end error statistically."""
    import copy
    from vita_b200.model.mixtral import MixtralDecoder
    from vita_b200.model.vita_mixtral import VITAMixtralForCausalLM
    cfg = VitaConfig.full(num_hidden_layers=2)
    state = W.synthetic_state(cfg, 0, parts=("llm",))
    packed = W.pack_llm(state, cfg, "cuda")
    ids = torch.randint(0, cfg.llm.vocab_size, (1, 128), generator=torch.Generator().manual_seed(0))
    emb = state["model.embed_tokens.weight"].float()[ids]
    trace = []
    ref_logits, past, _ = O.mixtral_forward(state, cfg.llm, emb, trace=trace)

    # (1) layer by layer, each fed the oracle's (bf16-rounded) layer input
    sub_cfg = copy.deepcopy(cfg.llm)
    sub_cfg.num_hidden_layers = 1
    for l in range(2):
        sub = dict(packed)
        sub["layers"] = [packed["layers"][l]]
        dec = MixtralDecoder(sub_cfg, sub, "cuda", max_seq_len=256, max_new_tokens=8)
        h = trace[l]["h_in"][0].to(torch.bfloat16).cuda().contiguous()
        dec.prefill(h, slot=0)                      # the residual stream is updated in place
        got, want = h.float().cpu(), trace[l]["h_out"][0]
        stable = _routing_stable([trace[l]])
        err = (got - want).abs().amax(-1) / want.abs().max()
        print(f"layer {l}: {int(stable.sum())}/128 tokens with clear routing; max rel err stable "
              f"{err[stable].max():.3e}, unstable {err[~stable].max() if (~stable).any() else 0:.3e}")
        assert stable.float().mean() > 0.8
        assert err[stable].max() < 2.5e-2, "decoder layer (attention + MoE) on clearly-routed tokens"

    # (2) end to end through the public surface
    model = VITAMixtralForCausalLM(cfg, {"llm": packed}, "cuda", max_seq_len=256, max_new_tokens=32)
    logits = model(input_ids=ids).logits[0].float().cpu()
    row_err = (logits - ref_logits[0]).abs().amax(-1) / ref_logits.abs().max()
    stable = _routing_stable(trace)
    print(f"prefill logits: median row err {row_err.median():.3e}, stable rows max {row_err[stable].max():.3e}, "
          f"all rows max {row_err.max():.3e}, stable rows {int(stable.sum())}/128")
    assert row_err.median() < 2e-2
    assert row_err[stable].max() < 6e-2     # includes second-order effects of flipped neighbours through attention

    # greedy decode, teacher-forced with the oracle's tokens; steps with a routing near-tie are not compared
    n_new = 8
    toks, rows, step_ok = [], [], []
    lg = ref_logits[:, -1:]
    for _ in range(n_new):
        nxt = int(lg[0, -1].argmax())
        toks.append(nxt)
        rows.append(lg[0, -1])
        tr = []
        lg, past, _ = O.mixtral_forward(state, cfg.llm, state["model.embed_tokens.weight"].float()[torch.tensor([[nxt]])], past=past,
                                        last_only=True, trace=tr)
        step_ok.append(bool(_routing_stable(tr)[0]))
    rows = torch.stack(rows)
    out = model(input_ids=ids)
    got = [out.logits[0, -1].float().cpu()]
    for t in toks[:-1]:
        out = model(input_ids=torch.tensor([[t]]), past_key_values=out.past_key_values)
        got.append(out.logits[0, -1].float().cpu())
    got = torch.stack(got)
    # row 0 comes out of the forward of the last prompt token, row i >= 1 out of decode step i-1: compare the rows
    # whose own forward had clear routing (flips of *other* tokens only reach them through attention, second order)
    usable = torch.tensor([bool(stable[-1])] + [step_ok[i - 1] for i in range(1, n_new)])
    err = (got - rows).abs().amax(-1) / rows.abs().max()
    print(f"decode: usable steps {usable.tolist()}, rel err {err.tolist()}")
    assert usable.float().mean() >= 0.5
    if usable.any():
        assert err[usable].max() < 6e-2
        clear = _margin_ok(rows, 0.05 * rows.abs().max()) & usable
        assert torch.equal(got.argmax(-1)[clear], rows.argmax(-1)[clear])


def _decode_logits(model, ids, toks):
    out = model(input_ids=ids)
    rows = [out.logits[0, -1].float().cpu()]
    for t in toks:
        out = model(input_ids=torch.tensor([[t]]), past_key_values=out.past_key_values)
        rows.append(out.logits[0, -1].float().cpu())
    return torch.stack(rows)


def test_batched_decode_matches_single_sequence_decode(tiny):
    """BASELINE configs[4] shape (concurrent requests, paged KV, one token per request per step): the batched step
    (GEMM path, M = B rows at B different positions) must reproduce each request's own greedy decode."""
    from vita_b200.model.vita_mixtral import VITAMixtralForCausalLM
    cfg, state, _ = tiny
    model = VITAMixtralForCausalLM(cfg, W.pack(state, cfg, "cuda"), "cuda", max_batch=3, max_new_tokens=16,
                                   shuffle_pages=True)
    g = torch.Generator().manual_seed(21)
    reqs = [{"input_ids": torch.randint(0, cfg.llm.vocab_size, (1, n), generator=g)} for n in (9, 17, 30)]
    singles, rows = [], []
    for r in reqs:
        out = model.generate(r["input_ids"], max_new_tokens=6, output_scores=True, use_graph=False)
        singles.append(out.sequences[0, r["input_ids"].shape[1]:].tolist())
        rows.append(torch.cat([s.float().cpu() for s in out.scores]))
    for use_graph in (False, True):
        batch = model.generate_batch(reqs, max_new_tokens=6, use_graph=use_graph)
        for b in range(3):
            clear = _margin_ok(rows[b], 0.03 * rows[b].abs().max())
            n = 0
            while n < 6 and bool(clear[n]):
                n += 1
            assert batch[b][:n] == singles[b][:n], (b, use_graph, batch[b], singles[b], clear.tolist())


def test_load_pretrained_model_from_safetensors_checkpoint(tiny, golden, tmp_path):
    """Checkpoint ingestion through the reference-shaped entry point: HF-style config.json + safetensors shards with
    the reference's parameter names -> same logits as the directly packed weights."""
    import json
    from safetensors.torch import save_file
    from vita_b200.model.builder import load_pretrained_model
    cfg, state, model = tiny
    c, v, a = cfg.llm, cfg.vision, cfg.audio
    hf = {"text_config": {"vocab_size": c.vocab_size, "hidden_size": c.hidden_size,
                          "intermediate_size": c.intermediate_size, "num_hidden_layers": c.num_hidden_layers,
                          "num_attention_heads": c.num_attention_heads, "num_key_value_heads": c.num_key_value_heads,
                          "num_local_experts": 8, "num_experts_per_tok": 2, "rms_norm_eps": c.rms_norm_eps,
                          "rope_theta": c.rope_theta, "max_position_embeddings": c.max_position_embeddings},
          "vision_config": {"hidden_size": v.hidden_size, "intermediate_size": v.intermediate_size,
                            "num_hidden_layers": v.num_hidden_layers, "num_attention_heads": v.num_attention_heads,
                            "image_size": v.image_size, "patch_size": 14, "layer_norm_eps": v.layer_norm_eps},
          "audio_config": {"num_mel_bins": 80, "hidden_size": a.hidden_size, "num_attention_heads": a.num_attention_heads,
                           "intermediate_size": a.linear_units, "num_hidden_layers": a.num_blocks},
          "tokenizer_model_max_length": c.tokenizer_model_max_length}
    (tmp_path / "config.json").write_text(json.dumps(hf))
    names = sorted(state)
    half = len(names) // 2
    save_file({k: state[k].contiguous() for k in names[:half]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k: state[k].contiguous() for k in names[half:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    with pytest.raises(ValueError):
        load_pretrained_model(str(tmp_path), None, "vita", model_type="qwen2p5_instruct")
    tok, loaded, proc, ctx_len = load_pretrained_model(str(tmp_path), None, "vita", model_type="mixtral-8x7b",
                                                       max_new_tokens=16)
    inp, _ = golden
    ids = _t(inp["text_ids"])
    a_logits = loaded(input_ids=ids).logits.float().cpu()
    b_logits = model(input_ids=ids).logits.float().cpu()
    assert torch.equal(a_logits, b_logits)
