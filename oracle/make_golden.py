"""Mint golden vectors from the REFERENCE implementation and pin the oracle against them.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # writes tests/golden/*.npz

For each case the reference's own classes (vita.model.* on top of the installed transformers, CPU, fp32) are
executed with the synthetic weights of vita_b200.weights.synthetic_state(VitaConfig.tiny(), seed); the outputs are
(a) compared with oracle/vita_oracle.py (must agree to fp32 round-off, asserted here) and (b) saved together with
the inputs, so tests/test_oracle_golden.py can re-check the oracle anywhere and the `-m gpu` tests can check the CUDA
path against numbers that came out of the reference itself.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_shim, vita_oracle as O  # noqa: E402
from vita_b200.config import VitaConfig, IMAGE_TOKEN_INDEX, AUDIO_TOKEN_INDEX  # noqa: E402
from vita_b200 import weights as W  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
SEED = 0


# ------------------------------------------------------------------------------------------------ reference builders
def to_hf5_llm_names(state, lcfg):
    """4.41 checkpoint names -> the parameter names of the installed transformers (SURVEY.md Appendix C iii)."""
    out = {}
    for k, v in state.items():
        if not (k.startswith("model.layers.") or k in ("model.embed_tokens.weight", "model.norm.weight", "lm_head.weight")):
            continue
        if ".block_sparse_moe.experts." in k:
            continue
        out[k.replace(".block_sparse_moe.gate.", ".mlp.gate.")] = v.float()
    for l in range(lcfg.num_hidden_layers):
        p = f"model.layers.{l}.block_sparse_moe.experts."
        gu = [torch.cat([state[p + f"{e}.w1.weight"], state[p + f"{e}.w3.weight"]], 0) for e in range(lcfg.num_local_experts)]
        dn = [state[p + f"{e}.w2.weight"] for e in range(lcfg.num_local_experts)]
        out[f"model.layers.{l}.mlp.experts.gate_up_proj"] = torch.stack(gu).float()
        out[f"model.layers.{l}.mlp.experts.down_proj"] = torch.stack(dn).float()
    return out


def build_reference(cfg: VitaConfig, state):
    ref_shim.install()
    import yaml
    from vita.model.language_model.vita_mixtral import VITAMixtralConfig, VITAMixtralForCausalLM
    from vita.model.multimodal_encoder.builder import build_audio_encoder
    from vita.model.multimodal_encoder.internvit.configuration_intern_vit import InternVisionConfig
    from vita.model.multimodal_encoder.internvit.internvit_encoder import InternViTVisionTower
    from vita.model.multimodal_encoder.internvit.modeling_intern_vit import InternVisionModel
    from vita.model.multimodal_projector.builder import build_vision_projector

    c, v, a = cfg.llm, cfg.vision, cfg.audio
    hf_cfg = VITAMixtralConfig(
        vocab_size=c.vocab_size, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
        num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
        num_key_value_heads=c.num_key_value_heads, head_dim=c.head_dim, num_local_experts=c.num_local_experts,
        num_experts_per_tok=c.num_experts_per_tok, rms_norm_eps=c.rms_norm_eps, rope_theta=c.rope_theta,
        max_position_embeddings=c.max_position_embeddings, sliding_window=None, attention_dropout=0.0,
        tie_word_embeddings=False)
    hf_cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    m = VITAMixtralForCausalLM(hf_cfg).float().eval()
    missing, unexpected = m.load_state_dict(to_hf5_llm_names(state, c), strict=False)
    assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
    m.config.tokenizer_model_max_length = c.tokenizer_model_max_length

    # vision tower (SURVEY.md section 8c "Encoder construction without checkpoints")
    vit = InternVisionModel(InternVisionConfig(
        hidden_size=v.hidden_size, image_size=v.image_size, intermediate_size=v.intermediate_size,
        num_attention_heads=v.num_attention_heads, num_hidden_layers=v.num_hidden_layers, patch_size=v.patch_size,
        qk_normalization=False, qkv_bias=True, norm_type="layer_norm", initializer_factor=1.0, use_flash_attn=False,
        layer_norm_eps=v.layer_norm_eps, drop_path_rate=0.0)).float().eval()
    vit.load_state_dict({k[len(W.PREFIX_VISION):]: t.float() for k, t in state.items() if k.startswith(W.PREFIX_VISION)})
    tower = InternViTVisionTower.__new__(InternViTVisionTower)
    torch.nn.Module.__init__(tower)
    tower.is_loaded = True
    tower.select_layer = -1
    tower.scale_pix_shuffle = 0.5
    tower.vision_tower = vit
    m.model.vision_tower = tower
    m.config.mm_hidden_size = v.out_dim
    m.config.mm_projector_type = "mlp2x_gelu"
    m.model.mm_projector = build_vision_projector(m.config).float().eval()
    m.model.mm_projector.load_state_dict({k[len(W.PREFIX_PROJ):]: t.float() for k, t in state.items()
                                          if k.startswith(W.PREFIX_PROJ)})

    # audio encoder: reconstructed train.yaml, dynamic chunks OFF (parity note 8)
    d = tempfile.mkdtemp()
    C = a.hidden_size
    conf = {
        "input_dim": a.input_dim, "is_json_cmvn": True,
        "encoder_conf": {
            "overview_conf": {"encoder-layer-config": "subsampling-transformer", "encoder-input-dim": a.input_dim,
                              "encoder-output-dim": C},
            "para_conf": {
                "subsampling": {"subsampling-rate": 4, "subsampling-input-dim": a.input_dim,
                                "subsampling-output-dim": C, "subsampling-dropout-rate": 0.1},
                "transformer": {"transformer-input-dim": C, "transformer-output-dim": C,
                                "transformer-attention-dim": C, "transformer-attention-heads": a.num_attention_heads,
                                "transformer-linear-units": a.linear_units, "transformer-num-blocks": a.num_blocks,
                                "transformer-dropout-rate": 0.1, "transformer-attention-dropout-rate": 0.0,
                                "transformer-positional-dropout-rate": 0.1, "transformer-input-layer": "linear",
                                "transformer-pos-enc-class": "rel-enc", "transformer-normalize-before": True,
                                "transformer-concat-after": False, "transformer-positionwise-layer-type": "linear",
                                "transformer-chunk_size": -1, "transformer-left_chunks": -1,
                                "transformer-dynamic-chunks": False}}},
        "model_conf": {"llm_path": "", "enc_out_dim": C, "llm_embed_dim": c.hidden_size,
                       "kernel_size": a.adapter_kernel, "adpter_type": "subsampling", "activation_func": "gelu",
                       "norm": "layer"},
        "dataset_conf": {"resample_conf": {"resample_rate": 16000},
                         "fbank_conf": {"num_mel_bins": 80, "frame_length": 25, "frame_shift": 10, "dither": 0.0}},
    }
    with open(os.path.join(d, "train.yaml"), "w") as f:
        yaml.safe_dump(conf, f)
    with open(os.path.join(d, "global_cmvn"), "w") as f:
        json.dump({"mean_stat": [0.0] * a.input_dim, "var_stat": [1.0] * a.input_dim, "frame_num": 1}, f)
    enc = build_audio_encoder(SimpleNamespace(mm_audio_encoder=d)).float().eval()
    enc.load_state_dict({k[len(W.PREFIX_AUDIO):]: t.float() for k, t in state.items() if k.startswith(W.PREFIX_AUDIO)})
    m.model.audio_encoder = enc
    return m


def ref_greedy(m, input_ids, n_new, images=None, audios=None):
    """Manual greedy loop over the reference forward (HF generate() cannot run here: SURVEY.md section 8c shim 2)."""
    with torch.no_grad():
        out = m(input_ids=input_ids, images=images, audios=audios, use_cache=True)
        prefill_logits = out.logits
        pkv = out.past_key_values
        toks, rows = [], []
        logits = out.logits
        for _ in range(n_new):
            row = logits[0, -1]
            rows.append(row)
            nxt = row.argmax(-1, keepdim=True)[None]
            toks.append(int(nxt))
            out = m(input_ids=nxt, past_key_values=pkv, use_cache=True)
            pkv, logits = out.past_key_values, out.logits
    return toks, torch.stack(rows), prefill_logits


# ------------------------------------------------------------------------------------------------ cases
def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def make_inputs(cfg: VitaConfig):
    g = torch.Generator().manual_seed(1234)
    v = cfg.vision
    images = bf16_round(torch.randn(2, 3, v.image_size, v.image_size, generator=g))
    feats = bf16_round(torch.randn(2, 100, cfg.audio.input_dim, generator=g) * 2.0)
    lengths = torch.tensor([100, 77])
    text = torch.randint(0, cfg.llm.vocab_size, (1, 24), generator=g)
    omni = torch.randint(0, cfg.llm.vocab_size, (1, 14), generator=g)
    omni[0, 2] = IMAGE_TOKEN_INDEX
    omni[0, 9] = AUDIO_TOKEN_INDEX
    batch = torch.randint(0, cfg.llm.vocab_size, (3, 10), generator=g)
    batch[0, 1] = IMAGE_TOKEN_INDEX          # image only  -> consumes a dummy audio slot
    batch[1, 4] = AUDIO_TOKEN_INDEX          # audio only  -> consumes a dummy image slot
    batch[1, 7] = AUDIO_TOKEN_INDEX
    # row 2: text only -> consumes one dummy image and one dummy audio slot
    return dict(images=images, feats=feats, lengths=lengths, text_ids=text, omni_ids=omni, batch_ids=batch)


def close(a, b, tol, what):
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    print(f"  {what:36s} max|ref-oracle| = {err:.3e} (max|ref| = {ref:.3e})")
    assert err <= tol * max(1.0, ref), f"oracle disagrees with the reference on {what}: {err}"


def main():
    torch.set_num_threads(8)
    cfg = VitaConfig.tiny()
    state = W.synthetic_state(cfg, SEED)
    m = build_reference(cfg, state)
    inp = make_inputs(cfg)
    out = {}
    with torch.no_grad():
        # 1. vision tower + projector
        tower = m.get_vision_tower()(inp["images"])
        img_feat = m.encode_images(inp["images"])
        close(O.vision_tower(state, cfg.vision, inp["images"]), tower, 2e-5, "vision tower")
        close(O.encode_images(state, cfg, inp["images"]), img_feat, 2e-5, "encode_images")
        out.update(vision_tower=tower, image_features=img_feat)
        # 2. audio encoder + adapter
        aud = m.get_audio_encoder()(inp["feats"], inp["lengths"])
        oa = O.encode_audios(state, cfg, inp["feats"], inp["lengths"])
        close(oa["inputs_embeds"], aud["inputs_embeds"], 2e-5, "audio inputs_embeds")
        assert torch.equal(oa["attention_mask"], aud["attention_mask"])
        out.update(audio_embeds=aud["inputs_embeds"], audio_mask=aud["attention_mask"].to(torch.int32))
        # 3. text-only prefill + greedy
        toks, rows, pre = ref_greedy(m, inp["text_ids"], 8)
        o_logits, _, _ = O.forward(state, cfg, inp["text_ids"])
        close(o_logits, pre, 2e-5, "text prefill logits")
        o_toks, o_rows = O.greedy_generate(state, cfg, inp["text_ids"], max_new_tokens=8)
        close(o_rows, rows, 2e-5, "text decode logits")
        assert o_toks == toks, (o_toks, toks)
        out.update(text_prefill_logits=pre, text_greedy_tokens=torch.tensor(toks), text_decode_logits=rows)
        # 4. omni (image + audio + text) splice, prefill, greedy
        audios = {"audios": inp["feats"][:1], "lengths": inp["lengths"][:1]}
        images = inp["images"][:1]
        spliced = m.prepare_inputs_labels_for_multimodal(inp["omni_ids"], None, None, None, None, images, audios)[4]
        o_emb, _ = O.prepare_inputs_embeds(state, cfg, inp["omni_ids"], images, audios)
        close(o_emb, spliced, 2e-5, "omni spliced inputs_embeds")
        toks, rows, pre = ref_greedy(m, inp["omni_ids"], 6, images, audios)
        o_toks, o_rows = O.greedy_generate(state, cfg, inp["omni_ids"], images, audios, max_new_tokens=6)
        close(o_rows, rows, 2e-5, "omni decode logits")
        assert o_toks == toks, (o_toks, toks)
        out.update(omni_inputs_embeds=spliced, omni_last_logits=pre[:, -1], omni_greedy_tokens=torch.tensor(toks),
                   omni_decode_logits=rows)
        # 5. batched splice with missing modalities (vita_arch.py:240-251,309-316 dummy-slot bookkeeping)
        #    rows: image-only, audio-only (two clips), text-only -> 1+1+1 image slots, 1+2+1 audio slots
        b_images = torch.cat([inp["images"][:1], inp["images"][1:2], inp["images"][:1]])
        b_feats = torch.cat([inp["feats"][:1], inp["feats"][1:2], inp["feats"][:1], inp["feats"][1:2]])
        b_audios = {"audios": b_feats, "lengths": torch.tensor([100, 77, 100, 77])}
        spliced_b = m.prepare_inputs_labels_for_multimodal(inp["batch_ids"], None, None, None, None, b_images, b_audios)[4]
        o_emb_b, lens_b = O.prepare_inputs_embeds(state, cfg, inp["batch_ids"], b_images, b_audios)
        close(o_emb_b, spliced_b, 2e-5, "batched spliced inputs_embeds")
        out.update(batch_inputs_embeds=spliced_b, batch_lens=torch.tensor(lens_b))

    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / "tiny_inputs.npz", **{k: v.numpy() for k, v in inp.items()})
    np.savez_compressed(GOLDEN / "tiny_reference_outputs.npz", **{k: v.float().numpy() for k, v in out.items()})
    meta = {"seed": SEED, "config": cfg.to_dict(), "torch": torch.__version__,
            "generated_by": "oracle/make_golden.py (reference classes on CPU fp32 via oracle/ref_shim.py)"}
    (GOLDEN / "tiny_meta.json").write_text(json.dumps(meta, indent=1))
    print("wrote", [p.name for p in GOLDEN.iterdir()])


if __name__ == "__main__":
    main()
