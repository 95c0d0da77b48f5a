#!/usr/bin/env python
"""BASELINE configs[4] on one GPU: B concurrent audio queries (10 s audio + short prompt each), paged-KV batched
greedy decode; reports p50 / p99 inter-token latency and tokens/s.  (8 x B200 = 8 such replicas.)

    python scripts/bench_cfg5.py [--batch 16] [--new-tokens 256] [--layers 32]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vita_b200 import weights as W  # noqa: E402
from vita_b200.config import VitaConfig, AUDIO_TOKEN_INDEX  # noqa: E402
from vita_b200.model.vita_mixtral import VITAMixtralForCausalLM  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--new-tokens", type=int, default=256)
    ap.add_argument("--layers", type=int, default=32)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = VitaConfig.full(args.layers)
    B, NT = args.batch, args.new_tokens
    packed = W.random_packed(cfg, dev, seed=0, parts=("llm", "audio"))
    model = VITAMixtralForCausalLM(cfg, packed, dev, max_batch=B, max_seq_len=512 + NT, max_new_tokens=NT + 8)
    g = torch.Generator().manual_seed(0)
    reqs = []
    for b in range(B):
        ids = torch.randint(0, cfg.llm.vocab_size, (1, 140 + 2 * b), generator=g)
        ids[0, 3] = AUDIO_TOKEN_INDEX
        reqs.append({"input_ids": ids, "audios": {"audios": torch.randn(1, 998, 80, generator=g),
                                                   "lengths": torch.tensor([998])}})
    # audio-only requests: the splice needs the (dummy) image feature slot like the demo passes (video_audio_demo.py:227)
    for r in reqs:
        r["images"] = None
    def embeds(r):
        a = model.encode_audios(r["audios"]["audios"], r["audios"]["lengths"])["inputs_embeds"]
        ids = r["input_ids"][0].tolist()
        k = ids.index(AUDIO_TOKEN_INDEX)
        tok = lambda t: model._embeds_for(torch.tensor([t]), None, None)[0][0]
        return torch.cat([tok(ids[:k]), a[0], tok(ids[k + 1:])]).contiguous()
    model.llm.reset()
    events = []
    def run(record):
        model.llm.reset()
        for b, r in enumerate(reqs):
            model.llm.prefill(embeds(r), slot=b)
        for s in range(NT):
            model.llm.decode_step_batched(B, use_graph=True)
            if record:
                e = torch.cuda.Event(enable_timing=True); e.record(); events.append(e)
    run(False)
    torch.cuda.synchronize()
    start = torch.cuda.Event(enable_timing=True)
    model.llm.reset()
    for b, r in enumerate(reqs):
        model.llm.prefill(embeds(r), slot=b)
    start.record()
    for s in range(NT):
        model.llm.decode_step_batched(B, use_graph=True)
        e = torch.cuda.Event(enable_timing=True); e.record(); events.append(e)
    torch.cuda.synchronize()
    ts = [start.elapsed_time(e) for e in events]
    itl = sorted(b - a for a, b in zip([0.0] + ts[:-1], ts))
    toks = [model.llm.generated_tokens(b) for b in range(B)]
    assert all(len(t) == NT for t in toks)
    out = {"workload": f"configs[4] on 1 GPU: {B} concurrent audio queries (S0~264-294), {NT} decode steps, "
                       f"{args.layers} layers", "batch": B, "p50_itl_ms": itl[len(itl) // 2],
           "p99_itl_ms": itl[int(len(itl) * 0.99) - 1], "tokens_per_s": B * NT / (ts[-1] / 1e3),
           "launches_per_step": model.llm.launches_per_batched_step}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
