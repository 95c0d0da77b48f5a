"""Same-box, interleaved A/B of decode-step tunables (ms/token at bs=1, ctx 506.., full 32 layers).

Boxes and power states differ by ~1-2 %, so configurations are compared inside ONE process: the weights are built once,
then the configurations take turns (round robin, several rounds); each turn re-captures the CUDA graph under its
options and times NT decode steps with CUDA events.  Usage: python scripts/decode_ab.py [--rounds 3] [--new-tokens 128]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vita_b200 import ops, weights as W          # noqa: E402
from vita_b200.config import VitaConfig            # noqa: E402
from vita_b200.model.mixtral import MixtralDecoder  # noqa: E402

BASE = {"pdl": 1, "attn_early": 1, "attn_tagged": 1, "chain_wait": 1, "tc_prefetch_consts": 1, "tc_l2_ahead": 0,
        "tc_trigger_lead": 0, "tc_wide_route": 1, "smem_carveout_max": 0, "chain_counters": 0, "tc_park": 1}
CONFIGS = {
    # name: (library options on top of BASE, decoder attributes)
    "nopdl": ({"pdl": 0}, {}),
    "default": ({}, {}),
    "counters_on": ({"chain_counters": 1}, {}),
    "park_off": ({"tc_park": 0}, {}),
    "early_route_on": ({}, {"early_route": True}),
    "early_route_off": ({}, {"early_route": False}),
    "early_route_lead4": ({"tc_trigger_lead": 4}, {}),
    "carveout_default": ({"smem_carveout_max": 0}, {}),
    "lead4": ({"tc_trigger_lead": 4}, {}),
    "l2a4": ({"tc_l2_ahead": 4}, {}),
    "attn_tickets": ({"attn_tagged": 0}, {}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=506)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = VitaConfig.full(args.layers)
    packed = W.random_packed(cfg, dev, seed=0, parts=("llm",))
    llm = MixtralDecoder(cfg.llm, packed["llm"], dev, max_batch=1, max_seq_len=args.ctx + args.new_tokens + 64,
                         max_new_tokens=args.new_tokens + 16)
    g = torch.Generator(device=dev).manual_seed(1)
    emb = (torch.randn(args.ctx, cfg.llm.hidden_size, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    names = [n for n in CONFIGS if not args.only or n in args.only.split(",")]
    times = {n: [] for n in names}
    toks = {}
    for r in range(args.rounds + 1):          # round 0 = warm-up
        for n in names:
            opts, attrs = CONFIGS[n]
            opts = {**BASE, **opts}
            attrs = {"decode_splits": 16, "early_route": False, **attrs}
            for k, v in opts.items():
                ops.set_option(k, v)
            for k, v in attrs.items():
                setattr(llm, k, v)
            llm._graphs = {}
            llm.attn_ws.zero_()     # the two split-merge protocols use the same buffer differently
            llm.reset()
            llm.prefill(emb.clone(), slot=0)
            llm.decode_step(1, use_graph=True)   # capture + first token
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.new_tokens):
                llm.decode_step(1, use_graph=True)
            b.record()
            b.synchronize()
            if r > 0:
                times[n].append(a.elapsed_time(b) / args.new_tokens)
            t = llm.token_log[0, : args.new_tokens + 1].tolist()
            toks.setdefault(n, t)
            assert toks[n] == t, f"{n}: tokens changed between rounds"
    ref = toks[names[0]]
    out = {}
    for n in names:
        ts = sorted(times[n])
        out[n] = {"ms_per_token_min": round(ts[0], 4), "ms_per_token_med": round(ts[len(ts) // 2], 4),
                  "tokens_equal_to_first_config": toks[n] == ref}
        print(f"{n:24s} min {ts[0]:.4f}  med {ts[len(ts) // 2]:.4f}  all {['%.4f' % t for t in times[n]]}  "
              f"same-tokens {toks[n] == ref}", flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
