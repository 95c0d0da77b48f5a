"""Per-kernel timeline of one decode step under programmatic dependent launch (no nsys in this image).

Needs the instrumentation build of the library (-DVITA_TRACE, see scripts/build_trace_lib.sh) selected through
VITA_B200_LIB.  Each decode-chain kernel stamps %globaltimer into a 32-word record: CTA 0 in words 1-8, the last CTA in
words 17-24 (1 entry, 2 setup done, 3 dependency wait returned, 4 prologue done, 5 first MMA, 6 last weight tile issued
= dependent-launch trigger, 7 last accumulator ready, 8 exit).  The graph is captured once, replayed a few times, and
the records of the last replay are printed relative to the first kernel's entry (microseconds).
"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vita_b200 import _lib, weights as W                # noqa: E402
from vita_b200.config import VitaConfig                   # noqa: E402
from vita_b200.model.mixtral import MixtralDecoder        # noqa: E402

NAMES = {1: "qkv", 2: "attn", 3: "oproj", 4: "gate_up", 5: "down", 6: "lm_head"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--ctx", type=int, default=506)
    ap.add_argument("--show-layer", type=int, default=3)
    ap.add_argument("--opt", action="append", default=[], help="name=value library option (repeatable)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    if not hasattr(lib, "vita_debug_trace"):
        raise SystemExit("this library was not built with -DVITA_TRACE")
    from vita_b200 import ops
    for o in args.opt:
        k, v = o.split("=")
        ops.set_option(k, int(v))
    cfg = VitaConfig.full(args.layers)
    packed = W.random_packed(cfg, dev, seed=0, parts=("llm",))
    llm = MixtralDecoder(cfg.llm, packed["llm"], dev, max_batch=1, max_seq_len=args.ctx + 128, max_new_tokens=64)
    g = torch.Generator(device=dev).manual_seed(1)
    emb = (torch.randn(args.ctx, cfg.llm.hidden_size, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    llm.prefill(emb, slot=0)
    llm.decode_step(1, use_graph=False)            # warm-up (function attributes)
    torch.cuda.synchronize()
    n_rec = 4096
    buf = torch.zeros(n_rec, 32, dtype=torch.int64, device=dev)
    lib.vita_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int64]
    lib.vita_debug_trace(buf.data_ptr(), n_rec)     # serial restarts: the captured launches use records 0..
    llm._graphs = {}
    for _ in range(4):
        llm.decode_step(1, use_graph=True)          # first call: eager warm-up + capture, then replays
    torch.cuda.synchronize()
    lib.vita_debug_trace(None, 0)
    rec = buf.cpu()
    rows = [r for r in rec.tolist() if r[0] in NAMES and r[1] > 0]
    # the eager warm-up inside decode_step used the first records; keep the last full chain
    per_step = 5 * args.layers + 1
    rows = rows[-per_step:]
    t0 = rows[0][1]
    us = lambda t: (t - t0) / 1000.0 if t else float("nan")
    print(f"{'kernel':10s} {'entry':>8s} {'setup':>8s} {'waited':>8s} {'prolog':>8s} {'1stMMA':>8s} {'trigger':>8s} "
          f"{'lastAcc':>8s} {'exit':>8s} | last CTA: {'entry':>8s} {'waited':>8s} {'exit':>8s}")
    lo, hi = 5 * args.show_layer, 5 * (args.show_layer + 1) + 1
    for r in rows[lo:hi] + rows[-1:]:
        f = [us(r[i]) for i in range(1, 9)]
        l = [us(r[17]), us(r[19]), us(r[24])]
        print(f"{NAMES[r[0]]:10s} " + " ".join(f"{x:8.2f}" for x in f) + "  |           " +
              " ".join(f"{x:8.2f}" for x in l))
    for r in rows[lo:hi]:
        if r[0] == 2:
            arr = [us(r[9 + i]) for i in range(16)]
            print("attn: split arrival times:", " ".join(f"{x:.2f}" for x in arr))
            print(f"attn: collector start {us(r[25]):.2f}  gathered {us(r[27]):.2f}  exit {us(r[8]):.2f}")
    # per-layer summary: time from one qkv entry to the next
    q = [r[1] for r in rows if r[0] == 1]
    if len(q) > 2:
        d = [(b - a) / 1000.0 for a, b in zip(q, q[1:])]
        print("layer-to-layer (qkv entry to next qkv entry), us:", " ".join(f"{x:.1f}" for x in d))


if __name__ == "__main__":
    main()
