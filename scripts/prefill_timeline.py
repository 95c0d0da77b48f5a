"""Per-op timeline of the Mixtral prefill (CUDA events between the launches, so inter-kernel gaps are charged to the op
that follows): where do the microseconds of a layer go at S = 506 / 4096?   usage: python scripts/prefill_timeline.py [S] [layers]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vita_b200 import ops, weights as W            # noqa: E402
from vita_b200.config import VitaConfig              # noqa: E402
from vita_b200.model import mixtral as M             # noqa: E402


class Timed:
    def __init__(self, real):
        self.real, self.marks = real, []

    def __getattr__(self, name):
        f = getattr(self.real, name)
        if not callable(f):
            return f

        def g(*a, **k):
            r = f(*a, **k)
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            out = k.get("out")
            self.marks.append((f"{name}[N={out.shape[-1]}]" if name == "linear" and out is not None else name, e))
            return r
        return g


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 506
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    dev = torch.device("cuda", 0)
    cfg = VitaConfig.full(L)
    packed = W.random_packed(cfg, dev, seed=0, parts=("llm",))
    llm = M.MixtralDecoder(cfg.llm, packed["llm"], dev, max_batch=1, max_seq_len=S + 64, max_new_tokens=8)
    emb = (torch.randn(S, cfg.llm.hidden_size, device=dev) * 0.05).to(torch.bfloat16)
    for _ in range(3):
        llm.reset(); llm.prefill(emb.clone(), slot=0)
    torch.cuda.synchronize()
    timed = Timed(ops)
    M.ops = timed
    agg = collections.OrderedDict()
    reps = 5
    tot = 0.0
    for _ in range(reps):
        llm.reset()
        x = emb.clone()
        timed.marks = []
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        llm.prefill(x, slot=0)
        torch.cuda.synchronize()
        prev = e0
        for name, e in timed.marks:
            d = agg.setdefault(name, [0, 0.0])
            d[0] += 1; d[1] += prev.elapsed_time(e) * 1e3
            prev = e
        tot += e0.elapsed_time(prev) * 1e3
    M.ops = ops
    # untimed reference (no events between the ops)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    llm.reset(); x = emb.clone(); a.record(); llm.prefill(x, slot=0); b.record(); torch.cuda.synchronize()
    print(f"prefill S={S} L={L}: {tot / reps:.1f} us with events, {a.elapsed_time(b) * 1e3:.1f} us without; per layer "
          f"{tot / reps / L:.1f} us")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {name:28s} {n // reps:4d} calls  {us / n:9.2f} us/call  {us / reps / L:9.2f} us/layer  {100 * us / tot:5.1f}%")


if __name__ == "__main__":
    main()
