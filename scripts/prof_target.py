#!/usr/bin/env python
"""Short, deterministic launch sequences for `ncu --set full` captures (one GPU, a handful of launches).

    python scripts/prof_target.py gemm      # S = 4096 prefill GEMMs: qkv, o-proj, MoE gate/up + SiLU, MoE down (3 rounds)
    python scripts/prof_target.py flash     # S = 4096 causal GQA FlashAttention, 8 x 1025 ViT attention (3 rounds)
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vita_b200 import ops  # noqa: E402

BF16 = torch.bfloat16


def gemm():
    dev, S, H, I, E = "cuda", 4096, 4096, 14336, 8
    x = torch.randn(S, H, device=dev).to(BF16)
    wqkv = (torch.randn(6144, H, device=dev) * 0.02).to(BF16)
    wo = (torch.randn(H, H, device=dev) * 0.02).to(BF16)
    qkv = torch.empty(S, 6144, device=dev, dtype=BF16)
    o = torch.empty(S, H, device=dev, dtype=BF16)
    rows = 2 * S
    w13 = torch.empty(E, 2 * I, H, device=dev, dtype=BF16).normal_(0, 0.02)
    w2 = torch.empty(E, H, I, device=dev, dtype=BF16).normal_(0, 0.02)
    cnt = torch.full((E,), rows // E, dtype=torch.int64)
    cnt[0] += 37; cnt[1] -= 37
    offs = torch.cat([torch.zeros(1, dtype=torch.int64), cnt.cumsum(0)]).to(torch.int32).to(dev)
    xp = torch.randn(rows, H, device=dev).to(BF16)
    act = torch.empty(rows, I, device=dev, dtype=BF16)
    yp = torch.empty(rows, H, device=dev, dtype=BF16)
    rw = torch.rand(rows, device=dev)
    for _ in range(3):      # 4 GEMM launches per round: qkv, o-proj, gate/up, down
        ops.linear(x, wqkv, out=qkv)
        ops.linear(x, wo, residual=x, out=o)
        ops.moe_gate_up(xp, w13, act, offs, rows)
        ops.moe_down(act, w2, yp, offs, rw, rows)
    torch.cuda.synchronize()


def flash():
    dev = "cuda"
    S, nq, nkv, D = 4096, 32, 8, 128
    W = (nq + 2 * nkv) * D
    qkv = torch.randn(S, W, device=dev).to(BF16)
    out = torch.empty(S, nq * D, dtype=BF16, device=dev)
    N, Sv, nh, Dv = 8, 1025, 16, 64
    Hv = nh * Dv
    vq = torch.randn(N, Sv, 3 * Hv, device=dev).to(BF16)
    vo = torch.empty(N, Sv, Hv, dtype=BF16, device=dev)
    for _ in range(3):      # 2 attention launches per round
        ops.attention(qkv, qkv[:, nq * D:], qkv[:, (nq + nkv) * D:], out, (0, W, D), (0, W, D), (0, W, D),
                      (0, nq * D, D), 1, nq, nkv, S, S, D, D, None, True, D ** -0.5)
        ops.attention(vq, vq[..., Hv:], vq[..., 2 * Hv:], vo, (Sv * 3 * Hv, 3 * Hv, Dv), (Sv * 3 * Hv, 3 * Hv, Dv),
                      (Sv * 3 * Hv, 3 * Hv, Dv), (Sv * Hv, Hv, Dv), N, nh, nh, Sv, Sv, Dv, Dv, None, False, Dv ** -0.5)
    torch.cuda.synchronize()


if __name__ == "__main__":
    {"gemm": gemm, "flash": flash}[sys.argv[1]]()
