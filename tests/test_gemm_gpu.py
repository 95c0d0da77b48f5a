"""tcgen05 GEMM (vita_gemm_bf16 + grouped MoE variants) against the oracle's `linear` on the same bf16 inputs."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import randn, to_dev, assert_close, bf16_round
from oracle import vita_oracle as O

pytestmark = pytest.mark.gpu


def _run(M, N, K, bias=False, act=0, colscale=False, residual=False, lda_pad=0, seed=0):
    from vita_b200 import ops
    x = randn((M, K + lda_pad), seed, 1.0)
    w = randn((N, K), seed + 1, 0.05)
    b = randn((N,), seed + 2, 0.5) if bias else None
    cs = randn((N,), seed + 3, 1.0) if colscale else None
    r = randn((M, N), seed + 4, 1.0) if residual else None
    xd = to_dev(x)
    xv = xd[:, :K] if lda_pad else xd
    y = ops.linear(xv, to_dev(w), None if b is None else to_dev(b), act, None if cs is None else to_dev(cs),
                   None if r is None else to_dev(r))
    torch.cuda.synchronize()
    ref = O.linear(x[:, :K], w, b)
    if act == 1:
        ref = F.gelu(ref)
    elif act == 2:
        ref = F.relu(ref)
    if cs is not None:
        ref = ref * cs
    if r is not None:
        ref = ref + r
    assert_close(y, ref, what=f"gemm M={M} N={N} K={K} bias={bias} act={act} cs={colscale} res={residual}")


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 128), (256, 512, 256), (300, 520, 200), (1, 128, 64),
                                   (77, 8, 512), (129, 1000, 72)])
def test_gemm_plain_shapes(M, N, K):
    _run(M, N, K)


def test_gemm_epilogues():
    _run(260, 384, 320, bias=True, act=1)
    _run(260, 384, 320, bias=True, act=2)
    _run(260, 392, 320, bias=True, colscale=True, residual=True)
    _run(200, 256, 128, residual=True, lda_pad=64)


def test_gemm_model_shapes():
    _run(1025, 3072, 1024, bias=True)            # InternViT qkv
    _run(512, 6144, 4096)                        # Mixtral qkv, S=512 (BLOCK_N=128 path)
    _run(2048, 4096, 4096, residual=True)        # o_proj with residual (BLOCK_N=256 path)
    _run(256, 4096, 4096, bias=True, act=1)      # projector


def test_gemm_tail_split_shapes():
    # tiles of the last partial wave are cut along N (gemm_sm100.cu Sched): 148-SM tile counts with a short tail
    _run(2048, 5120, 512, bias=True)             # 256-wide: 320 tiles = 2 waves + 24 -> pieces of 64 columns
    _run(2048, 3328, 512, residual=True)         # 208 tiles = 1 wave + 60 -> pieces of 128 columns
    _run(1025, 3072, 256, bias=True, act=1)      # 128-wide: 216 tiles = 1 wave + 68 -> pieces of 64 columns
    _run(4096, 4096, 1024, residual=True)        # o_proj at S=4096: 512 tiles = 3 waves + 68


def test_gemm_tail_split_is_bit_identical_to_whole_tiles(monkeypatch):
    import os
    from vita_b200 import ops
    x, w = to_dev(randn((2048, 1024), 5, 1.0)), to_dev(randn((5120, 1024), 6, 0.05))
    monkeypatch.setenv("VITA_B200_GEMM_TAIL_SPLIT", "0")
    whole = ops.linear(x, w).clone()
    monkeypatch.setenv("VITA_B200_GEMM_TAIL_SPLIT", "1")
    cut = ops.linear(x, w)
    torch.cuda.synchronize()
    assert torch.equal(whole, cut)


def test_gemm_many_tiles_persistent():
    # more tiles than SMs -> every CTA loops (smem ring phase wrap, both TMEM accumulator stages)
    _run(1536, 8192, 512, bias=True)


def _moe_case(H, I, E, counts, seed=0):
    from vita_b200 import ops
    rows = sum(counts)
    offs = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32)
    x = randn((rows, H), seed, 1.0)
    w13 = randn((E, 2 * I, H), seed + 1, 0.05)
    w2 = randn((E, H, I), seed + 2, 0.05)
    rw = torch.rand(rows, generator=torch.Generator().manual_seed(seed + 3))
    xd, w13d, w2d = to_dev(x), to_dev(w13), to_dev(w2)
    offs_d = offs.cuda()
    act = torch.full((rows, I), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.moe_gate_up(xd, w13d, act, offs_d, rows)
    y = torch.full((rows, H), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.moe_down(act, w2d, y, offs_d, rw.cuda(), rows)
    torch.cuda.synchronize()
    ref_act = torch.zeros(rows, I)
    ref_y = torch.zeros(rows, H)
    for e in range(E):
        a, b = int(offs[e]), int(offs[e + 1])
        if a == b:
            continue
        gu = O.linear(x[a:b], w13[e])
        ref_act[a:b] = F.silu(gu[:, :I]) * gu[:, I:]
        ref_y[a:b] = O.linear(bf16_round(ref_act[a:b]), w2[e]) * rw[a:b, None]
    assert_close(act, ref_act, what="moe gate_up silu")
    assert_close(y, ref_y, rel=2.5e-2, what="moe down")


def test_moe_grouped_small():
    _moe_case(512, 1024, 8, [100, 0, 160, 129, 1, 128, 70, 56])


def test_moe_grouped_full_width():
    _moe_case(4096, 14336, 8, [130, 90, 128, 0, 260, 64, 127, 225])
