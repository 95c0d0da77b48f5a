"""Small decode-step kernels (embed bookkeeping, stand-alone router) vs the oracle's functions; the GEMVs themselves:
tests/test_decode_tc_gpu.py."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import randn, to_dev, assert_close, bf16_round, BF16
from oracle import vita_oracle as O

pytestmark = pytest.mark.gpu


def _gpu_randn(shape, seed, scale):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(BF16)


@pytest.mark.parametrize("H", [512, 4096])
def test_decode_router(H):
    """stand-alone post-attention RMSNorm + router kernel (the decode chain uses the form fused into the gate/up GEMV,
    tests/test_decode_tc_gpu.py)"""
    from vita_b200 import ops
    B, E = 2, 8
    h = randn((B, H), 1, 1.5)
    nw = randn((H,), 2)
    gw = randn((E, H), 3, 0.05)
    hd = to_dev(h)
    xn = torch.empty(B, H, dtype=BF16, device="cuda")
    ids = torch.empty(B, 2, dtype=torch.int32, device="cuda")
    tw = torch.empty(B, 2, dtype=torch.float32, device="cuda")
    ops.decode_router(hd, to_dev(nw), to_dev(gw), xn, ids, tw, 1e-5)
    ref_xn = bf16_round(O.rmsnorm(h, nw, 1e-5))
    assert_close(xn, ref_xn, rel=8e-3, what="decode xn")
    _, top_v, top_i = O.router_topk(ref_xn, gw)
    assert torch.equal(ids.cpu().long(), top_i), "router ids (random logits have wide margins at this scale)"
    assert (tw.cpu() - top_v).abs().max() < 5e-3


def test_decode_embed_bookkeeping():
    from vita_b200 import ops
    B, H, V = 3, 256, 100
    embed = randn((V, H), 1)
    toks = [7, 99, 0]
    best = torch.tensor([(123 << 32) | (0xFFFFFFFF - t) for t in toks], dtype=torch.int64).cuda()
    log = torch.full((B, 4), -1, dtype=torch.int32, device="cuda")
    cnt = torch.tensor([0, 1, 3], dtype=torch.int32, device="cuda")
    clen = torch.tensor([10, 20, 30], dtype=torch.int32, device="cuda")
    cur = torch.zeros(B, dtype=torch.int32, device="cuda")
    h = torch.zeros(B, H, dtype=BF16, device="cuda")
    ops.decode_embed(best, log, cnt, clen, cur, to_dev(embed), h, 31)    # KV capacity 31: the third sequence is full - 1
    assert torch.equal(h.float().cpu(), embed[toks])
    assert cur.cpu().tolist() == [10, 20, 30] and clen.cpu().tolist() == [11, 21, 31]
    # a full cache saturates instead of indexing past the block-table row / rope table
    best.copy_(torch.tensor([(1 << 32) | (0xFFFFFFFF - 1)] * B, dtype=torch.int64))
    ops.decode_embed(best, log, cnt, clen, cur, to_dev(embed), h, 31)
    assert cur.cpu().tolist() == [11, 21, 30] and clen.cpu().tolist() == [12, 22, 31]
    cnt.copy_(torch.tensor([1, 2, 4], dtype=torch.int32)); clen.copy_(torch.tensor([11, 21, 31], dtype=torch.int32))
    assert cnt.cpu().tolist() == [1, 2, 4] and best.cpu().tolist() == [0, 0, 0]
    # both calls logged their token (the second one token 1 at counts 1, 2; count 4 is past the 4-entry log)
    assert log.cpu().tolist() == [[7, 1, -1, -1], [-1, 99, 1, -1], [-1, -1, -1, 0]]
