"""Decode-step kernels (weight-streaming GEMVs with fused epilogues) vs the oracle's functions."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import randn, to_dev, assert_close, bf16_round, BF16
from oracle import vita_oracle as O

pytestmark = pytest.mark.gpu


def _gpu_randn(shape, seed, scale):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(BF16)


@pytest.mark.parametrize("H,nq,nkv", [(512, 4, 1), (4096, 32, 8)])
def test_decode_qkv_rope(H, nq, nkv):
    from vita_b200 import ops, weights
    B, D, page, max_pages = 2, 128, 16, 8
    h = randn((B, H), 1, 1.5)
    nw = randn((H,), 2)
    w = _gpu_randn(((nq + 2 * nkv) * D, H), 3, 0.03)
    table = weights.rope_table(256, D, 1e6)
    cur = torch.tensor([37, 5], dtype=torch.int32)
    bt = torch.tensor([[3, 1, 4, 0, 5, 2, 6, 7], [15, 14, 13, 12, 11, 10, 9, 8]], dtype=torch.int32)
    q = torch.zeros(B, nq * D, dtype=BF16, device="cuda")
    kc = torch.zeros(16 * page, nkv, D, dtype=BF16, device="cuda")
    vc = torch.zeros_like(kc)
    ops.decode_qkv_rope(to_dev(h), to_dev(nw), w, table.cuda(), cur.cuda(), bt.cuda(), q, kc, vc, nq, nkv, D, page, 1e-5)
    xn = bf16_round(O.rmsnorm(h, nw, 1e-5))
    qkv = bf16_round(O.linear(xn, w.float().cpu()))
    for b in range(B):
        qq = qkv[b, : nq * D].view(1, nq, 1, D)
        kk = qkv[b, nq * D: (nq + nkv) * D].view(1, nkv, 1, D)
        vv = qkv[b, (nq + nkv) * D:].view(nkv, D)
        cos, sin = O.rope_cos_sin(cur[b].view(1, 1).long(), D, 1e6)
        qr, kr = O.apply_rope(qq, kk, cos, sin)
        slot = int(bt[b, int(cur[b]) // page]) * page + int(cur[b]) % page
        assert_close(q[b], qr.reshape(-1), rel=2e-2, what="decode q")
        assert_close(kc[slot], kr.reshape(nkv, D), rel=2e-2, what="decode k cache")
        assert_close(vc[slot], vv, rel=2e-2, what="decode v cache")
    assert int((kc.float().abs().sum(-1).sum(-1) > 0).sum()) == B  # exactly one slot written per sequence


@pytest.mark.parametrize("N,K", [(512, 512), (4096, 4096)])
def test_decode_oproj(N, K):
    from vita_b200 import ops
    B = 2
    x, h = randn((B, K), 1), randn((B, N), 2)
    w = _gpu_randn((N, K), 3, 0.03)
    hd = to_dev(h)
    ops.decode_oproj(to_dev(x), w, hd)
    assert_close(hd, h + O.linear(x, w.float().cpu()), rel=1.2e-2, what="decode oproj")


@pytest.mark.parametrize("H,I", [(512, 1024), (4096, 14336)])
def test_decode_moe(H, I):
    from vita_b200 import ops
    B, E = 2, 8
    h = randn((B, H), 1, 1.5)
    nw = randn((H,), 2)
    gw = randn((E, H), 3, 0.05)
    w13 = _gpu_randn((E, 2 * I, H), 4, 0.03)
    w2 = _gpu_randn((E, H, I), 5, 0.03)
    hd = to_dev(h)
    xn = torch.empty(B, H, dtype=BF16, device="cuda")
    ids = torch.empty(B, 2, dtype=torch.int32, device="cuda")
    tw = torch.empty(B, 2, dtype=torch.float32, device="cuda")
    ops.decode_router(hd, to_dev(nw), to_dev(gw), xn, ids, tw, 1e-5)      # stand-alone router kernel
    ref_xn = bf16_round(O.rmsnorm(h, nw, 1e-5))
    assert_close(xn, ref_xn, rel=8e-3, what="decode xn")
    _, top_v, top_i = O.router_topk(ref_xn, gw)
    assert torch.equal(ids.cpu().long(), top_i), "router ids (random logits have wide margins at this scale)"
    assert (tw.cpu() - top_v).abs().max() < 5e-3
    act = torch.empty(B, 2, I, dtype=BF16, device="cuda")
    ids2 = torch.full((B, 2), -1, dtype=torch.int32, device="cuda")
    tw2 = torch.zeros(B, 2, dtype=torch.float32, device="cuda")
    ops.decode_moe_gate_up(hd, to_dev(nw), to_dev(gw), w13, ids2, tw2, act, 1e-5)   # router fused into the GEMV
    # the fused router takes its logits from the un-rounded normalised activations
    _, top_v2, top_i2 = O.router_topk(O.rmsnorm(h, nw, 1e-5), gw)
    assert torch.equal(ids2.cpu().long(), top_i2) and torch.equal(top_i2, top_i)
    assert (tw2.cpu() - top_v2).abs().max() < 5e-3
    ids, tw = ids2, tw2
    ops.decode_moe_down(act, w2, ids, tw, hd)
    want = h.clone()
    for b in range(B):
        for k in range(2):
            e = int(top_i[b, k])
            gu = O.linear(ref_xn[b], w13[e].float().cpu())
            a = bf16_round(F.silu(gu[:I]) * gu[I:])
            assert_close(act[b, k], a, rel=2e-2, what="decode act")
            want[b] += top_v[b, k] * O.linear(a, w2[e].float().cpu())
    assert_close(hd, want, rel=1.5e-2, what="decode moe out")


@pytest.mark.parametrize("H,V", [(512, 2047), (4096, 51760)])
def test_lm_head_argmax(H, V):
    from vita_b200 import ops
    B = 2
    hrows = randn((B, 3, H), 1, 1.5)     # strided rows: take row 2 of each
    nw = randn((H,), 2)
    w = _gpu_randn((V, H), 3, 0.03)
    hd = to_dev(hrows)
    logits = torch.empty(B, V, dtype=BF16, device="cuda")
    best = torch.zeros(B, dtype=torch.int64, device="cuda")
    ops.lm_head_argmax(hd[:, 2], 3 * H, to_dev(nw), w, logits, best, B, 1e-5)
    xn = bf16_round(O.rmsnorm(hrows[:, 2], nw, 1e-5))
    ref = O.linear(xn, w.float().cpu())
    assert_close(logits, ref, what="logits")
    got_idx = (0xFFFFFFFF - (best.cpu() & 0xFFFFFFFF)).long()
    # arg-max of the kernel's own bf16 logits, first index on ties (torch.argmax semantics on CPU)
    lg = logits.float().cpu()
    for b in range(B):
        mx = lg[b].max()
        first = int((lg[b] == mx).nonzero()[0])
        assert int(got_idx[b]) == first
        # and it is the oracle's arg-max whenever the oracle margin exceeds the bf16 rounding of the logits
        top2 = ref[b].topk(2).values
        if float(top2[0] - top2[1]) > 2e-2 * float(top2[0].abs()):
            assert int(got_idx[b]) == int(ref[b].argmax())


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
    assert log.cpu().tolist() == [[7, -1, -1, -1], [-1, 99, -1, -1], [-1, -1, -1, 0]]
