"""Row kernels (norms, RoPE + paged KV write, gather/scatter, router, align, combine, conv glue) vs the oracle."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.util import randn, to_dev, assert_close, bf16_round, BF16
from oracle import vita_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,H", [(3, 512), (70, 4096), (5, 1024)])
def test_rmsnorm(rows, H):
    from vita_b200 import ops
    x, w = randn((rows, H), 1, 2.0), randn((H,), 2, 1.0)
    y = ops.rmsnorm(to_dev(x), to_dev(w), 1e-5)
    assert_close(y, O.rmsnorm(x, w, 1e-5), what="rmsnorm")


@pytest.mark.parametrize("rows,H,act,scale,eps", [(9, 1024, 0, 1.0, 1e-6), (33, 128, 2, math.sqrt(128), 1e-5),
                                                  (12, 2048, 1, 1.0, 1e-3), (4, 256, 0, 1.0, 1e-5)])
def test_layernorm(rows, H, act, scale, eps):
    from vita_b200 import ops
    x, w, b = randn((rows, H), 1, 2.0), randn((H,), 2, 1.0), randn((H,), 3, 0.3)
    y = ops.layernorm(to_dev(x), to_dev(w), to_dev(b), eps, act, scale)
    ref = F.layer_norm(x, (H,), w, b, eps)
    ref = F.gelu(ref) if act == 1 else F.relu(ref) if act == 2 else ref
    assert_close(y, ref * scale, what="layernorm")


def test_row_copy_gather_scatter():
    from vita_b200 import ops
    table = randn((50, 256), 1)
    ids = torch.tensor([3, 49, -200, 0, 7, -500, 3], dtype=torch.int32)
    dst = torch.tensor([6, 5, 4, 3, 2, 1, 0], dtype=torch.int32)
    out = torch.zeros(7, 256, dtype=BF16, device="cuda")
    ops.row_copy(to_dev(table), ids.cuda(), dst.cuda(), out, 7)
    ref = torch.zeros(7, 256)
    for i, (s, d) in enumerate(zip(ids.tolist(), dst.tolist())):
        if s >= 0:
            ref[d] = table[s]
    assert torch.equal(out.float().cpu(), ref)


def test_rope_kv_write():
    from vita_b200 import ops, weights
    n_q, n_kv, D, S = 4, 1, 128, 37
    qkv = randn((S, (n_q + 2 * n_kv) * D), 1)
    pos = torch.arange(5, 5 + S, dtype=torch.int32)
    page = 16
    n_pages = 8
    perm = torch.randperm(n_pages, generator=torch.Generator().manual_seed(0))
    slots = torch.tensor([int(perm[(p // page)]) * page + p % page for p in range(S)], dtype=torch.int32)
    table = weights.rope_table(256, D, 1e6)
    kc = torch.zeros(n_pages * page, n_kv, D, dtype=BF16, device="cuda")
    vc = torch.zeros_like(kc)
    qd = to_dev(qkv)
    ops.rope_kv_write(qd, pos.cuda(), slots.cuda(), table.cuda(), kc, vc, n_q, n_kv, D)
    q = qkv[:, : n_q * D].view(S, n_q, D).transpose(0, 1)[None]
    k = qkv[:, n_q * D: (n_q + n_kv) * D].view(S, n_kv, D).transpose(0, 1)[None]
    v = qkv[:, (n_q + n_kv) * D:].view(S, n_kv, D)
    cos, sin = O.rope_cos_sin(pos[None].long(), D, 1e6)
    qr, kr = O.apply_rope(q, k, cos, sin)
    got = qd.float().cpu()
    assert_close(got[:, : n_q * D].view(S, n_q, D), qr[0].transpose(0, 1), rel=1e-2, what="rope q")
    assert_close(got[:, n_q * D: (n_q + n_kv) * D].view(S, n_kv, D), kr[0].transpose(0, 1), rel=1e-2, what="rope k")
    assert_close(kc.float().cpu()[slots.long()], kr[0].transpose(0, 1), rel=1e-2, what="k cache")
    assert torch.equal(vc.float().cpu()[slots.long()], v)


def test_router_align_combine():
    from vita_b200 import ops
    T, H, E = 301, 512, 8
    h = randn((T, H), 1, 1.5)
    nw = randn((H,), 2, 1.0)
    gw = randn((E, H), 3, 0.3)
    xn = torch.empty(T, H, dtype=BF16, device="cuda")
    ids = torch.empty(T, 2, dtype=torch.int32, device="cuda")
    tw = torch.empty(T, 2, dtype=torch.float32, device="cuda")
    ops.moe_router(to_dev(h), to_dev(nw), to_dev(gw), xn, ids, tw, 1e-5)
    ref_xn = O.rmsnorm(h, nw, 1e-5)
    assert_close(xn, ref_xn, what="router xn")
    probs, top_v, top_i = O.router_topk(bf16_round(ref_xn), gw)
    ids_c, tw_c = ids.cpu().long(), tw.cpu()
    # exact expert agreement wherever the oracle's decision margin exceeds the bf16 noise floor
    srt = probs.sort(dim=-1, descending=True).values
    safe = ((srt[:, 1] - srt[:, 2]) > 2e-3) & ((srt[:, 0] - srt[:, 1]) > 2e-3)
    assert safe.float().mean() > 0.5
    assert torch.equal(ids_c[safe], top_i[safe])
    assert (tw_c[safe] - top_v[safe]).abs().max() < 5e-3
    assert ((tw_c.sum(-1) - 1).abs() < 1e-5).all()

    offs = torch.empty(E + 1, dtype=torch.int32, device="cuda")
    perm = torch.empty(T * 2, dtype=torch.int32, device="cuda")
    rtok = torch.empty(T * 2, dtype=torch.int32, device="cuda")
    rw = torch.empty(T * 2, dtype=torch.float32, device="cuda")
    ops.moe_align(ids, tw, offs, perm, rtok, rw, T, E)
    flat = ids_c.reshape(-1)
    counts = torch.bincount(flat, minlength=E)
    assert offs.cpu().tolist() == [0] + counts.cumsum(0).tolist()
    # stable counting sort == torch.sort(stable=True)
    order = torch.sort(flat, stable=True).indices
    want_perm = torch.empty(T * 2, dtype=torch.long)
    want_perm[order] = torch.arange(T * 2)
    assert torch.equal(perm.cpu().long(), want_perm)
    assert torch.equal(rtok.cpu().long(), order // 2)
    assert torch.equal(rw.cpu(), tw_c.reshape(-1)[order])

    yp = randn((T * 2, H), 7)
    hres = to_dev(h)
    nxt = randn((H,), 8)
    xn2 = torch.empty(T, H, dtype=BF16, device="cuda")
    ops.moe_combine(hres, to_dev(yp), perm, to_dev(nxt), xn2, 1e-5)
    want_h = h + yp[want_perm.view(T, 2)[:, 0]] + yp[want_perm.view(T, 2)[:, 1]]
    assert_close(hres, want_h, rel=8e-3, what="combine")
    assert_close(xn2, O.rmsnorm(bf16_round(want_h), nxt, 1e-5), what="combine rmsnorm")


def test_vit_glue(golden):
    from vita_b200 import ops
    n, C, HW, P, H = 2, 3, 112, 14, 128
    img = randn((n, C, HW, HW), 1)
    g = HW // P
    kpad = 592
    col = torch.empty(n * g * g, kpad, dtype=BF16, device="cuda")
    ops.vit_im2col(to_dev(img), col, P, kpad)
    ref = F.unfold(img, P, stride=P).transpose(1, 2).reshape(n * g * g, C * P * P)
    got = col.float().cpu()
    assert torch.equal(got[:, : C * P * P], ref) and (got[:, C * P * P:] == 0).all()

    patches, cls, pos = randn((n * g * g, H), 2), randn((H,), 3), randn((g * g + 1, H), 4)
    out = torch.empty(n, g * g + 1, H, dtype=BF16, device="cuda")
    ops.vit_assemble(to_dev(patches), to_dev(cls), to_dev(pos), out, n, g * g, H)
    want = torch.cat([cls.expand(n, 1, H), patches.view(n, g * g, H)], 1) + pos
    assert_close(out, want, rel=8e-3, what="vit assemble")

    hfull = randn((n, g * g + 1, H), 5)
    ps = torch.empty(n, g * g // 4, 4 * H, dtype=BF16, device="cuda")
    ops.vit_pixel_shuffle(to_dev(hfull), ps, n, g, H, 0.5)
    want = O.pixel_shuffle(hfull[:, 1:].reshape(n, g, g, H) * 0.5).reshape(n, -1, 4 * H)
    assert torch.equal(ps.float().cpu(), want)


def test_whale_glue():
    from vita_b200 import ops
    B, T, Fd, C = 2, 61, 80, 128
    feat = randn((B, T, Fd), 1, 2.0)
    mean, istd = randn((Fd,), 2, 0.5), 1.0 + randn((Fd,), 3, 0.1).abs()
    w1, b1 = randn((C, 1, 3, 3), 4, 0.3), randn((C,), 5, 0.1)
    T1, F1 = (T - 1) // 2, (Fd - 1) // 2
    out1 = torch.empty(B, T1, F1, C, dtype=BF16, device="cuda")
    ops.whale_conv1(feat.cuda(), mean.cuda(), istd.cuda(), to_dev(w1.reshape(C, 9)), to_dev(b1), out1)
    ref1 = F.relu(F.conv2d(((feat - mean) * istd).unsqueeze(1), w1, b1, stride=2))      # [B, C, T1, F1]
    assert_close(out1, ref1.permute(0, 2, 3, 1), rel=2e-2, what="whale conv1")

    x1 = bf16_round(ref1.permute(0, 2, 3, 1).contiguous())
    T2, F2 = (T1 - 1) // 2, (F1 - 1) // 2
    col = torch.empty(B * T2 * F2, 9 * C, dtype=BF16, device="cuda")
    ops.whale_im2col2(to_dev(x1), col, B, T1, F1, C)
    w2 = randn((C, C, 3, 3), 6, 0.05)
    y = O.linear(col.float().cpu(), w2.permute(0, 2, 3, 1).reshape(C, 9 * C))
    ref2 = F.conv2d(x1.permute(0, 3, 1, 2), w2, stride=2)                                # [B, C, T2, F2]
    assert_close(y.view(B, T2, F2, C), ref2.permute(0, 2, 3, 1), rel=1e-4, what="whale im2col2")

    heads, dk = 2, 64
    qkv, p = randn((B * T2, 3 * C), 7), randn((T2, C), 8)
    bu, bv = randn((C,), 9, 0.2), randn((C,), 10, 0.2)
    q2 = torch.empty(B * T2, heads, 2 * dk, dtype=BF16, device="cuda")
    k2 = torch.empty_like(q2)
    ops.whale_qk_prep(to_dev(qkv), to_dev(p), to_dev(bu), to_dev(bv), q2, k2, B, T2, heads, dk)
    q = qkv[:, :C].view(B * T2, heads, dk)
    want_q = torch.cat([q + bu.view(heads, dk), q + bv.view(heads, dk)], -1)
    want_k = torch.cat([qkv[:, C:2 * C].view(B * T2, heads, dk), p.repeat(B, 1).view(B * T2, heads, dk)], -1)
    assert_close(q2, want_q, rel=8e-3, what="whale q2")
    assert torch.equal(k2.float().cpu(), want_k)

    x = randn((B, T2, C), 11)
    lens = torch.tensor([T2, T2 - 5], dtype=torch.int32)
    T3 = (T2 - 1) // 2 + 1
    colA = torch.empty(B * T3, 5 * C, dtype=BF16, device="cuda")
    ops.whale_adapter_im2col(to_dev(x), lens.cuda(), colA, B, T2, C, 5)
    wc = randn((2 * C, C, 5), 12, 0.05)
    y = O.linear(colA.float().cpu(), wc.permute(0, 2, 1).reshape(2 * C, 5 * C))
    mask = (torch.arange(T2)[None] < lens[:, None]).unsqueeze(1)
    xm = x.transpose(1, 2).masked_fill(~mask, 0.0)
    ref = F.conv1d(F.pad(xm, (0, 4)), wc, stride=2).transpose(1, 2)
    assert_close(y.view(B, T3, 2 * C), ref, rel=1e-4, what="whale adapter im2col")


@pytest.mark.parametrize("T", [1, 7, 512, 513, 4096, 4600])
def test_moe_align_is_a_stable_counting_sort(T):
    from vita_b200 import ops
    E = 8
    g = torch.Generator().manual_seed(T)
    ids_c = torch.stack([torch.randperm(E, generator=g)[:2] for _ in range(T)]).to(torch.int32)
    if T > 100:
        ids_c[T // 3: T // 2, 0] = 5          # a crowded expert and (below) an empty one
        ids_c[ids_c == 2] = 6
        ids_c[:, 1] = torch.where(ids_c[:, 1] == ids_c[:, 0], (ids_c[:, 0] + 1) % E, ids_c[:, 1])
    tw_c = torch.rand(T, 2, generator=g)
    ids, tw = ids_c.cuda(), tw_c.cuda()
    offs = torch.empty(E + 1, dtype=torch.int32, device="cuda")
    perm = torch.empty(T * 2, dtype=torch.int32, device="cuda")
    rtok = torch.empty(T * 2, dtype=torch.int32, device="cuda")
    rw = torch.empty(T * 2, dtype=torch.float32, device="cuda")
    ra = torch.empty(T * 2, dtype=torch.int32, device="cuda")
    ops.moe_align(ids, tw, offs, perm, rtok, rw, T, E, row_assign=ra)
    flat = ids_c.reshape(-1).long()
    assert offs.cpu().tolist() == [0] + torch.bincount(flat, minlength=E).cumsum(0).tolist()
    order = torch.sort(flat, stable=True).indices
    want_perm = torch.empty(T * 2, dtype=torch.long)
    want_perm[order] = torch.arange(T * 2)
    assert torch.equal(perm.cpu().long(), want_perm)
    assert torch.equal(rtok.cpu().long(), order // 2)
    assert torch.equal(ra.cpu().long(), order)
    assert torch.equal(rw.cpu(), tw_c.reshape(-1)[order])
