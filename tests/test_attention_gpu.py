"""FlashAttention-style prefill kernel and the paged decode kernel vs explicit fp32 softmax attention."""
import pytest
import torch

from tests.util import randn, to_dev, assert_close, BF16

pytestmark = pytest.mark.gpu


def _ref_attn(q, k, v, scale, causal, kv_lens=None):
    """q [B,Hq,Sq,D], k [B,Hkv,Skv,D], v [B,Hkv,Skv,Dv] -> [B,Hq,Sq,Dv] (fp32)."""
    B, Hq, Sq, _ = q.shape
    Hkv, Skv = k.shape[1], k.shape[2]
    k = k.repeat_interleave(Hq // Hkv, 1)
    v = v.repeat_interleave(Hq // Hkv, 1)
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    if causal:
        s = s.masked_fill(torch.arange(Skv)[None, :] > torch.arange(Sq)[:, None], float("-inf"))
    if kv_lens is not None:
        s = s.masked_fill((torch.arange(Skv)[None, :] >= kv_lens[:, None])[:, None, None, :], float("-inf"))
    return torch.matmul(torch.softmax(s, -1), v)


@pytest.mark.parametrize("S", [1, 63, 64, 65, 128, 129, 200, 513, 1025, 4096])
def test_mixtral_causal_gqa(S):
    from vita_b200 import ops
    nq, nkv, D = 8, 2, 128
    qkv = randn((S, (nq + 2 * nkv) * D), S)
    qd = to_dev(qkv)
    out = torch.empty(S, nq * D, dtype=BF16, device="cuda")
    W = (nq + 2 * nkv) * D
    ops.attention(qd, qd[:, nq * D:], qd[:, (nq + nkv) * D:], out, (0, W, D), (0, W, D), (0, W, D), (0, nq * D, D),
                  1, nq, nkv, S, S, D, D, None, True, D ** -0.5)
    q = qkv[:, : nq * D].view(S, nq, D).transpose(0, 1)[None]
    k = qkv[:, nq * D: (nq + nkv) * D].view(S, nkv, D).transpose(0, 1)[None]
    v = qkv[:, (nq + nkv) * D:].view(S, nkv, D).transpose(0, 1)[None]
    ref = _ref_attn(q, k, v, D ** -0.5, True)[0].transpose(0, 1).reshape(S, nq * D)
    assert_close(out, ref, rel=2e-2, what=f"causal gqa S={S}")


def test_mixtral_causal_gqa_long_128_row_ctas():
    """Long prefill shape: enough CTAs that the 128-query-row variant of the kernel is selected."""
    from vita_b200 import ops
    nq, nkv, D, S = 32, 8, 128, 1300
    qkv = randn((S, (nq + 2 * nkv) * D), 5)
    qd = to_dev(qkv)
    out = torch.empty(S, nq * D, dtype=BF16, device="cuda")
    W = (nq + 2 * nkv) * D
    ops.attention(qd, qd[:, nq * D:], qd[:, (nq + nkv) * D:], out, (0, W, D), (0, W, D), (0, W, D), (0, nq * D, D),
                  1, nq, nkv, S, S, D, D, None, True, D ** -0.5)
    q = qkv[:, : nq * D].view(S, nq, D).transpose(0, 1)[None]
    k = qkv[:, nq * D: (nq + nkv) * D].view(S, nkv, D).transpose(0, 1)[None]
    v = qkv[:, (nq + nkv) * D:].view(S, nkv, D).transpose(0, 1)[None]
    ref = _ref_attn(q, k, v, D ** -0.5, True)[0].transpose(0, 1).reshape(S, nq * D)
    assert_close(out, ref, rel=2e-2, what="causal gqa, 128-row CTAs")


@pytest.mark.parametrize("N,S", [(2, 65), (1, 1025)])
def test_vit_noncausal_packed_qkv(N, S):
    from vita_b200 import ops
    nh, D = 4, 64
    H = nh * D
    qkv = randn((N, S, 3 * H), 3)
    qd = to_dev(qkv)
    out = torch.empty(N, S, H, dtype=BF16, device="cuda")
    ops.attention(qd, qd[..., H:], qd[..., 2 * H:], out, (S * 3 * H, 3 * H, D), (S * 3 * H, 3 * H, D),
                  (S * 3 * H, 3 * H, D), (S * H, H, D), N, nh, nh, S, S, D, D, None, False, D ** -0.5)
    t = qkv.view(N, S, 3, nh, D).permute(2, 0, 3, 1, 4)
    ref = _ref_attn(t[0], t[1], t[2], D ** -0.5, False).transpose(1, 2).reshape(N, S, H)
    assert_close(out, ref, rel=2e-2, what="vit attention")


def test_whale_two_term_with_padding():
    from vita_b200 import ops
    B, T, nh, dk = 2, 70, 2, 64
    q2, k2 = randn((B, T, nh, 2 * dk), 1), randn((B, T, nh, 2 * dk), 2)
    qkv = randn((B, T, 3 * nh * dk), 3)
    lens = torch.tensor([70, 41], dtype=torch.int32)
    out = torch.empty(B, T, nh * dk, dtype=BF16, device="cuda")
    qd, kd, vd = to_dev(q2), to_dev(k2), to_dev(qkv)
    ops.attention(qd, kd, vd[..., 2 * nh * dk:], out, (T * nh * 2 * dk, nh * 2 * dk, 2 * dk),
                  (T * nh * 2 * dk, nh * 2 * dk, 2 * dk), (T * 3 * nh * dk, 3 * nh * dk, dk), (T * nh * dk, nh * dk, dk),
                  B, nh, nh, T, T, 2 * dk, dk, lens.cuda(), False, dk ** -0.5)
    v = qkv[..., 2 * nh * dk:].view(B, T, nh, dk).transpose(1, 2)
    ref = _ref_attn(q2.transpose(1, 2), k2.transpose(1, 2), v, dk ** -0.5, False, lens.long())
    assert_close(out, ref.transpose(1, 2).reshape(B, T, nh * dk), rel=2e-2, what="whale attention")


@pytest.mark.parametrize("ctx,splits", [(1, 4), (17, 4), (130, 8), (700, 16)])
def test_decode_paged(ctx, splits):
    from vita_b200 import ops
    B, nq, nkv, D, page = 2, 8, 2, 128, 16
    n_pages = (ctx + page - 1) // page + 1
    k = randn((B, ctx, nkv, D), 1)
    v = randn((B, ctx, nkv, D), 2)
    q = randn((B, nq, D), 3)
    kc = torch.zeros(B * n_pages * page, nkv, D)
    vc = torch.zeros_like(kc)
    bt = torch.zeros(B, n_pages, dtype=torch.int32)
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(B * n_pages, generator=g)
    for b in range(B):
        for pg in range(n_pages):
            bt[b, pg] = perm[b * n_pages + pg]
        cb = ctx if b == 0 else max(1, ctx - 3)
        for t in range(cb):
            slot = int(bt[b, t // page]) * page + t % page
            kc[slot], vc[slot] = k[b, t], v[b, t]
    cur = torch.tensor([ctx - 1, max(1, ctx - 3) - 1], dtype=torch.int32)
    out = torch.empty(B, nq * D, dtype=BF16, device="cuda")
    ws = ops.decode_attention_workspace(B, nkv, splits, "cuda")
    for _ in range(2):  # second call checks the self-resetting tickets
        ops.decode_attention(to_dev(q), to_dev(kc), to_dev(vc), bt.cuda(), cur.cuda(), out, ws, nq, nkv, D, page,
                             splits, D ** -0.5)
    for b in range(B):
        cb = int(cur[b]) + 1
        ref = _ref_attn(q[b][None, :, None], k[b, :cb].transpose(0, 1)[None], v[b, :cb].transpose(0, 1)[None],
                        D ** -0.5, False)[0, :, 0].reshape(-1)
        assert_close(out[b], ref, rel=2e-2, what=f"decode attention ctx={cb}")
