"""Decode-attention workspace sharing: the tagged all-to-all merge and the ticket merge alternate on one buffer."""
import pytest
import torch

from tests.util import randn, to_dev, assert_close, BF16
from tests.test_attention_gpu import _ref_attn

pytestmark = pytest.mark.gpu


def test_decode_paged_variants_share_one_workspace():
    """The two split-merge protocols (tagged all-to-all words / ticket + last CTA) alternate on ONE workspace with
    changing split counts: each launch must hand the buffer back all-zero, and every result must match the oracle."""
    from vita_b200 import ops
    B, nq, nkv, D, page, ctx = 2, 8, 2, 128, 16, 300
    n_pages = (ctx + page - 1) // page + 1
    k, v, q = randn((B, ctx, nkv, D), 11), randn((B, ctx, nkv, D), 12), randn((B, nq, D), 13)
    kc = torch.zeros(B * n_pages * page, nkv, D)
    vc = torch.zeros_like(kc)
    bt = torch.arange(B * n_pages, dtype=torch.int32).view(B, n_pages)
    for b in range(B):
        for t in range(ctx):
            slot = int(bt[b, t // page]) * page + t % page
            kc[slot], vc[slot] = k[b, t], v[b, t]
    cur = torch.tensor([ctx - 1, ctx - 1], dtype=torch.int32)
    ws = ops.decode_attention_workspace(B, nkv, 16, "cuda")
    want = [_ref_attn(q[b][None, :, None], k[b].transpose(0, 1)[None], v[b].transpose(0, 1)[None], D ** -0.5, False)[0, :, 0]
            .reshape(-1) for b in range(B)]
    before = ops.get_option("attn_tagged")
    try:
        # (tagged?, splits): 12 splits always takes the ticket kernel; attn_tagged = 0 forces it for 8 and 16
        for tagged, splits in [(1, 16), (1, 12), (1, 8), (0, 16), (1, 16), (0, 8), (1, 4), (1, 12), (1, 16)]:
            ops.set_option("attn_tagged", tagged)
            out = torch.zeros(B, nq * D, dtype=BF16, device="cuda")
            ops.decode_attention(to_dev(q), to_dev(kc), to_dev(vc), bt.cuda(), cur.cuda(), out, ws, nq, nkv, D, page,
                                 splits, D ** -0.5)
            for b in range(B):
                assert_close(out[b], want[b], rel=2e-2, what=f"decode attention tagged={tagged} splits={splits}")
            assert int(ws.count_nonzero()) == 0, f"workspace not handed back zeroed (tagged={tagged}, splits={splits})"
    finally:
        ops.set_option("attn_tagged", before)
