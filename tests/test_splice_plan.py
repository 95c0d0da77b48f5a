"""Host-side splice index arithmetic (plan_splice) vs the oracle's restatement of the reference's Python loop."""
import pytest
import torch

from oracle import vita_oracle as O
from vita_b200.config import VitaConfig, IMAGE_TOKEN_INDEX as IMG, AUDIO_TOKEN_INDEX as AUD
from vita_b200.model.vita_mixtral import plan_splice


def _apply(plan, embed, img, aud, B):
    H = embed.shape[1]
    out = torch.zeros(B * plan.max_len, H)
    out[plan.text_dst] = embed[plan.text_src]
    if plan.img_src:
        out[plan.img_dst] = img.reshape(-1, H)[plan.img_src]
    if plan.aud_src:
        out[plan.aud_dst] = aud.reshape(-1, H)[plan.aud_src]
    return out.view(B, plan.max_len, H)


CASES = [
    [[5, IMG, 7, 8, AUD, 9]],
    [[IMG, 1, 2], [3, AUD, 4, AUD], [6, 7, 8, 9, 10]],
    [[1, 2, 3]],
    [[IMG, IMG, AUD, 4], [AUD, IMG, 1, 2]],
    [[AUD], [IMG]],
]


@pytest.mark.parametrize("ids", CASES)
@pytest.mark.parametrize("max_len", [None, 9])
def test_plan_matches_oracle_splice(ids, max_len):
    cfg = VitaConfig.tiny()
    cfg.llm.tokenizer_model_max_length = max_len
    H, V, TI, TA = 16, 64, 4, 3
    g = torch.Generator().manual_seed(0)
    n_img = sum(r.count(IMG) for r in ids) + sum(IMG not in r for r in ids)
    n_aud = sum(r.count(AUD) for r in ids) + sum(AUD not in r for r in ids)
    embed = torch.randn(V, H, generator=g)
    img = torch.randn(n_img, TI, H, generator=g)
    aud = torch.randn(n_aud, TA, H, generator=g)
    L = max(len(r) for r in ids)
    assert all(len(r) == L for r in ids) or len(ids) == 1 or True
    plan = plan_splice(ids, n_img, TI, n_aud, TA, max_len)
    got = _apply(plan, embed, img, aud, len(ids))
    # oracle consumes rectangular id tensors: run it row by row with the matching feature slices
    state = {"model.embed_tokens.weight": embed}
    ii = ai = 0
    for b, row in enumerate(ids):
        ni = max(1, row.count(IMG))
        na = max(1, row.count(AUD))
        emb, lens = O.prepare_inputs_embeds(state, cfg, torch.tensor([row]), None, None, img[ii:ii + ni], aud[ai:ai + na])
        ii += ni
        ai += na
        assert lens[0] == plan.lengths[b]
        assert torch.equal(got[b, : lens[0]], emb[0])
        assert (got[b, lens[0]:] == 0).all()


def test_plan_rejects_feature_count_mismatch():
    with pytest.raises(AssertionError):
        plan_splice([[1, IMG, 2]], 2, 4, 1, 3, None)
    with pytest.raises(AssertionError):
        plan_splice([[1, AUD, 2]], 1, 4, 3, 3, None)
