"""CPU-side consistency of the PR1 fixture (tests/golden/pr1_l4.npz) with the code that the GPU test builds its model
from: prompt generator, router scale, recorded margins.  (The fixture's token ids themselves come from the reference's
own classes: oracle/make_golden_pr1.py mint --reference.)"""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLDEN = Path(__file__).resolve().parent / "golden" / "pr1_l4.npz"


@pytest.mark.skipif(not GOLDEN.exists(), reason="tests/golden/pr1_l4.npz not minted yet")
def test_pr1_fixture_matches_the_generators():
    from oracle import pr1
    g = np.load(GOLDEN)
    cfg = pr1.config()
    assert cfg.llm.num_hidden_layers == pr1.LAYERS == 4 and cfg.llm.hidden_size == 4096
    assert float(g["gate_scale"]) == pr1.GATE_SCALE and float(g["head_gain_sigma"]) == pr1.HEAD_GAIN_SIGMA
    ids = pr1.prompt(int(g["prompt_seed"]), cfg.llm.vocab_size)
    assert ids.shape == (1, pr1.PROMPT_LEN) and np.array_equal(ids.numpy(), g["input_ids"])
    toks = g["tokens"]
    assert toks.shape == (pr1.NEW_TOKENS,) and np.array_equal(g["top2_indices"][:, 0], toks)   # greedy = arg-max row by row
    # the recorded margins are the ones the seed was qualified on
    assert float(g["logit_rel_gaps"].min()) >= 0.05
    v = g["top2_values"]
    assert np.allclose((v[:, 0] - v[:, 1]) / np.abs(v[:, 0]), g["logit_rel_gaps"], rtol=1e-4, atol=1e-6)
    assert "tokens equal" in str(g["note"])                 # minted with --reference: the reference's classes agreed


def test_pr1_margin_bookkeeping():
    """pr1.margins: top-1 / top-2 logit gap relative to the top logit; router decisions count as harmless when the second
    expert's renormalised weight is <= 2 %, else by their rank-2 / rank-3 gap over the logits' spread."""
    from oracle import pr1
    rows = torch.tensor([[10.0, 9.0, 1.0], [4.0, -2.0, 3.0]])
    sharp = torch.softmax(torch.tensor([9.0, 2.0, 1.9, 0, 0, 0, 0, 0]), -1)      # second weight ~1e-3: harmless
    tied = torch.softmax(torch.tensor([1.0, 0.9, 0.89, 0, 0, 0, 0, 0]), -1)      # rank-2 / rank-3 gap 0.01
    norms = [torch.ones(8)]
    m = pr1.margins(rows, [[sharp], [tied]], norms)
    assert abs(m["logit_rel_gaps"][0] - 0.1) < 1e-6 and abs(m["logit_rel_gaps"][1] - 0.25) < 1e-6
    assert m["logit_rel_gap_min"] == min(m["logit_rel_gaps"])
    assert m["router_gaps"][0] == 1.0 and abs(m["router_gaps"][1] - 0.01) < 1e-4
    assert m["router_gap_min"] == min(m["router_gaps"]) and 0 < m["weight_noise_max"] < 0.01
