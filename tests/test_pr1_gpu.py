"""PR1 / BASELINE configs[0]: text-only Mixtral-8x7B geometry (full layer width, depth 4), 128-token prompt, bs = 1,
32 FREE-RUNNING greedy tokens: every token id must equal the fp32 oracle's (which equals the reference's own classes,
see the `note` field of the fixture) -- no margin gating, no teacher forcing.  The weight scaling and the prompt seed
that make this well-posed are described in oracle/pr1.py; the fixture is minted by oracle/make_golden_pr1.py."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden" / "pr1_l4.npz"


@pytest.mark.skipif(not GOLDEN.exists(), reason="tests/golden/pr1_l4.npz not minted yet")
def test_pr1_free_running_greedy_token_ids_exact():
    from oracle import pr1
    from vita_b200 import weights as W
    from vita_b200.model.vita_mixtral import VITAMixtralForCausalLM
    g = np.load(GOLDEN)
    cfg = pr1.config()
    ids = pr1.prompt(int(g["prompt_seed"]), cfg.llm.vocab_size)
    assert np.array_equal(ids.numpy(), g["input_ids"]), "prompt generator drifted from the fixture"
    state = pr1.build_state(cfg)
    model = VITAMixtralForCausalLM(cfg, {"llm": W.pack_llm(state, cfg, "cuda")}, "cuda", max_seq_len=256,
                                   max_new_tokens=pr1.NEW_TOKENS)
    del state
    want = g["tokens"].tolist()
    out = model.generate(ids.cuda(), max_new_tokens=pr1.NEW_TOKENS, output_scores=True)
    got = out.sequences[0, pr1.PROMPT_LEN:].tolist()
    print("oracle margins: logit gap min %.3f, router gap min %.3f" % (g["logit_rel_gaps"].min(), g["router_gaps"].min()))
    assert got == want, (got, want)
    # the logits the tokens were chosen from: top-2 values of every step within bf16 tolerance of the oracle's
    rows = torch.cat(list(out.scores)).float().cpu()
    idx = torch.from_numpy(g["top2_indices"])
    ref = torch.from_numpy(g["top2_values"])
    rel = ((rows.gather(1, idx) - ref).abs() / ref[:, :1].abs()).max().item()
    print(f"top-2 logit values: max rel err {rel:.3e}")
    assert rel < 4e-2
    # CUDA-graph replay, eager launches and a different read-back cadence give the same ids
    again = model.generate(ids.cuda(), max_new_tokens=pr1.NEW_TOKENS, use_graph=False, sync_every=5)
    assert again.sequences[0, pr1.PROMPT_LEN:].tolist() == want
