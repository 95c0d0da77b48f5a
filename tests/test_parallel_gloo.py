"""world_size-2 gloo test of the request-parallel plumbing (the N>1 path of bench.py / serving)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    from vita_b200 import parallel
    w, r, _ = parallel.init("gloo")
    assert (w, r) == (world, rank)
    owned = parallel.shard_requests(5, rank, world)
    toks = [[100 * i + k for k in range(i + 1)] for i in owned]          # stand-in for generated tokens
    everything = parallel.gather_token_lists(toks, owned, 5)
    tmax = parallel.reduce_max(10.0 + rank)
    tsum = parallel.reduce_sum(float(len(owned)))
    q.put((rank, owned, everything, tmax, tsum))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_request_sharding_and_reductions_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    want = [[100 * i + k for k in range(i + 1)] for i in range(5)]
    for _, _, everything, tmax, tsum in res:
        assert everything == want
        assert tmax == 11.0 and tsum == 5.0


def test_single_process_degenerates():
    from vita_b200 import parallel
    assert parallel.shard_requests(3, 0, 1) == [0, 1, 2]
    assert parallel.reduce_max(3.5) == 3.5
    assert parallel.gather_token_lists([[1], [2, 3]], [0, 1], 2) == [[1], [2, 3]]


def _ep_worker(rank, world, port, q):
    """Expert-parallel decomposition of the MoE block: sum over ranks of the local experts' weighted outputs
    (what MixtralDecoder.prefill all-reduces) equals the full MixtralSparseMoeBlock of the oracle."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    import torch.nn.functional as F
    from oracle import vita_oracle as O
    from vita_b200 import parallel, weights as W
    from vita_b200.config import VitaConfig
    parallel.init("gloo")
    cfg = VitaConfig.tiny()
    state = W.synthetic_state(cfg, 0, parts=("llm",))
    lo, hi = W.expert_range(cfg.llm.num_local_experts, rank, world)
    xn = torch.randn(37, cfg.llm.hidden_size, generator=torch.Generator().manual_seed(5))
    full, top_i, top_v = O.sparse_moe(state, cfg.llm, 0, xn)
    part = torch.zeros_like(xn)
    for e in range(lo, hi):
        tok, kpos = torch.where(top_i == e)
        if tok.numel():
            w1, w3, w2 = O.expert_weights(state, cfg.llm, 0, e)
            y = O.linear(F.silu(O.linear(xn[tok], w1)) * O.linear(xn[tok], w3), w2) * top_v[tok, kpos, None]
            part.index_add_(0, tok, y)
    torch.distributed.all_reduce(part)
    q.put((rank, (lo, hi), float((part - full).abs().max())))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_expert_parallel_decomposition_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ep_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [(0, 4), (4, 8)]
    assert all(r[2] < 1e-5 for r in res)
