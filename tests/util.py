"""Shared helpers for the parity tests (the oracle is imported here as the checker only)."""
import torch

BF16 = torch.bfloat16


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(BF16).to(torch.float32)


def randn(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return bf16_round(torch.randn(shape, generator=g) * scale)


def to_dev(t: torch.Tensor) -> torch.Tensor:
    return t.to(device="cuda", dtype=BF16).contiguous()


def assert_close(got: torch.Tensor, want: torch.Tensor, rel=1.6e-2, what=""):
    """bf16 tolerance: max |got - want| <= rel * max|want| (fp32 accumulate, one or two bf16 roundings)."""
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    assert got.shape == want.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    assert err <= rel * max(ref, 1e-6), f"{what}: max err {err:.4e} vs max|ref| {ref:.4e} (rel {err / max(ref, 1e-6):.3e})"
    return err / max(ref, 1e-6)
