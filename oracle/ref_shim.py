"""Import shim for the *reference* implementation (VITA-MLLM/VITA mounted read-only at /root/reference).

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py (run in the build container, where /root/reference exists)
to execute the reference's own Python classes on CPU and mint the golden vectors committed under tests/golden/.
Nothing in the product (vita_b200/), bench.py's GPU arm or the `-m gpu` tests imports this file, and it is never
used on the GPU box (the reference tree does not exist there).

The reference needs two soft dependencies that are absent from this image (timm, xformers); they are replaced by
inert stubs exactly as documented in SURVEY.md Appendix C.  Nothing under /root/reference is modified.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("VITA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vita", "model"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m


class _DropPath(torch.nn.Module):  # identity in eval; the reference only instantiates it for drop_path_rate > 0
    def __init__(self, p: float = 0.0):
        super().__init__()

    def forward(self, x):
        return x


_installed = False


def install():
    """Make `import vita.model...` work.  Must run after `import transformers`."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import transformers  # noqa: F401  (must be imported before the stubs are registered)
    from transformers import MixtralForCausalLM  # noqa: F401

    common = dict(DropPath=_DropPath, drop_path=lambda x, *a, **k: x, to_2tuple=lambda x: (x, x),
                  trunc_normal_=torch.nn.init.trunc_normal_)
    for name in ("xformers", "xformers.ops", "timm", "timm.models"):
        if name not in sys.modules:
            _stub(name)
    _stub("timm.models.layers", **common)
    _stub("timm.layers", **common)
    _stub("timm.layers.norm_act", LayerNormAct2d=torch.nn.Identity)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
