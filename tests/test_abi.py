"""The C-ABI shared library builds, loads and exports every symbol include/vita_b200.h declares (no GPU needed)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "vita_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vita_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from vita_b200 import build
    return ctypes.CDLL(str(build.build()))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vita_b200.h but not exported"


def test_ctypes_table_matches_header():
    from vita_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_loader_attaches_prototypes_and_reports_errors(lib):
    from vita_b200 import _lib
    l = _lib.load()
    assert l.vita_version() >= 100
    # argument validation happens before any CUDA call: a bad shape is VITA_ERR_INVALID with a message
    rc = l.vita_gemm_bf16(None, 8, None, None, 8, 16, 16, 12, None, 0, None, None, 0, None)
    assert rc == -1 and b"multiples of 8" in l.vita_last_error()
    with pytest.raises(_lib.VitaB200Error):
        _lib.call("vita_rmsnorm", None, None, None, 1, 7, 1e-5, None)


def test_no_cpu_fallback_in_ops():
    import torch
    from vita_b200 import ops, _lib
    with pytest.raises(_lib.VitaB200Error):
        ops.rmsnorm(torch.zeros(2, 8, dtype=torch.bfloat16), torch.ones(8, dtype=torch.bfloat16), 1e-5)


def test_product_never_imports_the_oracle():
    for p in (ROOT / "vita_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


def test_library_options_round_trip():
    """vita_set_option / vita_get_option: environment-independent defaults, overrides, unknown names (no GPU needed)."""
    from vita_b200 import _lib, ops
    for name in ("pdl", "attn_early", "attn_tagged", "chain_wait", "tc_prefetch_consts", "tc_wide_route"):
        assert ops.get_option(name) in (0, 1)
    before = ops.get_option("tc_l2_ahead")
    ops.set_option("tc_l2_ahead", 5)
    assert ops.get_option("tc_l2_ahead") == 5
    ops.set_option("tc_l2_ahead", before)
    import pytest
    with pytest.raises(_lib.VitaB200Error):
        ops.set_option("no_such_option", 1)
