"""CPU-side checks of the drop-in boundary: the library loads without a GPU, exports every symbol the header
declares, the ctypes table covers the header, and the product refuses CPU tensors (no fallback)."""
import os
import re

import pytest
import torch

from tair_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tair_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tair_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.lib()


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for s in ("tair_gemm_bf16", "tair_conv3x3_bf16", "tair_attention_bf16", "tair_groupnorm_nhwc", "tair_layernorm",
              "tair_sampler_update", "tair_msda_forward", "tair_blend_tiles"):
        assert s in syms


def test_library_exports_every_header_symbol(lib):
    for s in header_symbols():
        assert hasattr(lib, s), f"libtair_b200.so does not export {s}"


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_abi_version_and_error_string(lib):
    assert lib.tair_abi_version() == 4
    assert isinstance(lib.tair_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # NULL operands are rejected before anything touches the device
    rc = lib.tair_gemm_bf16(None, 8, None, 8, 1, 1, 8, None, None)
    assert rc == -1 and b"NULL" in lib.tair_last_error()
    rc = lib.tair_attention_bf16(1, 64, 1, 64, 1, 64, 1, 64, 1, 1, 1, 1, 32, 1.0, 0, None)
    assert rc == -1 and b"head_dim" in lib.tair_last_error()


def test_no_cpu_fallback():
    from tair_b200 import ops
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(ops.TairError):
        ops.gemm(a, a)
    with pytest.raises(ops.TairError):
        ops.layernorm(a, torch.ones(8), torch.zeros(8))


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tair_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} imports the oracle"
