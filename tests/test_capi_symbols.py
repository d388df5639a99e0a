"""CPU test: the C-ABI library builds, loads, and exports every symbol include/radtts_b200.h declares.
No compute call is made here (there is no GPU in the build container)."""
import os
import subprocess

from radtts_b200 import _lib, build


def test_library_builds_and_exports_declared_symbols():
    path = build.build()
    assert os.path.exists(path)
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert "radtts_mas_forward" in names and len(names) >= 5
    for n in names:
        assert getattr(L, n) is not None
    assert L.radtts_abi_version() >= 1
    assert L.radtts_error_string(-2).decode().startswith("radtts_b200")


def test_sass_is_sm100_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "sm_90" not in out and "sm_80" not in out


def test_product_fails_loudly_without_cuda_tensor():
    import pytest
    import torch
    from radtts_b200 import alignment
    with pytest.raises(_lib.RadttsB200Error):
        alignment.mas_forward(torch.rand(1, 1, 4, 3), [3], [4])
