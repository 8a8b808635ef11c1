"""The C-ABI library loads without a GPU and exports every symbol include/pt_core.h declares."""
import ctypes as C
import os
import re

import pytest

import conftest

core = conftest.core


def declared_symbols():
    text = open(os.path.join(conftest.ROOT, "include", "pt_core.h")).read()
    return sorted(set(re.findall(r"PT_API\s+[\w\s\*]+?\b(pt_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(core.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(core.LIB_PATH):
        core.build()
    lib = C.CDLL(core.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in pt_core.h but not exported"


def test_test_mode_strides_match_oracle(oracle_mod):
    L = core.lib()
    for mode in range(len(oracle_mod.TEST_IN)):
        assert L.pt_test_input_stride(mode) == oracle_mod.TEST_IN[mode]
        assert L.pt_test_output_stride(mode) == oracle_mod.TEST_OUT[mode]
    assert L.pt_test_input_stride(99) == 0


def test_no_cpu_fallback_without_device():
    """Without a CUDA device pt_context_create must fail with PT_ERR_NO_DEVICE, not fall back."""
    if conftest._has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(core.PtError) as e:
        core.Renderer(0)
    assert e.value.status == -2


def test_null_arguments_are_rejected():
    L = core.lib()
    assert L.pt_context_create(0, None) == -1
    assert L.pt_render_begin(None, 4, 4) == -1
    assert L.pt_get_stats(None, None) == -1
