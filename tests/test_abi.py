"""CPU: the C-ABI library builds/loads and exports every symbol include/cbas_b200.h declares (no compute)."""
import ctypes
import os
import re

from cbas_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cbas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cbas_b200_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 15
    lib = _lib.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_error_channel():
    lib = _lib.lib()
    assert lib.cbas_b200_abi_version() == _lib.ABI_VERSION == 3
    assert isinstance(lib.cbas_b200_launch_count(), int)
    # argument validation happens before any CUDA call, so this is safe without a GPU
    rc = lib.cbas_b200_encoder_create(None, None, None)
    assert rc != 0 and b"null" in lib.cbas_b200_last_error()


def test_struct_layouts_match_header():
    # sizes the C compiler produces for the header's structs (LP64): guards against field drift
    assert ctypes.sizeof(_lib.EncoderCfg) == 14 * 4
    assert ctypes.sizeof(_lib.LayerWeights) == 10 * 8
    assert ctypes.sizeof(_lib.EncoderWeights) == 13 * 8
    assert ctypes.sizeof(_lib.HeadCfg) == 9 * 4
    assert ctypes.sizeof(_lib.HeadWeights) == 28 * 8 + 8 + 8 * 8  # + the second LSTM layer


def test_option_constants_match_header():
    """The per-handle option numbers the Python mirror passes to cbas_b200_encoder_set_option are the header's."""
    src = open(os.path.join(ROOT, "include", "cbas_b200.h")).read()
    want = {k: int(v) for k, v in re.findall(r"#define\s+(CBAS_OPT_\w+)\s+(\d+)", src)}
    assert want == {"CBAS_OPT_ATTENTION_IMPL": _lib.OPT_ATTENTION_IMPL, "CBAS_OPT_PRUNE_LAST_LAYER": _lib.OPT_PRUNE_LAST_LAYER,
                    "CBAS_OPT_RESIZE_KERNEL": _lib.OPT_RESIZE_KERNEL, "CBAS_OPT_LN_FUSION": _lib.OPT_LN_FUSION,
                    "CBAS_OPT_SERPENTINE": _lib.OPT_SERPENTINE}
    lib = _lib.lib()
    # validation of the option value happens before any CUDA call
    assert lib.cbas_b200_encoder_set_option(None, _lib.OPT_LN_FUSION, 2) != 0 and b"null" in lib.cbas_b200_last_error()
