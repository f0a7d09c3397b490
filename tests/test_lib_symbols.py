"""CPU: the C-ABI library loads and exports exactly the symbols include/vmb200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

from videomamba_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vmb200.h")).read()
    return set(re.findall(r"^VMB_API\s+[\w\s\*]+?\b(vmb_\w+)\s*\(", text, flags=re.M))


def test_header_and_binding_agree():
    names = _declared()
    assert len(names) >= 13
    assert names == set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(_lib.LIB_PATH), "run `python -m videomamba_b200.build` (or build()) first"
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.vmb_abi_version() == _lib.ABI_VERSION
    assert lib.vmb_last_error() is not None


def test_struct_sizes_match_header_layout():
    # natural alignment, no packing: spot-check against sizes computed from the C declarations
    assert ctypes.sizeof(_lib.ScanArgs) % 8 == 0
    assert ctypes.sizeof(_lib.MixerArgs) % 8 == 0
    assert _lib.ScanArgs.B.offset - _lib.ScanArgs.h_last.offset == 8
    assert _lib.MixerArgs.B.offset - _lib.MixerArgs.workspace_bytes.offset == 8


def test_invalid_arguments_are_reported_not_crashed():
    lib = _lib.load()
    rc = lib.vmb_linear_fwd(None, 0, None, 0, None, None, 0, 4, 4, 4, 0, None)
    assert rc == -1
    assert b"null" in lib.vmb_last_error()
    assert lib.vmb_mixer_workspace_bytes(2, 16, 8, 16, 4, 1, 7) == -1
    assert lib.vmb_mixer_workspace_bytes(2, 16, 8, 16, 4, 1, 0) > 0
