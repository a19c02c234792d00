"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/oscen_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from oscen_b200 import build
    build.build()
    from oscen_b200 import abi
    return abi.load()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "oscen_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)        # drop comments
    declared = set(re.findall(r"\b(oscen_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    from oscen_b200 import abi
    assert declared == set(abi.SYMBOLS), declared ^ set(abi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_input_counts(lib):
    assert lib.oscen_b200_abi_version() == 1
    assert [lib.oscen_b200_n_inputs(k) for k in range(8)] == [28, 3, 2, 2, 9, 4, 8, 0]


def test_no_cpu_fallback(lib):
    from oscen_b200 import abi
    import conftest
    if conftest.HAS_GPU:
        pytest.skip("GPU present")
    assert lib.oscen_b200_device_count() == 0
    with pytest.raises(abi.EngineError, match="no CUDA device"):
        abi.Engine(abi.GRAPH_FM_VOICE, 128, 8)


def test_bad_config_rejected(lib):
    from oscen_b200 import abi
    cfg = abi.Config(99, 128, 8, 48000.0, 0, 0, 0, 0)
    h = C.c_void_p()
    assert lib.oscen_b200_create(C.byref(cfg), C.byref(h)) == -1
    cfg = abi.Config(0, 100, 8, 48000.0, 0, 0, 0, 0)         # n_voices % vpi != 0
    assert lib.oscen_b200_create(C.byref(cfg), C.byref(h)) == -1
    assert b"bad config" in lib.oscen_b200_last_error(None)


def test_event_layout_matches_oracle_event():
    from oscen_b200 import abi
    import oracle_lib
    assert C.sizeof(abi.Event) == C.sizeof(oracle_lib.Event) == 16
    assert abi.EVENT_GATE == oracle_lib.EV_GATE
    assert abi.EVENT_DTYPE.itemsize == 16


def test_product_does_not_touch_oracle():
    """the product tree must not import / link / open anything under oracle/"""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "oscen_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle_lib|liboscen_oracle|oscen_oracle\.h|/oracle/", txt):
                    bad.append(f)
    assert not bad, bad


def test_rust_sys_crate_declares_every_abi_symbol():
    """rust/oscen-b200-sys cannot be compiled here (no cargo): at least its extern block must name every entry point of
    the header, with the same number of parameters, and nothing else"""
    hdr = open(os.path.join(ROOT, "include", "oscen_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    c_protos = {m.group(1): m.group(2) for m in re.finditer(r"\b(oscen_b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr)}
    rs = open(os.path.join(ROOT, "rust", "oscen-b200-sys", "src", "lib.rs")).read()
    rs_protos = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (oscen_b200_[a-z0-9_]+)\(([^)]*)\)", rs)}
    assert set(rs_protos) == set(c_protos), set(rs_protos) ^ set(c_protos)

    def n_params(s):
        s = s.strip()
        return 0 if s in ("", "void") else s.count(",") + 1
    for name in c_protos:
        assert n_params(c_protos[name]) == n_params(rs_protos[name]), name
    # #[repr(C)] mirrors exist for every struct the header defines
    for struct in re.findall(r"\}\s*(oscen_b200_[a-z_]+)\s*;", hdr):
        assert f"pub struct {struct}" in rs, struct
