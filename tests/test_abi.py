"""The C-ABI library loads and exports exactly what include/rbod.h declares (no compute, CPU only)."""
import ctypes
import os
import re

import pytest

from retrieval_based_object_detection_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rbod.h"), encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(rbod_[a-z0-9_]+)\s*\(", text))


def test_header_symbols_exported_and_bound():
    declared = _declared()
    assert {"rbod_create", "rbod_upsert", "rbod_segment_mean", "rbod_search", "rbod_merge_topk"} <= declared
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rbod.h but not exported by librbod.so"
    assert declared == set(_native.SIGNATURES), "ctypes table and header disagree"


def test_loads_and_reports_errors_without_gpu():
    lib = _native.load()
    assert lib.rbod_abi_version() == 1
    h = ctypes.c_void_p()
    assert lib.rbod_create(0, 0, 0, 0, 0, ctypes.byref(h)) == _native.RBOD_E_INVAL
    assert b"dim" in lib.rbod_last_error()
    assert lib.rbod_create(512, 9, 0, 0, 0, ctypes.byref(h)) == _native.RBOD_E_INVAL
    assert lib.rbod_upsert(None, None, 1, None, None, 0, None) == _native.RBOD_E_INVAL
    assert lib.rbod_search(None, None, 1, 1, None, None, None, None, None, None) == _native.RBOD_E_INVAL
    assert lib.rbod_destroy(None) == 0
    # the split-search entry points reject missing handles / buffers before touching the device
    assert lib.rbod_search_begin(None, None, 1, 1, 1, None, None, None, None) == _native.RBOD_E_INVAL
    assert b"rbod_search_begin" in lib.rbod_last_error()
    assert lib.rbod_search_end(None, None, 1, 1, None, None, None, None, None) == _native.RBOD_E_INVAL
    assert lib.rbod_global_cut(None, 2, 1, 1, 1, None, None) == _native.RBOD_E_INVAL
    assert lib.rbod_merge_topk_certified(None, None, 2, 1, 1, None, None, None, None, None, None) == _native.RBOD_E_INVAL
    assert lib.rbod_last_k3_ms(None, None) == _native.RBOD_E_INVAL


def test_product_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from retrieval_based_object_detection_b200 import Gallery

    with pytest.raises(_native.RbodError) as e:
        Gallery(512)
    assert e.value.code == _native.RBOD_E_IO and "no CPU path" in str(e.value)


def test_product_never_imports_oracle():
    bad = []
    for pkg in ("retrieval_based_object_detection_b200", "qdrant_client", "clip"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, fn), encoding="utf-8").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "liboracle" in src:
                        bad.append(os.path.join(dirpath, fn))
    assert not bad, f"product code references the oracle: {bad}"


def test_header_constants_match_the_binding():
    """Every #define RBOD_* value in include/rbod.h that the ctypes layer mirrors has the same value there, and the
    metric / dtype / delegate tables of the Python layer cover exactly the codes the header defines."""
    text = open(os.path.join(ROOT, "include", "rbod.h"), encoding="utf-8").read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"^#define\s+(RBOD_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", text, flags=re.M)}
    for name in ("RBOD_OK", "RBOD_E_IO", "RBOD_E_NOMEM", "RBOD_E_INVAL", "RBOD_E_RANGE", "RBOD_E_OVERFLOW",
                 "RBOD_E_UNSUPPORTED", "RBOD_F32", "RBOD_BF16", "RBOD_F16", "RBOD_COSINE", "RBOD_DOT", "RBOD_EUCLID",
                 "RBOD_MANHATTAN", "RBOD_UPSERT_RAW"):
        assert defs[name] == getattr(_native, name), name
    assert defs["RBOD_ABI_VERSION"] == 1
    assert set(_native.METRICS.values()) == {defs[n] for n in ("RBOD_COSINE", "RBOD_DOT", "RBOD_EUCLID", "RBOD_MANHATTAN")}
    assert set(_native.DTYPES.values()) == {defs[n] for n in ("RBOD_F32", "RBOD_BF16", "RBOD_F16")}
    assert _native.DELEGATE_KINDS == {"average": defs["RBOD_DELEGATE_AVERAGE"], "centroid": defs["RBOD_DELEGATE_CENTROID"],
                                      "weighted": defs["RBOD_DELEGATE_WEIGHTED"], "medoid": defs["RBOD_DELEGATE_MEDOID"]}
    # struct sizes the ctypes mirrors must agree with (checked through a round trip of rbod_info's error path only:
    # no GPU here) -- field counts guard against silent drift
    assert len(_native.GalleryInfo._fields_) == 11 and len(_native.SearchStats._fields_) == 10
    # every distance of the reference's menu (util/qdrant_manager.py:61-66) maps to a metric of the library
    from qdrant_client.models import Distance
    from retrieval_based_object_detection_b200 import store

    src = open(store.__file__, encoding="utf-8").read()
    for d in Distance:
        assert f'"{d.value}"' in src, d
