"""CPU-only: libse3tn.so builds/loads and exports every symbol include/se3tn.h declares, and the
ctypes table mirrors the header.  No compute calls (there is no GPU on the CPU runner)."""
import ctypes, importlib, os, re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'se3tn.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(se3tn_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for must in ('se3tn_create', 'se3tn_destroy', 'se3tn_load_weights', 'se3tn_preprocess', 'se3tn_forward',
                 'se3tn_pose_update', 'se3tn_so3_log', 'se3tn_track_batch', 'se3tn_last_error', 'se3tn_workspace_bytes'):
        assert must in syms


def test_library_exports_every_declared_symbol():
    L = importlib.import_module('iros20-6d-pose-tracking_b200._lib')
    lib = L.load()
    raw = ctypes.CDLL(L.LIB_PATH)
    for name in header_symbols():
        assert hasattr(raw, name), 'libse3tn.so does not export ' + name
    assert sorted(L.SIGNATURES) == header_symbols(), 'ctypes table and header disagree'


def test_workspace_bytes_and_blob_constant():
    L = importlib.import_module('iros20-6d-pose-tracking_b200._lib')
    W = importlib.import_module('iros20-6d-pose-tracking_b200.weights')
    lib = L.load()
    assert lib.se3tn_workspace_bytes(0) == 0
    b1, b64 = lib.se3tn_workspace_bytes(1), lib.se3tn_workspace_bytes(64)
    assert 11_000_000 < b1 < 12_000_000 and abs(b64 - 64 * b1) < 64 * 16 * 1024
    hdr = open(os.path.join(ROOT, 'include', 'se3tn.h')).read()
    assert int(re.search(r'SE3TN_WEIGHT_BLOB_FLOATS\s+(\d+)u', hdr).group(1)) == W.BLOB_FLOATS == L.WEIGHT_BLOB_FLOATS


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    pkg = importlib.import_module('iros20-6d-pose-tracking_b200')
    with pytest.raises(RuntimeError):
        pkg.Engine(max_batch=1)
    # the C ABI itself reports an error rather than computing anything
    L = importlib.import_module('iros20-6d-pose-tracking_b200._lib')
    lib = L.load()
    ctx = ctypes.c_void_p()
    assert lib.se3tn_create(0, 1, None, ctypes.byref(ctx)) < 0
    assert lib.se3tn_last_error(None)
