"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol that
include/t2i_b200.h declares (no compute calls here); the ctypes table mirrors the header."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from t2i_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return _lib.LIB_PATH


def header_functions():
    text = open(os.path.join(ROOT, "include", "t2i_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"\b(?:int|long long|const char\*)\s+(t2i_\w+)\s*\(", text)


def test_header_declares_the_expected_surface():
    names = header_functions()
    assert len(names) == len(set(names)) and len(names) >= 30
    for must in ("t2i_conv_gemm", "t2i_wgrad_gemm", "t2i_adam_tf", "t2i_gp_penalty", "t2i_bn_stats", "t2i_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in header_functions():
        assert hasattr(lib, name), "libt2i_b200.so does not export %s" % name


def test_ctypes_table_matches_header(lib_path):
    from t2i_b200 import _lib
    declared = set(header_functions())
    bound = set(_lib.SIGNATURES) | set(_lib.OTHER_SYMBOLS)
    assert declared == bound, (declared - bound, bound - declared)
    lib = _lib.load()
    assert lib.t2i_version() >= 1 and lib.t2i_launch_count() == 0
    # argument counts of the binding equal the header's parameter counts
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "t2i_b200.h")).read(), flags=re.S)
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip()]) == len(argtypes), name


def test_product_path_has_no_cpu_fallback(lib_path):
    """Without a CUDA device the public model must refuse to construct (fail loudly), and nothing in
    the package may import the oracle or the test restatement of the kernels."""
    import torch
    from t2i_b200.models.wgancls.model import WGanCls
    from t2i_b200.utils.config import config_from_yaml
    cfg = config_from_yaml(os.path.join(ROOT, "text-to-image_b200", "models", "wgancls", "cfg", "flowers.yml"))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            WGanCls(cfg)
    pkg = os.path.join(ROOT, "text-to-image_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "fake_kernels" not in src.replace(
                    "tests/fake_kernels.py", ""), os.path.join(dirpath, f)


def test_config_from_yaml_attribute_access():
    from t2i_b200.utils.config import config_from_yaml
    cfg = config_from_yaml(os.path.join(ROOT, "text-to-image_b200", "models", "wgancls", "cfg", "flowers.yml"))
    assert cfg.MODEL.GF_DIM == 128 and cfg.TRAIN.COEFF.KL == 1.0 and cfg.MODEL.IMAGE_SHAPE.D == 3
    assert cfg.TRAIN.BETA1 == 0.0 and cfg.TRAIN.N_CRITIC == 1
