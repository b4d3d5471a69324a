"""CPU-side checks of the boundary: libb2c.so loads without a GPU and exports every symbol that
include/b2c.h declares; the Python mirror keeps the reference's API surface."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from iclr_17_compression_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b2c.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2c_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    from iclr_17_compression_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libb2c.so does not export {n}"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.b2c_version() >= 100
    assert isinstance(lib.b2c_last_error(), bytes)


def test_no_cpu_fallback(lib):
    import iclr_17_compression_b200 as P
    from iclr_17_compression_b200._lib import B2CError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.b2c_check_device() != 0          # no device -> error code, message set
    assert b"no CPU fallback" in lib.b2c_last_error()
    m = P.ImageCompressor(32).eval()
    with pytest.raises(B2CError):
        m(torch.rand(1, 3, 32, 32))


def test_product_never_imports_oracle():
    """the oracle is test infrastructure; the product path must not route through it."""
    pkg = os.path.join(ROOT, "iclr_17_compression_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f
                assert "balle17_oracle" not in src, f


def test_reference_api_surface():
    """constructor, attribute names, state_dict keys/shapes, star-import names."""
    import iclr_17_compression_b200 as P
    from iclr_17_compression_b200 import model as M
    from oracle import balle17_oracle as O
    torch.manual_seed(234)
    m = P.ImageCompressor(out_channel_N=64)
    assert isinstance(m.Encoder, P.Analysis_net_17) and isinstance(m.Decoder, P.Synthesis_net_17)
    assert isinstance(m.bitEstimator, P.BitEstimator) and m.out_channel_N == 64
    sd = m.state_dict()
    ref = O.init_state_dict(64, seed=234)
    assert list(sd.keys()) == list(ref.keys()) and len(sd) == 30
    for k in sd:
        assert sd[k].dtype == torch.float32 and torch.equal(sd[k], ref[k]), k
    for name in ["ImageCompressor", "load_model", "save_model", "logging", "np", "nn", "F", "os", "math", "torch",
                 "GDN", "Analysis_net_17", "Synthesis_net_17", "BitEstimator"]:
        assert hasattr(M, name), name


def test_save_load_roundtrip(tmp_path):
    import iclr_17_compression_b200 as P
    torch.manual_seed(1)
    a = P.ImageCompressor(16)
    P.save_model(a, 7, str(tmp_path))
    f = tmp_path / "iter_7.pth.tar"
    assert f.exists()
    torch.manual_seed(2)
    b = P.ImageCompressor(16)
    extra = dict(torch.load(f))
    extra["not.a.key"] = torch.zeros(1)            # unknown keys are dropped silently (model.py:26)
    torch.save(extra, f)
    assert P.load_model(b, str(f)) == 0
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k])
    P.load_model(b.Decoder, str(f))               # sub-modules are individually loadable (filters by key)
