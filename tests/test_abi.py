"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol
include/pv2.h declares; the host modules keep the reference's state_dict keys; ops refuse CPU tensors."""
import os
import re

import pytest
import torch

import pranet_v2_b200 as P
from oracle import templates

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "pv2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pv2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = P._lib.load()
    syms = _header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pv2.h but not exported"
    # and the ctypes table binds exactly the header's set
    assert sorted(P._lib.declared_symbols()) == syms
    assert lib.pv2_version() >= 100


def test_no_cpu_fallback():
    x = torch.zeros(1, 1, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        P.structure_loss(x, x, x, x)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        P.interpolate_bilinear(x, scale_factor=2)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        P.dsra_fuse(x, x, x)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pranet-v2_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{fn} imports oracle"


@pytest.mark.parametrize("name,kw,ch,v1", [
    ("PraNet_V2", dict(num_class=1), (512, 1024, 2048), False),
    ("PraNet_V2", dict(num_class=3), (512, 1024, 2048), False),
    ("PVT_PraNet_V2", dict(num_class=1), (128, 320, 512), False),
    ("PraNet", dict(), (512, 1024, 2048), True),
    ("PVT_PraNet", dict(), (128, 320, 512), True),
])
def test_head_state_dict_keys(name, kw, ch, v1):
    m = getattr(P, name)(**kw)
    ours = {k: v for k, v in m.state_dict().items() if not k.startswith(("backbone.", "resnet.", "conv."))}
    tmpl = templates.pranet_head(ch, 32, kw.get("num_class", 1), v1=v1)
    assert set(ours) == set(tmpl)
    for k in tmpl:
        assert tuple(ours[k].shape) == tuple(tmpl[k].shape), k
    if name == "PraNet_V2" and kw["num_class"] == 1:
        assert len(ours) + 7 == 425      # 425 non-backbone keys incl. the 1->3 stem


def test_reference_import_surface():
    from pranet_v2_b200.lib.pranet import BasicConv2d, PraNet_V2, PVT_PraNet_V2, RFB_modified, aggregation  # noqa: F401
    from pranet_v2_b200.lib.PraNet_Res2Net import PraNet, PVT_PraNet  # noqa: F401
    import inspect
    sig = inspect.signature(PraNet_V2.__init__)
    assert list(sig.parameters)[1:] == ["channel", "num_class", "sem_downsample", "use_softmax"]
    assert [p.default for p in list(sig.parameters.values())[1:]] == [32, 3, 1, True]
    assert list(inspect.signature(PraNet_V2.forward).parameters)[1:] == ["x", "segSize"]
    assert list(inspect.signature(P.structure_loss).parameters)[:4] == ["pred", "pred_bg", "mask_fg", "mask_bg"]


def test_ctypes_table_matches_header_arity():
    """Every prototype of include/pv2.h has as many parameters as its ctypes signature in _lib._SIGS (a drifted binding would
    otherwise only show up as garbage arguments on the GPU)."""
    src = open(os.path.join(ROOT, "include", "pv2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    protos = dict(re.findall(r"\b(pv2_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S))
    assert set(protos) == set(P._lib._SIGS)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(P._lib._SIGS[name][1]), f"{name}: header has {n} parameters, ctypes table {len(P._lib._SIGS[name][1])}"
