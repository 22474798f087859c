"""Import shim for the UPSTREAM reference (test infrastructure, builder container only).

This file is part of the ORACLE / TEST INFRASTRUCTURE.  It is used by
``oracle/make_golden.py`` (run once, in the builder container where
``/root/reference`` is mounted) to import the reference's own modules so that
golden vectors can be frozen under ``tests/golden/``.  Nothing on the product
path, in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports it: the GPU box
has no ``/root/reference``.

What blocks a plain import of the reference and how each is
handled here:

* ``timm`` is absent -> stub modules providing the 7 symbols the reference pulls
  (``binary_seg/lib/pvtv2.py:6-9``, ``multiclass_seg/EMCAD/lib/decoders.py:7-8``).
* hard-coded checkpoint paths (``binary_seg/lib/Res2Net_v1b.py:198``,
  ``binary_seg/lib/pranet.py:147-148``) -> ``torch.load`` patched to return ``{}``
  while a model is constructed and Res2Net built with ``pretrained=False``.
* ``import MyTrain_med`` pulls ``thop`` -> ``structure_loss`` is extracted by AST
  from ``binary_seg/MyTrain_med.py:19-38`` and exec'd with {torch, F}.
"""
from __future__ import annotations

import ast
import contextlib
import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

def _root() -> str:
    """The mounted reference (builder container) or the byte-for-byte staged copy oracle/_ref (oracle/build_ref.py; GPU box)."""
    env = os.environ.get("PV2_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/binary_seg/lib"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REF = _root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "binary_seg", "lib"))


def _install_timm_stub() -> None:
    if "timm" in sys.modules and not getattr(sys.modules["timm"], "_pv2_stub", False):
        return
    if "timm" in sys.modules:
        return

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = float(drop_prob or 0.0)

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            mask = x.new_empty(shape).bernoulli_(keep)
            return x * mask / keep

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

    def trunc_normal_tf_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        with torch.no_grad():
            nn.init.trunc_normal_(t, 0.0, 1.0, a, b)
            t.mul_(std).add_(mean)
        return t

    def register_model(fn):
        return fn

    def _cfg(**kw):
        return dict(kw)

    def named_apply(fn, module, name="", depth_first=True, include_root=False):
        if not depth_first and include_root:
            fn(module=module, name=name)
        for cn, cm in module.named_children():
            cn = ".".join((name, cn)) if name else cn
            named_apply(fn, cm, cn, depth_first, True)
        if depth_first and include_root:
            fn(module=module, name=name)
        return module

    timm = types.ModuleType("timm")
    timm._pv2_stub = True
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    registry = types.ModuleType("timm.models.registry")
    vit = types.ModuleType("timm.models.vision_transformer")
    helpers = types.ModuleType("timm.models.helpers")
    layers.DropPath, layers.to_2tuple = DropPath, to_2tuple
    layers.trunc_normal_, layers.trunc_normal_tf_ = trunc_normal_, trunc_normal_tf_
    registry.register_model = register_model
    vit._cfg = _cfg
    helpers.named_apply = named_apply
    timm.models = models
    models.layers, models.registry, models.vision_transformer, models.helpers = layers, registry, vit, helpers
    for m in (timm, models, layers, registry, vit, helpers):
        sys.modules[m.__name__] = m


@contextlib.contextmanager
def _app_on_path(app_dir: str):
    """The reference apps import ``lib.*`` / ``utils.*`` absolutely; each app vendors its own copy."""
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k == "lib" or k.startswith("lib.") or k == "utils" or k.startswith("utils.")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, app_dir)
    try:
        yield
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.") or k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


@contextlib.contextmanager
def _no_checkpoints():
    real_load = torch.load
    torch.load = lambda *a, **k: {}
    try:
        yield
    finally:
        torch.load = real_load


_cache: dict = {}


def binary_lib():
    """Returns (pranet_module, pranet_v1_module) of binary_seg/lib with checkpoints disabled."""
    if "binary" in _cache:
        return _cache["binary"]
    _install_timm_stub()
    with _app_on_path(os.path.join(REF, "binary_seg")):
        pr = importlib.import_module("lib.pranet")
        v1 = importlib.import_module("lib.PraNet_Res2Net")
        r2n = importlib.import_module("lib.Res2Net_v1b")
        # Res2Net_v1b.py:198 loads a checkpoint from a hard-coded relative path when pretrained=True
        for mod in (pr, v1):
            mod.res2net50_v1b_26w_4s = lambda pretrained=False, **kw: r2n.res2net50_v1b_26w_4s(pretrained=False, **kw)
    _cache["binary"] = (pr, v1)
    return pr, v1


def build_binary(name: str, **kw):
    """name in {PraNet_V2, PVT_PraNet_V2, PraNet, PVT_PraNet}; random init, no checkpoint IO."""
    pr, v1 = binary_lib()
    cls = getattr(pr, name, None) or getattr(v1, name)
    with _no_checkpoints():
        return cls(**kw)


def structure_loss():
    """The reference's structure_loss, extracted by AST from binary_seg/MyTrain_med.py:19-38."""
    if "sl" in _cache:
        return _cache["sl"]
    src = open(os.path.join(REF, "binary_seg", "MyTrain_med.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "structure_loss")
    ns = {"torch": torch, "F": F}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "MyTrain_med.py", "exec"), ns)
    _cache["sl"] = ns["structure_loss"]
    return _cache["sl"]


def _load_file(modname: str, path: str):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def emcad_decoders():
    if "emcad" not in _cache:
        _install_timm_stub()
        _cache["emcad"] = _load_file("_ref_emcad_decoders", os.path.join(REF, "multiclass_seg/EMCAD/lib/decoders.py"))
    return _cache["emcad"]


def emcad_networks():
    if "emcad_net" not in _cache:
        _install_timm_stub()
        with _app_on_path(os.path.join(REF, "multiclass_seg", "EMCAD")):
            _cache["emcad_net"] = importlib.import_module("lib.networks")
    return _cache["emcad_net"]


def merit_decoders():
    if "merit" not in _cache:
        _cache["merit"] = _load_file("_ref_merit_decoders", os.path.join(REF, "multiclass_seg/MERIT/lib/decoders.py"))
    return _cache["merit"]


def mist_cam():
    if "mist" not in _cache:
        _cache["mist"] = _load_file("_ref_mist", os.path.join(REF, "multiclass_seg/MIST/lib/MIST.py"))
    return _cache["mist"]


def emcad_loss_pieces():
    """(powerset, DiceLoss, convert_labels_to_one_hot_masks) extracted by AST.

    EMCAD/utils/utils.py:20-30,102-138 and EMCAD/trainer.py:22-29 (the modules themselves
    import medpy/thop/h5py which are absent)."""
    if "mcl" in _cache:
        return _cache["mcl"]
    ns = {"torch": torch, "nn": nn, "F": F}
    src = open(os.path.join(REF, "multiclass_seg/EMCAD/utils/utils.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name == "powerset") or (isinstance(n, ast.ClassDef) and n.name == "DiceLoss")]
    exec(compile(ast.Module(body=keep, type_ignores=[]), "utils.py", "exec"), ns)
    src = open(os.path.join(REF, "multiclass_seg/EMCAD/trainer.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "convert_labels_to_one_hot_masks"]
    exec(compile(ast.Module(body=keep, type_ignores=[]), "trainer.py", "exec"), ns)
    _cache["mcl"] = (ns["powerset"], ns["DiceLoss"], ns["convert_labels_to_one_hot_masks"])
    return _cache["mcl"]


def _stmts_between(path: str, first: int, last: int):
    """The innermost statement list of `path` that covers source lines [first, last], restricted to that range."""
    tree = ast.parse(open(path).read())
    best = None
    for node in ast.walk(tree):
        for field in ("body", "orelse", "finalbody"):
            stmts = getattr(node, field, None)
            if not isinstance(stmts, list) or not stmts or not isinstance(stmts[0], ast.stmt):
                continue
            inside = [s for s in stmts if s.lineno >= first and s.end_lineno <= last]
            if inside and inside[0].lineno == first and inside[-1].end_lineno == last:
                best = inside
    if best is None:
        raise RuntimeError(f"{path}: no statement list spans lines {first}-{last}")
    return best


def binary_test_tail():
    """The reference's binary test-time post-processing, run from its own source: the statements of
    binary_seg/MyTest_med.py:35-42 (forward, p2+p3+p4+p5, resize to gt.shape, sigmoid, min-max, uint8) compiled as they
    stand.  Returns f(model, image, gt) -> output_uint8 (numpy)."""
    stmts = _stmts_between(os.path.join(REF, "binary_seg", "MyTest_med.py"), 35, 42)
    code = compile(ast.Module(body=stmts, type_ignores=[]), "MyTest_med.py", "exec")

    def run(model, image, gt):
        import numpy as np
        ns = {"torch": torch, "F": F, "np": np, "model": model, "image": image, "gt": gt}
        exec(code, ns)
        return ns["output_uint8"]
    return run


def binary_train_loss_block():
    """The loss block of the reference's training loop, run from its own source: binary_seg/MyTrain_med.py:74 (bg_mask = 1-gts)
    and :78-82 (the four structure_loss calls and their sum), compiled as they stand with the reference's structure_loss (:19-38).
    Returns f(outs, gts) -> (loss, (loss2, loss3, loss4, loss5)) where `outs` is the 8-tuple the model returns (:76)."""
    path = os.path.join(REF, "binary_seg", "MyTrain_med.py")
    stmts = _stmts_between(path, 74, 74) + _stmts_between(path, 78, 82)
    code = compile(ast.Module(body=stmts, type_ignores=[]), "MyTrain_med.py", "exec")
    sl = structure_loss()

    def run(outs, gts):
        ns = {"torch": torch, "F": F, "structure_loss": sl, "gts": gts}
        (ns["lateral_map_2_fg"], ns["lateral_map_3_fg"], ns["lateral_map_4_fg"], ns["lateral_map_5_fg"],
         ns["lateral_map_2_bg"], ns["lateral_map_3_bg"], ns["lateral_map_4_bg"], ns["lateral_map_5_bg"]) = outs
        exec(code, ns)
        return ns["loss"], (ns["loss2"], ns["loss3"], ns["loss4"], ns["loss5"])
    return run


def multiclass_val_tail():
    """The dual-branch prediction rule of EMCAD/utils/utils.py:285-296 (val_single_volume, 2-D branch: P = net(input)[:4],
    P_bg = net(input)[-4:], outputs = sum (P - P_bg), argmax softmax), compiled from the reference source as it stands.
    Returns f(net, input) -> label map tensor (H, W)."""
    stmts = _stmts_between(os.path.join(REF, "multiclass_seg/EMCAD/utils/utils.py"), 285, 296)
    code = compile(ast.Module(body=stmts, type_ignores=[]), "utils.py", "exec")

    def run(net, inp):
        ns = {"torch": torch, "net": net, "input": inp, "use_dual": True}
        exec(code, ns)
        return ns["out"]
    return run
