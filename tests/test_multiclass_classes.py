"""The multiclass drop-in carriers (pranet_v2_b200.multiclass: EMCAD_dual, CASCADE_Add_dual, CAM, EMCADNet) expose the reference's
constructor signatures and EXACTLY its state_dict keys and shapes (multiclass_seg/EMCAD/lib/decoders.py:407-444,
MERIT/lib/decoders.py:289-322, MIST/lib/MIST.py:368-412, EMCAD/lib/networks.py:18-95), so checkpoints load with strict=True.
The key / shape lists were frozen from the unmodified reference by `python -m oracle.make_golden mcdec`; when the reference
is mounted (builder container) the live classes are checked too.  CPU only: constructing the modules runs no kernel."""
import inspect
import json
import os

import pytest

import pranet_v2_b200 as P
from oracle import golden_cases as G
from oracle import ref_import as R

KEYS = json.load(open(os.path.join(G.GOLDEN_DIR, "mc_state_dict_keys.json")))
BUILD = {
    "EMCAD_dual": lambda: P.EMCAD_dual(channels=[512, 320, 128, 64], num_class=9),
    "CASCADE_Add_dual": lambda: P.CASCADE_Add_dual(channels=[768, 384, 192, 96], num_class=4),
    "CAM": lambda: P.CAM("SSS", channels=[768, 384, 192, 96], n_class=9),
    "EMCADNet": lambda: P.EMCADNet(num_classes=9, encoder="pvt_v2_b2", pretrain=False, dual=True),
}


@pytest.mark.parametrize("name", list(BUILD))
def test_state_dict_keys_and_shapes(name):
    if name not in KEYS:
        pytest.skip("not frozen")
    ours = {k: list(v.shape) for k, v in BUILD[name]().state_dict().items()}
    ref = KEYS[name]
    assert sorted(ours) == sorted(ref), (sorted(set(ref) - set(ours))[:8], sorted(set(ours) - set(ref))[:8])
    bad = [k for k in ref if ours[k] != ref[k]]
    assert not bad, [(k, ours[k], ref[k]) for k in bad[:8]]


@pytest.mark.skipif(not R.available(), reason="reference not mounted")
def test_constructor_signatures_match_reference():
    pairs = [(P.EMCAD_dual, R.emcad_decoders().EMCAD_dual), (P.CASCADE_Add_dual, R.merit_decoders().CASCADE_Add_dual)]
    for ours, ref in pairs:
        a, b = inspect.signature(ours.__init__), inspect.signature(ref.__init__)
        assert list(a.parameters) == list(b.parameters), (ours.__name__, list(a.parameters), list(b.parameters))
        for k in a.parameters:
            assert a.parameters[k].default == b.parameters[k].default, (ours.__name__, k)
    # CAM(args, **kwargs) with channels= / n_class= keywords; forward(skip1..skip4)
    assert list(inspect.signature(P.CAM.forward).parameters) == list(inspect.signature(R.mist_cam().CAM.forward).parameters)
    assert list(inspect.signature(P.EMCAD_dual.forward).parameters) == ["self", "x", "skips"]
    assert list(inspect.signature(P.CASCADE_Add_dual.forward).parameters) == ["self", "x", "skips"]


@pytest.mark.skipif(not R.available(), reason="reference not mounted")
def test_reference_checkpoints_load_strict():
    """A state_dict produced by the reference class loads into ours with strict=True, and vice versa."""
    for ours, ref in ((P.EMCAD_dual(channels=[512, 320, 128, 64], num_class=9), R.emcad_decoders().EMCAD_dual(channels=[512, 320, 128, 64], num_class=9)),
                      (P.CASCADE_Add_dual(channels=[768, 384, 192, 96], num_class=4), R.merit_decoders().CASCADE_Add_dual(channels=[768, 384, 192, 96], num_class=4)),
                      (P.CAM("SSS", channels=[768, 384, 192, 96], n_class=9), R.mist_cam().CAM("SSS", channels=[768, 384, 192, 96], n_class=9))):
        ours.load_state_dict(ref.state_dict(), strict=True)
        ref.load_state_dict(ours.state_dict(), strict=True)
