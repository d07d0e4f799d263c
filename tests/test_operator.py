"""The scene pre-pass in front of the redistribution path: the adaptor's `lentil_operator` node and plugin entry point
(adaptor/lentil_b200_operator.cpp, lentil_b200_loader.cpp) against the COMPILED REFERENCE's
(/root/reference/src/lentil_operator.cpp:19-191, lentil_loader.cpp:20-28 in oracle/_ref/libref.so).

Both libraries run the same scenario code (oracle/ref_harness.cpp: ref_operator_cook, ref_node_loader) on the `options.outputs`
lists and filter / driver nodes of the reference's own test scenes (/root/reference/tests/*/*.ass, extracted by
tests/golden/make_operator_scenes.py into tests/golden/operator_scenes.json together with the reference's answers) and on
hand-written scenes covering every branch of lentil_operator.cpp:44-99, and dump what every node DECLARES to the renderer
(ref_node_interface: parameters, metadata, required AOVs, filter width / output types, render hints).  Compared: cook's return value, every field of every
OperatorData entry (name, type, remembered filter kind, duplicate flag, the five output tokens + HALF, resolved driver), the
nodes and links the cook creates, options.aov_shaders, and what the camera's rebuild + sanitize step turns the list into (the
framebuffers lb_filter_begin is asked for).  Host-side string work: exact equality.  No GPU needed."""
import json
import os

import pytest

from oracle import ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "operator_scenes.json")
with open(GOLDEN) as _f:
    FIX = json.load(_f)
SCENES = sorted(FIX["scenes"])


@pytest.fixture(scope="module")
def adaptor():
    if not ref.adaptor_available():
        pytest.skip("adaptor/_build/libadaptor.so not built")
    return ref.adaptor_lib()


@pytest.fixture(scope="module")
def reference():
    if not ref.available():
        pytest.skip("oracle/_ref/libref.so not built")
    L = ref.lib()
    ref._declare_scenario_api(L)
    return L


def test_fixture_covers_the_reference_scenes():
    real = [s for s in SCENES if s.endswith(".ass")]
    assert len(real) >= 10 and any(len(FIX["scenes"][s]["cook1"]) > 40 for s in real)
    cooked = [s for s in SCENES if "cook0=1" in FIX["scenes"][s]["cook1"]]
    refused = [s for s in SCENES if "cook0=0" in FIX["scenes"][s]["cook1"]]
    assert len(cooked) >= 16 and len(refused) >= 8  # scenes of older camera node names are left alone (lentil_operator.cpp:30-33)


def test_node_loader_matches_reference(adaptor):
    """same four nodes, names, node types, output types and version string, each with its own method table"""
    got = ref.node_loader(adaptor)
    assert got == FIX["loader"]
    assert [ln.split()[1] for ln in got] == ["lentil_camera", "lentil_filter", "imager_lentil", "lentil_operator"]
    assert [ln.split()[-1] for ln in got] == ["methods=camera", "methods=filter", "methods=imager", "methods=operator"]


def test_node_interface_matches_reference(adaptor):
    """what the four nodes declare to the renderer -- the 29 camera parameters with their types, defaults and enum strings
    (lentil_camera.cpp:19-52: a scene file written for the reference must load on the adaptor's node), node metadata, the AOVs the
    filter node requires (lentil_filter.cpp:16-25), its width with and without an OIDN imager (:33-38), filter_output_type over
    the data types (:45-63), the imager's `enable` parameter, subtype and render hints (lentil_imager.cpp:19-38)"""
    got = ref.node_interface(adaptor)
    assert got == FIX["interface"]
    cam = got[got.index("node lentil_camera") + 1:got.index("node lentil_filter")]
    assert len([ln for ln in cam if ln.startswith("  param ")]) == 29
    assert "  param enum lens_model default=16 values=" in "\n".join(cam) and "  requires FLOAT lentil_bidir_ignore" in got


def test_golden_is_the_compiled_reference(reference):
    assert ref.node_loader(reference) == FIX["loader"]
    assert ref.node_interface(reference) == FIX["interface"]
    for name in SCENES:
        sc = FIX["scenes"][name]
        assert ref.operator_cook(reference, sc["scene"], 1) == sc["cook1"], name
        assert ref.operator_cook(reference, sc["scene"], 2) == sc["cook2"], name


@pytest.mark.parametrize("name", SCENES)
def test_operator_cook_matches_reference(adaptor, name):
    sc = FIX["scenes"][name]
    got = ref.operator_cook(adaptor, sc["scene"], 1)
    assert got == sc["cook1"], "\n".join(x for x in got if x not in sc["cook1"])
    # a re-cooked scene: the list keeps growing and the second pass's entries are flagged duplicate (lentil_operator.cpp:90-97)
    assert ref.operator_cook(adaptor, sc["scene"], 2) == sc["cook2"]


def test_operator_result_feeds_filter_begin(adaptor):
    """the framebuffer list derived from the operator's result: one per distinct redistributable AOV, closest kinds kept"""
    lines = ref.operator_cook(adaptor, FIX["scenes"]["hand_branches"]["scene"], 1)
    fbs = [ln.split()[2:] for ln in lines if ln.startswith("framebuffer ")]
    assert fbs == [["RGBA", "gaussian_filter"], ["Z", "closest_filter"], ["N", "variance_filter"], ["diffuse", "gaussian_filter"],
                   ["lentil_debug", "closest_filter"], ["lentil_time", "gaussian_filter"], ["lentil_raydir", "gaussian_filter"]]


def test_scene_without_outputs_is_refused(adaptor):
    """the reference indexes aovs[0] of an empty list there (lentil_operator.cpp:107, undefined behaviour); the adaptor's node
    reports the scene and cooks nothing"""
    lines = ref.operator_cook(adaptor, "node gauss gaussian_filter", 1)
    assert "cook0=0" in lines and not any(ln.startswith("aov ") for ln in lines)


def test_operator_driven_frame_equals_the_harness_list(reference):
    """filter_pixel + driver_process_bucket of the compiled reference with the AOV list cooked by ITS OWN lentil_operator node
    from the scene's outputs, against the same frame with the list the harness writes by hand (oracle/ref_harness.cpp:
    build_operator_aovs) -- the stand-in every other reference-vs-oracle test relies on.  Bit-identical user AOVs; the operator's
    lentil_debug output equals the oracle's lentil_debug AOV."""
    import numpy as np

    from oracle import orc
    from pota_b200 import workloads
    from tests.util import po_params

    p = po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6)
    aovs = [("RGBA", 0, 1), ("light0", 0, 0), ("N", 1, 0)]
    W, H, spp = 96, 54, 9
    hand, cooked, o = ref.RefCamera(p), ref.RefCamera(p), orc.OracleCamera(p)
    cooked.use_operator(True)
    fr = workloads.highlight_frame(W, H, spp, o.state.tan_fov, "cpu", n_extra_aov=1)
    vals = [None, fr["aov_values"][0].numpy(), None]
    args = (fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp)
    for cam in (hand, cooked):
        cam.filter_begin(W, H, aovs, spp=spp)
        cam.filter_accumulate(*args, aov_values=vals)
    o.filter_begin(W, H, aovs + [("lentil_debug", 1, 2)])
    o.filter_accumulate(*args, aov_values=vals + [None])
    assert o.filter_stats()["redistributed"] > 20
    for i in range(len(aovs)):
        (bh, wh), (bc, wc) = hand.buffers(i), cooked.buffers(i)
        np.testing.assert_array_equal(bh, bc)
        np.testing.assert_array_equal(wh, wc)
        np.testing.assert_array_equal(hand.resolve(i), cooked.resolve(i))
    np.testing.assert_array_equal(cooked.resolve(len(aovs)), o.resolve(len(aovs)))  # lentil_debug, made by the operator
    cooked.use_operator(False)  # and back to the hand-written list on the same camera
    cooked.filter_begin(W, H, aovs, spp=spp)
    cooked.filter_accumulate(*args, aov_values=vals)
    np.testing.assert_array_equal(cooked.resolve(0), hand.resolve(0))
