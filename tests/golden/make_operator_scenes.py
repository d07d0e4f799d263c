"""Generates tests/golden/operator_scenes.json: the `options.outputs` lists, filter / driver node declarations and camera node
entries of the reference's own test scenes (/root/reference/tests/*/*.ass) plus a few hand-written edge cases, each with what the
COMPILED REFERENCE's lentil_operator node (oracle/_ref/libref.so: /root/reference/src/lentil_operator.cpp behind oracle/shims/)
leaves behind after operator_cook -- the expected result tests/test_operator.py holds the adaptor's operator node to.

Run here (needs /root/reference and oracle/_ref/libref.so):  python -m tests.golden.make_operator_scenes
"""
import glob
import json
import os
import re

from oracle import ref

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TESTS = "/root/reference/tests"

BLOCK = re.compile(r"^([A-Za-z_][\w]*)\s*\n\{\s*\n(.*?)^\}", re.M | re.S)


def scene_from_ass(path: str):
    """-> scene text for ref_operator_cook: every named node as `node <name> <entry>`, the camera's entry, the outputs"""
    text = open(path, errors="replace").read()
    nodes, outputs, camera_name, aov_shaders = [], [], None, []
    for m in BLOCK.finditer(text):
        entry, body = m.group(1), m.group(2)
        if entry == "options":
            lines = body.splitlines()
            for i, ln in enumerate(lines):
                t = ln.strip()
                if t.startswith("outputs "):
                    one = re.match(r'outputs\s+"(.*)"\s*$', t)
                    if one:
                        outputs = [one.group(1)]
                    else:
                        n = int(t.split()[1])
                        outputs = [lines[i + 1 + k].strip().strip('"') for k in range(n)]
                elif t.startswith("camera "):
                    camera_name = t.split(None, 1)[1].strip('"')
                elif t.startswith("aov_shaders "):
                    rest = t.split()[1:]
                    aov_shaders = [r.strip('"') for r in rest if not r.isdigit() and r != "NODE"]
            continue
        nm = re.search(r"^\s*name\s+(\S+)", body, re.M)
        if nm:
            nodes.append((nm.group(1), entry))
    cam_entry = next((e for n, e in nodes if n == camera_name), "lentil_camera")
    keep = {tok for o in outputs for tok in o.split()}  # only the nodes the outputs name (filters, drivers) + aov shaders
    lines = [f"node {n} {e}" for n, e in nodes if n in keep or n in aov_shaders]
    lines.append(f"camera {cam_entry}")
    lines += [f"aov_shader {n}" for n in aov_shaders if any(n == x for x, _ in nodes)]
    lines += [f"output {o}" for o in outputs]
    return "\n".join(lines), len(outputs)


HAND = {
    # every branch of lentil_operator.cpp:44-99: unsupported filter kind, unsupported data type, unranked / ranked cryptomatte
    # outputs, camera token, HALF flag, the same AOV on two drivers, an output already on the shared filter
    "hand_branches": "\n".join([
        "node gauss gaussian_filter", "node closest closest_filter", "node variance variance_filter", "node box box_filter",
        "node crypto_flt cryptomatte_filter", "node exr driver_exr", "node exr2 driver_exr", "node png driver_png",
        "node cryptomatte_shader cryptomatte", "aov_shader cryptomatte_shader",
        "output RGBA RGBA gauss exr",
        "output persp RGBA RGBA gauss exr2 HALF",
        "output Z FLOAT closest exr",
        "output N VECTOR variance exr HALF",
        "output diffuse RGB box exr",
        "output ID UINT closest exr",
        "output crypto_material RGB crypto_flt exr",
        "output crypto_material00 FLOAT crypto_flt exr",
        "output crypto_object01 FLOAT crypto_flt exr",
        "output lentil_replaced_filter RGBA gauss png",
        "output persp Z FLOAT closest png",
        "output short two tokens",
    ]),
    "hand_not_lentil_camera": "node gauss gaussian_filter\nnode exr driver_exr\ncamera persp_camera\noutput RGBA RGBA gauss exr",
    "hand_existing_shared_filter": "node lentil_replaced_filter lentil_filter\nnode gauss gaussian_filter\nnode exr driver_exr\n"
                                   "output RGBA RGBA lentil_replaced_filter exr\noutput albedo RGB gauss exr",
    "hand_missing_filter_node": "node exr driver_exr\noutput RGBA RGBA nowhere exr",
}


def main():
    L = ref.lib()
    ref._declare_scenario_api(L)
    scenes = {}
    for path in sorted(glob.glob(os.path.join(REF_TESTS, "*", "*.ass"))):
        scene, n = scene_from_ass(path)
        if n == 0:
            continue
        key = os.path.relpath(path, REF_TESTS)
        scenes[key] = scene
        # scenes exported with the plugin's earlier camera node names (lentil_thinlens, pota, ...) are left alone by the operator;
        # the same output lists under today's camera node are the variant that exercises the cook
        lines = scene.splitlines()
        if "camera lentil_camera" not in lines:
            scenes[key + "+lentil_camera"] = "\n".join("camera lentil_camera" if ln.startswith("camera ") else ln for ln in lines)
    scenes.update(HAND)
    out = {"loader": ref.node_loader(L), "interface": ref.node_interface(L), "scenes": {}}
    for name, scene in scenes.items():
        out["scenes"][name] = {"scene": scene, "cook1": ref.operator_cook(L, scene, 1), "cook2": ref.operator_cook(L, scene, 2)}
    with open(os.path.join(HERE, "operator_scenes.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(f"{len(out['scenes'])} scenes, loader: {out['loader']}")


if __name__ == "__main__":
    main()
