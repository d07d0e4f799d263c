"""Generates tests/golden/reference_scene_cameras.json: the `lentil_camera` node of every reference test scene that carries one
under today's node name (/root/reference/tests/*/*.ass) -- the parameters it sets (everything else stays at the defaults of
lentil_camera.cpp:19-52) and its camera-to-world matrix -- so that the parity tests can run the parameter sets the reference's own
scenes were rendered with.  Parameters the current node does not declare (optical_vignetting_radius, optical_vignetting_distance,
cryptomatte: leftovers of earlier plugin versions in the exported files) are listed under "ignored", as Arnold would warn.

Run here (needs /root/reference):  python -m tests.golden.make_scene_cameras
"""
import glob
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TESTS = "/root/reference/tests"
# lentil_camera.cpp:19-52
DECLARED = {"camera_type", "bidir_sample_mult", "units", "sensor_width", "enable_dof", "fstop", "focus_dist", "aperture_blades_lentil", "exp",
            "lens_model", "wavelength", "extra_sensor_shift", "focal_length_lentil", "optical_vignetting", "abb_spherical", "abb_distortion",
            "abb_coma", "abb_chromatic", "abb_chromatic_type", "bokeh_circle_to_square", "bokeh_anamorphic", "bokeh_enable_image",
            "bokeh_image_path", "vignetting_retries", "bidir_add_energy", "bidir_add_energy_minimum_luminance", "bidir_add_energy_transition",
            "enable_bidir_transmission", "enable_skydome"}
ARNOLD_CAMERA = {"name", "matrix", "near_clip", "far_clip", "shutter_start", "shutter_end", "shutter_type", "rolling_shutter",
                 "rolling_shutter_duration", "motion_start", "motion_end", "exposure", "declare", "dcc_name", "screen_window_min",
                 "screen_window_max", "filtermap", "uv_remap", "handedness", "time_samples", "shutter_curve", "position", "look_at", "up"}


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(REF_TESTS, "*", "*.ass"))):
        text = open(path, errors="replace").read()
        blocks = re.findall(r"^lentil_camera\s*\n\{\s*\n(.*?)^\}", text, re.M | re.S)
        if not blocks:
            continue
        o = re.search(r"^options\s*\n\{\s*\n(.*?)^\}", text, re.M | re.S)
        active = re.search(r'^\s*camera\s+"?([^"\n]+)"?', o.group(1), re.M).group(1) if o else None
        # the camera the scene renders through (options.camera) when several lentil_camera nodes exist
        m = next((b for b in blocks if re.search(r"^\s*name\s+%s\s*$" % re.escape(active or ""), b, re.M)), blocks[0])
        m = re.match(r"(.*)", m, re.S)
        lines = [ln.strip() for ln in m.group(1).splitlines() if ln.strip()]
        params, ignored, matrix, options = {}, {}, None, {}
        i = 0
        while i < len(lines):
            tok = lines[i].split(None, 1)
            key, val = tok[0], (tok[1] if len(tok) > 1 else "")
            if key == "matrix":
                matrix = [[float(x) for x in lines[i + 1 + r].split()] for r in range(4)]
                i += 5
                continue
            if key in DECLARED:
                v = val.strip('"')
                params[key] = (1 if v == "on" else 0 if v == "off" else float(v) if re.fullmatch(r"[-+0-9.eE]+", v) else v)
            elif key not in ARNOLD_CAMERA:
                ignored[key] = val
            i += 1
        for ln in (o.group(1).splitlines() if o else []):
            t = ln.split()
            if len(t) == 2 and t[0] in ("xres", "yres", "AA_samples"):
                options[t[0]] = int(t[1])
        out[os.path.relpath(path, REF_TESTS)] = {"camera": active, "params": params, "ignored": ignored, "camera_to_world": matrix, "options": options}
    with open(os.path.join(HERE, "reference_scene_cameras.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, v["params"], v["options"], "ignored:", sorted(v["ignored"]))


if __name__ == "__main__":
    main()
