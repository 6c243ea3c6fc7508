"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/art_hotpath.h declares; without a GPU it refuses to compute (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

import art_b200
from art_b200 import api


def test_header_declares_symbols():
    assert "art_hp_create" in api.ABI_SYMBOLS and "art_hp_demosaic_bayer" in api.ABI_SYMBOLS
    assert len(api.ABI_SYMBOLS) >= 14


def test_library_exports_every_declared_symbol():
    lib = art_b200.load_library()
    out = subprocess.check_output(["nm", "-D", "--defined-only", art_b200.lib_path()], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in api.ABI_SYMBOLS if s not in exported]
    assert not missing, "declared in include/art_hotpath.h but not exported: %s" % missing
    for s in api.ABI_SYMBOLS:
        assert getattr(lib, s) is not None
    assert lib.art_hp_abi_version() == 5


def test_header_is_plain_c():
    """The boundary header must compile as C (no C++ or torch types in the signatures)."""
    src = '#include "art_hotpath.h"\nint main(void){return art_hp_abi_version()==0;}\n'
    inc = os.path.join(os.path.dirname(art_b200.lib_path()), "..", "include")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c", "-"],
                   input=src, text=True, check=True)


def test_no_cpu_fallback_without_gpu():
    lib = art_b200.load_library()
    if lib.art_hp_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(art_b200.HotPathError) as e:
        art_b200.HotPath(0)
    assert e.value.code == 2   # ART_HP_ERR_NO_DEVICE


def test_product_does_not_reference_oracle():
    """Nothing under art_b200/ may import, link or execute oracle/ (tier rule 3)."""
    root = os.path.dirname(art_b200.lib_path())
    bad = []
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "libartoracle", "libartref", "oracle/"):
                    if needle in text and f != "__init__.py":
                        bad.append((f, needle))
    assert not bad, bad
    out = subprocess.check_output(["ldd", art_b200.lib_path()], text=True)
    assert "artoracle" not in out and "artref" not in out


def test_ctypes_structs_match_the_header(tmp_path):
    """The Python mirror's ctypes structures have the size and field offsets the C compiler gives the header's structs."""
    import ctypes
    pairs = {
        "art_hp_denoise_params": (api._DenoiseParamsC, ["luminance", "luminanceDetail", "luminanceDetailThreshold", "chrominance", "gamma", "scale",
                                                         "colorSpace", "noiseCCurve", "noiseCCurveSum", "wprof_inverse", "chrominanceAutoFactor"]),
        "art_hp_develop_params": (api._DevelopParamsC, ["method", "filters", "initialGain", "border", "mul", "doClip", "cam2work", "denoise",
                                                         "nlStrength", "fattal_enabled", "fattal_satcontrol", "wprof", "sharpen", "chain", "xtrans", "rgb_cam",
                                                         "full_frame", "guidedChromaRadius", "denoise_expcomp", "tran", "hr_blend", "hlmax"]),
        "art_hp_chain_params": (api._ChainParamsC, ["exposure_enabled", "exp_scale", "black", "saturation_enabled", "vibrance", "tonecurve_mode",
                                                     "tonecurve_lut", "rcurve", "bcurve", "lab_enabled", "lab_lcurve", "lab_bcurve", "lab_chroma", "ws", "iws",
                                                     "tonecurve_whitept", "tonecurve_stages", "tonecurve_nstages", "neutral_to_out", "neutral_to_work", "satcurve_lut", "softlight_lut"]),
        "art_hp_bw_params": (api._BwParamsC, ["bwr", "bwg", "bwb", "kcorec", "gamma_r", "gamma_g", "gamma_b", "ulut", "vlut", "ws"]),
        "art_hp_toneeq_params": (api._ToneEqParamsC, ["bands", "regularization", "pivot", "scale", "ws"]),
        "art_hp_flat_curve": (api._FlatCurveC, ["n", "poly_x", "poly_y", "dy_by_dx"]),
        "art_hp_hsl_params": (api._HslParamsC, ["hcurve", "scurve", "lcurve", "coeff", "smoothing", "scale", "ws"]),
        "art_hp_curve_stage": (api._CurveStageC, ["kind", "poly_x", "poly_y", "n", "a", "b", "w"]),
        "art_hp_sharpen_params": (api._SharpenParamsC, ["contrast", "radius", "amount", "threshold", "edgesonly", "halocontrol", "halocontrol_amount", "scale", "method", "deconvradius", "deconvamount", "deconvCornerBoost", "deconvCornerLatitude", "offset_x", "full_height", "edges_radius", "edges_tolerance"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "art_hotpath.h"', 'int main(void){']
    for s, (_, fields) in pairs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (s, s))
        for f in fields:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (s, f, s, f))
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(art_b200.lib_path()), "..", "include")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for s, (cls, fields) in pairs.items():
        assert int(out[s]) == ctypes.sizeof(cls), s
        for f in fields:
            assert int(out["%s.%s" % (s, f)]) == getattr(cls, f).offset, (s, f)
