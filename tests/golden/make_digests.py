#!/usr/bin/env python3
"""Generate tests/golden/digests.json: SHA-256 of what the REFERENCE (oracle/_ref/libartref_det.so, the reference's own function
bodies compiled where they lie under /root/reference) returns for seeded inputs, for the ports whose outputs are too many to
commit as planes: X-Trans demosaic, unsharp mask, the colour / curve chain, Fattal tone mapping, RGB_denoise.

Run in the build container (needs /root/reference):  python tests/golden/make_digests.py
tests/test_oracle_golden_digests.py recomputes the same inputs, runs the plain-C port and compares digests -- that test needs
neither /root/reference nor oracle/_ref.  The FFTW call sites of Fattal and RGB_denoise run the same double-precision stand-in on
both sides (fftw3f is absent), so those two digests pin everything around the transforms, not the transforms themselves."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def digest(planes):
    h = hashlib.sha256()
    for p in planes:
        h.update(np.ascontiguousarray(p, dtype=np.float32).tobytes())
    return h.hexdigest()


def cases(which):
    """name -> callable returning the output planes; `which` is "ref" or "port"."""
    import oracle
    from art_b200 import synth
    import test_oracle_xtrans as tx
    import test_oracle_usm as tu
    import test_oracle_chain as tc
    import test_oracle_fattal as tf
    import test_oracle_denoise as td
    import test_chain_gpu as tcg
    ref = which == "ref"
    lib = oracle.ref().lib if ref else oracle.port().lib
    pre = "artref_" if ref else "artoracle_"
    out = {}
    for passes, lab, dy, dx, W, H in [(3, 1, 0, 0, 233, 217), (1, 0, 2, 5, 131, 140), (3, 0, 4, 1, 120, 125)]:
        def f(passes=passes, lab=lab, dy=dy, dx=dx, W=W, H=H):
            xt = synth.xtrans_matrix(dy, dx)
            raw = synth.xtrans_frame(W, H, xt, seed=W + dy)
            return tx.ref_xtrans(raw, xt, passes, lab) if ref else tx.port_xtrans(raw, xt, passes, lab)
        out["xtrans_p%d_lab%d_o%d%d_%dx%d" % (passes, lab, dy, dx, W, H)] = f
    for k, kw in enumerate([dict(), dict(radius=0.9, amount=350), dict(contrast=55.0, radius=2.4, amount=80, thr=(10, 40, 1500, 600))]):
        def f(k=k, kw=kw):
            return tu.run(lib, pre + "usm", tu.scene(301, 203, 40 + k, wild=bool(k & 1)), **kw)[0]
        out["usm_case%d_301x203" % k] = f
    for name in ("all_std", "all_film"):
        def f(name=name):
            planes = tc.image(77, 130, 5)
            if not ref:
                return tcg.oracle_chain(planes, **tcg.STAGES[name])
            # the same sequence through the reference's loops
            kw = dict(tcg.STAGES[name])
            o = planes
            if "exposure" in kw:
                ev, black = kw["exposure"]
                o = tc.call(lib, "artref_chain_expcomp", o, tc.F(np.float32(2.0) ** np.float32(ev)), tc.F(np.float32(black) * np.float32(2000.0)))
            if "saturation" in kw and (kw["saturation"][0] or kw["saturation"][1]):
                o = tc.call(lib, "artref_chain_saturation", o, kw["saturation"][0], kw["saturation"][1], tc.PROPHOTO.ctypes.data_as(tc.dp))
            if "tonecurve" in kw:
                o = tc.call(lib, "artref_chain_tonecurve", o, kw["tonecurve"][0], kw["tonecurve"][1].ctypes.data_as(tc.fp), tc.F(1.0))
            if "rgbcurves" in kw:
                o = tc.call(lib, "artref_chain_rgbcurves", o, *[c.ctypes.data_as(tc.fp) if c is not None else None for c in kw["rgbcurves"]])
            if "lab" in kw:
                lab = kw["lab"]
                o = tc.call(lib, "artref_chain_lab", o, lab[0].ctypes.data_as(tc.fp), lab[1].ctypes.data_as(tc.fp), lab[2].ctypes.data_as(tc.fp), tc.F(lab[3]),
                            tc.PROPHOTO.ctypes.data_as(tc.dp), tc.PROPHOTO_INV.ctypes.data_as(tc.dp))
            return o
        out["chain_%s_130x77" % name] = f

    def f():
        return tf.fattal(lib, pre + "fattal", tf.scene(203, 301, 7), 30, 20, 1)
    out["fattal_301x203_30_20_sat"] = f

    def f():
        planes = td.rgb_frame(200, 264, seed=21)
        return td.run(lib, pre + "rgb_denoise", planes, (30, 50, 0, 15, 0, 0, 1.7, 1.0), td.noise_ccurve(), with_inverse=ref)[0]
    out["rgb_denoise_264x200_lum30_curve"] = f

    # later additions: rld (recursive GAUSS_DIV / GAUSS_MULT above sigma 1.15, corner boost), edges-only unsharp mask, recursive div / mult,
    # the automatic chroma estimator, the Lanczos resampler, green equilibration
    import test_oracle_denoise_auto as tda
    import test_oracle_resize as tr
    import test_oracle_greeneq as tg
    for k, kw in enumerate([dict(radius=0.75), dict(radius=1.8, amount=70)]):
        def f(k=k, kw=kw):
            return tu.run_rld(lib, pre + "rld", tu.scene(301, 203, 60 + k), **kw)[0]
        out["rld_case%d_301x203" % k] = f

    def f():
        return tu.run_rld_ex(lib, pre + "rld_ex", tu.scene(130, 77, 71, wild=True), boost=0.5, radius=1.0)
    out["rld_boost_130x77"] = f
    for k, kw in enumerate([dict(edges_radius=0.9, edges_tolerance=400), dict(edges_radius=2.2, halo=1)]):
        def f(k=k, kw=kw):
            return tu.run_edges(lib, pre + "usm_ex", tu.scene(130, 77, 80 + k), **kw)
        out["usm_edges_case%d_130x77" % k] = f
    for kind in ("mult", "div"):
        def f(kind=kind):
            import test_oracle_gauss as tgs
            src = tgs.image(53, 67, seed=120) - 9000.0
            dst = np.random.default_rng(5).uniform(-2.0, 3.0, size=(53, 67)).astype(np.float32)
            divb = (tgs.image(53, 67, seed=7) - 3000.0).astype(np.float32)
            return list((oracle.ref() if ref else oracle.port()).gauss_iir(src, dst, divb, 2.5, kind))
        out["gauss_iir_%s_67x53" % kind] = f

    def f():
        rows = [tda.info(lib, pre + "denoise_info", tda.crop(200 + 8 * k, 260 - 6 * k, 31 + k, k % 3), aggressive=k & 1) for k in range(9)]
        stats = np.ascontiguousarray(np.stack(rows), dtype=np.float32)
        return [stats, tda.auto(lib, pre + "denoise_auto_params", stats, 1, 0), tda.auto(lib, pre + "denoise_auto_params", stats, 0, 1)]
    out["denoise_auto_9crops"] = f

    def f():
        return tr.lanczos(lib, pre + "lanczos", tr.planes3(131, 203, 9), 156, 101, 0.77)
    out["lanczos_203x131_0.77"] = f

    def f():
        raw = tg.mosaic(203, 301, 17)
        import ctypes
        a = raw.copy()
        getattr(lib, pre + "green_equilibrate_global")(a.ctypes.data_as(tg.fp), 301, 203, ctypes.c_uint(0x94949494), 4)
        getattr(lib, pre + "green_equilibrate")(a.ctypes.data_as(tg.fp), 301, 203, ctypes.c_uint(0x94949494), ctypes.c_float(0.05), None)
        return [a]
    out["greeneq_301x203_rggb"] = f

    import test_oracle_dual as tdu
    import test_oracle_pack as tp

    def f():
        raw = synth.bayer_frame(301, 203, 0x94949494, seed=31)
        planes = list(oracle.port().amaze(raw, 0x94949494))       # the first demosaicer's planes are an input here (AMaZE has its own pins)
        if ref:
            o, blend, _ = tdu.ref_chain(raw, 0x94949494, planes, 20.0)
        else:
            o, blend = tdu.port_chain(raw, 0x94949494, planes, 20.0)
        return o + [blend]
    out["dual_amaze_bilinear_301x203_c20"] = f

    def f():
        lum = tdu.lum_scene(333, 251, 334, 1)
        import ctypes
        if ref:
            thr = ctypes.c_float(0.2)
            blend = np.zeros_like(lum)
            lib.artref_blend_mask_ex(tdu.P(lum), tdu.P(blend), 333, 251, ctypes.byref(thr), tdu.F(1.0), 1, tdu.F(2.0))
            return [np.array([thr.value], np.float32), blend]
        fn = lib.artoracle_auto_contrast_threshold
        fn.restype = ctypes.c_float
        t = fn(tdu.P(lum), 333, 251, tdu.F(0.2), tdu.F(1.0))
        blend = np.zeros_like(lum)
        lib.artoracle_blend_mask(tdu.P(lum), tdu.P(blend), 333, 251, tdu.F(t), tdu.F(1.0), tdu.F(2.0))
        return [np.array([t], np.float32), blend]
    out["auto_contrast_333x251"] = f

    def f():
        import test_oracle_vng4 as tv
        raw = synth.bayer_frame(130, 77, 0x61616161, seed=9)
        return tv.vng4(lib, pre + "vng4", raw, 0xe1e1e1e1, *([1] if ref else []))
    out["vng4_130x77_grbg"] = f

    def f():
        import test_oracle_smoothing as tsm
        return tsm.run(lib, pre + "denoise_guided_smoothing", td.rgb_frame(200, 264, seed=4, noise=2500.0), 3, 1.0)
    out["guided_smoothing_264x200_r3"] = f

    def f():
        import test_oracle_hlblend as thl
        hl = (124000.0, 65535.0, 98000.0)
        return thl.run(lib, pre + "hl_blend", thl.line(4001, 5, hl), hl)
    out["hl_blend_4001"] = f

    def f():
        planes = tp.frame(35, 67, 102)
        return [tp.scan(lib, pre + "scanlines", planes, bps, fl).view(np.uint8).astype(np.float32) for bps, fl in ((8, 0), (16, 0), (16, 1))]
    out["scanlines_67x35"] = f
    return out


def main():
    d = {name: digest(fn()) for name, fn in cases("ref").items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "digests.json")
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, len(d), "digests")


if __name__ == "__main__":
    main()
