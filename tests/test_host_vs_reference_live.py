"""Build-container only: the host-side helpers of the path (decode, file collation, fit, shuffles, C-vector
statistics) executed side by side with the unmodified reference on the same files and
arrays.  Skipped where /root/reference is absent (the GPU box); the committed golden vectors cover that side."""
import contextlib
import importlib
import io
import os
import sys

import cv2
import numpy as np
import pytest

import pylc_oracle as orc
import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    ref_harness.load()
    sys.path.insert(0, ref_harness.REF_ROOT)
    try:
        mods = {n: importlib.import_module(m) for n, m in (("tools", "utils.tools"), ("profile", "utils.profile"),
                                                            ("metrics", "utils.metrics"), ("config", "config"))}
    finally:
        sys.path.remove(ref_harness.REF_ROOT)
    return type("Ref", (), mods)


def run(fn, *a, **k):
    """(outcome, value, printed text): the reference signals errors with print + exit(1)."""
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            val = fn(*a, **k)
        return "ok", val, buf.getvalue()
    except SystemExit as e:
        return "exit", e.code, buf.getvalue()
    except AssertionError:
        return "assert", None, buf.getvalue()


@pytest.fixture(scope="module")
def files(tmp_path_factory, palettes):
    d = tmp_path_factory.mktemp("live")
    os.makedirs(d / "img")
    os.makedirs(d / "mask")
    os.makedirs(d / "odd")
    pal = palettes["a"]
    for i, (name, ext) in enumerate([("a", ".tif"), ("b", ".png"), ("c", ".jpg")]):
        assert cv2.imwrite(str(d / "img" / (name + ext)), orc.synth_image(i, 300, 200, 3)[..., ::-1])
        assert cv2.imwrite(str(d / "mask" / (name + ".png")), np.ascontiguousarray(orc.synth_mask(i, 300, 200, pal)[..., ::-1]))
    (d / "img" / "notes.txt").write_text("not an image")
    assert cv2.imwrite(str(d / "odd" / "only.png"), orc.synth_image(5, 64, 64, 3))
    assert cv2.imwrite(str(d / "gray.tif"), orc.synth_image(9, 300, 200, 1))
    return d


def test_collate_and_load_files(ref, files):
    from pylc_b200.utils import tools
    d = str(files)
    for args in ((d + "/img", d + "/mask"), (d + "/img",), (d + "/img/a.tif", d + "/mask/a.png"),
                 (d + "/img", d + "/odd"), (d + "/missing",)):
        want, got = run(ref.tools.collate, *args), run(tools.collate, *args)
        assert want[:2] == got[:2], args
    assert run(ref.tools.load_files, d + "/img", [".tif", ".png"])[:2] == run(tools.load_files, d + "/img", [".tif", ".png"])[:2]
    for p in ("/x/y/abc.def.tif", "rel/name.png", "noext"):
        assert ref.tools.get_fname(p) == tools.get_fname(p)


@pytest.mark.parametrize("name,ch,scale,interp", [
    ("img/a.tif", 3, None, cv2.INTER_AREA), ("img/b.png", 3, 0.5, cv2.INTER_AREA), ("img/c.jpg", 3, None, cv2.INTER_AREA),
    ("mask/a.png", 3, 0.4, cv2.INTER_NEAREST), ("gray.tif", 1, None, cv2.INTER_AREA), ("gray.tif", 1, 0.7, cv2.INTER_AREA),
    ("gray.tif", 3, None, cv2.INTER_AREA),      # a gray file asked for as colour: print + exit(1) in both
    ("img/a.tif", 1, None, cv2.INTER_AREA),     # a colour file asked for as gray (the reference prompts; stdin is empty)
])
def test_get_image(ref, files, name, ch, scale, interp, monkeypatch):
    from pylc_b200.utils import tools
    monkeypatch.setattr("builtins.input", lambda *a: "y")
    path = str(files / name)
    want = run(ref.tools.get_image, path, ch, scale=scale, interpolate=interp)
    got = run(tools.get_image, path, ch, scale=scale, interpolate=interp)
    assert want[0] == got[0]
    if want[0] == "ok":
        assert np.array_equal(want[1][0], got[1][0]) and tuple(want[1][1:]) == tuple(got[1][1:])


def test_adjust_to_tile_coshuffle_and_statistics(ref):
    from pylc_b200.utils import metrics, tools
    for seed, (w, h, ch) in enumerate([(700, 500, 1), (650, 430, 3), (512, 512, 1)]):
        img = orc.synth_image(seed, w, h, ch)
        want, got = ref.tools.adjust_to_tile(img, 128, 64, ch), tools.adjust_to_tile(img, 128, 64, ch)
        assert np.array_equal(want[0], got[0]) and tuple(want[1:]) == tuple(got[1:])
    a, b = np.arange(40).reshape(10, 4), np.arange(10)
    np.random.seed(5)
    want = ref.tools.coshuffle(a.copy(), b.copy())
    np.random.seed(5)
    got = tools.coshuffle(a.copy(), b.copy())
    assert all(np.array_equal(x, y) for x, y in zip(want, got))
    rng = np.random.default_rng(0)
    for C in (2, 9, 11):
        p = rng.random(C)
        p /= p.sum()
        q = np.full(C, 1.0 / C)
        assert ref.metrics.jsd(p, q) == metrics.jsd(p, q) and ref.metrics.m2(p, C) == metrics.m2(p, C)
    for c in ([0, 0, 0], [255, 128, 7]):
        assert ref.tools.rgb2hex(c) == tools.rgb2hex(c)
