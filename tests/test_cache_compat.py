"""Feature-cache interoperability with the reference's cached runner (refc:84-192).

The expected values in tests/golden/cache_golden.json and the bundle tests/golden/cache_ref_bundle.pt were produced by
the REFERENCE's own functions (scripts/make_golden.py::gold_cache, which also verified that the reference loads a bundle
written by our module instead of recomputing).  No GPU: the extractor is only asked for its ``to_dict()``; feature
computation is replaced by a stub so the cache logic can be driven on the CPU.
"""
import json
import os
import shutil

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "cache_golden.json")) as f:
        return json.load(f)


@pytest.fixture()
def fixture_files(gold):
    """Recreate the fake recordings at the absolute paths the golden keys were computed for."""
    root = gold["fixture_dir"]
    shutil.rmtree(root, ignore_errors=True)
    os.makedirs(os.path.join(root, "audio"))
    for f in gold["files"]:
        p = os.path.join(root, "audio", f["name"])
        with open(p, "wb") as fh:
            fh.write(bytes((i * 37 + 11) & 0xFF for i in range(f["size"])))
        os.utime(p, (f["mtime"], f["mtime"]))
    yield root
    shutil.rmtree(root, ignore_errors=True)


def _fx(name):
    from zenker_audio_detection_b200 import synth
    from zenker_audio_detection_b200.fx import ZenkerASTFeatureExtractor

    if name == "stage1_default":
        return ZenkerASTFeatureExtractor(mean=synth.STAGE1_MEAN, std=synth.STAGE1_STD)
    return ZenkerASTFeatureExtractor(max_length=16, mean=synth.STAGE2_MEAN, std=synth.STAGE2_STD)


class StubFx:
    """An extractor with the small extractor's identity whose features are a recognisable constant."""

    model_input_names = ["input_values"]

    def __init__(self, value=7.0):
        self.inner, self.value, self.calls = _fx("small"), value, 0

    def to_dict(self):
        return self.inner.to_dict()

    def __call__(self, batch, sampling_rate=None, return_tensors=None, **kw):
        self.calls += 1
        assert sampling_rate == 16000 and return_tensors == "pt"
        return {"input_values": torch.full((len(batch), 16, 128), self.value)}


def test_fingerprint_and_dict_match_the_reference(gold):
    from zenker_audio_detection_b200 import cache

    for name in ("stage1_default", "small"):
        fx = _fx(name)
        assert fx.to_dict() == gold["fx_dicts"][name]
        assert cache.get_fx_fingerprint(fx) == gold["fingerprints"][name]


def test_cache_path_and_metadata_match_the_reference(gold, fixture_files):
    from zenker_audio_detection_b200 import cache

    cache_dir = os.path.join(fixture_files, "cache")
    for k in gold["keys"]:
        path = os.path.join(fixture_files, "audio", k["file"])
        fp = gold["fingerprints"][k["fx"]]
        assert cache.build_cache_path(cache_dir, path, k["window_sec"], k["hop_sec"], 16000, fp) == k["cache_path"]
        meta = cache.build_base_metadata(path, k["window_sec"], k["hop_sec"], k["num_windows"], 16000, fp)
        assert meta == k["base_metadata"]
        assert json.loads(json.dumps(meta)) == meta  # plain ints / floats / strings only


def test_bundle_written_by_the_reference_is_loaded(gold, fixture_files):
    from zenker_audio_detection_b200 import cache

    rb = gold["ref_bundle"]
    os.makedirs(os.path.dirname(rb["cache_path"]))
    shutil.copy(os.path.join(GOLD, rb["file"]), rb["cache_path"])
    fx, lines = StubFx(), []
    path = os.path.join(fixture_files, "audio", "rec_A.wav")
    windows = [np.zeros(16000, np.float32)] * rb["num_windows"]
    feats = cache.load_or_compute_features(path, windows, fx, 1.0, 0.5, 2, os.path.dirname(rb["cache_path"]), False, False,
                                           "stage1", log=lines.append)
    assert fx.calls == 0 and lines == [f"[cache:stage1] Loaded {rb['cache_path']}"]
    assert list(feats.shape) == rb["feature_shape"] and feats.dtype == torch.float32 and not feats.is_cuda
    assert float(feats.double().sum()) == rb["feature_sum"]


def test_mismatch_refresh_disable_and_corrupt_files(gold, fixture_files):
    from zenker_audio_detection_b200 import cache

    rb = gold["ref_bundle"]
    cdir = os.path.dirname(rb["cache_path"])
    os.makedirs(cdir)
    shutil.copy(os.path.join(GOLD, rb["file"]), rb["cache_path"])
    path = os.path.join(fixture_files, "audio", "rec_A.wav")
    fx, lines = StubFx(3.0), []
    # (1) a different window count: metadata mismatch -> recompute and overwrite (refc:166-168,174-181)
    five = [np.zeros(16000, np.float32)] * 5
    feats = cache.load_or_compute_features(path, five, fx, 1.0, 0.5, 2, cdir, False, False, "stage1", log=lines.append)
    assert fx.calls == 3 and tuple(feats.shape) == (5, 16, 128) and float(feats[0, 0, 0]) == 3.0
    assert lines == [f"[cache:stage1] Metadata mismatch for {rb['cache_path']}; recomputing.",
                     f"[cache:stage1] Saved {rb['cache_path']}"]
    bundle = torch.load(rb["cache_path"], map_location="cpu")
    ref_bundle = torch.load(os.path.join(GOLD, rb["file"]), map_location="cpu")
    assert set(bundle) == set(ref_bundle) == {"metadata", "features"}
    assert set(bundle["metadata"]) == set(ref_bundle["metadata"])
    for k, v in ref_bundle["metadata"].items():
        assert type(bundle["metadata"][k]) is type(v), k
    assert bundle["metadata"]["feature_shape"] == [5, 16, 128] and bundle["metadata"]["num_windows"] == 5
    assert bundle["features"].dtype == torch.float32 and not bundle["features"].is_cuda
    # (2) now it loads
    lines.clear()
    again = cache.load_or_compute_features(path, five, fx, 1.0, 0.5, 2, cdir, False, False, "stage1", log=lines.append)
    assert fx.calls == 3 and torch.equal(again, feats) and lines[0].startswith("[cache:stage1] Loaded")
    # (3) --refresh-cache recomputes and saves, --disable-cache computes and touches nothing (refc:149-151,159)
    lines.clear()
    cache.load_or_compute_features(path, five, fx, 1.0, 0.5, 4, cdir, False, True, "stage2", log=lines.append)
    assert fx.calls == 5 and lines == [f"[cache:stage2] Saved {rb['cache_path']}"]
    before = os.path.getmtime(rb["cache_path"])
    lines.clear()
    cache.load_or_compute_features(path, five, fx, 1.0, 0.5, 5, cdir, True, False, "stage1", log=lines.append)
    cache.load_or_compute_features(path, five, fx, 1.0, 0.5, 5, None, False, False, "stage1", log=lines.append)
    assert fx.calls == 7 and lines == ["[cache:stage1] Computing features (cache disabled)."] * 2
    assert os.path.getmtime(rb["cache_path"]) == before
    # (4) an unreadable bundle is reported and recomputed (refc:171-172)
    with open(rb["cache_path"], "wb") as f:
        f.write(b"not a torch file")
    lines.clear()
    cache.load_or_compute_features(path, five, fx, 1.0, 0.5, 5, cdir, False, False, "stage1", log=lines.append)
    assert lines[0].startswith(f"[cache:stage1] Failed to load {rb['cache_path']}:") and lines[0].endswith("; recomputing.")
    assert lines[1] == f"[cache:stage1] Saved {rb['cache_path']}"
    # (5) no windows: the reference's error (refc:135-136)
    with pytest.raises(RuntimeError, match="yielded no data"):
        cache.compute_features(fx, [], 4)


def test_window_audio_matches_the_reference_geometry():
    from zenker_audio_detection_b200 import cache

    with open(os.path.join(GOLD, "glue_windows.json")) as f:
        cases = json.load(f)
    cases = cases["cases"] if isinstance(cases, dict) else cases
    checked = 0
    for c in cases:
        L, w, h = c["L"], c["window_sec"], c["hop_sec"]
        wins = cache.window_audio(np.arange(L, dtype=np.float32), w, h)
        assert wins.shape == (c["n"], int(w * 16000))
        assert [int(x) for x in wins[:, 0]] == ([0] if L == 0 else c["starts"])
        checked += 1
    assert checked >= 50


def test_empty_feature_tensor_gives_the_reference_empty_array():
    from zenker_audio_detection_b200 import cache

    out = cache.forward_probs_from_features(object(), torch.zeros((0, 16, 128)), 8)  # refc:208
    assert out.shape == (0, 0)
