"""Host-side mirror of the reference's two call contracts (no GPU needed): feature-extractor (de)serialisation,
argument validation / error behaviour, model directory loading, the transformers swap, sharding and the
world-size-2 gloo gather."""
import json
import os

import numpy as np
import pytest
import torch

from zenker_audio_detection_b200 import cascade, compat, dist as zdist, synth
from zenker_audio_detection_b200.fx import ZenkerASTFeatureExtractor
from zenker_audio_detection_b200.model import ZenkerASTForAudioClassification


def test_fx_roundtrip_and_hf_interop(tmp_path):
    from transformers import ASTFeatureExtractor

    hf = ASTFeatureExtractor(mean=synth.STAGE1_MEAN, std=synth.STAGE1_STD)
    hf.save_pretrained(str(tmp_path / "hf"))
    ours = ZenkerASTFeatureExtractor.from_pretrained(str(tmp_path / "hf"))
    assert ours.to_dict() == hf.to_dict()
    assert ours.model_input_names[0] == "input_values"
    ours.mean, ours.std = -2.0, 4.0
    ours.save_pretrained(str(tmp_path / "ours"))
    back = ASTFeatureExtractor.from_pretrained(str(tmp_path / "ours"))
    assert back.mean == -2.0 and back.std == 4.0 and back.max_length == 1024
    json.dumps(ours.to_dict(), sort_keys=True)  # refc:84-86 fingerprints this
    with pytest.raises(OSError):
        ZenkerASTFeatureExtractor.from_pretrained(str(tmp_path / "missing"))


def test_fx_argument_errors_match_hf():
    fx = ZenkerASTFeatureExtractor()
    with pytest.raises(ValueError, match="sampling rate"):
        fx(np.zeros(16000, np.float32), sampling_rate=8000)
    with pytest.raises(ValueError, match="mono-channel"):
        fx(np.zeros((2, 3, 100), np.float32), sampling_rate=16000)


def test_model_from_pretrained_reads_hf_directory(tmp_path):
    from safetensors.torch import save_file
    from transformers import ASTConfig

    sd = synth.random_state_dict(1)
    root = tmp_path / "m"
    root.mkdir()
    ASTConfig(num_labels=2).save_pretrained(str(root))
    save_file({k: v.contiguous() for k, v in sd.items()}, str(root / "model.safetensors"))
    cfg = ASTConfig.from_pretrained(str(root))
    cfg.label2id = {"Idle": 0, "Swallow": 1}
    cfg.id2label = {0: "Idle", 1: "Swallow"}
    m = ZenkerASTForAudioClassification.from_pretrained(str(root), config=cfg)
    assert m.eval() is m and m.num_labels == 2 and m.max_length == 1024 and m.ln_eps == 1e-12
    assert m.config.id2label[1] == "Swallow"
    assert len(m.state_dict()) == 203
    with pytest.raises(ValueError, match="input_values"):
        m(None)
    m2 = ZenkerASTForAudioClassification.from_pretrained(str(root))  # config.json read directly
    assert m2.max_length == 1024
    bad = ASTConfig(num_labels=2, hidden_size=384)
    with pytest.raises(Exception):
        ZenkerASTForAudioClassification(bad, sd)


def test_patch_transformers_swaps_the_two_names():
    import sys

    import transformers  # noqa: F401

    transformers = sys.modules["transformers"]
    orig = transformers.ASTFeatureExtractor
    compat.patch_transformers()
    try:
        from transformers import ASTConfig, ASTFeatureExtractor, ASTForAudioClassification

        assert ASTFeatureExtractor is ZenkerASTFeatureExtractor
        assert ASTForAudioClassification is ZenkerASTForAudioClassification
        assert ASTConfig is transformers.models.audio_spectrogram_transformer.ASTConfig
    finally:
        compat.unpatch_transformers()
    assert transformers.ASTFeatureExtractor is orig


def test_shard_recordings_is_balanced_and_deterministic():
    rng = np.random.default_rng(0)
    lengths = rng.integers(8 * 60, 12 * 60, size=200).tolist()
    for ws in (1, 2, 4, 8):
        shards = zdist.shard_recordings(lengths, ws)
        assert sorted(i for s in shards for i in s) == list(range(200))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)
        assert shards == zdist.shard_recordings(lengths, ws)
    assert zdist.shard_recordings([5], 4) == [[0], [], [], []]


def test_shard_window_ranges_cover_every_window_once_in_equal_runs():
    rng = np.random.default_rng(2)
    counts = rng.integers(150, 1300, size=64).tolist() + [1, 0, 7]
    for ws in (1, 2, 3, 8):
        shards = zdist.shard_window_ranges(counts, ws)
        per = [sum(b - a for _, a, b in s) for s in shards]
        assert sum(per) == sum(counts) and max(per) - min(per) <= 1          # equal runs of windows
        seen = {}
        for s in shards:
            assert s == sorted(s)
            for i, a, b in s:
                seen.setdefault(i, []).append((a, b))
        for i, c in enumerate(counts):                                       # every window exactly once, in order
            pieces = sorted(seen.get(i, []))
            assert (pieces == []) if c == 0 else (pieces[0][0] == 0 and pieces[-1][1] == c)
            assert all(pieces[k][1] == pieces[k + 1][0] for k in range(len(pieces) - 1))
        assert sum(1 for i in seen if len(seen[i]) > 1) <= ws - 1            # at most one cut between neighbouring ranks
        assert shards == zdist.shard_window_ranges(counts, ws)
    assert zdist.shard_window_ranges([1], 4) == [[], [], [], [(0, 0, 1)]]
    # proportional runs: a rank that measured 10 % more windows/s gets 10 % more windows; same coverage rules
    w = [1.0, 1.1, 0.9, 1.0]
    shards = zdist.shard_window_ranges(counts, 4, weights=w)
    per = [sum(b - a for _, a, b in s) for s in shards]
    assert sum(per) == sum(counts)
    assert all(abs(p - sum(counts) * wi / sum(w)) <= 1.0 for p, wi in zip(per, w))
    assert zdist.shard_window_ranges(counts, 4, weights=[2.0] * 4) == zdist.shard_window_ranges(counts, 4)
    with pytest.raises(ValueError):
        zdist.shard_window_ranges(counts, 4, weights=[1.0, 0.0, 1.0, 1.0])


def test_records_of_a_split_recording_merge_back():
    rng = np.random.default_rng(3)
    s1 = rng.random((37, 2)).astype(np.float32)
    idx = np.array([0, 5, 6, 30, 36], dtype=np.int64)
    s2 = rng.random((5, 2)).astype(np.float32)
    whole = zdist.pack_records(7, s1, idx, s2)
    cut = 6   # windows 0..5 on one rank, 6..36 on the next; Stage-2 indices are relative to the chunk
    a = zdist.pack_records(7, s1[:cut], idx[idx < cut], s2[idx < cut])
    b = zdist.pack_records(7, s1[cut:], idx[idx >= cut] - cut, s2[idx >= cut], window_base=cut)
    assert np.array_equal(np.concatenate([a, b]), whole)
    got = zdist.unpack_records(np.concatenate([b, a]))[7]
    assert np.array_equal(got[0], s1) and np.array_equal(got[1], idx) and np.array_equal(got[2], s2)


def test_record_pack_unpack_roundtrip():
    rng = np.random.default_rng(1)
    s1 = rng.random((37, 2)).astype(np.float32)
    idx = np.array([0, 5, 6, 30], dtype=np.int64)
    s2 = rng.random((4, 2)).astype(np.float32)
    rec = zdist.pack_records(7, s1, idx, s2)
    rec0 = zdist.pack_records(3, s1[:0], idx[:0], s2[:0])
    out = zdist.unpack_records(np.concatenate([rec[::-1], rec0]))
    a, b, c = out[7]
    assert np.array_equal(a, s1) and np.array_equal(b, idx) and np.array_equal(c, s2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [1199, 600, 900, 1199, 300]
    mine = zdist.shard_recordings(lengths, world)[rank]
    blocks = []
    for rid in mine:
        g = np.random.default_rng(100 + rid)  # scores depend on the recording only, never on the rank
        n = lengths[rid]
        s1 = g.random((n, 2)).astype(np.float32)
        idx = np.where(s1[:, 1] > 0.7)[0]
        s2 = g.random((len(idx), 2)).astype(np.float32)
        blocks.append(zdist.pack_records(rid, s1, idx, s2))
    local = np.concatenate(blocks) if blocks else np.zeros((0, zdist.RECORD_WIDTH))
    allrec = zdist.all_gather_records(local, torch.device("cpu"))
    per = zdist.unpack_records(allrec)
    assert allrec.dtype == np.int32 and allrec.shape[1] == 6  # 24 bytes per window on the wire
    summ = {rid: cascade.summarize_stage_outputs(*per[rid], 0.5) for rid in sorted(per)}
    docs = zdist.patient_documents(per, [[0, 1], [2, 3]], 0.5)  # ref:361-382: two files per patient
    assert docs[0]["aggregate"] == cascade.aggregate_patient(docs[0]["per_file"], ["rec0", "rec1"])
    q.put((rank, json.dumps(summ, sort_keys=True)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_all_gather_records_gloo(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(set(outs.values())) == 1  # every rank reconstructs the identical per-recording summaries
    # and they equal the single-process result
    lengths = [1199, 600, 900, 1199, 300]
    ref = {}
    for rid, n in enumerate(lengths):
        g = np.random.default_rng(100 + rid)
        s1 = g.random((n, 2)).astype(np.float32)
        idx = np.where(s1[:, 1] > 0.7)[0]
        s2 = g.random((len(idx), 2)).astype(np.float32)
        ref[rid] = cascade.summarize_stage_outputs(s1, idx, s2, 0.5)
    assert json.loads(outs[0]) == json.loads(json.dumps({str(k): v for k, v in ref.items()}, sort_keys=True))


def _plan_worker(rank, world, port, root, q):
    """batch.plan_patients under two ranks where rank 1 arrives late and rank 0 has ALREADY written a result file: the
    plan is rank 0's, broadcast, so both ranks shard the same list (per-rank os.path.exists plans diverge)."""
    import time

    from zenker_audio_detection_b200 import batch

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    args = batch.build_arg_parser().parse_args(["--fold", "1", "--ids-root", os.path.join(root, "ids"), "--long-audio-root",
                                                os.path.join(root, "long"), "--output-dir", os.path.join(root, "out")])
    if rank == 0:
        todo, _ = batch.global_plan(args)  # what the plan looks like before anything was written
        q.put(("before", sorted(p for p, _, _ in todo)))
    else:
        time.sleep(1.0)  # a slow starter: by now rank 0's first result exists on disk
    if rank == 0:
        with open(os.path.join(root, "out", "11_2stage.json"), "w") as f:
            f.write("{}")
    mine, _ = batch.plan_patients(args, rank, world)
    q.put((rank, sorted(p for p, _ in mine)))
    import torch.distributed as dist

    dist.destroy_process_group()


def test_batch_plan_is_made_once_and_broadcast(tmp_path):
    import torch.multiprocessing as mp

    from zenker_audio_detection_b200 import wavio

    root = tmp_path
    (root / "ids").mkdir()
    (root / "out").mkdir()
    pids = ["11", "22", "33", "44"]
    (root / "ids" / "test_ids_fold1.txt").write_text("".join(f"Healthy/{p}\n" for p in pids))
    for i, pid in enumerate(pids):
        d = root / "long" / "Healthy" / pid
        d.mkdir(parents=True)
        for k in range(2):
            wavio.write_pcm16(str(d / f"r{k}.wav"), np.zeros((1, 400 * (i + 1)), dtype=np.float32), 16000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_plan_worker, args=(r, 2, port, str(root), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # rank 0 planned AFTER writing 11's result in this test, so 11 is skipped for everyone -- the point is that both
    # ranks hold the SAME plan: disjoint shards whose union is exactly rank 0's todo list
    assert got["before"] == sorted(pids)
    assert sorted(got[0] + got[1]) == ["22", "33", "44"] and not set(got[0]) & set(got[1])


def test_stats_aggregate_matches_the_reference_function(golden_dir):
    """stats.aggregate_stats / stats.finish against tests/golden/stats_golden.json, produced by EXECUTING the reference's
    own aggregate_stats (utils/compute_ast_normalization_stats.py:98-113) and its accumulation loop (:62-95)."""
    from zenker_audio_detection_b200 import stats

    g = json.load(open(os.path.join(golden_dir, "stats_golden.json")))
    agg = stats.aggregate_stats(g["per_fold"])
    assert agg.keys() == g["aggregate"].keys() and agg["total_count"] == g["aggregate"]["total_count"]
    assert agg["mean"] == g["aggregate"]["mean"] and agg["std"] == g["aggregate"]["std"]  # same float64 arithmetic
    assert stats.aggregate_stats([]) == {"mean": 0.0, "std": 0.0, "total_count": 0}
    assert stats.finish(0.0, 0.0, 0) == {"mean": 0.0, "std": 0.0, "count": 0}
    f = stats.finish(10.0, 30.0, 4)
    assert f["mean"] == 2.5 and abs(f["std"] - ((30.0 / 4 - 6.25) * 4 / 3) ** 0.5) < 1e-15


REF_SRC = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference checkout only exists in the build container")
@pytest.mark.parametrize("script", ["test_long_audio_windows_2stage.py"])  # (the cached script shares load_stage_model; ~45 s each)
def test_launcher_runs_the_unmodified_reference_script_up_to_the_device(tmp_path, script):
    """`python -m zenker_audio_detection_b200.run <reference script>` in the one place the reference sources exist (no
    GPU here; tests/test_gpu_pipeline.py holds the GPU half, which needs both): the script is executed unmodified -- its
    own argparse takes its own flags -- and its `load_stage_model` (ref:86-98) reaches OUR classes through
    `from transformers import ...` and stops at their loud no-CPU-path error, not inside HF's CPU model."""
    import subprocess
    import sys

    from safetensors.torch import save_file
    from transformers import ASTConfig

    from oracle import thirdparty
    from zenker_audio_detection_b200 import synth, wavio

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(REF_SRC, script)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    root = tmp_path / "s"
    root.mkdir()
    ASTConfig(num_labels=2).save_pretrained(str(root))
    save_file({k: v.contiguous() for k, v in synth.random_state_dict(11).items()}, str(root / "model.safetensors"))
    thirdparty.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD).save_pretrained(str(root))
    files = []
    for i in range(2):
        p = tmp_path / f"f{i}.wav"
        wavio.write_pcm16(str(p), synth.recording(2.0, 16000, seed=60 + i)[None], 16000)
        files.append(str(p))
    r = subprocess.run([sys.executable, "-m", "zenker_audio_detection_b200.run", path, "--stage1-model-root", str(root),
                        "--stage2-model-root", str(root), "--file-a", files[0], "--file-b", files[1], "--output-json",
                        str(tmp_path / "out.json")], cwd=repo, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and not os.path.exists(tmp_path / "out.json")
    assert "ZkError" in r.stderr and "load_stage_model" in r.stderr, r.stderr[-2000:]


def test_shard_window_ranges_properties():
    """Property test (hypothesis): for any pool and any positive weights every window is dealt exactly once, in order,
    with at most one cut between neighbouring ranks, and a split recording's records merge back to the unsplit ones."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.lists(st.integers(0, 60), min_size=1, max_size=12), st.integers(1, 9), st.data())
    def check(counts, world, data):
        weights = data.draw(st.one_of(st.none(), st.lists(st.floats(0.2, 5.0), min_size=world, max_size=world)))
        shards = zdist.shard_window_ranges(counts, world, weights=weights)
        assert len(shards) == world
        flat = [c for s in shards for c in s]
        assert flat == sorted(flat)                                   # rank order == recording / window order
        covered = {}
        for i, a, b in flat:
            assert 0 <= a < b <= counts[i]
            assert covered.get(i, 0) == a                             # contiguous, no gap, no overlap
            covered[i] = b
        assert all(covered.get(i, 0) == c for i, c in enumerate(counts))
        per = [sum(b - a for _, a, b in s) for s in shards]
        if weights is None:
            assert max(per) - min(per) <= 1
        else:
            tot, wsum = sum(counts), sum(weights)
            assert all(abs(p - tot * w / wsum) <= 1.0 + 1e-6 for p, w in zip(per, weights))
        # records of the pieces == records of the whole
        rng = np.random.default_rng(len(flat))
        blocks, whole = [], []
        for i, c in enumerate(counts):
            s1 = rng.random((c, 2)).astype(np.float32)
            idx = np.where(s1[:, 1] > 0.6)[0].astype(np.int64)
            s2 = rng.random((len(idx), 2)).astype(np.float32)
            whole.append(zdist.pack_records(i, s1, idx, s2))
            for j, a, b in flat:
                if j == i:
                    sel = (idx >= a) & (idx < b)
                    blocks.append(zdist.pack_records(i, s1[a:b], idx[sel] - a, s2[sel], window_base=a))
        w = np.concatenate(whole) if whole else np.zeros((0, 6), np.int32)
        p = np.concatenate(blocks) if blocks else np.zeros((0, 6), np.int32)
        assert np.array_equal(w, p)

    check()
