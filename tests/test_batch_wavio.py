"""Ingest + orchestration around the hot path: the WAV reader that replaces torchaudio.load /
torchaudio.info (TorchCodec is absent in this image) and the in-process batch runner that replaces
run_batch_simple_2stage.py.  CPU-only checks here; the end-to-end run is in test_gpu_pipeline.py."""
import io
import json
import os
import struct
import wave

import numpy as np
import pytest

from zenker_audio_detection_b200 import batch, wavio


def _tone(n, ch, seed):
    r = np.random.default_rng(seed)
    return (r.standard_normal((ch, n)) * 0.2).clip(-1, 1).astype(np.float32)


def test_wav_pcm16_matches_the_stdlib_reader(tmp_path):
    x = _tone(4801, 2, 1)
    p = str(tmp_path / "a.wav")
    wavio.write_pcm16(p, x, 48000)
    data, sr = wavio.read(p)
    assert sr == 48000 and data.dtype == np.int16 and data.shape == (4801, 2)
    with wave.open(p, "rb") as w:
        assert (w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()) == (48000, 2, 2, 4801)
        ref = np.frombuffer(w.readframes(4801), dtype="<i2").reshape(4801, 2)
    assert np.array_equal(data, ref)
    assert np.abs(data.T / 32768.0 - x).max() <= 0.5 / 32768 + 1e-7
    wi = wavio.info(p)
    assert (wi.sample_rate, wi.num_frames, wi.num_channels, wi.bits_per_sample, wi.encoding) == (48000, 4801, 2, 16, "PCM_S")


@pytest.mark.parametrize("tag,bits,dtype,scale", [(1, 8, np.uint8, 128.0), (1, 24, None, 8388608.0), (1, 32, "<i4", 2147483648.0),
                                                  (3, 32, "<f4", 1.0), (3, 64, "<f8", 1.0)])
def test_wav_other_encodings_are_normalised_like_torchaudio(tmp_path, tag, bits, dtype, scale):
    x = _tone(1000, 1, bits)[0]
    if tag == 3:
        raw = x.astype(dtype).tobytes()
        want = x.astype(dtype).astype(np.float32)
    elif bits == 8:
        q = np.clip(np.round(x * 128.0) + 128, 0, 255).astype(np.uint8)
        raw, want = q.tobytes(), (q.astype(np.float32) - 128.0) / 128.0
    elif bits == 24:
        q = np.clip(np.round(x * scale), -scale, scale - 1).astype(np.int32)
        raw = b"".join(struct.pack("<i", int(v))[:3] for v in q)
        want = q.astype(np.float32) / np.float32(scale)
    else:
        q = np.clip(np.round(x.astype(np.float64) * scale), -scale, scale - 1).astype(np.int64).astype(dtype)
        raw, want = q.tobytes(), q.astype(np.float32) / np.float32(scale)
    p = tmp_path / "b.wav"
    with open(p, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + 14 + len(raw)) + b"WAVE")
        f.write(b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\0")  # an odd-sized chunk before fmt: word alignment
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, tag, 1, 16000, 16000 * bits // 8, bits // 8, bits))
        f.write(b"data" + struct.pack("<I", len(raw)) + raw)
    data, sr = wavio.read(str(p))
    assert sr == 16000 and data.dtype == np.float32 and data.shape == (1, 1000)
    assert np.array_equal(data[0], want)


def test_wav_read_delivers_into_the_callers_buffer(tmp_path):
    """read(path, alloc=...): the samples arrive in the array the caller hands out (the batch runner hands out page-locked
    memory, batch.read_pinned), same values as the plain read; an array of another shape or dtype is refused."""
    x = _tone(3001, 2, 7)
    p = str(tmp_path / "a.wav")
    wavio.write_pcm16(p, x, 48000)
    plain, _ = wavio.read(p)
    given = []

    def alloc(shape, dtype):
        backing = np.zeros(int(np.prod(shape)) * np.dtype(dtype).itemsize + 64, dtype=np.uint8)  # a larger foreign buffer
        given.append(backing[:int(np.prod(shape)) * np.dtype(dtype).itemsize].view(dtype).reshape(shape))
        return given[-1]

    data, sr = wavio.read(p, alloc=alloc)
    assert sr == 48000 and data is given[0] and np.array_equal(data, plain)
    f = tmp_path / "f.wav"
    raw = x[0].astype("<f4").tobytes()
    with open(f, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE")
        fh.write(b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, 16000, 64000, 4, 32))
        fh.write(b"data" + struct.pack("<I", len(raw)) + raw)
    data, sr = wavio.read(str(f), alloc=alloc)
    assert data is given[1] and data.shape == (1, 3001) and np.array_equal(data[0], x[0])
    with pytest.raises(wavio.WavError):
        wavio.read(p, alloc=lambda shape, dtype: np.empty(shape, dtype=np.float32))
    with pytest.raises(wavio.WavError):
        wavio.read(str(f), alloc=lambda shape, dtype: np.empty((shape[1], shape[0]), dtype=dtype))


def test_wav_rejects_garbage(tmp_path):
    p = tmp_path / "c.wav"
    p.write_bytes(b"RIFF\x00\x00\x00\x00WAVEfmt ")
    with pytest.raises(wavio.WavError):
        wavio.read(str(p))
    p.write_bytes(b"not a wav file at all")
    with pytest.raises(wavio.WavError):
        wavio.info(str(p))


def _tree(tmp_path):
    # ids are matched as SUBSTRINGS of the directory path (ref:123), so none of them may occur in pytest's tmp path (pytest-NN)
    root = tmp_path / "long"
    lens = {"Healthy/224": [1600, 800, 2400], "Zenker/301": [1200, 900], "Zenker/9017": [500]}
    for rel, ns in lens.items():
        d = root / rel
        d.mkdir(parents=True)
        for i, n in enumerate(ns):
            wavio.write_pcm16(str(d / f"rec{i}.wav"), _tone(n, 1, n), 16000)
        (d / "notes.txt").write_text("x")
    ids = tmp_path / "ids"
    ids.mkdir()
    (ids / "test_ids_fold3.txt").write_text("Healthy/224\n\nZenker/301\nZenker/9017\n")
    return root, ids


def test_discover_two_files_follows_the_reference(tmp_path):
    root, _ = _tree(tmp_path)
    two = batch.discover_two_files(str(root), "301", "*.wav")
    assert [os.path.basename(p) for p in two] == ["rec0.wav", "rec1.wav"]  # sorted (ref:128)
    longest = batch.discover_two_files(str(root), "224", "*.wav")  # > 2 files: the two with most frames (ref:129-137)
    assert [os.path.basename(p) for p in longest] == ["rec2.wav", "rec0.wav"]
    with pytest.raises(ValueError, match="Expected exactly 2 files for patient 9017, found 1"):
        batch.discover_two_files(str(root), "9017", "*.wav")


def test_ids_thresholds_and_plan(tmp_path, capsys):
    root, ids = _tree(tmp_path)
    assert batch.read_ids(str(ids / "test_ids_fold3.txt")) == ["224", "301", "9017"]
    cfg = {"folds": {"3": {"stage1": {"threshold": 0.61}, "stage2": {"threshold": 0.35}}},
           "thresholds": {"stage1": {"threshold": 0.9}}}
    assert batch.resolve_thresholds(cfg, 3) == (0.61, 0.35)       # per-fold block wins (ref batch:97-108)
    assert batch.resolve_thresholds({"thresholds": {"stage2": {"threshold": 0.4}}}, 3) == (None, 0.4)
    assert batch.resolve_thresholds(None, 3) == (None, None)
    out = tmp_path / "out"
    out.mkdir()
    (out / "301_2stage.json").write_text("{}")
    args = batch.build_arg_parser().parse_args(["--fold", "3", "--ids-root", str(ids), "--long-audio-root", str(root),
                                                "--output-dir", str(out), "--dry-run"])
    todo, _ = batch.global_plan(args)
    plans = [batch.shard_plan(todo, r, 2) for r in range(2)]
    assert sorted(pid for pl in plans for pid, _ in pl) == ["224"]  # 301 exists -> skipped, 17 has one file -> error
    assert batch.run(args, 0, 1) == 0
    printed = capsys.readouterr().out
    assert "[SKIP] 301" in printed and "[ERROR] patient 9017" in printed and "[RUN] rank 0: 224" in printed
    args.force = True
    assert sorted(pid for pid, _ in batch.plan_patients(args, 0, 1)[0]) == ["224", "301"]


def _plan_worker(rank, world, port, argv, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    args = batch.build_arg_parser().parse_args(argv)
    # rank 1's own view of the work differs from rank 0's (as if it had looked at the output directory at another
    # time): planning on its own it would also take 224, whose result file exists
    args.force = rank == 1
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine, _ = batch.plan_patients(args, rank, world)
    q.put((rank, [pid for pid, _ in mine]))
    dist.destroy_process_group()


def test_plan_is_made_on_rank_0_and_broadcast(tmp_path):
    """every rank must shard the SAME plan (gloo, world size 2): the union of the shares is rank 0's plan
    and no patient is dropped or duplicated, whatever the other ranks see in the output directory."""
    import torch.multiprocessing as mp

    root, ids = _tree(tmp_path)
    out = tmp_path / "out"
    out.mkdir()
    (out / "224_2stage.json").write_text("{}")
    argv = ["--fold", "3", "--ids-root", str(ids), "--long-audio-root", str(root), "--output-dir", str(out), "--dry-run"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_plan_worker, args=(r, 2, port, argv, q)) for r in range(2)]
    for p in procs:
        p.start()
    shares = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = sorted(shares[0] + shares[1])
    assert len(got) == len(set(got))                       # nobody is processed twice
    assert got == ["301"], shares                          # rank 0's plan (224 skipped), not rank 1's own view


def test_batch_flags_are_the_launchers():
    """ref batch:145-211 (+ the per-patient thresholds / batch size of refc:303-376 it forwards)."""
    flags = {a.option_strings[0] for a in batch.build_arg_parser()._actions if a.option_strings}
    for f in ("--fold", "--ids-root", "--long-audio-root", "--pattern", "--window-sec", "--hop-sec", "--plot",
              "--output-dir", "--threshold-config", "--stage1-model-root", "--stage2-model-root",
              "--stage1-forward-min-prob", "--stage2-argmax", "--extra", "--force", "--dry-run"):
        assert f in flags, f


def test_patched_torchaudio_load_and_info(tmp_path):
    """compat.patch_torchaudio: ``load_audio`` (ref:53-59) and ``discover_two_files`` (ref:132) keep working without
    TorchCodec: float32 (channels, frames) in [-1, 1), sample rate, ``info(...).num_frames``."""
    import torch
    import torchaudio

    from zenker_audio_detection_b200 import compat

    x = _tone(3000, 2, 5)
    p = str(tmp_path / "d.wav")
    wavio.write_pcm16(p, x, 44100)
    compat.patch_torchaudio()
    try:
        wav, sr = torchaudio.load(p)
        assert sr == 44100 and wav.dtype == torch.float32 and tuple(wav.shape) == (2, 3000)
        assert float((wav - torch.from_numpy(x)).abs().max()) <= 0.5 / 32768 + 1e-7
        assert torchaudio.info(p).num_frames == 3000
        mono = wav.mean(dim=0, keepdim=True)  # ref:55-56
        assert tuple(mono.shape) == (1, 3000)
    finally:
        compat.unpatch_torchaudio()


def test_wav_prefetcher_keeps_order_and_reraises():
    from zenker_audio_detection_b200.batch import WavPrefetcher

    calls = []

    def reader(path):
        calls.append(path)
        if path == "bad":
            raise ValueError("not a RIFF file")
        return (path.upper(), 16000)

    pf = WavPrefetcher(["a", "bad", "c", "d"], reader=reader)
    assert pf.get("a") == ("A", 16000)
    with pytest.raises(ValueError, match="RIFF"):
        pf.get("bad")            # the failure of one file surfaces at that file ...
    assert pf.get("c") == ("C", 16000)  # ... and the next ones still come
    assert pf.get("zzz") == ("ZZZ", 16000)  # out of order: read synchronously, prefetch of "d" untouched
    assert pf.get("d") == ("D", 16000)
    assert pf.get("a") == ("A", 16000)  # exhausted: synchronous
    pf.close()
    assert calls.count("d") == 1 and calls[:3] == ["a", "bad", "c"]


def test_wav_prefetcher_extend_keeps_prefetching():
    from zenker_audio_detection_b200.batch import WavPrefetcher

    seen = []

    def reader(path):
        seen.append(path)
        return path.upper(), 48000

    pf = WavPrefetcher([], reader=reader)       # the dynamic schedule starts empty and learns its files claim by claim
    pf.extend(["a", "b"])
    assert pf.get("a") == ("A", 48000)
    pf.extend(["c"])
    assert pf.get("b") == ("B", 48000) and pf.get("c") == ("C", 48000)
    pf.extend(["d"])
    assert pf.get("d") == ("D", 48000)
    pf.close()
    assert seen == ["a", "b", "c", "d"]          # each file decoded once, in order


def _claim_worker(rank, world, port, n, q):
    import random
    import time

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    got = []
    rnd = random.Random(rank)
    for i in batch.claim_indices(n, rank, world, "dynamic"):
        got.append(i)
        time.sleep((0.03 if rank else 0.005) * (1 + rnd.random()))  # rank 1 is the slow GPU
    dist.barrier()
    q.put((rank, got))
    dist.destroy_process_group()


def test_dynamic_schedule_claims_every_patient_exactly_once():
    """`batch --schedule dynamic`: ranks take the next unclaimed patient from a counter in the process group's store, so
    the queue is covered exactly once whatever the ranks' speeds, and the slow rank ends up with fewer patients."""
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 23
    procs = [ctx.Process(target=_claim_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(outs[0] + outs[1]) == list(range(n))
    assert outs[0] == sorted(outs[0]) and outs[1] == sorted(outs[1])   # each rank walks the queue front to back
    assert len(outs[0]) > len(outs[1]) > 0
    assert list(batch.claim_indices(5, 0, 1)) == [0, 1, 2, 3, 4]
    assert [list(batch.claim_indices(5, r, 2, "static", sizes=[9, 1, 8, 2, 7])) for r in range(2)] == [[0, 1, 3], [2, 4]]


def test_all_folds_launcher_follows_the_shell_script(tmp_path, capsys, monkeypatch):
    """python -m zenker_audio_detection_b200.folds = src/run_all_folds_simple_batch.sh in one process: LONG_AUDIO_ROOT from
    the environment / the project's .env / the fallback (:22-42), flags in any order with unknown ones warned about
    (:52-82), per-fold model roots, output directory and threshold config (:93-121), folds 1..5 in order."""
    from zenker_audio_detection_b200 import folds

    root = tmp_path / "proj"
    long_root, _ = _tree(tmp_path)
    ids = root / "data_ast_stage2"
    ids.mkdir(parents=True)
    for f in folds.FOLDS:
        (ids / f"test_ids_fold{f}.txt").write_text("Healthy/224\n" if f % 2 else "Zenker/301\n")
    (root / ".env").write_text(f"# paths\nexport OTHER=1\nLONG_AUDIO_ROOT=\"{long_root}\"\n")
    (root / "runs").mkdir()
    (root / "runs" / "optimal_thresholds_per_fold_both_stages.json").write_text(
        json.dumps({"folds": {"1": {"stage1": {"threshold": 0.7}, "stage2": {"threshold": 0.2}}}}))
    monkeypatch.delenv("LONG_AUDIO_ROOT", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    assert folds.read_env_file(str(root / ".env")) == {"OTHER": "1", "LONG_AUDIO_ROOT": str(long_root)}

    cwd = os.getcwd()
    assert folds.main(["--stage2-argmax", "runs", "--bogus", "--dry-run", "--stage1-forward-min-prob", "0.8",
                       "--project-root", str(root)]) == 0
    assert os.getcwd() == cwd
    cap = capsys.readouterr()
    out = cap.out
    assert "Warning: Unknown option --bogus" in cap.err
    assert f"Long audio directory: {long_root}" in out and "Using models from: runs" in out
    assert f"Found threshold config: {root / 'runs' / 'optimal_thresholds_per_fold_both_stages.json'}" in out
    assert (root / "runs" / "results" / "patient_inference").is_dir()
    marks = [out.index(f"================ Fold {f} ================") for f in folds.FOLDS]
    assert marks == sorted(marks) and out.rstrip().endswith("All folds completed.")
    assert out.count("[RUN] rank 0: 224") == 3 and out.count("[RUN] rank 0: 301") == 2   # folds 1,3,5 / 2,4

    opts = folds.parse(["exp/v1", "--no-threshold-config"])
    argv = folds.fold_argv(opts, 4, str(root), "/data/long")
    a = batch.build_arg_parser().parse_args(argv)
    assert a.fold == 4 and a.long_audio_root == "/data/long" and a.pattern == "*.wav" and a.plot and not a.dry_run
    assert a.stage1_model_root == str(root / "exp/v1" / "ast_classifier_stage1" / "fold4" / "best")
    assert a.stage2_model_root == str(root / "exp/v1" / "ast_classifier_stage2" / "fold4" / "best")
    assert a.output_dir == str(root / "exp/v1" / "results" / "patient_inference")
    assert a.threshold_config is None and a.stage1_forward_min_prob is None and not a.stage2_argmax
    b = batch.build_arg_parser().parse_args(folds.fold_argv(folds.parse(["--stage1-forward-min-prob", "0.8", "--stage2-argmax"]),
                                                            1, str(root), "/data/long"))
    assert b.threshold_config == str(root / "runs" / "optimal_thresholds_per_fold_both_stages.json")
    assert b.stage1_forward_min_prob == 0.8 and b.stage2_argmax

    # precedence of LONG_AUDIO_ROOT: environment, then .env, then the script's fallback
    monkeypatch.setenv("LONG_AUDIO_ROOT", "/from/env")
    folds.main(["--dry-run", "--project-root", str(root)])
    assert "Long audio directory: /from/env" in capsys.readouterr().out
    monkeypatch.delenv("LONG_AUDIO_ROOT")
    (root / ".env").unlink()
    folds.main(["--dry-run", "--project-root", str(root)])
    out = capsys.readouterr().out
    assert "Warning: LONG_AUDIO_ROOT not set" in out and f"Using fallback: {folds.FALLBACK_LONG_AUDIO_ROOT}" in out


REF_EXTRACT = "/root/reference/utils/extract_thresholds_per_fold.py"


@pytest.mark.skipif(not os.path.exists(REF_EXTRACT), reason="the reference checkout only exists in the build container")
def test_threshold_config_written_by_the_reference_tool_is_read_per_fold(tmp_path):
    """Upstream of the path: `utils/extract_thresholds_per_fold.py`, unmodified, writes
    optimal_thresholds_per_fold_both_stages.json from validation ROC/PR metrics (:93-122); `batch.resolve_thresholds`
    must pick each fold's own pair out of exactly that file (run_batch:97-118), and `folds.fold_argv` must find it."""
    import subprocess
    import sys

    from zenker_audio_detection_b200 import folds

    def metrics(base):
        return {"fold_reports": [{"fold": f, "best_f1_threshold": base + 0.01 * f, "best_f1": 0.9, "best_f1_precision": 0.8,
                                  "best_f1_recall": 0.95} for f in (1, 2, 3, 4, 5)],
                "aggregate": {"best_f1_threshold": base, "best_f1": 0.88, "best_f1_precision": 0.8, "best_f1_recall": 0.9}}

    (tmp_path / "m1.json").write_text(json.dumps(metrics(0.60)))
    (tmp_path / "m2.json").write_text(json.dumps(metrics(0.30)))
    runs = tmp_path / "runs"
    runs.mkdir()
    cfg_path = runs / "optimal_thresholds_per_fold_both_stages.json"
    r = subprocess.run([sys.executable, REF_EXTRACT, "--stage1-metrics", str(tmp_path / "m1.json"), "--stage2-metrics",
                        str(tmp_path / "m2.json"), "--output-config", str(cfg_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and cfg_path.exists(), r.stderr[-1500:]
    cfg = json.load(open(cfg_path))
    for f in (1, 2, 3, 4, 5):
        t1, t2 = batch.resolve_thresholds(cfg, f)
        assert t1 == pytest.approx(0.60 + 0.01 * f) and t2 == pytest.approx(0.30 + 0.01 * f)
    assert batch.resolve_thresholds(cfg, 9) == (None, None)      # a fold the file does not know keeps the CLI defaults
    argv = folds.fold_argv(folds.parse(["runs"]), 2, str(tmp_path), "/data/long")
    assert batch.build_arg_parser().parse_args(argv).threshold_config == str(cfg_path)
