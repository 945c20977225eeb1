"""In-process replacement of ``src/run_batch_simple_2stage.py`` + ``src/test_long_audio_windows_2stage_cache.py``:

    python -m zenker_audio_detection_b200.batch --fold 1 --long-audio-root data/long --output-dir outputs
    torchrun --nproc-per-node 8 -m zenker_audio_detection_b200.batch ...        # patients sharded over the GPUs

The reference launcher spawns one Python process per patient (ref batch:282-284), each of which reloads both models
from disk; here both models are loaded once per rank, every rank takes its share of the fold's patients (longest
first, ``dist.shard_recordings``), and each patient's ``<pid>_2stage.json`` is written in the schema of the script the
launcher runs (refc:570-601), so ``utils/aggregate_2stage_results.py`` consumes the output directory unchanged.  The
flags are the launcher's own (ref batch:145-211); ``--plot`` / ``--extra`` are accepted and ignored.  Patients are
independent, so there is no collective at all on this path.
"""
from __future__ import annotations

import argparse
import fnmatch
import json
import os
import sys
from typing import Dict, List, Optional, Sequence, Tuple

from . import wavio


def read_ids(ids_path: str) -> List[str]:
    """ref batch:48-57: one ``Class/ID`` per line, the leaf is the patient id."""
    patients = []
    with open(ids_path, "r") as f:
        for line in f:
            line = line.strip()
            if line:
                patients.append(line.split("/")[-1])
    return patients


def resolve_thresholds(config: Optional[dict], fold: int) -> Tuple[Optional[float], Optional[float]]:
    """ref batch:97-118: per-fold thresholds (``folds[str(fold)]``) win over the single ``thresholds`` block."""
    if not config:
        return None, None
    block = config.get("folds", {}).get(str(fold)) if config.get("folds") else None
    if block is None:
        block = config.get("thresholds", {})
    s1 = block.get("stage1", {}).get("threshold") if "stage1" in block else None
    s2 = block.get("stage2", {}).get("threshold") if "stage2" in block else None
    return s1, s2


def discover_two_files(root: str, patient_id: str, pattern: str) -> List[str]:
    """ref:119-142: every file matching ``pattern`` below a directory whose path contains the patient id, sorted; with
    more than two, the two with the most frames (header only); anything but two is an error."""
    base = os.path.abspath(root)
    matches = []
    for dirpath, _, filenames in os.walk(base):
        if patient_id not in dirpath:
            continue
        for fn in filenames:
            if fnmatch.fnmatch(fn, pattern):
                matches.append(os.path.join(dirpath, fn))
    matches = sorted(matches)
    if len(matches) > 2:
        lengths = []
        for p in matches:
            try:
                lengths.append((p, wavio.info(p).num_frames))
            except Exception:  # noqa: BLE001 - like the reference: unreadable files sort last
                lengths.append((p, 0))
        matches = [p for p, _ in sorted(lengths, key=lambda x: x[1], reverse=True)[:2]]
    if len(matches) != 2:
        raise ValueError(f"Expected exactly 2 files for patient {patient_id}, found {len(matches)}: {matches}")
    return matches


def default_model_root(stage: int, fold: int) -> str:
    """refc:397-405 resolves ``<project_root>/runs/ast_classifier_stage<k>/fold<F>/best``; the project root here is the
    working directory (this package does not live inside the reference checkout)."""
    return os.path.join(os.getcwd(), "runs", f"ast_classifier_stage{stage}", f"fold{fold}", "best")


def read_pinned(path: str):
    """``wavio.read`` into page-locked memory -> ``(torch tensor, sample_rate)``: the file's bytes land where the H2D
    copy starts from, instead of in pageable memory that ``resample_to_device`` would first have to copy into a pinned
    staging buffer on the main thread (115 MB for a 10-minute stereo recording: 4.9 -> 2.4 ms from "samples in host
    memory" to "16 kHz mono on the GPU", ``scripts/ingest_time.py``).  The tensor comes from torch's caching host
    allocator, which also keeps the buffer alive until the copy has run."""
    import numpy as np
    import torch

    held = []

    def alloc(shape, dtype):
        t = torch.empty(tuple(shape), dtype=torch.int16 if np.dtype(dtype) == np.dtype("<i2") else torch.float32,
                        pin_memory=True)
        held.append(t)
        return t.numpy()

    _, sr = wavio.read(path, alloc=alloc)
    return held[0], sr


class WavPrefetcher:
    """Decodes the next recording on a host thread while the GPU works on the current one (a 10-minute stereo PCM16
    file is 115 MB: ~60 ms from NVMe, ~12 % of the 0.46 s the cascade takes).  ``get(path)`` returns what
    ``wavio.read(path)`` returns and re-raises its exception; paths must be requested in the order given."""

    def __init__(self, paths: Sequence[str], reader=wavio.read):
        from concurrent.futures import ThreadPoolExecutor

        self._paths, self._reader, self._next = list(paths), reader, 0
        self._pool = ThreadPoolExecutor(max_workers=1)
        self._pending = None
        self._submit()

    def _submit(self) -> None:
        self._pending = None
        if self._next < len(self._paths):
            self._pending = (self._paths[self._next], self._pool.submit(self._reader, self._paths[self._next]))
            self._next += 1

    def extend(self, paths: Sequence[str]) -> None:
        """Append more files to the plan (the dynamic schedule learns its patients one claim at a time)."""
        self._paths.extend(paths)
        if self._pending is None:
            self._submit()

    def get(self, path: str):
        if self._pending is not None and self._pending[0] == path:
            fut = self._pending[1]
            try:
                return fut.result()
            finally:
                self._submit()
        # Out of order: a patient failed on its first file and the caller moved on, so the pending decode is for a file
        # nobody will ask for.  Drop it (and the buffer it holds), read this one synchronously and re-arm the
        # prefetch on whatever follows `path` in the plan, so that the remaining reads overlap again.
        if self._pending is not None:
            self._pending[1].cancel()
            self._pending = None
        try:
            return self._reader(path)
        finally:
            if path in self._paths:
                self._next = self._paths.index(path) + 1
            self._submit()

    def close(self) -> None:
        self._pool.shutdown(wait=False, cancel_futures=True)


def build_arg_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="Two-stage window inference over every patient of a fold (B200, in process).")
    ap.add_argument("--fold", type=int, required=True)
    ap.add_argument("--ids-root", default=None, help="Directory containing test_ids_fold<fold>.txt (default: ./data_ast_stage2)")
    ap.add_argument("--long-audio-root", required=True)
    ap.add_argument("--pattern", default="*.wav")
    ap.add_argument("--window-sec", type=float, default=1.0)
    ap.add_argument("--hop-sec", type=float, default=0.5)
    ap.add_argument("--batch-size", type=int, default=128)
    ap.add_argument("--plot", action="store_true", help="accepted for compatibility; ignored")
    ap.add_argument("--output-dir")
    ap.add_argument("--threshold-config")
    ap.add_argument("--stage1-model-root")
    ap.add_argument("--stage2-model-root")
    ap.add_argument("--stage1-threshold", type=float, default=0.5)
    ap.add_argument("--stage2-threshold", type=float, default=0.5)
    ap.add_argument("--stage1-forward-min-prob", type=float)
    ap.add_argument("--stage2-argmax", action="store_true")
    ap.add_argument("--feature-cache-dir", default=None,
                    help="reuse / fill the reference's per-recording feature bundles (refc:84-192); default: no cache, "
                         "the fused path recomputes the fbank")
    ap.add_argument("--disable-cache", action="store_true")
    ap.add_argument("--refresh-cache", action="store_true")
    ap.add_argument("--extra", help="accepted for compatibility; ignored")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--schedule", default="dynamic", choices=["dynamic", "static"],
                    help="under torchrun: ranks claim patients from a shared longest-first queue (default), or take a fixed "
                         "longest-first partition")
    return ap


def global_plan(args) -> Tuple[List[Tuple[str, List[str], int]], List[str]]:
    """-> (``[(patient, [file_a, file_b], bytes)]`` still to do, log lines).  Depends on the file system (which result
    files already exist), so under torchrun it is computed ONCE, on rank 0, and broadcast (``plan_patients``)."""
    ids_root = args.ids_root or os.path.join(os.getcwd(), "data_ast_stage2")
    ids_path = os.path.join(ids_root, f"test_ids_fold{args.fold}.txt")
    if not os.path.exists(ids_path):
        raise FileNotFoundError(f"IDs file not found: {ids_path}")
    out_dir = args.output_dir or "outputs"
    log, todo = [], []
    for pid in read_ids(ids_path):
        expected = os.path.join(out_dir, f"{pid}_2stage.json")
        if os.path.exists(expected) and not args.force:
            log.append(f"[SKIP] {pid} (exists: {expected})")  # ref batch:274-277
            continue
        try:
            files = discover_two_files(args.long_audio_root, pid, args.pattern)
        except ValueError as e:
            log.append(f"[ERROR] patient {pid}: {e}")  # ref batch:286-289: a failing patient does not stop the batch
            continue
        todo.append((pid, files, sum(os.path.getsize(p) for p in files)))
    return todo, log


def plan_queue(args, rank: int, world: int) -> Tuple[List[Tuple[str, List[str]]], List[str]]:
    """-> (the WHOLE work queue ``[(patient, [file_a, file_b])]``, longest first, identical on every rank; log lines):
    what the dynamic schedule claims from (``claim_indices``).  Made on rank 0 and broadcast like ``plan_patients``."""
    todo, log = _global_plan_everywhere(args, rank, world)
    return [(pid, files) for pid, files, _ in sorted(todo, key=lambda t: (-t[2], t[0]))], log


def plan_patients(args, rank: int, world: int) -> Tuple[List[Tuple[str, List[str]]], List[str]]:
    """-> (this rank's ``[(patient, [file_a, file_b])]``, log lines): the static longest-first partition.

    With more than one rank the plan is made by rank 0 alone and broadcast: a rank that started late would otherwise see
    result files an early rank has already written, skip those patients, shard a DIFFERENT list and drop or duplicate
    work.  The exchange goes over a gloo group (host objects); ranks then shard the identical list longest-first."""
    todo, log = _global_plan_everywhere(args, rank, world)
    return shard_plan(todo, rank, world), log


def _global_plan_everywhere(args, rank: int, world: int):
    if world > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group("gloo", rank=rank, world_size=world)
        box = [global_plan(args) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        todo, log = box[0]
    else:
        todo, log = global_plan(args)
    return todo, log


def claim_indices(n: int, rank: int, world: int, schedule: str = "dynamic", sizes: Optional[Sequence[int]] = None,
                  key: str = "zk_batch_next"):
    """Yields the positions of a queue of ``n`` work items that THIS rank is to process.

    ``dynamic`` (more than one rank): self-scheduling -- every rank takes the next unclaimed position from a shared
    counter in the process group's key-value store (``store.add`` is atomic), so a GPU that runs at a lower power-capped
    clock, or that drew the long recordings, simply claims fewer patients; with the queue ordered longest-first the
    ranks finish within one patient of each other.  No collective, no rank waits for another.  ``static``: the fixed
    longest-first partition of ``dist.shard_recordings`` over ``sizes`` (what a dry run prints)."""
    if world <= 1:
        yield from range(n)
        return
    if schedule == "static":
        from .dist import shard_recordings

        yield from shard_recordings(list(sizes) if sizes is not None else [1] * n, world)[rank]
        return
    import torch.distributed as dist

    store = dist.distributed_c10d._get_default_store()
    while True:
        i = int(store.add(key, 1)) - 1
        if i >= n:
            return
        yield i


def shard_plan(todo, rank: int, world: int) -> List[Tuple[str, List[str]]]:
    """Rank ``rank``'s share of one global plan ``[(patient, files, bytes)]``: patients dealt longest-first
    (``dist.shard_recordings`` on the byte counts), so every rank derives the same partition from the same list."""
    from .dist import shard_recordings

    mine = shard_recordings([b for _, _, b in todo], world)[rank] if todo else []
    return [(todo[i][0], todo[i][1]) for i in mine]


_RUN_SEQ = 0  # run() calls in this process; every rank makes the same calls, so it names the claim counter of a call


def run(args, rank: int = 0, world: int = 1) -> int:
    global _RUN_SEQ
    _RUN_SEQ += 1
    cfg = None
    if args.threshold_config and os.path.exists(args.threshold_config):
        with open(args.threshold_config, "r") as f:
            cfg = json.load(f)
    t1, t2 = resolve_thresholds(cfg, args.fold)
    thr1 = args.stage1_threshold if t1 is None else float(t1)
    thr2 = args.stage2_threshold if t2 is None else float(t2)
    s1_root = args.stage1_model_root or default_model_root(1, args.fold)
    s2_root = args.stage2_model_root or default_model_root(2, args.fold)
    out_dir = args.output_dir or "outputs"
    os.makedirs(out_dir, exist_ok=True)
    dynamic = getattr(args, "schedule", "dynamic") == "dynamic" and world > 1
    mine, log = plan_queue(args, rank, world) if dynamic else plan_patients(args, rank, world)
    if rank == 0:
        for line in log:
            print(line)
    if dynamic:
        queue = mine
        if rank == 0:
            for pid, files in queue:
                print(f"[QUEUE] {pid}  A: {files[0]}  B: {files[1]}")
    else:
        queue = mine
        for pid, files in mine:
            print(f"[RUN] rank {rank}: {pid}  A: {files[0]}  B: {files[1]}")
    if args.dry_run or not queue:
        return 0

    import torch
    from transformers import ASTConfig

    from . import cache as zcache, results
    from .fx import ZenkerASTFeatureExtractor
    from .model import ZenkerASTForAudioClassification
    from .pipeline import TwoStagePipeline

    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)

    def load(root: str, labels: Sequence[str]):  # ref:86-98
        fx = ZenkerASTFeatureExtractor.from_pretrained(root)
        cfg_ = ASTConfig.from_pretrained(root)
        cfg_.label2id = {l: i for i, l in enumerate(labels)}
        cfg_.id2label = {i: l for i, l in enumerate(labels)}
        return fx, ZenkerASTForAudioClassification.from_pretrained(root, config=cfg_).to(device).eval()

    fx1, m1 = load(s1_root, ["Idle", "Swallow"])
    fx2, m2 = load(s2_root, ["Healthy", "Zenker"])
    pipe = TwoStagePipeline(m1, fx1, m2, fx2, batch_size=args.batch_size, window_sec=args.window_sec, hop_sec=args.hop_sec,
                            stage1_threshold=thr1, stage2_threshold=thr2,
                            stage1_forward_min_prob=args.stage1_forward_min_prob, stage2_argmax=args.stage2_argmax,
                            device=device)
    cache_dir = None  # refc:428-430
    if not args.disable_cache and args.feature_cache_dir:
        cache_dir = os.path.abspath(args.feature_cache_dir)
    failures = 0
    # This rank's patients in the order it takes them: its static share, or claims from the shared queue.  One patient is
    # claimed AHEAD of the one being processed so that its first file decodes on the host while the GPU works.
    claims = (claim_indices(len(queue), rank, world, "dynamic", key=f"zk_batch_next/{_RUN_SEQ}/fold{args.fold}") if dynamic
              else iter(range(len(queue))))
    wavs = WavPrefetcher([], reader=read_pinned)
    if dynamic:  # start claiming together: a rank whose models loaded first would otherwise drain a short queue alone
        import torch.distributed as dist

        dist.barrier()

    def take():
        i = next(claims, None)
        if i is None:
            return None
        wavs.extend(queue[i][1])
        return queue[i]

    ahead = take()
    while ahead is not None:
        (pid, files), ahead = ahead, take()
        if dynamic:
            print(f"[RUN] rank {rank}: {pid}  A: {files[0]}  B: {files[1]}")
        try:
            summaries = []
            for path in files:
                samples, sr = wavs.get(path)
                if cache_dir:  # refc:433-507 on the reference's feature bundles
                    audio = pipe.resample_to_device(samples, sr)
                    windows = zcache.window_audio(audio, args.window_sec, args.hop_sec)
                    summaries.append(zcache.run_recording_cached(pipe, path, windows, cache_dir, False,
                                                                 args.refresh_cache).summary)
                else:
                    summaries.append(pipe.run_waveform(samples, sr).summary)
            doc = results.build_document(s1_root, s2_root, args.window_sec, args.hop_sec, args.batch_size, thr1, files,
                                         summaries, variant="cached", stage2_threshold=thr2,
                                         stage1_forward_min_prob=args.stage1_forward_min_prob,
                                         stage2_argmax=args.stage2_argmax, feature_cache_dir=cache_dir,
                                         disable_cache=cache_dir is None)
            results.write_json(doc, os.path.join(out_dir, f"{pid}_2stage.json"))
            print(f"[DONE] {pid} OK")
        except Exception as e:  # noqa: BLE001 - ref batch:286-289: log and continue with the next patient
            failures += 1
            print(f"[ERROR] patient {pid}: {type(e).__name__}: {e}")
    wavs.close()
    if dynamic:  # rank 0 hosts the store the others claim from: nobody leaves before the queue is empty everywhere
        import torch.distributed as dist

        dist.barrier()
    print("Batch complete." if world == 1 else f"Batch complete (rank {rank}).")
    return failures


def main(argv: Optional[Sequence[str]] = None) -> None:
    args = build_arg_parser().parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    failures = run(args, rank, world)
    if failures:
        sys.exit(1)


if __name__ == "__main__":
    main()
