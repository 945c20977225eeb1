"""Dataset normalisation statistics on the GPU: drop-in for ``utils/compute_ast_normalization_stats.py``.

    python -m zenker_audio_detection_b200.stats --stage stage1            # same flags, same output files

For every fold the reference runs the HF extractor with ``do_normalize = False`` over the training snippets and keeps a
float64 running sum / sum of squares of every element of the zero-padded ``(B, 1024, 128)`` features
(utils/compute_ast_normalization_stats.py:55-95).  Here the sums are an EPILOGUE of the feature kernel
(``zk_fx_stats_f32``): the log-mel values are added up in fp64 while they are computed and never written to memory.
Decoding is RIFF/WAVE through ``wavio`` (the reference uses soundfile/librosa, absent from this image); a file that is
not at 16 kHz goes through the GPU resampler (torchaudio's sinc kernel, not librosa's), so only 16 kHz material is
guaranteed to reproduce the reference's numbers to rounding.

Outputs in ``--output-dir``: ``stats_per_fold.json``, ``stats_aggregate.json``, ``stats_all.npz`` (same keys).
"""
from __future__ import annotations

import argparse
import json
import os
from collections import defaultdict
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

NUM_FOLDS = 5
SAMPLING_RATE = 16000


def finish(total_sum: float, total_sq: float, count: int) -> Dict[str, float]:
    """ref stats:82-95: mean, unbiased std from float64 sums."""
    if count == 0:
        return {"mean": 0.0, "std": 0.0, "count": 0}
    mean = total_sum / count
    var_pop = max(total_sq / count - mean * mean, 0.0)
    var = var_pop * (count / (count - 1)) if count > 1 else 0.0
    return {"mean": float(mean), "std": float(var ** 0.5), "count": int(count)}


def aggregate_stats(per_fold: Sequence[Dict[str, Any]]) -> Dict[str, Any]:
    """ref stats:98-113: count-weighted mean; pooled variance sum((n_k-1) s_k^2 + n_k (mu_k - mu)^2) / (N - 1)."""
    total = sum(d["count"] for d in per_fold)
    if total == 0:
        return {"mean": 0.0, "std": 0.0, "total_count": 0}
    wmean = sum(d["mean"] * d["count"] for d in per_fold) / total
    num = 0.0
    for d in per_fold:
        n = d["count"]
        if n < 2:
            continue
        num += (n - 1) * (d["std"] ** 2) + n * (d["mean"] - wmean) ** 2
    var = num / (total - 1) if total > 1 else 0.0
    return {"mean": float(wmean), "std": float(var ** 0.5), "total_count": int(total)}


def compute_fold_stats(data_dir: str, fold: int, batch_size: int, max_length: int = 1024, device=None) -> Dict[str, Any]:
    """ref stats:55-95 for one fold; returns ``{"fold", "mean", "std", "count"}``."""
    import torch

    from . import ops, wavio

    path = os.path.join(data_dir, f"train_x_fold{fold}.npy")
    if not os.path.exists(path):
        raise FileNotFoundError(f"Missing fold {fold} train data. Expected {path}")
    files: List[str] = np.load(path, allow_pickle=True).tolist()
    if len(files) == 0:
        return {"fold": fold, "mean": 0.0, "std": 0.0, "count": 0}
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(dev):
        plan = ops.FbankPlan("hanning", 128)
        stats = ops.FeatureStats(dev)
        for start in range(0, len(files), batch_size):
            by_len = defaultdict(list)
            for p in files[start:start + batch_size]:
                samples, sr = wavio.read(p)  # (channels, n) float32 or (n, channels) int16
                t = torch.from_numpy(np.ascontiguousarray(samples)).to(dev)
                audio = ops.resample(t, int(sr), SAMPLING_RATE)  # channel mean (+ resample when sr != 16 kHz)
                by_len[int(audio.numel())].append(audio)
            for n, group in by_len.items():  # one launch per distinct length (the kernel takes equal-length rows)
                if n < 400:  # shorter than one 25 ms frame: all-zero features, only the count grows
                    stats.count += len(group) * max_length * 128
                    continue
                stats.update_from_waveforms(plan, torch.stack(group), max_length)
        out = stats.result()
    return {"fold": fold, "mean": out["mean"], "std": out["std"], "count": out["count"]}


def main(argv: Optional[Sequence[str]] = None) -> None:
    ap = argparse.ArgumentParser(description="Compute AST normalization stats across CV folds (B200, stats as an fbank epilogue)")
    ap.add_argument("--data-dir", default="data_ast_cv")
    ap.add_argument("--folds", type=int, default=NUM_FOLDS)
    ap.add_argument("--output-dir", default="data_ast_cv")
    ap.add_argument("--batch-size", type=int, default=16)
    ap.add_argument("--stage", choices=["stage1", "stage2"])
    args = ap.parse_args(argv)
    if args.stage:  # ref stats:124-133 (the project root here is the working directory)
        mapped = os.path.join(os.getcwd(), f"data_ast_{args.stage}")
        print(f"[Info] Using stage alias '{args.stage}' -> data/output dir '{mapped}'")
        args.data_dir = args.output_dir = mapped
    per_fold = []
    for fold in range(1, args.folds + 1):
        print(f"Computing stats for fold {fold} (batch_size={args.batch_size}) ...")
        st = compute_fold_stats(args.data_dir, fold, args.batch_size)
        print(f"  Fold {fold}: mean={st['mean']:.6f} std={st['std']:.6f} (count={st['count']})")
        per_fold.append(st)
    agg = aggregate_stats(per_fold)
    print("\nWeighted aggregate (training folds, with repetition):")
    print(f"  mean={agg['mean']:.6f} std={agg['std']:.6f} (total_count={agg['total_count']})")
    os.makedirs(args.output_dir, exist_ok=True)
    with open(os.path.join(args.output_dir, "stats_per_fold.json"), "w") as f:
        json.dump(per_fold, f, indent=2)
    with open(os.path.join(args.output_dir, "stats_aggregate.json"), "w") as f:
        json.dump(agg, f, indent=2)
    np.savez(os.path.join(args.output_dir, "stats_all.npz"), per_fold=per_fold, aggregate=agg)
    print(f"\nSaved per-fold and aggregate stats to {args.output_dir}")


if __name__ == "__main__":
    main()
