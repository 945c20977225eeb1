"""Result documents of the two-stage path: the ``{config, per_file, aggregate}`` JSON the reference writes per
patient (ref:384-410; cached variant refc:570-601), so ``utils/aggregate_2stage_results.py`` can consume ours."""
from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional, Sequence

from . import cascade


def build_document(stage1_model_root: str, stage2_model_root: str, window_sec: float, hop_sec: float, batch_size: int,
                   stage1_threshold: float, files: Sequence[str], summaries: Sequence[Dict[str, Any]],
                   variant: str = "plain", stage2_threshold: float = 0.5,
                   stage1_forward_min_prob: Optional[float] = None, stage2_argmax: bool = False,
                   feature_cache_dir: Optional[str] = None, disable_cache: bool = True,
                   refresh_cache: bool = False) -> Dict[str, Any]:
    """``variant`` "plain" -> test_long_audio_windows_2stage.py schema, "cached" -> ..._cache.py schema."""
    files = list(files)
    per_file = {f"file_{i}": {"path": p, **s} for i, (p, s) in enumerate(zip(files, summaries))}
    config: Dict[str, Any] = {
        "stage1_model_root": stage1_model_root, "stage2_model_root": stage2_model_root, "window_sec": window_sec,
        "hop_sec": hop_sec, "batch_size": batch_size, "stage1_threshold": stage1_threshold,
    }
    if variant == "cached":
        config.update({"stage1_forward_min_prob": stage1_forward_min_prob, "stage2_threshold": stage2_threshold,
                       "stage2_argmax": stage2_argmax, "files": files, "feature_cache_dir": feature_cache_dir,
                       "disable_cache": disable_cache, "refresh_cache": refresh_cache})
    elif variant == "plain":
        config["files"] = files
    else:
        raise ValueError(f"unknown variant {variant!r}")
    return {"config": config, "per_file": per_file, "aggregate": cascade.aggregate_patient(per_file, files)}


def default_output_path(patient_id: str, variant: str = "plain", out_dir: str = "outputs") -> str:
    """ref:399-402 (``<pid>_2stage.json``) / refc:590-593 (``<pid>_2stage_cached.json``)."""
    return os.path.join(out_dir, f"{patient_id}_2stage{'_cached' if variant == 'cached' else ''}.json")


def write_json(doc: Dict[str, Any], path: str) -> None:
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "w") as f:
        json.dump(doc, f, indent=2)  # NaN is written bare, like the reference
