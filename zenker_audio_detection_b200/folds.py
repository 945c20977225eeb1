"""All five folds in one process: the drop-in for ``src/run_all_folds_simple_batch.sh`` (the outermost caller
of the two-stage path).  The shell script starts ``run_batch_simple_2stage.py`` once per fold, which starts one Python
process per patient (run_batch:282-284); here the interpreter, torch and the CUDA context are paid once, each fold's two
models are loaded once (``batch.run``), and under ``torchrun`` the GPUs of the box share every fold's patients.

    python -m zenker_audio_detection_b200.folds [MODEL_DIR] [--no-threshold-config] [--stage1-forward-min-prob V]
                                                [--stage2-argmax] [--dry-run]
    python -m torch.distributed.run --nproc-per-node 8 -m zenker_audio_detection_b200.folds runs

Same conventions as the script (run_all_folds_simple_batch.sh:22-152): ``LONG_AUDIO_ROOT`` from the environment or the
project root's ``.env`` file, else ``datasets/New_SwallowSet/Long``; models at
``<root>/<MODEL_DIR>/ast_classifier_stage{1,2}/fold<F>/best``; results in ``<root>/<MODEL_DIR>/results/patient_inference``;
``<root>/<MODEL_DIR>/optimal_thresholds_per_fold_both_stages.json`` is used when it exists unless
``--no-threshold-config``; unknown ``--flags`` are warned about and ignored; the last bare word is MODEL_DIR.  The
project root is the working directory (``--project-root`` overrides; the script uses the parent of its own directory).
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List, Optional, Sequence

from . import batch

FOLDS = (1, 2, 3, 4, 5)
FALLBACK_LONG_AUDIO_ROOT = "datasets/New_SwallowSet/Long"


def read_env_file(path: str) -> Dict[str, str]:
    """``KEY=VALUE`` lines of a ``.env`` file (the script sources it, :29-34): comments, ``export`` prefixes and quotes
    are handled; anything fancier than plain assignments is ignored."""
    out: Dict[str, str] = {}
    if not os.path.isfile(path):
        return out
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#") or "=" not in line:
                continue
            if line.startswith("export "):
                line = line[len("export "):].lstrip()
            key, val = line.split("=", 1)
            val = val.strip()
            if len(val) >= 2 and val[0] == val[-1] and val[0] in "\"'":
                val = val[1:-1]
            out[key.strip()] = val
    return out


def parse(argv: Sequence[str]):
    """The script's own argument loop (:43-82): flags in any order, unknown ``--flags`` warned about, bare word = MODEL_DIR."""
    opts = {"model_dir": "runs", "no_threshold_config": False, "stage2_argmax": False, "dry_run": False,
            "stage1_forward_min_prob": None, "project_root": None, "schedule": None}
    args = list(argv)
    i = 0
    while i < len(args):
        a = args[i]
        if a == "--no-threshold-config":
            opts["no_threshold_config"] = True
        elif a == "--stage2-argmax":
            opts["stage2_argmax"] = True
        elif a == "--dry-run":
            opts["dry_run"] = True
        elif a in ("--stage1-forward-min-prob", "--project-root", "--schedule"):
            if i + 1 >= len(args):
                raise SystemExit(f"Error: {a} requires a value")
            opts[a[2:].replace("-", "_")] = args[i + 1]
            i += 1
        elif a.startswith("--"):
            print(f"Warning: Unknown option {a}", file=sys.stderr)
        else:
            opts["model_dir"] = a
        i += 1
    return opts


def fold_argv(opts, fold: int, root: str, long_audio_root: str) -> List[str]:
    """The command line the script builds for one fold (:113-146), for ``batch.build_arg_parser``."""
    model_dir = os.path.join(root, opts["model_dir"])
    argv = ["--fold", str(fold), "--long-audio-root", long_audio_root, "--pattern", "*.wav",
            "--stage1-model-root", os.path.join(model_dir, "ast_classifier_stage1", f"fold{fold}", "best"),
            "--stage2-model-root", os.path.join(model_dir, "ast_classifier_stage2", f"fold{fold}", "best"),
            "--output-dir", os.path.join(model_dir, "results", "patient_inference")]
    cfg = os.path.join(model_dir, "optimal_thresholds_per_fold_both_stages.json")
    if not opts["no_threshold_config"] and os.path.isfile(cfg):
        argv += ["--threshold-config", cfg]
    if opts["stage1_forward_min_prob"] is not None:
        argv += ["--stage1-forward-min-prob", str(opts["stage1_forward_min_prob"])]
    if opts["stage2_argmax"]:
        argv.append("--stage2-argmax")
    if opts["dry_run"]:
        argv.append("--dry-run")
    if opts["schedule"]:
        argv += ["--schedule", opts["schedule"]]
    return argv + ["--plot"]


def main(argv: Optional[Sequence[str]] = None) -> int:
    opts = parse(sys.argv[1:] if argv is None else argv)
    root = os.path.abspath(opts["project_root"] or os.getcwd())
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    say = print if rank == 0 else (lambda *a, **k: None)
    long_root = os.environ.get("LONG_AUDIO_ROOT") or read_env_file(os.path.join(root, ".env")).get("LONG_AUDIO_ROOT", "")
    if not long_root:
        say("Warning: LONG_AUDIO_ROOT not set. Please set it as environment variable or in .env file")
        say(f"Using fallback: {FALLBACK_LONG_AUDIO_ROOT}")
        long_root = FALLBACK_LONG_AUDIO_ROOT
    say(f"Long audio directory: {long_root}")
    say(f"Using models from: {opts['model_dir']}")
    out_base = os.path.join(root, opts["model_dir"], "results", "patient_inference")
    os.makedirs(out_base, exist_ok=True)
    say(f"Output directory: {out_base}")
    cfg = os.path.join(root, opts["model_dir"], "optimal_thresholds_per_fold_both_stages.json")
    if opts["no_threshold_config"]:
        say("Threshold config disabled (--no-threshold-config), will use default 0.5 threshold")
    elif os.path.isfile(cfg):
        say(f"Found threshold config: {cfg}")
    else:
        say(f"No threshold config found in {opts['model_dir']}, will use default 0.5 threshold")
    failures = 0
    cwd = os.getcwd()
    try:
        os.chdir(root)  # batch.run resolves ./data_ast_stage2 (run_batch:40-45) against the project root
        for fold in FOLDS:
            say(f"================ Fold {fold} ================")
            args = batch.build_arg_parser().parse_args(fold_argv(opts, fold, root, long_root))
            failures += batch.run(args, rank, world)
            say(f"\nDone fold {fold}\n")
    finally:
        os.chdir(cwd)
    say("All folds completed.")
    return failures


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
