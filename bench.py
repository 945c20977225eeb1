#!/usr/bin/env python
"""bench.py -- the two-stage sliding-window cascade on B200 (BASELINE.json metric "2-stage windows/s").

One "step" = one pass of the hot path over one synthetic 10-minute 48 kHz recording (BASELINE.json configs[1]):
channel mean + 48->16 kHz resample, continuous Kaldi fbank, Stage-1 AST forward over all 1199 sliding windows,
softmax + threshold gate + compaction, Stage-2 AST forward on the forwarded windows, scores back to the host.
windows/s counts Stage-1 windows; the Stage-2 work is inside the time but not in the count.

  value  : whole-job windows/s with the waveform already resident in HBM (CUDA events, max over ranks)
  e2e    : the same through the public API with the waveform in pinned HOST memory (H2D + D2H inside the timing)
  roofline: the dominant kernel class (the fc1 GEMM) timed live with CUDA events inside the timed region
  cpu_baseline / --impl reference: the reference's CPU stack (installed transformers + torchaudio, fp32, all host
           threads) through the oracle's restatement of forward_probs (ref:104-113) on a bounded sample of windows.

N > 1 (torchrun): every rank runs its own recording per step (weak scaling, recordings are independent) and the
per-window score records are all-gathered over NCCL inside the timed region, as the path does after its last stage.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_WINDOW = 261.028  # dense MMA flops of one AST-base forward at 1214 tokens
# The last encoder layer only feeds tokens 0/1 to the classifier, so it runs K/V for every token and the rest for two
# rows per window (zk_model.cu): 242.211 GFLOP per window are actually executed.
GFLOP_EXECUTED_PER_WINDOW = 242.211
FULL_FC1_LAYERS = 11
TOKENS, HID, MLP = 1214, 768, 3072
METRIC, UNIT = "two_stage_windows_per_s", "windows/s"


def load_traffic():
    """DRAM bytes per full-batch fc1 launch from the committed ncu capture (profiles/roofline_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        d = json.load(open(p))["gemm_fc1"]
        return int(d["dram_bytes_read"] + d["dram_bytes_write"]), d["rows"]
    except Exception:
        return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_burst": d.get("bf16_tflops", 1590.0),
                "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2]
        return {"sm_mhz": busy, "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_reference_windows_per_s(windows, stage2_fraction: float, seconds_budget: float = 20.0):
    """HF ASTFeatureExtractor + ASTForAudioClassification on the host cores (fp32, all threads): Stage 1 on the
    sample, Stage 2 on the same fraction of it the GPU cascade forwards.  Returns (windows/s, cores, description)."""
    from oracle import thirdparty
    from zenker_audio_detection_b200 import synth

    cores = thirdparty.set_threads(os.cpu_count())
    fx1 = thirdparty.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD)
    fx2 = thirdparty.hf_feature_extractor(synth.STAGE2_MEAN, synth.STAGE2_STD)
    m1 = thirdparty.hf_model_from_state_dict(synth.random_state_dict(11))
    m2 = thirdparty.hf_model_from_state_dict(synth.random_state_dict(22))
    bs = 8
    thirdparty.forward_probs(m1, fx1, windows[:2], bs)  # warm the thread pool / allocator
    t0 = time.perf_counter()
    thirdparty.forward_probs(m1, fx1, windows[:bs], bs)
    per_batch = time.perf_counter() - t0
    n1 = int(max(bs, min(len(windows), bs * max(1, int(seconds_budget / max(per_batch, 1e-3) / (1 + stage2_fraction) - 1)))))
    n1 = (n1 // bs) * bs
    n2 = int(round(n1 * stage2_fraction))
    t0 = time.perf_counter()
    thirdparty.forward_probs(m1, fx1, windows[:n1], bs)
    if n2:
        thirdparty.forward_probs(m2, fx2, windows[:n2], bs)
    dt = time.perf_counter() - t0
    desc = (f"{n1} stage-1 + {n2} stage-2 one-second windows of the cfg2 recording through HF ASTFeatureExtractor + "
            f"ASTForAudioClassification fp32 on {cores} host threads (batch {bs}, oracle.thirdparty.forward_probs = ref:104-113)")
    return n1 / dt, cores, desc, dt


def run_reference(args):
    from oracle import glue, thirdparty
    from zenker_audio_detection_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rec = synth.recording(60.0, 48000, seed=2002)  # a 1-minute slice is plenty for the bounded sample
    audio = thirdparty.resample(rec, 48000, 16000)
    windows = glue.window_audio(audio, 1.0, 0.5)
    vals = []
    total = 0.0
    for i in range(args.warmup + args.steps):
        v, cores, desc, dt = cpu_reference_windows_per_s(windows, args.stage2_fraction, seconds_budget=args.cpu_seconds)
        if i >= args.warmup:
            vals.append(v)
            total += dt
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: two-stage cascade over a synthetic 10-min 48 kHz recording (bounded sample of its windows)",
                   "stage2_fraction": args.stage2_fraction},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- our arm
def build_pipeline(args, device):
    from zenker_audio_detection_b200 import ops, synth
    from zenker_audio_detection_b200.fx import ZenkerASTFeatureExtractor
    from zenker_audio_detection_b200.model import ZenkerASTForAudioClassification
    from zenker_audio_detection_b200.pipeline import TwoStagePipeline

    fx1 = ZenkerASTFeatureExtractor(mean=synth.STAGE1_MEAN, std=synth.STAGE1_STD)
    fx2 = ZenkerASTFeatureExtractor(mean=synth.STAGE2_MEAN, std=synth.STAGE2_STD)
    sd1, sd2 = synth.random_state_dict(11), synth.random_state_dict(22)
    m1 = ZenkerASTForAudioClassification({"max_length": 1024}, sd1).to(device)
    m2 = ZenkerASTForAudioClassification({"max_length": 1024}, sd2).to(device)
    return TwoStagePipeline(m1, fx1, m2, fx2, batch_size=args.batch_size), sd1


def calibrate_gate(pipe, sd1, wave_dev, fraction, device):
    """Random-init weights give an arbitrary gate split; shift the Stage-1 head bias (as the oracle conditioning in
    SURVEY.md 8c does) so that ~`fraction` of this recording's windows are forwarded to Stage 2."""
    from zenker_audio_detection_b200.model import ZenkerASTForAudioClassification

    res = pipe.run_waveform(wave_dev, 48000)
    p = np.clip(res.s1_probs.astype(np.float64), 1e-12, 1.0)
    d = np.log(p[:, 1]) - np.log(p[:, 0])  # = l1 - l0
    shift = -float(np.quantile(d, 1.0 - fraction))
    sd = dict(sd1)
    b = sd["classifier.dense.bias"].clone()
    b[1] += shift
    sd["classifier.dense.bias"] = b
    pipe.m1 = ZenkerASTForAudioClassification({"max_length": 1024}, sd).to(device)
    return shift


def cpu_secondary(windows):
    """the reference's front end on the host cores, bounded samples (a few seconds in all):
    HF ASTFeatureExtractor over 64 windows, torchaudio kaldi.fbank over 10 min of 16 kHz audio (1/6 of cfg3),
    torchaudio resample 48 -> 16 kHz over 2 min (1/5 of cfg2); GB/s are the algorithmic bytes of SURVEY.md 8d."""
    from oracle import thirdparty
    from zenker_audio_detection_b200 import synth

    out = {"cores": torch.get_num_threads()}
    fx = thirdparty.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD)
    w = [np.ascontiguousarray(x) for x in windows[:64]]
    t0 = time.perf_counter()
    fx(w, sampling_rate=16000, return_tensors="pt")
    out["fx_windows_per_s"] = len(w) / (time.perf_counter() - t0)
    n = 9_600_000
    wave = (np.random.default_rng(3003).standard_normal(n) * 0.05).astype(np.float32)
    t0 = time.perf_counter()
    fb = thirdparty.kaldi_fbank(wave)
    dt = time.perf_counter() - t0
    out["fbank_gb_per_s"] = (4.0 * n + 512.0 * fb.shape[0]) / dt / 1e9
    rec = synth.recording(120.0, 48000, seed=2002)
    t0 = time.perf_counter()
    a = thirdparty.resample(rec, 48000, 16000)
    dt = time.perf_counter() - t0
    out["resample_gb_per_s"] = (4.0 * rec.shape[-1] + 4.0 * a.shape[-1]) / dt / 1e9
    return out


def secondary_metrics(device, peaks, engine=None):
    """The other two quantities BASELINE.json's metric names, measured outside the timed region on rank 0: the
    continuous fbank over 1 h of 16 kHz audio (cfg3; algorithmic bytes 4 n + 512 m) and the 48 -> 16 kHz
    resampler over a 10-minute recording (cfg2), as achieved GB/s against the measured HBM copy bandwidth."""
    from zenker_audio_detection_b200 import ops

    def best_ms(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):  # back to back: a single 40 us launch would be dominated by the host-side call
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {}
    g = torch.Generator(device=device).manual_seed(3003)
    plan = ops.FbankPlan()
    wave = torch.randn(57_600_000, device=device, generator=g) * 0.05
    m = plan.num_frames(wave.numel())
    ms = best_ms(lambda: plan.fbank(wave))
    gbs = (4.0 * wave.numel() + 512.0 * m) / ms / 1e6
    out["fbank_cfg3"] = {"ms": ms, "gb_per_s": gbs, "frac_hbm_peak": gbs / peaks["hbm_gbs"], "frames": int(m)}
    del wave
    rec = torch.randn(28_800_000, device=device, generator=g) * 0.1
    ms = best_ms(lambda: ops.resample(rec, 48000, 16000))
    gbs = (4.0 * 28_800_000 + 4.0 * 9_600_000) / ms / 1e6
    out["resample_cfg2"] = {"ms": ms, "gb_per_s": gbs, "frac_hbm_peak": gbs / peaks["hbm_gbs"]}
    if engine is not None:
        # cfg5: one AST forward over (32, 1024, 128) features, 8.353 TFLOP dense -> 6.01 ms at the
        # sustained bf16 peak; the last-layer pruning executes 7.751 TFLOP of it
        feats = torch.randn(32, 1024, 128, device=device, generator=torch.Generator(device=device).manual_seed(5005)) * 0.5
        ms = best_ms(lambda: engine.forward_features(feats), reps=5)
        out["ast_forward_cfg5"] = {"ms": ms, "tflops_dense_equivalent": 32 * GFLOP_PER_WINDOW / ms,
                                   "tflops_executed": 32 * GFLOP_EXECUTED_PER_WINDOW / ms,
                                   "frac_of_sustained_peak_executed": 32 * GFLOP_EXECUTED_PER_WINDOW / ms / peaks["bf16_sustained"]}
    out["note"] = ("10 back-to-back launches each, taken right after the timed region, i.e. at the power-capped clock "
                   "the clocks key reports; scripts/bench_kernels.py times the same kernels from a cold start")
    return out


def run_ours(args):
    import torch.distributed as dist

    from zenker_audio_detection_b200 import _lib, dist as zdist, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the zenker-b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.require_device()
    peaks = load_peaks()

    seconds = args.recording_seconds
    rec = synth.recording(seconds, 48000, seed=2002 + rank)
    host = torch.from_numpy(rec).pin_memory()
    wave_dev = host.to(device)
    pipe, sd1 = build_pipeline(args, device)
    calibrate_gate(pipe, sd1, wave_dev, args.stage2_fraction, device)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step(src):
        r = pipe.run_waveform(src, 48000)
        if world > 1:
            zdist.all_gather_records(zdist.pack_records(rank, r.s1_probs, r.swallow_indices, r.s2_probs), device)
        return r

    def timed(src, steps, time_launches):
        _lib.prof_collect()
        _lib.prof_enable(time_launches)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = k = 0
        for _ in range(steps):
            r = step(src)
            n += r.num_windows
            k += len(r.swallow_indices)
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        prof = _lib.prof_collect()
        _lib.prof_enable(False)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            c = torch.tensor([n, k], dtype=torch.int64, device=device)
            dist.all_reduce(c)
            n, k = int(c[0].item()), int(c[1].item())
        return ms, n, k, prof

    for _ in range(args.warmup):
        step(wave_dev)
    with ClockSampler(local) as clk:
        ms, n, k, prof = timed(wave_dev, args.steps, True)
    clocks = clk.summary()
    for _ in range(min(args.warmup, 1)):
        step(host)
    ms_e2e, n_e2e, _, _ = timed(host, args.steps, False)

    value = n / (ms / 1000.0)
    e2e = n_e2e / (ms_e2e / 1000.0)
    launches = sum(v[1] for v in prof.values())
    # dominant kernel class: the fc1 GEMM ([B*1214 x 768] x [768 x 3072] + bias + GELU), 2*M*768*3072 flops / launch
    windows_fwd = (n + k) // max(1, world)  # windows one rank pushed through an AST forward in the timed region
    fc1_ms, fc1_n = prof["gemm_fc1"]
    full_fc1 = 12 if os.environ.get("ZK_FULL_LAST_LAYER", "0") not in ("", "0") else FULL_FC1_LAYERS
    flops_per_launch = 2.0 * (windows_fwd * TOKENS) * HID * MLP * full_fc1 / max(1, fc1_n)
    achieved = flops_per_launch / (fc1_ms / max(1, fc1_n) * 1e-3) / 1e12 if fc1_ms > 0 else None
    peak = peaks["bf16_sustained"]
    traffic, traffic_rows = load_traffic()
    gemm_ms = sum(prof[c][0] for c in ("gemm_qkv", "gemm_out", "gemm_fc1", "gemm_fc2", "gemm_patch"))
    breakdown = {c: round(v[0] / args.steps, 3) for c, v in prof.items() if v[1]}
    model_tflops = (n + k) / max(1, world) * GFLOP_PER_WINDOW / 1e3 / (ms / 1000.0)
    exec_gflop = GFLOP_PER_WINDOW if full_fc1 == 12 else GFLOP_EXECUTED_PER_WINDOW
    model_tflops_exec = (n + k) / max(1, world) * exec_gflop / 1e3 / (ms / 1000.0)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"cfg2: full two-stage cascade over one synthetic {seconds:.0f}-s 48 kHz recording per GPU per step "
                               f"({n // max(1, world) // args.steps} sliding 1-s windows, hop 0.5 s; Stage 2 on the compacted swallow windows)",
                   "batch_size": args.batch_size, "stage2_fraction": round(k / max(1, n), 4),
                   "weights": "random-init AST-base x2 (conditioned, SURVEY.md 8c)", "parallelism": f"recordings sharded over {world} GPU(s)",
                   "l2": "activation working set ~18.7 MB/window x batch >> 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(host.numel() * 4),
                "d2h_bytes_per_step": int((n_e2e // max(1, world) // args.steps) * 12 + 4 + (k // max(1, world) // args.steps) * 12)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "pair::gemm_pair_kernel<BIAS_GELU> (fc1, M=batch*1214, N=3072, K=768)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "traffic_note": (f"dram read+write of one {traffic_rows}-row launch, ncu --set full "
                                                          "(profiles/roofline_traffic.json)") if traffic else None,
                     "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step)",
                     "flops_per_launch": flops_per_launch, "launches": fc1_n},
        "kernel_ms_per_step": breakdown,
        "model_tflops_dense_equivalent": model_tflops, "model_tflops_executed": model_tflops_exec,
        "model_executed_frac_of_peak": model_tflops_exec / peak,
        "gemm_share_of_step": gemm_ms / ms if ms else None,
    }
    if rank == 0:
        line["secondary"] = secondary_metrics(device, peaks, pipe.m2.engine)
        if args.cpu_seconds > 0:
            from oracle import glue, thirdparty

            audio = thirdparty.resample(synth.recording(60.0, 48000, seed=2002), 48000, 16000)
            v, cores, desc, _ = cpu_reference_windows_per_s(glue.window_audio(audio, 1.0, 0.5), k / max(1, n),
                                                            seconds_budget=args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
            line["cpu_baseline"]["secondary"] = cpu_secondary(glue.window_audio(audio, 1.0, 0.5))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-size", type=int, default=128)
    ap.add_argument("--recording-seconds", type=float, default=600.0)
    ap.add_argument("--stage2-fraction", type=float, default=0.3)
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the CPU baseline sample (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
