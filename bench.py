#!/usr/bin/env python
"""bench.py - frames/s of the streamed DINOv3 ViT-B/16 encode (BASELINE.json configs[1]) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one 512-frame chunk (the reference's CHUNK_SIZE, cbas.py:48) of a synthetic 10-min 30-fps clip
(18 000 frames, 256x256 uint8 RGB) through the whole hot path: fused resize-to-224 / normalise / patchify,
the 12-block ViT, CLS pooling.  `value` is device-timed with the frames resident in HBM; `e2e` is the public call a
user makes - `cbas_b200.cbas.encode_file` on a clip file (a .npy of the same frames: no video decoder on the box),
i.e. PAGEABLE host frames -> pinned staging -> H2D -> encode -> D2H -> `_cls.h5` on disk, all inside the timed region.
Work shards by clip/chunk across ranks (one process per GPU, no collective on the data path): weak scaling.

Besides the contract's keys the line carries
  roofline / forward.kernels  per-kernel device time with EXECUTED FLOPs (summed over the launches, the pruned last
                              block included as what it is) against the measured bf16 peaks, sustained and burst;
  gpu_eager_baseline          the reference's own GPU path (transformers eager, fp16 autocast, SDPA) timed on the same
                              B200 (oracle/eager_gpu.py) - N = 1 only;
  cpu_baseline                the reference's CPU path on the host cores (bounded sample) - N = 1 only;
  other_workloads             BASELINE configs[3] (head, 1 M frames) and configs[4] (backlog, scaled) in short form -
                              N = 1 only; `--workload head|backlog` prints their full lines.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = 512
CLIP_FRAMES = 18000          # 10 min x 30 fps
SRC_HW = (256, 256)          # recording geometry (cbas.py:733)
SIDE = 224
METRIC = "frames_per_sec_encoded_dinov3_vitb16_224px"


def metric_name(arch):
    """BASELINE.json's metric for the default arch; the same pattern for the others."""
    return METRIC.replace("dinov3_vitb16", arch.replace("-", "_") if arch.startswith("dinov2") else f"dinov3_{arch}")


def flops_per_frame(D, L, I, side, patch=16):
    """Dense forward FLOPs: 2*[Np*3*P^2*D + L*(4*N*D^2 + 2*N^2*D + 2*N*D*I)], N = Np + 5
    (P = 16: the patch term is Np*768*D)."""
    Np = (side // patch) ** 2
    N = Np + 5
    return 2.0 * (Np * 3 * patch * patch * D + L * (4 * N * D * D + 2 * N * N * D + 2 * N * D * I))


def executed_work(a, side, n, preprocess, pruned=True):
    """Per profile tag: FLOPs (GEMMs, attention) or algorithmic HBM bytes (the rest) one n-frame step EXECUTES.
    With last-block pruning the final block projects K/V for every token but runs Q, attention, proj and the MLP on the
    n CLS rows only (cbas_b200/csrc/encoder.cu::encoder_last_layer_cls_only)."""
    D, L, I, P = a["hidden_size"], a["num_hidden_layers"], a["intermediate_size"], a.get("patch_size", 16)
    Np = (side // P) ** 2
    T, heads = Np + 5, D // 64
    M = n * T
    Kp = ((3 * P * P if preprocess == "processor" else P * P) + 63) // 64 * 64
    full = L - 1 if pruned else L
    last = 1 if pruned else 0
    return {
        "flops": {
            "patch_gemm": 2.0 * n * Np * Kp * D,
            "qkv_gemm": full * 2.0 * M * 3 * D * D + last * (2.0 * M * 2 * D * D + 2.0 * n * D * D),
            "proj_gemm": full * 2.0 * M * D * D + last * 2.0 * n * D * D,
            "up_gemm": full * 2.0 * M * I * D + last * 2.0 * n * I * D,
            "down_gemm": full * 2.0 * M * D * I + last * 2.0 * n * D * I,
            "attention": full * 4.0 * n * heads * T * T * 64 + last * 4.0 * n * heads * T * 64,
        },
        "bytes": {
            "preprocess": n * (side * side * 3 if preprocess != "processor" else SRC_HW[0] * SRC_HW[1] * 3) + n * Np * Kp * 2.0,
            # standalone norm1 / norm2 kernels (the default): fp32 h in, bf16 out, two per full block, norm1 over every
            # row + norm2 over the CLS rows in the pruned last block
            "layernorm": (2 * full + last) * M * D * (4 + 2.0) + last * n * D * (4 + 2.0),
            "layernorm_fused": M * D * (4 + 2.0),  # CBAS_B200_LN_FUSION=1: only the statistics pass after the embedding
            "final_ln": n * D * (4 + 4.0),
        },
    }


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(tf_sustained=float(j.get("bf16_tflops_sustained", 1400.0)), tf_burst=float(j["bf16_tflops"]),
                    hbm=float(j["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(tf_sustained=1400.0, tf_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, threading.Event(), [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_fps(frames_per_step, steps, warmup, threads=None, min_seconds=0.0):
    """The reference's own CPU implementation of the path: transformers' DINOv3ViTModel (what cbas.py:657,676
    runs) in fp32 behind the HF processor arithmetic, restated in oracle/encoder.py, on all host threads.
    Runs `steps` steps (more, until min_seconds of work have accumulated)."""
    from oracle import encoder as oenc
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = oenc.build_hf_model("vitb16", seed=0)
    frames = np.random.default_rng(0).integers(0, 256, (frames_per_step, *SRC_HW, 3), dtype=np.uint8)
    for _ in range(warmup):
        oenc.encode(model, frames, mode="processor", size=SIDE, batch=frames_per_step)
    done = 0
    t0 = time.perf_counter()
    while done < steps or time.perf_counter() - t0 < min_seconds:
        oenc.encode(model, frames, mode="processor", size=SIDE, batch=frames_per_step)
        done += 1
    dt = time.perf_counter() - t0
    return frames_per_step * done / dt, dt, torch.get_num_threads(), done


def run_reference(args, rank):
    if rank != 0:
        return
    fpst = 4
    fps, dt, threads, done = cpu_reference_fps(fpst, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": metric_name(args.arch), "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DINOv3 ViT-B/16 224px streamed encode of a synthetic 10-min 30fps 256x256 clip "
                               "(BASELINE configs[1]); each step a 4-frame sample on the host CPU"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{fpst} frames/step x {args.steps} steps, transformers DINOv3ViTModel fp32 + "
                                   "HF-processor preprocessing (oracle/encoder.py), random-init weights"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ head (configs[3])
HEAD_FRAMES = 1_000_000


def run_head(args, rank, world, local_rank):
    """LSTM classifier head over 1 M precomputed ViT-B embeddings, 9 behaviours, window 31 (BASELINE configs[3]).
    A step = the whole 1 M-frame array through cbas_b200_head_infer (what infer_file calls)."""
    import torch.distributed as dist
    from cbas_b200 import _lib
    from cbas_b200.classifier_head import ClassifierLSTMDeltas, synthetic_head_state_dict
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(3, args.warmup), max(1, min(args.steps, 10))
    sd = synthetic_head_state_dict(768, 9, seed=0, scale=2.0)
    head = ClassifierLSTMDeltas(768, 9, seq_len=31)
    head.load_state_dict(sd)
    head = head.to(dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    emb = torch.randn(HEAD_FRAMES, 768, device=dev, generator=g).half()
    for _ in range(W):
        probs = head.infer_embeddings(emb)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    _lib.profile_enable(True)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        probs = head.infer_embeddings(emb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    clocks = sampler.result()
    assert bool(torch.isfinite(probs).all())
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * K * HEAD_FRAMES / (ms_max / 1000.0)
    # end to end: f16 embeddings in pinned host memory -> H2D -> head -> probabilities D2H
    host = torch.empty(HEAD_FRAMES, 768, dtype=torch.float16).pin_memory()
    host.copy_(emb)
    out_host = torch.empty(HEAD_FRAMES, 9, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        d = host.to(dev, non_blocking=True)
        out_host.copy_(head.infer_embeddings(d), non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    total_ms = sum(v[0] for v in prof.values())
    breakdown = {k: {"ms_per_step": v[0] / K, "share": v[0] / total_ms} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    peaks = load_peaks()
    alg_bytes = HEAD_FRAMES * (768 * 2 + 9 * 4)
    # Tensor-core work the pipeline executes per frame (fp32-accurate products as three bf16 passes, hi*hi + lo*hi +
    # hi*lo): bottleneck projections once per frame, lin0 on all 31 steps of the frame's window, the input gates on the
    # 21 steps each direction runs, the recurrence (64 x 256 per step and direction).  The GEMM stages are bound by
    # this, not by HBM: it also sets the floor of the whole head (total / sustained bf16 peak).
    f_proj, f_lin0 = 2 * 3 * 768 * 384, 2 * 3 * 31 * 384 * 256
    f_ih, f_rec = 2 * 3 * 42 * 256 * 256, 2 * 3 * 42 * 64 * 256
    def _tf(tag, flop):
        ms = breakdown.get(tag, {}).get("ms_per_step")
        return None if not ms else {"tflops": HEAD_FRAMES * flop / (ms / 1000.0) / 1e12,
                                    "frac_of_sustained": HEAD_FRAMES * flop / (ms / 1000.0) / 1e12 / peaks["tf_sustained"]}
    tensor = {"flop_per_frame": f_proj + f_lin0 + f_ih + f_rec,
              "floor_ms_per_1M_frames_at_sustained_peak": (f_proj + f_lin0 + f_ih + f_rec) * 1e6 / (peaks["tf_sustained"] * 1e12) * 1e3,
              "head_lin0_gemm": _tf("head_lin0_gemm", f_lin0), "head_ih_gemm": _tf("head_ih_gemm", f_ih),
              "head_proj_gemm": _tf("head_proj_gemm", f_proj)}
    cpu_baseline = gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        try:
            from oracle import eager_gpu
            gpu_eager = eager_gpu.time_eager_head(sd, 768, 9, dev, frames=60000)
            gpu_eager["speedup_of_e2e"] = world * K * HEAD_FRAMES / e2e_s / gpu_eager["value"]
        except Exception as exc:
            gpu_eager = {"unavailable": f"{type(exc).__name__}: {exc}"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import head as ohead
        n = 2048
        e = np.random.default_rng(0).standard_normal((n, 768)).astype(np.float16)
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 10.0:
            ohead.infer_windows(e, sd, seq_len=31)
            reps += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": n * reps / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{n * reps} windows in {dt:.1f} s: oracle/head.py (restated ClassifierLSTMDeltas + "
                                  "infer_file window loop, batch 512, fp32)"}
    if rank == 0:
        print(json.dumps({
            "metric": "frames_per_sec_lstm_head_inference", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (split-bf16 tensor-core GEMMs, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "LSTM classifier head over 1M precomputed ViT-B embeddings, 9 behaviours, window 31 "
                                   "(BASELINE configs[3])", "frames": HEAD_FRAMES},
            "roofline": {"bound": "hbm", "kernel": "head pipeline (all stages)", "achieved": alg_bytes / (ms_max / K / 1000.0) / 1e9,
                         "peak": peaks["hbm"], "unit": "GB/s", "frac": alg_bytes / (ms_max / K / 1000.0) / 1e9 / peaks["hbm"],
                         "traffic": None, "note": "algorithmic bytes = 1536 B in + 36 B out per frame; the pipeline is "
                                                  "bound by its intermediates and the recurrence, not by this traffic"},
            "stages": breakdown, "tensor": tensor, "cpu_baseline": cpu_baseline, "gpu_eager_baseline": gpu_eager,
            "e2e": {"value": world * K * HEAD_FRAMES / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": HEAD_FRAMES * 768 * 2,
                    "d2h_bytes_per_step": HEAD_FRAMES * 9 * 4},
            "clocks": clocks, "gpu_launches": int(launches)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ backlog (configs[4])
def run_backlog(args, rank, world, local_rank):
    """End-to-end recording backlog (BASELINE configs[4]): one camera per GPU, `--hours` of 10-fps 256x256 recording
    in the reference's 600-s segments (cbas.py:732-734) -> streamed ViT-B/16 encode from pinned host memory -> f16
    embeddings (as `_cls.h5` stores them) -> LSTM head over each segment -> per-behaviour actogram bins (30-min bins)
    -> the bin vectors of all cameras summed across ranks (the one optional collective, parallel.allreduce_bins).
    Frames are synthetic and one pinned 512-frame chunk is reused per H2D copy (no decoder, no 24-h clip in RAM);
    every copy, kernel and the D2H of the embeddings / probabilities is inside the timed region."""
    import torch.distributed as dist
    from cbas_b200 import _lib, parallel
    from cbas_b200.classifier_head import ClassifierLSTMDeltas, actogram_bins, synthetic_head_state_dict
    from cbas_b200.encoder import DinoEncoder
    from cbas_b200.pipeline import StreamedEncoder
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fps, seg_s = 10, 600
    seg_frames = fps * seg_s
    n_seg = max(1, int(round(args.hours * 3600 / seg_s)))
    bin_frames = int(30 * fps * 60)
    with contextlib.redirect_stdout(sys.stderr):
        enc = DinoEncoder(f"synthetic:{args.arch}", dev, preprocess="processor", image_size=SIDE, max_frames=CHUNK)
    D = enc.hidden_size
    sd = synthetic_head_state_dict(D, 9, seed=0, scale=2.0)
    head = ClassifierLSTMDeltas(D, 9, seq_len=31)
    head.load_state_dict(sd)
    head = head.to(dev)
    g = torch.Generator().manual_seed(rank)
    host_chunks = [torch.randint(0, 256, (CHUNK, *SRC_HW, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(2)]
    pipe = StreamedEncoder(enc, SRC_HW, CHUNK, depth=2)
    emb_host = torch.empty(seg_frames, D, dtype=torch.float16).pin_memory()

    def segment():
        """one 600-s file: encode -> f16 rows on the host (the `_cls.h5` content) -> head -> probabilities"""
        pos = [0]

        def chunks():
            for i in range(0, seg_frames, CHUNK):
                yield host_chunks[(i // CHUNK) & 1][:min(CHUNK, seg_frames - i)]

        def sink(e):
            n = e.shape[0]
            emb_host[pos[0]:pos[0] + n].copy_(torch.from_numpy(e))  # float32 -> float16, like the store does
            pos[0] += n

        pipe.run(chunks(), sink)
        return head.infer_embeddings(emb_host.to(dev, non_blocking=True))

    segment()  # warm-up: library, workspaces, first launches
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    probs_all = [segment() for _ in range(n_seg)]
    probs = torch.cat(probs_all)
    bins = torch.stack([actogram_bins(probs, b, 0.5, bin_frames) for b in range(9)])  # [9, n_bins] int32
    names = [f"behaviour{b}" for b in range(9)]
    summed = parallel.allreduce_bins({k: bins[i] for i, k in enumerate(names)}, {k: int(bins.shape[1]) for k in names})
    total_bins = torch.stack([summed[k] for k in names])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = _lib.launch_count() - l0
    clocks = sampler.result()
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_max = float(t.item())
    frames = n_seg * seg_frames
    if rank == 0:
        print(json.dumps({
            "metric": "frames_per_sec_backlog_encode_head_actogram", "value": world * frames / dt_max, "unit": "frames/s",
            "n_gpus": world, "steps": n_seg, "warmup": 1, "ms_per_step": 1000.0 * dt_max / n_seg,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 encoder, f32 head",
            "data": "synthetic",
            "config": {"workload": f"recording backlog, one camera per GPU: {args.hours} h at {fps} fps in {seg_s}-s segments, "
                                   f"ViT-{args.arch} 224px encode + LSTM head + 30-min actogram bins (BASELINE configs[4], "
                                   "scaled by --hours); host wall clock, max over ranks, everything from pinned host "
                                   "frames to the reduced bin vectors inside",
                       "frames_per_camera": frames, "segments": n_seg, "bins_per_behaviour": int(total_bins.shape[1])},
            "realtime_factor": world * frames / dt_max / (world * fps),
            "e2e": {"value": world * frames / dt_max, "unit": "frames/s",
                    "h2d_bytes_per_step": seg_frames * (SRC_HW[0] * SRC_HW[1] * 3 + D * 2),
                    "d2h_bytes_per_step": seg_frames * (D * 4 + 0)},
            "clocks": clocks, "gpu_launches": int(launches)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=35)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="vitb16")
    ap.add_argument("--preprocess", default="processor", choices=["processor", "reference"])
    ap.add_argument("--hours", type=float, default=1.0, help="backlog workload: hours of recording per camera")
    ap.add_argument("--workload", default="encoder", choices=["encoder", "head", "backlog"],
                    help="encoder = BASELINE configs[1] (the headline); head = configs[3], 1M precomputed embeddings")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the torch-eager GPU baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the short head / backlog runs appended at N = 1")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong = ONE 18 000-frame clip (BASELINE configs[1]) split into contiguous spans over the GPUs "
                         "(SURVEY 8e); a step is one pass over the whole clip")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "backlog":
        run_backlog(args, rank, world, local_rank)
        return
    if args.workload == "head":
        run_head(args, rank, world, local_rank)
        return
    if args.scaling == "strong":
        run_strong(args, rank, world, local_rank)
        return

    import torch.distributed as dist
    from cbas_b200 import _lib
    from cbas_b200.encoder import ARCHITECTURES, DinoEncoder
    from cbas_b200.pipeline import StreamedEncoder

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (cbas_b200 has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)

    side = SIDE
    src_hw = SRC_HW if args.preprocess == "processor" else (SIDE, SIDE)
    with contextlib.redirect_stdout(sys.stderr):  # stdout carries exactly one JSON line
        enc = DinoEncoder(f"synthetic:{args.arch}", dev, preprocess=args.preprocess, image_size=side, max_frames=CHUNK)
    a = ARCHITECTURES[args.arch]
    F = flops_per_frame(a["hidden_size"], a["num_hidden_layers"], a["intermediate_size"], side, a.get("patch_size", 16))
    peaks = load_peaks()

    # the clip: K distinct chunks when memory allows (inputs >> L2), at most the 36 chunks of the 10-min clip
    n_chunks = min(K, (CLIP_FRAMES + CHUNK - 1) // CHUNK)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    clip = torch.randint(0, 256, (n_chunks * CHUNK, *src_hw, 3), dtype=torch.uint8, device=dev, generator=g)
    out = torch.empty(CHUNK, enc.hidden_size, device=dev, dtype=torch.float32)

    def step(i):
        c = i % n_chunks
        enc.encode_u8(clip[c * CHUNK:(c + 1) * CHUNK], out=out)

    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    # Per-kernel breakdown: the SAME K steps again, back to back with the timed ones (same clocks, same power state),
    # this time with a CUDA-event pair around every launch.  The events are instrumentation (they cost ~1 % and
    # serialise what programmatic dependent launch overlaps), so `value` is timed without them.
    _lib.profile_enable(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(K):
        step(W + K + i)
    p1.record()
    torch.cuda.synchronize()
    ms_profiled = p0.elapsed_time(p1)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    clocks = sampler.result()
    checksum = float(out.double().abs().sum().item())  # D2H of the result: the work was really done
    if not np.isfinite(checksum) or checksum == 0.0:
        raise SystemExit("bench: encoder produced a non-finite or all-zero result")

    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * K * CHUNK / (ms_max / 1000.0)

    # ---- per-kernel roofline: EXECUTED work summed over the launches of a tag / summed device time of the tag
    total_prof_ms = sum(v[0] for v in prof.values())
    work = executed_work(a, side, CHUNK, args.preprocess)
    traffic_all = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic_all = json.load(open(tp))
    breakdown = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        e = {"ms_per_step": v[0] / K, "launches_per_step": v[1] / K, "share": v[0] / total_prof_ms}
        sec = v[0] / K / 1000.0
        if k in work["flops"]:
            tf = work["flops"][k] / sec / 1e12
            e.update({"tflops_executed": tf, "frac_of_sustained": tf / peaks["tf_sustained"], "frac_of_burst": tf / peaks["tf_burst"]})
        elif k in work["bytes"]:
            by = work["bytes"]["layernorm_fused"] if k == "layernorm" and v[1] / K < 3 else work["bytes"][k]
            gbs = by / sec / 1e9
            e.update({"gbs_algorithmic": gbs, "frac_of_hbm": gbs / peaks["hbm"]})
        breakdown[k] = e
    dom = max(prof, key=lambda k: prof[k][0])
    d = breakdown[dom]
    if dom in work["flops"]:
        roofline = {"bound": "tensor", "kernel": dom, "achieved": d["tflops_executed"], "peak": peaks["tf_sustained"],
                    "unit": "TFLOP/s", "frac": d["frac_of_sustained"], "frac_of_burst_peak": d["frac_of_burst"],
                    "burst_peak": peaks["tf_burst"], "traffic": traffic_all.get(dom),
                    "peak_source": peaks["source"] + "; `peak` = sustained bf16 (the kernel is timed inside a long, "
                                   "power-capped step), burst figure beside it",
                    "how": "executed FLOPs of every launch of this tag in a step (the pruned last block counted as what "
                           "it executes) / summed CUDA-event time of those launches, over the event-instrumented repeat of "
                           "the timed steps (forward.kernels_measured_in)",
                    "share_of_step": d["share"]}
    else:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": d.get("gbs_algorithmic"), "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": d.get("frac_of_hbm"), "traffic": traffic_all.get(dom), "share_of_step": d["share"]}
    F_dense = F
    Fx = sum(work["flops"].values()) / CHUNK
    per_gpu = value / world
    forward = {"gflop_per_frame_dense": F_dense / 1e9, "gflop_per_frame_executed": Fx / 1e9,
               "note": "tflops_dense uses the DENSE count of SURVEY 8d (what BASELINE.md's 60 % target is drawn on); executed "
                       "is lower because the last block only computes what the pooled CLS row needs",
               "tflops_dense": per_gpu * F_dense / 1e12, "tflops_executed": per_gpu * Fx / 1e12,
               "frac_of_bf16_burst_peak": per_gpu * F_dense / 1e12 / peaks["tf_burst"],
               "frac_of_bf16_sustained_peak": per_gpu * F_dense / 1e12 / peaks["tf_sustained"],
               "target_frames_per_s_at_60pct_of_burst": 0.6 * peaks["tf_burst"] * 1e12 / F_dense,
               "kernels_measured_in": f"{K} more steps run straight after the timed ones with a CUDA-event pair around every "
                                      f"launch: {ms_profiled / K:.3f} ms/step against {ms / K:.3f} without the events",
               "kernels": breakdown}

    # ---- end to end through the public API: encode_file on a clip file (pageable frames), `_cls.h5` written
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e_encode_file(enc, clip, n_chunks, W, K, world, dev, dist if world > 1 else None)

    cpu_baseline = gpu_eager = other = None
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            fps_cpu, dt, threads, done = cpu_reference_fps(8, 2, 1, min_seconds=12.0)
            cpu_baseline = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": "port",
                            "sample": f"{8 * done} frames ({done} batches of 8) of the same workload in {dt:.1f} s: "
                                      "transformers DINOv3ViTModel fp32 + HF-processor preprocessing (oracle/encoder.py)"}
        if not args.no_gpu_baseline and not args.arch.startswith("dinov2"):
            del clip
            torch.cuda.empty_cache()
            try:
                from oracle import eager_gpu
                gpu_eager = eager_gpu.time_eager_encoder(args.arch, dev, CHUNK, src_hw, side, steps=max(4, min(K, 12)))
                gpu_eager["speedup_of_value"] = value / gpu_eager["value"]
                if e2e:
                    gpu_eager["reference_loop"]["speedup_of_e2e"] = e2e["value"] / gpu_eager["reference_loop"]["value"]
            except Exception as exc:  # a baseline leg must never take the measurement down with it
                gpu_eager = {"unavailable": f"{type(exc).__name__}: {exc}"}
        if not args.no_extra:
            other = run_other_workloads(args)

    if rank == 0:
        line = {
            "metric": metric_name(args.arch), "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{'DINOv2-with-registers' if args.arch.startswith('dinov2') else 'DINOv3'} {args.arch} {side}px streamed encode of a synthetic 10-min 30fps "
                                   f"{src_hw[0]}x{src_hw[1]} uint8 clip, {CHUNK}-frame chunks (BASELINE configs[1])",
                       "preprocess": args.preprocess, "chunk_frames": CHUNK, "tokens_per_frame": (side // a.get("patch_size", 16)) ** 2 + 5,
                       "parallelism": f"dp{world} (one clip per GPU, no collective)",
                       "l2": f"{n_chunks} distinct {CHUNK * src_hw[0] * src_hw[1] * 3 >> 20} MiB input chunks and "
                             f">1 GiB of activations per step: far larger than the 126 MB L2",
                       "weights": "random-init (synthetic:%s, gated hub weights unavailable offline)" % args.arch},
            "roofline": roofline, "forward": forward, "cpu_baseline": cpu_baseline, "gpu_eager_baseline": gpu_eager,
            "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches), "other_workloads": other,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_strong(args, rank, world, local_rank):
    """Strong scaling of BASELINE configs[1]: the ONE 10-min 30-fps clip (18 000 frames) on N GPUs.  Rank r encodes
    the contiguous span parallel.split_frame_range gives it, 512-frame chunks (its last chunk is partial); a step is
    one pass over the whole clip; `value` = clip frames / max-over-ranks device time with the span resident in HBM,
    `e2e` the same through StreamedEncoder.run from pinned host frames with the f16 embeddings copied back."""
    import torch.distributed as dist
    from cbas_b200 import _lib, parallel
    from cbas_b200.encoder import DinoEncoder
    from cbas_b200.pipeline import StreamedEncoder
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (cbas_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(3, args.warmup), max(1, args.steps)
    src_hw = SRC_HW if args.preprocess == "processor" else (SIDE, SIDE)
    with contextlib.redirect_stdout(sys.stderr):
        enc = DinoEncoder(f"synthetic:{args.arch}", dev, preprocess=args.preprocess, image_size=SIDE, max_frames=CHUNK)
    span = parallel.split_frame_range(CLIP_FRAMES, world)[rank]
    n = len(span)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)  # a span-sized random tensor per rank: the same workload
    frames = torch.randint(0, 256, (n, *src_hw, 3), dtype=torch.uint8, device=dev, generator=g)
    out = torch.empty(n, enc.hidden_size, device=dev, dtype=torch.float32)

    def one_pass():
        for s in range(0, n, CHUNK):
            e = min(s + CHUNK, n)
            enc.encode_u8(frames[s:e], out=out[s:e])

    enc.encode_u8(frames[:CHUNK], out=out[:CHUNK])  # warm-up: W chunks, not W whole passes
    for _ in range(max(0, W - 1)):
        enc.encode_u8(frames[:min(CHUNK, n)], out=out[:min(CHUNK, n)])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.result()
    if not np.isfinite(float(out.double().abs().sum().item())):
        raise SystemExit("bench: encoder produced a non-finite result")
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    e2e = None
    if not args.no_e2e:
        host = torch.empty((n, *src_hw, 3), dtype=torch.uint8).pin_memory()
        host.copy_(frames)
        hn = host.numpy()
        pipe = StreamedEncoder(enc, src_hw, CHUNK, depth=2)
        emb = np.empty((n, enc.hidden_size), np.float16)
        pos = [0]

        def sink(e):
            emb[pos[0]:pos[0] + len(e)] = e
            pos[0] += len(e)

        def chunks():
            for s in range(0, n, CHUNK):
                yield hn[s:min(s + CHUNK, n)]

        pipe.run((hn[s:min(s + CHUNK, n)] for s in range(0, min(n, CHUNK), CHUNK)), lambda e: None)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        t0 = time.perf_counter()
        pipe.run(chunks(), sink)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        t2 = torch.tensor([wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": CLIP_FRAMES / float(t2.item()), "unit": "frames/s", "h2d_bytes_per_step": int(pipe.h2d_bytes),
               "d2h_bytes_per_step": int(pipe.d2h_bytes), "seconds": float(t2.item()),
               "api": "StreamedEncoder.run over this rank's span: pinned host frames -> H2D -> encode -> D2H -> f16 rows; "
                      "host wall clock of one pass, max over ranks (bytes are this rank's)"}
    if rank == 0:
        print(json.dumps({
            "metric": f"frames_per_sec_encoded_{'dinov3_' if not args.arch.startswith('dinov2') else ''}{args.arch.replace('-', '_')}_{SIDE}px",
            "value": CLIP_FRAMES * K / (ms_max / 1000.0), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"ONE synthetic 10-min 30fps {src_hw[0]}x{src_hw[1]} uint8 clip ({CLIP_FRAMES} frames, BASELINE "
                                   f"configs[1]) split into {world} contiguous spans, {CHUNK}-frame chunks, DINOv3 {args.arch} {SIDE}px",
                       "frames_per_rank": [len(r) for r in parallel.split_frame_range(CLIP_FRAMES, world)],
                       "parallelism": f"frame spans over {world} GPU(s), no collective (the head would read +-15 rows of "
                                      "context across each cut from the merged file)",
                       "l2": "every chunk of a pass is a distinct 96 MiB input; >1 GiB of activations per chunk"},
            "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e_encode_file(enc, clip_dev, n_chunks, W, K, world, dev, dist):
    """The call a user makes: cbas_b200.cbas.encode_file(encoder, path).  The clip is written to a .npy file first
    (outside the timed region); the timed call then reads PAGEABLE frames from the page cache, stages them through
    pinned memory, copies them to the device, encodes, copies the embeddings back and writes `<clip>_cls.h5`."""
    import shutil
    import tempfile
    from cbas_b200 import cbas, gui_state, store
    frames_total = min(K, n_chunks) * CHUNK
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 3 * frames_total * clip_dev[0].numel() else None
    td = tempfile.mkdtemp(prefix="cbas_b200_bench_", dir=base)
    try:
        path = os.path.join(td, f"clip_rank{os.environ.get('RANK', '0')}.npy")
        arr = np.lib.format.open_memmap(path, mode="w+", dtype=np.uint8, shape=(frames_total, *clip_dev.shape[1:]))
        for c in range(frames_total // CHUNK):
            arr[c * CHUNK:(c + 1) * CHUNK] = clip_dev[c * CHUNK:(c + 1) * CHUNK].cpu().numpy()
        arr.flush()
        del arr
        warm = os.path.join(td, "warm.npy")
        np.save(warm, clip_dev[:min(W, n_chunks) * CHUNK].cpu().numpy())
        saved_proj, gui_state.proj = gui_state.proj, None
        with contextlib.redirect_stdout(sys.stderr):
            cbas.encode_file(enc, warm)  # pipeline creation, first launches, file-system warm-up
            pipe = next(iter(enc.__dict__["_pipelines"].values()))
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            pipe.h2d_bytes = pipe.d2h_bytes = 0
            t0 = time.perf_counter()
            out = cbas.encode_file(enc, path)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
        gui_state.proj = saved_proj
        with store.EmbeddingReader(out) as r:
            assert r.shape[0] == frames_total, "encode_file wrote a short file"
        t2 = torch.tensor([wall], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        steps = frames_total // CHUNK
        return {"value": world * frames_total / float(t2.item()), "unit": "frames/s",
                "h2d_bytes_per_step": pipe.h2d_bytes // steps, "d2h_bytes_per_step": pipe.d2h_bytes // steps,
                "steps": steps, "seconds": float(t2.item()),
                "api": "cbas_b200.cbas.encode_file(encoder, '<clip>.npy') -> '<clip>_cls.h5': pageable frames from the page "
                       "cache -> pinned staging -> H2D -> encode -> D2H -> f16 HDF5 append+flush per chunk -> os.replace; "
                       "host wall clock, max over ranks",
                "store_backend": store.backend_name() if hasattr(store, "backend_name") else "native"}
    finally:
        shutil.rmtree(td, ignore_errors=True)


def run_other_workloads(args):
    """BASELINE configs[3] and configs[4] in short form, each in its own process (`--workload head|backlog` prints the
    full line): so the driver's default run records them too."""
    import subprocess
    out = {}
    for name, extra in (("head_1M_frames", ["--workload", "head", "--steps", "5", "--no-cpu-baseline"]),
                        ("backlog_1gpu_20min", ["--workload", "backlog", "--hours", "0.3334"])):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), *extra], capture_output=True, text=True, timeout=600)
            j = json.loads(r.stdout.strip().splitlines()[-1])
            keep = {k: j[k] for k in ("metric", "value", "unit", "ms_per_step", "gpu_launches", "clocks", "e2e") if k in j}
            keep["workload"] = j["config"]["workload"]
            if "roofline" in j:
                keep["roofline"] = {k: j["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic") if k in j["roofline"]}
            for k in ("gpu_eager_baseline", "realtime_factor", "stages"):
                if k in j:
                    keep[k] = j[k]
            out[name] = keep
        except Exception as exc:
            out[name] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    return out


if __name__ == "__main__":
    main()
