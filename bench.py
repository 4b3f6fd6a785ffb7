"""Benchmark of the diarization hot path: diarized audio-hours / second (embed + NME-SC), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload meeting_1h|telephonic_10min|...]

One "step" is one pass of the hot path over one recording per GPU: waveform -> multi-scale windows -> fused
log-mel + TitaNet-L embeddings -> multi-scale affinity -> NME-SC -> spectral clustering -> labels on the host.
N = 1 runs BASELINE config #3 (1-hour 8-speaker meeting, diar_infer_meeting.yaml), the configuration the headline
target is quoted on; N > 1 gives every rank its own recording of the same shape (multi-file batches shard by
recording, no data-path collective: weak scaling).  Audio is synthetic (whisper_nemo_b200.synth) and the
TitaNet-L weights are the fixed-seed random init (whisper_nemo_b200.checkpoint); speech regions come from the
ground-truth RTTM (oracle VAD) -- VAD, WAV decoding and RTTM text are outside the metric.

`value`   device-timed (CUDA events) with the waveform already resident in HBM.
`e2e`     the same call with the waveform in pinned HOST memory: H2D copy of the samples, all device work, and the
          D2H read of the labels inside the timed region.
`--impl reference` times the CPU restatement of NeMo's ClusteringDiarizer (oracle/, NeMo itself is not
installable here) on the box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (domain yaml, seconds, speakers, BASELINE.json config index)
    "clip_22s": ("telephonic", 22.58, 2, 0),
    "telephonic_10min": ("telephonic", 600.0, 2, 1),
    "meeting_1h": ("meeting", 3600.0, 8, 2),
    "general_10min": ("general", 600.0, 3, 3),
    "telephonic_4h": ("telephonic", 14400.0, 12, 4),
}
# BASELINE config #4: a batch of 10-minute recordings in ONE manifest, dealt to the ranks by sharding.assign_recordings
# (strong scaling over the batch).  --batch R sets the batch size (64 in BASELINE.json; smaller values keep the host-side
# synthesis short).
BATCH_WORKLOAD = "general_10min_batch"
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed ncu --set full capture of
# this same command (profiles/r01_ncu_full_gemm2cta_final.txt): gemm_tcgen05_2cta_kernel<bias>, M = 131 072 frames, N = K = 1024:
# 270.3 MB read + 231.1 MB written against 268 MB (A) + 2 MB (W) + 268 MB (out) algorithmic; 86.7 % tensor-pipe active, 192.5 us
NCU_GEMM_TRAFFIC = {"bytes_per_launch": 501163520, "launch": "M=131072 N=1024 K=1024 bias epilogue", "algorithmic_bytes": 538968064,
                    "tensor_pipe_active_pct": 85.4, "source": "profiles/r01_ncu_full_gemm2cta_r1end.txt"}
METRIC = "diarized audio-hours/sec (embed+NME-SC, device-timed)"
UNIT = "audio-hours/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (the recipe's clocks line)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


def _session(work_dir, workload, seed, batch=0, rank=0, world=1):
    """Synthetic recording(s) + manifest + config for one rank.  Returns (cfg, seconds of audio this rank diarizes)."""
    from tests.util import make_session_cfg

    if workload == BATCH_WORKLOAD:
        from whisper_nemo_b200 import config, sharding, synth

        mine = sharding.assign_recordings([600.0] * batch, world)[rank]
        entries = []
        for i in mine:
            wav_path, rttm_path, _, _ = synth.make_session(work_dir, f"rec{i:03d}", 600.0, 3, seed=1000 + i)
            entries.append({"audio_filepath": wav_path, "rttm_filepath": rttm_path})
        cfg = config.load_config("general")
        man = os.path.join(work_dir, "manifest.json")
        synth.write_manifest(man, entries)
        cfg.diarizer.manifest_filepath, cfg.diarizer.out_dir, cfg.diarizer.oracle_vad = man, work_dir, True
        return cfg, 600.0 * len(mine)
    domain, seconds, speakers, _ = WORKLOADS[workload]
    cfg, wav, turns = make_session_cfg(work_dir, domain, seconds, speakers, seed)
    return cfg, seconds


def titanet_flops(diar):
    """Algorithmic FLOPs of the TitaNet-L forward for the planned windows: 2 x (17.49 M MAC per frame + 4.59 M MAC per
    window) (pointwise 15.81 M + depthwise 0.10 M + decoder 1.57 M per frame; SE + projections per window)."""
    from whisper_nemo_b200.titanet import frames_of

    frames = windows = 0
    for plan in diar._scales.values():
        windows += len(plan["len"])
        frames += int(sum(frames_of(int(f)) for f in plan["fixed"]))
    return 2.0 * (17.49e6 * frames + 4.59e6 * windows), frames, windows


def run_b200(args):
    import torch

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from whisper_nemo_b200 import ClusteringDiarizer, _cabi, checkpoint

    workload = args.workload
    work_dir = os.path.join(tempfile.gettempdir(), f"b200d_bench_{workload}_r{rank}")
    cfg, seconds = _session(work_dir, workload, seed=100 + rank, batch=args.batch, rank=rank, world=world)
    batch_mode = workload == BATCH_WORKLOAD
    weights = checkpoint.calibrated(dev)
    t0 = time.perf_counter()
    diar = ClusteringDiarizer(cfg=cfg, speaker_model=weights).to("cuda")
    diar._prepare()
    prep_s = time.perf_counter() - t0
    wav_dev = diar._wav_host.to(dev)
    flops, frames, windows = titanet_flops(diar)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        diar.run_device(wav_dev=wav_dev, timers=False)
    # ---- timed: inputs resident in HBM (waveform 230 MB/h + GBs of activations per step: larger than the 126 MB L2)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = _cabi.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        labels = diar.run_device(wav_dev=wav_dev, timers=False)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = _cabi.launch_count - launches0
    # ---- timed: end to end from pinned host memory (H2D of the samples + device work + D2H of the labels)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        labels = diar.run_device(wav_dev=None, timers=False)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()
    # ---- untimed extras on rank 0: per-stage times, per-kernel roofline, CPU baseline sample
    if rank == 0 and args.profiling:
        print(json.dumps({"profiling_run": True, "ms_per_step": dev_ms / args.steps, "note": "not a bench value"}))
    elif rank == 0:
        diar.run_device(wav_dev=wav_dev, timers=True)
        stage_ms = dict(diar.stage_ms)
        from whisper_nemo_b200 import clustering as _cl

        _cl.spectral_log.clear()
        _cabi.start_profile()
        diar.run_device(wav_dev=wav_dev, timers=False)
        prof = _cabi.stop_profile()
        spectral = list(_cl.spectral_log)
        g = {"calls": 0, "ms": 0.0, "work": 0.0}      # the dominant kernel: gemm_tcgen05_2cta_kernel (all its launches)
        g_all = {"calls": 0, "ms": 0.0, "work": 0.0}  # every tcgen05 GEMM launch, small projections included
        for key, v in prof.items():
            if key.startswith("b200d_gemm_f16"):
                for f in g_all:
                    g_all[f] += v[f]
                if "|2cta" in key:
                    for f in g:
                        g[f] += v[f]
        peaks = _peaks()
        gemm_tflops = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        roofline = {
            "kernel": "gemm_tcgen05_2cta_kernel (CTA-pair tcgen05 GEMM: the TitaNet-L pointwise convs of the embedding stage)",
            "bound": "tensor", "achieved": round(gemm_tflops, 1), "peak": peaks["tflops"], "unit": "TFLOP/s",
            "frac": round(gemm_tflops / peaks["tflops"], 4), "traffic": NCU_GEMM_TRAFFIC, "peak_source": peaks["src"],
            "launches_per_step": g["calls"], "avg_launch_ms": round(g["ms"] / max(g["calls"], 1), 4),
            "share_of_step": round(g["ms"] / (dev_ms / args.steps), 3),
            "all_gemm_launches": {"calls": g_all["calls"], "ms": round(g_all["ms"], 3),
                                  "tflops": round(g_all["work"] / (g_all["ms"] * 1e-3) / 1e12, 1) if g_all["ms"] > 0 else 0.0},
            "step_titanet_tflops": round(flops / (stage_ms.get("embed", float("nan")) * 1e-3) / 1e12, 1),
        }
        kernels = {k: ({"calls": v["calls"], "ms": round(v["ms"], 3)} if not v["work"] else
                       {"calls": v["calls"], "ms": round(v["ms"], 3), "tflops": round(v["work"] / (v["ms"] * 1e-3) / 1e12, 1)})
                   for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        hours = seconds / 3600.0
        total_hours = (600.0 * args.batch / 3600.0) if batch_mode else world * hours
        value = total_hours * args.steps / (dev_ms * 1e-3)
        e2e_val = total_hours * args.steps / (e2e_ms * 1e-3)
        res = next(iter(diar.results.values()))
        cpu = cpu_baseline_sample(args, weights) if not args.no_cpu_baseline else None
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True, "scaling": "strong" if batch_mode else "weak", "vs_baseline": None,
            "dtype": "fp16", "dtype_detail": "fp16 tensor-core GEMMs with fp32 accumulation (TitaNet-L), fp32 featurizer / affinity / eigen-solvers, fp64 Sturm + Jacobi",
            "data": "synthetic",
            "config": {"workload": (f"{workload} (BASELINE.json configs[3]): batch of {args.batch} x 600 s synthetic 3-speaker recordings in one "
                                    f"manifest, diar_infer_general.yaml, dealt to {world} rank(s) by recording, oracle VAD, TitaNet-L random-init seed 1234")
                       if batch_mode else
                       (f"{workload} (BASELINE.json configs[{WORKLOADS[workload][3]}]): {seconds:.0f} s synthetic "
                        f"{WORKLOADS[workload][2]}-speaker 16 kHz recording per GPU, diar_infer_{WORKLOADS[workload][0]}.yaml, oracle VAD, "
                        "TitaNet-L random-init seed 1234"), "windows_per_recording": windows, "frames_per_recording": frames,
                       "base_scale_windows": int(len(res["labels"])), "speakers_found": int(res["debug"]["n_clusters"]),
                       "p_hat": int(res["debug"]["p_hat"]), "sharding": "one recording per GPU, no collective" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (waveform 230 MB/h, activations > 1 GB per step); no explicit flush"},
            "e2e": {"value": round(e2e_val, 4), "unit": UNIT, "h2d_bytes_per_step": int(diar._wav_host.numel() * 4),
                    "d2h_bytes_per_step": int(len(res["labels"]) * 8), "ms_per_step": round(e2e_ms / args.steps, 3)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()}, "kernels_ms_per_step": kernels, "host_prepare_s": round(prep_s, 3),
            "spectral_solver": spectral,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _oracle_run(seconds, domain, speakers, seed, weights, threads):
    """One pass of the CPU restatement over `seconds` of synthetic audio; returns (audio-hours/s, stage seconds)."""
    import torch

    from oracle.clustering_diarizer import OracleClusteringDiarizer
    from oracle.titanet import TitaNetL
    from tests.util import make_session_cfg

    torch.set_num_threads(threads)
    model = TitaNetL(compute_logits=False)
    model.load_state_dict(weights, strict=False)
    model.eval()
    with tempfile.TemporaryDirectory() as tmp:
        cfg, _, _ = make_session_cfg(tmp, domain, seconds, speakers, seed)
        d = OracleClusteringDiarizer(cfg, model)
        d.prepare()
        t0 = time.perf_counter()
        d.embed()
        d.cluster()
        dt = time.perf_counter() - t0
    return (seconds / 3600.0) / dt, dict(d.stage_seconds)


def _reference_weights():
    import torch

    if torch.cuda.is_available():
        from whisper_nemo_b200 import checkpoint

        return checkpoint.calibrated(torch.device("cuda", 0))
    from oracle.titanet import seeded_state_dict

    return seeded_state_dict(1234, compute_logits=False).state_dict()


def cpu_baseline_sample(args, weights, sample_s=150.0):
    domain, seconds, speakers, _ = WORKLOADS["general_10min" if args.workload == BATCH_WORKLOAD else args.workload]
    sample_s = min(sample_s, seconds)
    threads = os.cpu_count() or 1
    v, stages = _oracle_run(sample_s, domain, speakers, 100, weights, threads)
    return {"value": round(v, 6), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first-principles CPU restatement of NeMo ClusteringDiarizer (oracle/, NeMo not installable) on {sample_s:.0f} s of the same "
                      f"synthetic {domain} workload, fp32, {threads} torch threads; stage seconds {({k: round(x, 2) for k, x in stages.items()})}; "
                      "clustering cost grows super-linearly with length, so the full-length CPU throughput is lower than this sample's"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    domain, seconds, speakers, idx = WORKLOADS["general_10min" if args.workload == BATCH_WORKLOAD else args.workload]
    total = args.steps + args.warmup
    sample_s = min(seconds, 120.0 if total <= 8 else 60.0)
    threads = os.cpu_count() or 1
    weights = _reference_weights()
    for _ in range(args.warmup):
        _oracle_run(sample_s, domain, speakers, 100, weights, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, stages = _oracle_run(sample_s, domain, speakers, 100, weights, threads)
    dt = time.perf_counter() - t0
    value = args.steps * (sample_s / 3600.0) / dt
    cpu = {"value": round(value, 6), "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{sample_s:.0f} s of the synthetic {domain} workload per step (bounded sample of the {seconds:.0f} s recording), CPU restatement of "
                     f"NeMo ClusteringDiarizer (oracle/), fp32, {threads} torch threads"}
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"{args.workload} (BASELINE.json configs[{idx}]), bounded sample: {sample_s:.0f} s per step"},
            "cpu_baseline": cpu, "e2e": {"value": round(value, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="meeting_1h", choices=sorted(WORKLOADS) + [BATCH_WORKLOAD])
    ap.add_argument("--batch", type=int, default=64, help="recordings in the batch workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profiling", action="store_true", help="allow < 3 warm-up steps and skip the extras (only for runs under ncu; never a bench value)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200" and not args.profiling:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
