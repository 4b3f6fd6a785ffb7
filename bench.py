"""Benchmark of the diarization hot path: diarized audio-hours / second (embed + NME-SC), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload meeting_1h|telephonic_10min|...] [--shard]

One "step" is one pass of the hot path over one recording per GPU: waveform -> multi-scale windows -> fused
log-mel + TitaNet-L embeddings -> multi-scale affinity -> NME-SC -> spectral clustering -> labels on the host.
N = 1 runs BASELINE config #3 (1-hour 8-speaker meeting, diar_infer_meeting.yaml), the configuration the headline
target is quoted on; N > 1 gives every rank its own recording of the same shape (multi-file batches shard by
recording, no data-path collective: weak scaling) and adds two objects for BASELINE config #5, ONE 4-hour recording over
the N GPUs: `strong_4h` at the shipped knobs (embedding launch groups dealt to the ranks + NCCL all-gather, long-form
chunks dealt to the ranks) and `strong_4h_fullmatrix` with embeddings_per_chunk raised above the recording's length, i.e.
the N x N affinity, its graph and the eigensolver's products ROW-SHARDED (whisper_nemo_b200/rowshard.py: peer stores over
NVLink from the GEMM epilogue, device-side barriers).
`--shard` makes that single-recording sharding the main measurement of any workload (scaling "strong").
Audio is synthetic (tools/workload.py) and the TitaNet-L weights are the fixed-seed random init
(whisper_nemo_b200.checkpoint.seeded); speech regions come from the ground-truth RTTM (oracle VAD) -- VAD, WAV decoding
and RTTM text are outside the metric.

`value`         device-timed (CUDA events) with the waveform already resident in HBM.
`e2e`           the same call with the waveform in pinned HOST memory: H2D copy of the samples, all device work, and the
                D2H read of the labels inside the timed region.
`diarize_call`  wall clock of the reference-facing `ClusteringDiarizer(cfg).diarize()` itself: WAV read, manifests,
                windows, H2D, device work, D2H and the RTTM / label files (what a user of the reference sees).
`--impl reference` times the CPU restatement of NeMo's ClusteringDiarizer (oracle/, NeMo itself is not installable
here) on the box's host cores on the SAME configuration: one full pass over the recording, its dataloader batches dealt
round-robin to the W + K steps (each step a bounded sample of the workload), the clustering of the full recording
timed once after the last step.  That arm imports neither the product's engine nor libb200d.so.
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
METRIC = "diarized audio-hours/sec (embed+NME-SC, device-timed)"
UNIT = "audio-hours/s"
SEED = 100  # recording seed of rank 0 (rank r: SEED + r); the reference arm diarizes rank 0's recording


_T0 = time.perf_counter()


def _progress(msg):
    """Timestamped phase marker on stderr (the JSON line on stdout stays the only stdout output)."""
    sys.stderr.write(f"[bench r{os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:6.1f}s] {msg}\n")
    sys.stderr.flush()


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, written by
    `tools/ncu_summarize.py traffic` from the committed `ncu --set full` capture (profiles/ncu_traffic.json carries the
    capture's file name and the commit it was taken at); None when no capture has been summarised."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def workload_config(args, world):
    """The `config` object of the JSON line -- a pure function of the command line, identical for both arms."""
    if args.workload == BATCH_WORKLOAD:
        text = (f"{args.workload} (BASELINE.json configs[3]): batch of {args.batch} x 600 s synthetic 3-speaker recordings in one manifest, "
                f"diar_infer_general.yaml, oracle VAD, TitaNet-L random-init seed 1234")
        sharding = f"recordings dealt to {world} rank(s), no collective"
    else:
        domain, seconds, speakers, idx = WORKLOADS[args.workload]
        text = (f"{args.workload} (BASELINE.json configs[{idx}]): {seconds:.0f} s synthetic {speakers}-speaker 16 kHz recording"
                f"{'' if args.shard else ' per GPU'}, diar_infer_{domain}.yaml, oracle VAD, TitaNet-L random-init seed 1234")
        sharding = ("single GPU" if world == 1 else
                    f"ONE recording over {world} GPUs: windows, long-form chunks and spectral products sharded, NCCL all-gathers" if args.shard else
                    "one recording per GPU, no collective")
    return {"workload": text, "sharding": sharding,
            "l2": "inputs larger than L2 (waveform 230 MB/h, activations > 1 GB per step); no explicit flush"}


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
    from tools import workload as wl

    if workload == BATCH_WORKLOAD:
        from whisper_nemo_b200 import sharding

        mine = sharding.assign_recordings([600.0] * batch, world)[rank]
        entries = []
        for i in mine:
            wav_path, rttm_path, _, _ = wl.make_session(work_dir, f"rec{i:03d}", 600.0, 3, seed=1000 + i)
            entries.append({"audio_filepath": wav_path, "rttm_filepath": rttm_path})
        cfg = wl.load_domain_config("general")
        man = os.path.join(work_dir, "manifest.json")
        wl.write_manifest(man, entries)
        cfg.diarizer.manifest_filepath, cfg.diarizer.out_dir, cfg.diarizer.oracle_vad = man, work_dir, True
        return cfg, 600.0 * len(mine)
    domain, seconds, speakers, _ = WORKLOADS[workload]
    cfg, wav, turns = wl.make_session_cfg(work_dir, domain, seconds, speakers, seed)
    return cfg, seconds


def _shared_session(workload, seed, rank, barrier):
    """ONE recording for all ranks (single-recording sharding): rank 0 synthesises it into a directory every rank of
    the box can read, the others wait.  Returns (cfg with a rank-private out_dir, seconds)."""
    from tools import workload as wl

    domain, seconds, speakers, _ = WORKLOADS[workload]
    shared = os.path.join(tempfile.gettempdir(), f"b200d_bench_shared_{workload}_s{seed}")
    if rank == 0 and not os.path.exists(os.path.join(shared, "input_manifest.json")):
        wl.make_session_cfg(shared, domain, seconds, speakers, seed)
    barrier()
    cfg = wl.load_domain_config(domain)
    out = os.path.join(tempfile.gettempdir(), f"b200d_bench_shared_{workload}_s{seed}_out_r{rank}")
    os.makedirs(out, exist_ok=True)
    cfg.diarizer.manifest_filepath = os.path.join(shared, "input_manifest.json")
    cfg.diarizer.out_dir, cfg.diarizer.oracle_vad = out, True
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


def _time_steps(diar, wav_dev, steps, barrier, torch):
    """(device ms over `steps` passes with the waveform resident, wall ms over `steps` passes from pinned host memory)"""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        diar.run_device(wav_dev=wav_dev, timers=False)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        diar.run_device(wav_dev=None, timers=False)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    return dev_ms, e2e_ms


def run_b200(args):
    # stdout carries ONE JSON line: anything a library prints there (NCCL's version banner under NCCL_DEBUG=VERSION, ...) goes to
    # stderr instead; the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    import torch

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from whisper_nemo_b200 import ClusteringDiarizer, _cabi, checkpoint

    weights = checkpoint.seeded()
    _progress("weights ready")
    # the bounded CPU sample runs BEFORE the process group exists: the other ranks block in the rendezvous (no NCCL barrier
    # spinning on the host cores next to it, as round 1 had) and torchrun's OMP_NUM_THREADS=1 is overridden explicitly
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and not args.profiling:
        cpu = cpu_baseline_sample(args, weights)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
        # waits that can last seconds (another rank synthesising audio, rank 0 running a recording alone) go through a gloo
        # group: its barrier sleeps on a socket, while an NCCL barrier spins on a host core -- and on a box whose ranks share
        # one CPU allowance the spinning ranks starved the working one (round 1's wrong N > 1 cpu_baseline; a 4-hour synthesis
        # that took 30 s alone took > 4 min next to one spinning rank)
        host_group = dist.new_group(backend="gloo")
        _progress("process group up")

    def barrier():
        if world > 1:
            dist.barrier(group=host_group)
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(vals):
        if world == 1:
            return vals
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    workload = args.workload
    batch_mode = workload == BATCH_WORKLOAD
    shard = bool(args.shard) and world > 1 and not batch_mode
    if shard:
        cfg, seconds = _shared_session(workload, SEED, rank, barrier)
    else:
        work_dir = os.path.join(tempfile.gettempdir(), f"b200d_bench_{workload}_r{rank}")
        cfg, seconds = _session(work_dir, workload, seed=SEED + rank, batch=args.batch, rank=rank, world=world)
    t0 = time.perf_counter()
    diar = ClusteringDiarizer(cfg=cfg, speaker_model=weights, shard_windows=shard).to("cuda")
    diar._prepare()
    prep_s = time.perf_counter() - t0
    wav_dev = diar._wav_host.to(dev)
    flops, frames, windows = titanet_flops(diar)
    _progress(f"recording prepared ({windows} windows)")

    for _ in range(args.warmup):
        diar.run_device(wav_dev=wav_dev, timers=False)
    # ---- timed: inputs resident in HBM (waveform 230 MB/h + GBs of activations per step: larger than the 126 MB L2)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _cabi.total_launches()
    dev_ms, e2e_ms = _time_steps(diar, wav_dev, args.steps, barrier, torch)
    launches = (_cabi.total_launches() - launches0) // 2  # the two timed loops run the same launches
    clocks = sampler.stop() if rank == 0 else None
    dev_ms, e2e_ms = all_max([dev_ms, e2e_ms])
    labels_main = {u: r["labels"].copy() for u, r in diar.results.items()}
    _progress(f"timed steps done: {dev_ms / args.steps:.1f} ms per step")
    # ---- wall clock of the reference-facing call itself (prepare + device path + output files), same recording
    call_s = []
    for _ in range(0 if args.profiling else 3):
        barrier()
        t0 = time.perf_counter()
        diar.diarize()
        call_s.append(time.perf_counter() - t0)
    call_ms = all_max([1e3 * sum(call_s) / max(len(call_s), 1)])[0]
    host_split = dict(getattr(diar, "host_seconds", {}))
    # ---- N = 1: the same recording on the FULL-MATRIX path ("final solve 14 399^2 or long-form ... run and
    # report both"): embeddings_per_chunk raised above the recording's length, i.e. one N x N affinity instead of long-form chunks
    fullmatrix = None
    if world == 1 and not batch_mode and not args.profiling and seconds >= 3000:
        import copy

        cfg_f = copy.deepcopy(cfg)
        cfg_f.diarizer.clustering.parameters.embeddings_per_chunk = 10 ** 7
        diar_f = ClusteringDiarizer(cfg=cfg_f, speaker_model=weights).to("cuda")
        diar_f._prepare()
        for _ in range(2):
            diar_f.run_device(wav_dev=wav_dev, timers=False)
        f_dev_ms, f_e2e_ms = _time_steps(diar_f, wav_dev, 3, barrier, torch)
        diar_f.run_device(wav_dev=wav_dev, timers=True)
        res_f = next(iter(diar_f.results.values()))
        fullmatrix = {"what": "the same recording with embeddings_per_chunk raised above its length: one N x N affinity / graph / eigensolve instead of "
                              "long-form chunks (3 steps)", "ms_per_step": round(f_dev_ms / 3, 3), "value": round(seconds / 3600.0 * 3 / (f_dev_ms * 1e-3), 4),
                      "e2e_ms_per_step": round(f_e2e_ms / 3, 3), "unit": UNIT, "base_scale_windows": int(len(res_f["labels"])),
                      "speakers_found": int(res_f["debug"]["n_clusters"]), "p_hat": int(res_f["debug"]["p_hat"]),
                      "stage_ms": {k: round(v, 3) for k, v in diar_f.stage_ms.items()}}
        del diar_f
        torch.cuda.empty_cache()
        _progress(f"full-matrix variant done: {f_dev_ms / 3:.1f} ms per step")
    # ---- N > 1: BASELINE config #5, one 4-hour recording sharded over all ranks (strong scaling), next to the weak-scaling line
    _progress("diarize() calls done")
    strong = strong_full = None
    if world > 1 and not shard and not batch_mode and not args.no_strong and not args.profiling:
        # a failure of one of these side measurements (e.g. CUDA IPC refused on a box) must not take the main line with it: it is
        # recorded in the object instead (the product itself raises -- this guard is the benchmark's only)
        def guarded(**kw):
            try:
                return strong_single_recording(args, weights, rank, world, dev, barrier, all_max, torch, **kw)
            except Exception as exc:  # noqa: BLE001
                _progress(f"strong measurement failed: {exc!r}")
                return {"error": repr(exc)[:300]} if rank == 0 else None

        strong = guarded()
        strong_full = guarded(fullmatrix=True)
    # ---- untimed extras on rank 0: per-stage times, per-kernel roofline
    if args.profiling:
        if rank == 0:
            emit({"profiling_run": True, "ms_per_step": dev_ms / args.steps, "note": "not a bench value"})
    elif rank == 0 or shard:
        # (under --shard every rank has to take part in the collectives of these extra passes)
        diar.run_device(wav_dev=wav_dev, timers=False)  # the side measurements above released the allocator's cache: re-warm it
        diar.run_device(wav_dev=wav_dev, timers=True)
        stage_ms = dict(diar.stage_ms)
        from whisper_nemo_b200 import clustering as _cl

        _cl.spectral_log.clear()
        _cabi.start_profile()
        diar.run_device(wav_dev=wav_dev, timers=False)
        prof = _cabi.stop_profile()
        spectral = list(_cl.spectral_log)
    if rank == 0 and not args.profiling:
        g = {"calls": 0, "ms": 0.0, "work": 0.0}      # the dominant kernel: gemm_tcgen05_2cta_kernel (all its launches)
        g_all = {"calls": 0, "ms": 0.0, "work": 0.0}  # every tcgen05 GEMM launch, small projections included
        for key, v in prof.items():
            if key.startswith("b200d_gemm_f16") or key.startswith("b200d::gemm["):  # fine-grained calls / spans inside the composites
                for f in g_all:
                    g_all[f] += v[f]
                if "|2cta" in key:
                    for f in g:
                        g[f] += v[f]
        peaks = _peaks()
        gemm_tflops = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        traffic = _ncu_traffic()
        roofline = {
            "kernel": "gemm_tcgen05_2cta_kernel (CTA-pair tcgen05 GEMM: the TitaNet-L pointwise convs of the embedding stage)",
            "bound": "tensor", "achieved": round(gemm_tflops, 1), "peak": peaks["tflops"], "unit": "TFLOP/s",
            "frac": round(gemm_tflops / peaks["tflops"], 4), "traffic": traffic["bytes_per_launch"] if traffic else None,
            "traffic_detail": traffic, "peak_source": peaks["src"],
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
        total_hours = (600.0 * args.batch / 3600.0) if batch_mode else (hours if shard else world * hours)
        value = total_hours * args.steps / (dev_ms * 1e-3)
        e2e_val = total_hours * args.steps / (e2e_ms * 1e-3)
        res = next(iter(diar.results.values()))
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True, "scaling": "strong" if (batch_mode or shard) else "weak",
            "vs_baseline": None, "dtype": "fp16",
            "dtype_detail": "fp16 tensor-core GEMMs with fp32 accumulation (TitaNet-L), fp32 featurizer / affinity / eigen-solvers, fp64 Sturm + Jacobi",
            "data": "synthetic", "config": workload_config(args, world),
            "workload_stats": {"windows_per_recording": windows, "frames_per_recording": frames, "base_scale_windows": int(len(res["labels"])),
                               "speakers_found": int(res["debug"]["n_clusters"]), "p_hat": int(res["debug"]["p_hat"]),
                               "labels_stable_across_steps": all((labels_main[u] == r["labels"]).all() for u, r in diar.results.items())},
            "e2e": {"value": round(e2e_val, 4), "unit": UNIT, "h2d_bytes_per_step": int(diar._wav_host.numel() * 4),
                    "d2h_bytes_per_step": int(sum(len(r["labels"]) for r in diar.results.values()) * 8), "ms_per_step": round(e2e_ms / args.steps, 3)},
            "diarize_call": {"value": round(total_hours / (call_ms * 1e-3), 4), "unit": UNIT, "ms_per_call": round(call_ms, 1),
                             "host_seconds_last_call": {k: round(v, 4) for k, v in host_split.items()},
                             "what": "wall clock of ClusteringDiarizer(cfg).diarize(): WAV read + manifests + windows + H2D + device path + D2H + RTTM/label files"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()}, "kernels_ms_per_step": kernels, "host_prepare_s": round(prep_s, 3),
            "spectral_solver": spectral,
        }
        if fullmatrix is not None:
            line["fullmatrix_path"] = fullmatrix
        if strong is not None:
            line["strong_4h"] = strong
        if strong_full is not None:
            line["strong_4h_fullmatrix"] = strong_full
        emit(line)
    if world > 1:
        dist.barrier(group=host_group)
        dist.destroy_process_group()


def strong_single_recording(args, weights, rank, world, dev, barrier, all_max, torch, workload="telephonic_4h", steps=3, warmup=2, fullmatrix=False):
    """BASELINE config #5 on the N GPUs of this run: ONE 4-hour 12-speaker recording (telephonic YAML).  The embedding launch
    groups of every scale are dealt to the ranks (NCCL all-gather of the embeddings); the clustering is sharded either by
    long-form chunk (shipped knobs: 6 chunks of 10 000 windows) or -- fullmatrix: embeddings_per_chunk raised above the
    recording's length -- by ROWS of the 51 360^2 affinity / graph / eigensolver products (peer stores over NVLink).  Rank 0
    also diarizes the recording alone; the labels must be bit-identical."""
    from whisper_nemo_b200 import ClusteringDiarizer

    _progress(f"strong_4h{'_fullmatrix' if fullmatrix else ''}: synthesising / waiting for the shared recording")
    cfg, seconds = _shared_session(workload, 7, rank, barrier)
    _progress("shared recording ready")
    if fullmatrix:
        cfg.diarizer.clustering.parameters.embeddings_per_chunk = 10 ** 7
    diar = ClusteringDiarizer(cfg=cfg, speaker_model=weights, shard_windows=True).to("cuda")
    diar._prepare()
    wav_dev = diar._wav_host.to(dev)
    for _ in range(warmup):
        diar.run_device(wav_dev=wav_dev, timers=False)
    dev_ms, e2e_ms = _time_steps(diar, wav_dev, steps, barrier, torch)
    dev_ms, e2e_ms = all_max([dev_ms, e2e_ms])
    diar.run_device(wav_dev=wav_dev, timers=True)
    stage_ms = dict(diar.stage_ms)
    sharded = next(iter(diar.results.values()))
    _progress(f"sharded runs done: {dev_ms / steps:.1f} ms per step; rank 0 now runs it alone")
    out = None
    if rank == 0:
        solo = ClusteringDiarizer(cfg=cfg, speaker_model=weights, shard_windows=False).to("cuda")
        solo._prepare()
        solo.run_device(wav_dev=wav_dev, timers=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        solo.run_device(wav_dev=wav_dev, timers=False)
        e1.record()
        torch.cuda.synchronize()
        solo_ms = e0.elapsed_time(e1)
        alone = next(iter(solo.results.values()))
        solo.run_device(wav_dev=wav_dev, timers=True)
        solo_stage = dict(solo.stage_ms)
        hours = seconds / 3600.0
        how = ("embeddings_per_chunk raised: full-matrix path, affinity / graph / eigensolver products row-sharded" if fullmatrix else
               "shipped knobs: long-form path, chunks dealt to the ranks")
        out = {"workload": f"{workload} (BASELINE.json configs[4]): ONE {seconds:.0f} s synthetic 12-speaker recording, diar_infer_telephonic.yaml, "
                           f"sharded over {world} GPUs ({how})", "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warmup,
               "value": round(hours * steps / (dev_ms * 1e-3), 4), "unit": UNIT, "ms_per_step": round(dev_ms / steps, 3),
               "e2e": {"value": round(hours * steps / (e2e_ms * 1e-3), 4), "ms_per_step": round(e2e_ms / steps, 3)},
               "one_gpu_ms_per_step": round(solo_ms, 3), "speedup_vs_one_gpu": round(solo_ms / (dev_ms / steps), 3),
               "labels_identical_to_one_gpu": bool(len(alone["labels"]) == len(sharded["labels"]) and (alone["labels"] == sharded["labels"]).all()),
               "base_scale_windows": int(len(sharded["labels"])), "speakers_found": int(sharded["debug"]["n_clusters"]),
               "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()}, "one_gpu_stage_ms": {k: round(v, 3) for k, v in solo_stage.items()}}
    del diar
    torch.cuda.empty_cache()
    barrier()
    _progress("strong measurement done")
    return out


# ------------------------------------------------------------------------------------------ CPU arm (oracle/)
def _oracle_model(weights):
    from oracle.titanet import TitaNetL

    model = TitaNetL(compute_logits=False)
    model.load_state_dict(weights, strict=False)
    return model.eval()


def _oracle_run(seconds, domain, speakers, seed, weights, threads):
    """One pass of the CPU restatement over `seconds` of synthetic audio; returns (audio-hours/s, stage seconds)."""
    import torch

    from oracle.clustering_diarizer import OracleClusteringDiarizer
    from tools.workload import make_session_cfg

    torch.set_num_threads(threads)
    model = _oracle_model(weights)
    with tempfile.TemporaryDirectory() as tmp:
        cfg, _, _ = make_session_cfg(tmp, domain, seconds, speakers, seed)
        d = OracleClusteringDiarizer(cfg, model)
        d.prepare()
        t0 = time.perf_counter()
        d.embed()
        d.cluster()
        dt = time.perf_counter() - t0
    return (seconds / 3600.0) / dt, dict(d.stage_seconds)


def _host_threads():
    """Cores this process may run on (the GPU box pins the job to a subset of its CPUs)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline_sample(args, weights, sample_s=120.0):
    domain, seconds, speakers, _ = WORKLOADS["general_10min" if args.workload == BATCH_WORKLOAD else args.workload]
    sample_s = min(sample_s, seconds)
    threads = _host_threads()
    v, stages = _oracle_run(sample_s, domain, speakers, SEED, weights, threads)
    return {"value": round(v, 6), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first-principles CPU restatement of NeMo ClusteringDiarizer (oracle/, NeMo not installable) on a {sample_s:.0f} s recording of the "
                      f"same synthetic {domain} workload, end to end, fp32, {threads} torch threads; stage seconds {({k: round(x, 2) for k, x in stages.items()})}; "
                      "clustering cost grows super-linearly with length, so the full-length CPU throughput (bench.py --impl reference) is lower"}


def run_reference(args):
    """The reference arm: NeMo's ClusteringDiarizer cannot be installed here, so this times its CPU
    restatement (oracle/) on the SAME workload as the B200 arm -- ONE full pass over rank 0's recording:
      * the dataloader batches (64 windows each, all scales) are dealt round-robin to W + K steps; every step embeds its
        share (a bounded, stratified sample of the workload); the W warm-up steps are real work of the pass, just untimed;
      * after the K timed steps the clustering of the FULL recording (long-form path included) runs once, timed.
    value = audio hours / (timed embedding seconds x (W + K) / K + clustering seconds): the full-length CPU throughput,
    with the untimed warm-up share of the (linear, batch-independent) embedding cost filled in at the timed steps' rate."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch

    from oracle.clustering_diarizer import OracleClusteringDiarizer
    from tools.workload import make_session_cfg
    from whisper_nemo_b200 import checkpoint  # pure-Python weight initialiser + fixture; loads no native code

    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    domain, seconds, speakers, idx = WORKLOADS["general_10min" if args.workload == BATCH_WORKLOAD else args.workload]
    threads = _host_threads()
    torch.set_num_threads(threads)
    model = _oracle_model(checkpoint.seeded())
    W, K = max(args.warmup, 0), max(args.steps, 1)
    with tempfile.TemporaryDirectory() as tmp:
        cfg, _, _ = make_session_cfg(tmp, domain, seconds, speakers, SEED)
        d = OracleClusteringDiarizer(cfg, model)
        d.prepare()
        plan = d.plan_batches()
        step_s = []
        for step in range(W + K):
            t0 = time.perf_counter()
            for item in plan[step::W + K]:
                d.embed_batch(item)
            step_s.append(time.perf_counter() - t0)
        d.finish_embed()
        t0 = time.perf_counter()
        d.cluster()
        cluster_s = time.perf_counter() - t0
        res = next(iter(d.results.values()))
    timed_embed_s = sum(step_s[W:])
    full_pass_s = timed_embed_s * (W + K) / K + cluster_s
    value = (seconds / 3600.0) / full_pass_s
    sample = (f"ONE full pass of the CPU restatement of NeMo ClusteringDiarizer (oracle/, fp32, {threads} torch threads) over the {seconds:.0f} s "
              f"recording: {len(plan)} dataloader batches of 64 windows dealt round-robin to {W} warm-up + {K} timed steps "
              f"({timed_embed_s:.1f} s timed embedding, {sum(step_s[:W]):.1f} s untimed), then the clustering of the full recording once "
              f"({cluster_s:.1f} s, {len(res['labels'])} base windows, {res['debug']['n_clusters']} speakers); "
              f"value = hours / (timed embedding x {(W + K) / K:.3f} + clustering) = hours / {full_pass_s:.1f} s")
    cpu = {"value": round(value, 6), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round((timed_embed_s + cluster_s) / K * 1e3, 1), "higher_is_better": True,
            "scaling": "strong" if (args.workload == BATCH_WORKLOAD or args.shard) else "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": workload_config(args, world),
            "steps_effective": 1, "full_pass_seconds": round(full_pass_s, 1),
            "stage_seconds": {"embed_timed": round(timed_embed_s, 2), "embed_untimed_warmup": round(sum(step_s[:W]), 2), "cluster_full": round(cluster_s, 2)},
            "cpu_baseline": cpu, "e2e": {"value": round(value, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    import faulthandler

    # a run that is still going after this many seconds dumps every thread's stack to stderr (and keeps going)
    faulthandler.dump_traceback_later(int(os.environ.get("B200D_BENCH_WATCHDOG_S", "900")), repeat=True, file=sys.stderr)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="meeting_1h", choices=sorted(WORKLOADS) + [BATCH_WORKLOAD])
    ap.add_argument("--batch", type=int, default=64, help="recordings in the batch workload")
    ap.add_argument("--shard", action="store_true", help="N > 1: ONE recording sharded over all ranks (strong scaling) instead of one recording per rank")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the extra strong_4h measurement")
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
