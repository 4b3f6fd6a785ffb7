"""`ClusteringDiarizer`: the drop-in for the NeMo class the reference runs.

The reference calls `NeuralDiarizer(cfg=create_config(dir)).to(device).diarize()`
(diarize.py:200-201, nemo_process.py:31-32); NeMo's NeuralDiarizer builds a
`ClusteringDiarizer(cfg, speaker_model)` and runs its `diarize()` before the MSDD decoder.
This class keeps that constructor / `.to()` / `.diarize(paths2audio_files=None, batch_size=0)`
surface, the unchanged `diar_infer_*.yaml` schema (nemo_msdd_configs/*.yaml:7-56 after the
overrides of helpers.py:282-301) and the manifest-in / RTTM-out contract, and replaces
everything in between with the B200 path:

  WAV once -> HBM  ->  featurize + TitaNet-L over every window of every scale (titanet.py)
               ->  multi-scale affinity + NME-SC + spectral clustering (clustering.py / longform.py)
               ->  labels -> out_dir/pred_rttms/<stem>.rttm (+ the speaker_outputs/ files MSDD reads)

MarbleNet VAD is outside this path: speech regions come from `oracle_vad: True`
(+ `rttm_filepath` in the manifest) or `vad.external_vad_manifest`.  There is no CPU fallback.
"""
import json
import os
import pickle as pkl
import shutil
import time
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _cabi
from . import speaker_utils as su
from .config import as_config
from .titanet import TitaNetB200, frames_of


def _get(cfg, dotted, default=None):
    cur = cfg
    for k in dotted.split("."):
        if cur is None:
            return default
        cur = cur.get(k, None) if hasattr(cur, "get") else getattr(cur, k, None)
    return default if cur is None else cur


_PINNED: Dict[int, torch.Tensor] = {}


def _pinned_buffer(n: int) -> torch.Tensor:
    """float32 [n] view on a process-wide pinned staging buffer that only grows: cudaHostAlloc of 230 MB (one hour of
    audio) costs ~0.1 s, as much as the whole device path, and the reference builds a fresh diarizer object per recording."""
    have = _PINNED.get(0)
    if have is None or have.numel() < n:
        have = None
        _PINNED.pop(0, None)
        _PINNED[0] = have = torch.empty(max(n, 1), dtype=torch.float32).pin_memory()
    return have[:n]


class DeviceTimer:
    """Per-stage device time from CUDA events on the launching stream (no host sync until read)."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled
        self.spans = []

    def start(self, name):
        if not self.enabled:
            return None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.spans.append((name, e0, e1))
        return e1

    @staticmethod
    def stop(handle):
        if handle is not None:
            handle.record()

    def totals_ms(self) -> Dict[str, float]:
        torch.cuda.synchronize()
        out: Dict[str, float] = {}
        for name, e0, e1 in self.spans:
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out


class ClusteringDiarizer:
    def __init__(self, cfg, speaker_model=None, shard_windows: bool = False):
        """cfg: DiarConfig / dict / OmegaConf DictConfig with the diar_infer_*.yaml schema.
        shard_windows: under torch.distributed, split the windows of every scale across the ranks and all-gather
        the embeddings (one long recording on several GPUs); clustering then runs replicated on every rank.
        speaker_model: a TitaNetB200, a TitaNet-L state_dict, or None (-> `speaker_embeddings.model_path`:
        a checkpoint path, or the name `titanet_large`, which resolves to the fixed-seed random-init
        TitaNet-L of checkpoint.py -- there is no network for the NGC download)."""
        _cabi.require_device()
        self.cfg = as_config(cfg)
        self._cfg = self.cfg
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.sample_rate = int(_get(self.cfg, "sample_rate", 16000))
        if self.sample_rate != 16000:
            raise ValueError("TitaNet-L features are defined at 16 kHz")
        self.batch_size = int(_get(self.cfg, "batch_size", 64))
        self.verbose = bool(_get(self.cfg, "verbose", False))
        p = _get(self.cfg, "diarizer.speaker_embeddings.parameters")
        self.multiscale_args_dict = su.parse_scale_configs(_get(p, "window_length_in_sec"), _get(p, "shift_length_in_sec"),
                                                           _get(p, "multiscale_weights"))
        self._cluster_params = _get(self.cfg, "diarizer.clustering.parameters")
        self.shard_windows = bool(shard_windows)
        self._row_comm = None
        self._shard_loads: List[int] = []
        self._embed_streams: List[torch.cuda.Stream] = []
        self._cluster_streams: List[torch.cuda.Stream] = []
        self._speaker_model = self._init_speaker_model(speaker_model)
        self.keep_affinity = bool(int(os.environ.get("B200D_KEEP_AFFINITY", "0")))  # results[...]["fused_affinity"] (debug / parity tests)
        self._last_clusterers: Dict[str, object] = {}
        self.stage_ms: Dict[str, float] = {}
        self.results: Dict[str, dict] = {}
        self.embs_and_timestamps: Dict[str, dict] = {}

    # NeMo's `.to(device)` idiom (diarize.py:200): accepted; the path only exists on CUDA.
    def to(self, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("whisper_nemo_b200.ClusteringDiarizer runs on a B200 only (no CPU path)")
        return self

    def _init_speaker_model(self, speaker_model):
        if isinstance(speaker_model, TitaNetB200):
            return speaker_model
        if isinstance(speaker_model, dict):
            return TitaNetB200(speaker_model, self.device)
        if speaker_model is not None and hasattr(speaker_model, "state_dict"):
            return TitaNetB200(speaker_model.state_dict(), self.device)
        model_path = _get(self.cfg, "diarizer.speaker_embeddings.model_path")
        from . import checkpoint

        return TitaNetB200(checkpoint.resolve(model_path, self.device), self.device)

    # ------------------------------------------------------------------ untimed preparation
    def _prepare(self, paths2audio_files=None):
        cfg = self.cfg
        out_dir = _get(cfg, "diarizer.out_dir")
        if out_dir is None:
            raise ValueError("diarizer.out_dir is required")
        self._out_dir = out_dir
        self._speaker_dir = os.path.join(out_dir, "speaker_outputs")
        if os.path.exists(self._speaker_dir):
            shutil.rmtree(self._speaker_dir, ignore_errors=True)
        os.makedirs(self._speaker_dir)
        os.makedirs(os.path.join(out_dir, "pred_rttms"), exist_ok=True)
        manifest = _get(cfg, "diarizer.manifest_filepath")
        if paths2audio_files:
            manifest = os.path.join(out_dir, "paths2audio_filepath.json")
            with open(manifest, "w") as f:
                for path in paths2audio_files:
                    json.dump({"audio_filepath": path, "offset": 0, "duration": None, "label": "infer", "text": "-"}, f)
                    f.write("\n")
        if manifest is None:
            raise ValueError("diarizer.manifest_filepath is required")
        self.AUDIO_RTTM_MAP = su.audio_rttm_map(manifest)
        # every recording is decoded ONCE (memory-mapped) into one pinned host buffer (upstream re-opens the WAV per segment)
        wavs = {u: su.read_wav(m["audio_filepath"], self.sample_rate) for u, m in self.AUDIO_RTTM_MAP.items()}
        self._wav_offset, total = {}, 0
        for u, w in wavs.items():
            self._wav_offset[u] = (total, len(w))
            total += len(w)
        if total >= 2 ** 31:
            raise ValueError("more than 2^31 samples in one manifest; split the batch")
        self._wav_host = _pinned_buffer(total)
        host_np = self._wav_host.numpy()
        for u, w in wavs.items():
            o, n = self._wav_offset[u]
            host_np[o : o + n] = w
        del wavs
        durations = {u: n / self.sample_rate for u, (o, n) in self._wav_offset.items()}
        ext = _get(cfg, "diarizer.vad.external_vad_manifest")
        if _get(cfg, "diarizer.oracle_vad", False):
            speech_manifest = su.write_rttm2manifest(self.AUDIO_RTTM_MAP, os.path.join(self._speaker_dir, "oracle_vad_manifest.json"), durations)
        elif ext:
            speech_manifest = ext
        else:
            speech_manifest = self._run_vad(durations)
        regions = su.read_speech_regions(speech_manifest)
        self._plan_windows(regions, write_manifests=True)

    def _run_vad(self, durations):
        """`oracle_vad: False` without an external VAD manifest: upstream runs MarbleNet here.  A user-supplied callable
        (`vad_fn(uniq_id, waveform float32 numpy) -> [(start_s, end_s), ...]`, set as `diarizer.vad_fn` on this object) stands
        in for it; the MarbleNet network itself is outside the B200 path."""
        vad_fn = getattr(self, "vad_fn", None)
        if vad_fn is None:
            raise NotImplementedError(
                "MarbleNet VAD is outside the B200 hot path (embed + NME-SC): set diarizer.oracle_vad=True with rttm_filepath "
                "in the manifest, diarizer.vad.external_vad_manifest, or assign a callable to ClusteringDiarizer.vad_fn")
        path = os.path.join(self._speaker_dir, "vad_fn_manifest.json")
        host_np = self._wav_host.numpy()
        with open(path, "w") as out:
            for uniq_id, meta in self.AUDIO_RTTM_MAP.items():
                o, n = self._wav_offset[uniq_id]
                for a, b in su._merge_on_grid([[float(x), float(y)] for x, y in vad_fn(uniq_id, host_np[o : o + n])], 5):
                    a, b = max(a, 0.0), min(b, durations[uniq_id])
                    if b > a:
                        json.dump({"audio_filepath": meta["audio_filepath"], "offset": round(a, 5), "duration": round(b - a, 5), "label": "UNK",
                                   "uniq_id": uniq_id}, out)
                        out.write("\n")
        return path

    def _plan_windows(self, regions: List[dict], write_manifests: bool):
        """Window descriptors of every scale from the speech regions (vectorised `get_subsegments`), the once-per-recording
        log-mel stream plan (titanet.plan_mel_streams) and, per recording and scale, the row range of its windows in the
        scale's embedding matrix + the [start, end] stamps."""
        from .titanet import plan_mel_streams

        sr = self.sample_rate
        uniq_of = [r.get("uniq_id") or su.get_uniqname_from_filepath(r["audio_filepath"]) for r in regions]
        uniq_ids = list(self.AUDIO_RTTM_MAP.keys())
        uniq_index = {u: i for i, u in enumerate(uniq_ids)}
        reg_uniq = np.asarray([uniq_index[u] for u in uniq_of], dtype=np.int64)
        reg_off = np.asarray([self._wav_offset[u][0] for u in uniq_of], dtype=np.int64)
        reg_tot = np.asarray([self._wav_offset[u][1] for u in uniq_of], dtype=np.int64)
        offsets = np.asarray([r["offset"] for r in regions], dtype=np.float64)
        durs = np.asarray([r["duration"] for r in regions], dtype=np.float64)
        self._scales = {}
        for scale_idx, (window, shift) in self.multiscale_args_dict["scale_dict"].items():
            region, start_s, dur_s = su.subsegment_arrays(offsets, durs, window, shift)
            if write_manifests:
                su.write_subsegments_manifest(os.path.join(self._speaker_dir, f"subsegments_scale{scale_idx}.json"), regions, region, start_s, dur_s)
            s = (start_s * sr).astype(np.int64)  # int(offset * sr), as the dataset's slicing
            length = np.minimum((dur_s * sr).astype(np.int64), reg_tot[region] - s)
            if length.size and int(length.min()) < 1:
                bad = int(np.argmin(length))
                raise ValueError(f"empty window at {start_s[bad]} s in {uniq_ids[reg_uniq[region[bad]]]}")
            # the frame count each window is tiled up to: the max length inside its dataloader batch of `batch_size` windows
            # (label_models' fixed_seq collate)
            nb = -(-length.size // self.batch_size)
            padded = np.zeros(nb * self.batch_size, dtype=np.int64)
            padded[: length.size] = length
            fixed = np.repeat(padded.reshape(nb, self.batch_size).max(axis=1), self.batch_size)[: length.size]
            uniq_arr = reg_uniq[region]
            plan = {"n": int(length.size), "uniq_idx": uniq_arr, "start": reg_off[region] + s, "len": length, "fixed": fixed,
                    "t0": start_s, "t1": start_s + dur_s, "rows": {}, "stamps": {}}
            stamps32 = np.stack([plan["t0"], plan["t1"]], axis=1).astype(np.float32) if length.size else np.zeros((0, 2), np.float32)
            for u, ui in uniq_index.items():
                sel = np.nonzero(uniq_arr == ui)[0]
                if len(sel):
                    plan["rows"][u] = sel
                    plan["stamps"][u] = torch.from_numpy(stamps32[sel])  # float32, as upstream's torch.tensor(list of floats)
            self._scales[scale_idx] = plan
        plans = list(self._scales.values())
        cat = lambda key: np.concatenate([pl[key] for pl in plans]) if plans else np.zeros(0, np.int64)
        stream_start, stream_off, row0 = plan_mel_streams(cat("start"), cat("len"), cat("fixed"))
        pos = 0
        for pl in plans:
            pl["row0"] = row0[pos : pos + pl["n"]]
            pos += pl["n"]
        self._streams = (stream_start, stream_off)
        self._shard_loads = []

    # ------------------------------------------------------------------ device work
    def _extract_embeddings(self, plan: dict, wav_dev: torch.Tensor, logmel: torch.Tensor = None, gather: bool = True):
        """All windows of one scale -> float32 [n, 192] on device (manifest order).  `logmel`: the stream frames of
        `_mel_streams` (interior frames of every window are gathered from it instead of recomputed).

        shard_windows: the unit dealt to the ranks is the LAUNCH GROUP of the single-GPU run (the windows of one tiled-up
        length, in its order, cut at multiples of b200d_titanet_group_windows), so every window is computed by exactly the
        launches -- same neighbours, same position -- that compute it on one GPU and the all-gathered embeddings are bit for bit
        the single-GPU ones (a window's position in its launch decides the summation order of the SqueezeExcite column sums).
        gather=False returns this rank's block and a closure that all-gathers it: the caller embeds every scale first and
        gathers afterwards, so that only the TOTAL load of a rank has to balance, not its load within each scale."""
        n = plan["n"]
        out = torch.empty(n, 192, dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        fixed = plan["fixed"]
        rank, world = 0, 1
        if self.shard_windows:
            from . import sharding

            rank, world = sharding.rank_world()
        # per tiled-up length: the windows, those on a log-mel stream first (the featurizer handles the two kinds with different
        # kernels), cut into launch groups; index / descriptor arrays are uploaded once per plan
        key = (rank, world, str(self.device), self._speaker_model.max_frames)
        groups = plan.setdefault("_groups", {}).get(key)
        if groups is None:
            units = []  # (frames, fixed_len, windows on a stream, window indices)
            for fl in np.unique(fixed):
                idx = np.nonzero(fixed == fl)[0]
                on_stream = plan["row0"][idx] >= 0
                idx = np.concatenate([idx[on_stream], idx[~on_stream]])
                n_fast = int(on_stream.sum())
                step = len(idx) if world == 1 else max(1, self._speaker_model.group_windows(int(fl)))
                for c0 in range(0, len(idx), step):
                    part = idx[c0 : c0 + step]
                    units.append((len(part) * frames_of(int(fl)), int(fl), max(0, min(n_fast - c0, len(part))), part))
            owner = self._deal_launch_groups([u[0] for u in units], world)
            i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.int32))).to(self.device)
            mine, offs, pos = [], [0] * world, [None] * len(units)
            for j, (_, fl, n_fast, idx) in enumerate(units):
                r = owner[j]
                pos[j] = (r, offs[r])  # rows [offs, offs + len) of rank r's block
                offs[r] += len(idx)
                if r == rank:
                    mine.append((fl, n_fast, pos[j][1], len(idx), i32(plan["start"][idx]), i32(plan["len"][idx]), i32(plan["row0"][idx])))
            pad = max(offs)
            # where every window's embedding lies in the all-gathered [world * pad, 192] block
            where = np.empty(n, dtype=np.int64)
            for (r, o), (_, _, _, idx) in zip(pos, units):
                where[idx] = r * pad + o + np.arange(len(idx))
            groups = plan["_groups"][key] = (mine, offs[rank], pad, torch.from_numpy(where).to(self.device))
        mine, n_mine, pad, where = groups
        local = out if world == 1 else torch.zeros(pad, 192, dtype=torch.float32, device=self.device)
        for fl, n_fast, o, cnt, st, ln, r0 in mine:
            emb = self._speaker_model.embed_segments(wav_dev, st, ln, fl, logmel=logmel, seg_row0=r0 if logmel is not None else None,
                                                     n_on_stream=n_fast)
            local[o : o + cnt] = emb
        if world == 1:
            return out.index_select(0, where)

        def finish():
            gathered = torch.empty(world * pad, 192, dtype=torch.float32, device=self.device)
            torch.distributed.all_gather_into_tensor(gathered, local)
            return gathered.index_select(0, where)

        return finish() if gather else finish

    def _deal_launch_groups(self, frames: List[int], world: int) -> List[int]:
        """Owner rank of every launch group of one scale: largest first onto the least-loaded rank, the loads carried over the
        scales of the recording (self._shard_loads; every rank computes the same assignment)."""
        owner = [0] * len(frames)
        if world == 1:
            return owner
        loads = self._shard_loads
        if len(loads) != world:
            loads[:] = [0] * world
        for j in sorted(range(len(frames)), key=lambda i: (-frames[i], i)):
            r = min(range(world), key=lambda q: (loads[q], q))
            owner[j] = r
            loads[r] += frames[j]
        return owner

    def _mel_streams(self, wav_dev: torch.Tensor):
        """The recording's log-mel frames, each computed once (b200d_mel_stream); None when no window lies on a stream."""
        stream_start, stream_off = self._streams
        if stream_start.size == 0 or os.environ.get("B200D_NO_MEL_STREAMS") == "1":
            return None
        dev = getattr(self, "_streams_dev", None)
        if dev is None or dev[0].device != self.device or dev[2] is not self._streams:
            dev = self._streams_dev = (torch.from_numpy(stream_start).to(self.device), torch.from_numpy(stream_off).to(self.device), self._streams)
        return self._speaker_model.mel_stream(wav_dev, dev[0], dev[1], int(stream_off[-1]))

    def _embed_all_scales(self, wav_dev: torch.Tensor) -> Dict[int, torch.Tensor]:
        """Embeddings of every scale.  B200D_EMBED_STREAMS > 1 keeps several scales in flight on separate streams (own
        activation workspaces); off by default: the multi-stream region has to give up the CTA-pair GEMM (see
        _cabi.single_cta_gemms), which costs more than the overlap wins."""
        n_streams = min(len(self._scales), max(1, int(os.environ.get("B200D_EMBED_STREAMS", "1"))))
        logmel = self._mel_streams(wav_dev)
        if self.shard_windows:
            pending = {k: self._extract_embeddings(plan, wav_dev, logmel, gather=False) for k, plan in self._scales.items()}
            return {k: (v() if callable(v) else v) for k, v in pending.items()}
        if n_streams <= 1:
            return {k: self._extract_embeddings(plan, wav_dev, logmel) for k, plan in self._scales.items()}
        from concurrent.futures import ThreadPoolExecutor

        main = torch.cuda.current_stream()
        if len(self._embed_streams) < n_streams:
            self._embed_streams.extend(torch.cuda.Stream() for _ in range(n_streams - len(self._embed_streams)))
        streams = self._embed_streams[:n_streams]
        for st in streams:
            st.wait_stream(main)
        dev_index = self.device.index
        free = list(streams)

        def work(scale_idx, plan):
            torch.cuda.set_device(dev_index)
            st = free.pop()  # list.pop / append are atomic under the GIL; at most n_streams workers run
            try:
                with torch.cuda.stream(st), torch.no_grad(), _cabi.single_cta_gemms():
                    out = self._extract_embeddings(plan, wav_dev, logmel)
                    out.record_stream(main)
                    return scale_idx, out
            finally:
                free.append(st)

        main.synchronize()
        with _cabi.short_gil_switch(), ThreadPoolExecutor(max_workers=n_streams) as pool:
            results = dict(f.result() for f in [pool.submit(work, k, plan) for k, plan in self._scales.items()])
            for st in streams:
                st.synchronize()
        return results

    def _cluster_one(self, uniq_id: str, e: dict, chunk_streams: int = None):
        from .longform import LongFormSpeakerClustering

        clus = self._cluster_params
        if _get(clus, "oracle_num_speakers", False):
            num_speakers = self.AUDIO_RTTM_MAP[uniq_id].get("num_speakers", None)
            if num_speakers is None:
                raise ValueError("Provided option as oracle num of speakers but num_speakers in manifest is null")
        else:
            num_speakers = -1
        sc = LongFormSpeakerClustering(shard_chunks=self.shard_windows, chunk_streams=chunk_streams)
        sc.speaker_clustering.keep_affinity = self.keep_affinity
        if self.shard_windows:
            # a recording that takes the full-matrix path (<= embeddings_per_chunk windows): affinity, graph and the
            # eigensolver's products row-sharded over the ranks (rowshard.py); the long-form path deals its chunks instead
            from . import rowshard, sharding

            if sharding.is_distributed() and os.environ.get("B200D_NO_ROW_SHARDING") != "1":
                if self._row_comm is None:
                    self._row_comm = rowshard.DistComm()
                sc.speaker_clustering.row_comm = self._row_comm
        labels = sc.forward_infer(
            embeddings_in_scales=e["embeddings"],
            timestamps_in_scales=e["timestamps"],
            multiscale_segment_counts=e["multiscale_segment_counts"],
            multiscale_weights=e["multiscale_weights"],
            oracle_num_speakers=int(num_speakers),
            max_num_speakers=int(_get(clus, "max_num_speakers", 8)),
            max_rp_threshold=float(_get(clus, "max_rp_threshold", 0.25)),
            sparse_search_volume=int(_get(clus, "sparse_search_volume", 30)),
            chunk_cluster_count=_get(clus, "chunk_cluster_count", None),
            embeddings_per_chunk=_get(clus, "embeddings_per_chunk", None),
            scale_mapping=e.get("scale_mapping"),
        )
        return labels, sc

    def run_device(self, wav_dev: Optional[torch.Tensor] = None, timers: bool = True) -> Dict[str, np.ndarray]:
        """The timed region of the benchmark: waveform (host or already on device) -> labels on host."""
        timer = DeviceTimer(timers)
        h = timer.start("h2d")
        if wav_dev is None:
            wav_dev = self._wav_host.to(self.device, non_blocking=True)
        timer.stop(h)
        h = timer.start("embed")
        per_scale = {}
        scale_embs = self._embed_all_scales(wav_dev)
        for scale_idx, plan in self._scales.items():
            embs = scale_embs[scale_idx]
            e_by = {}
            for u, sel in plan["rows"].items():
                contiguous = len(sel) == int(sel[-1]) - int(sel[0]) + 1
                e_by[u] = embs[int(sel[0]) : int(sel[-1]) + 1] if contiguous else embs[torch.from_numpy(sel).to(self.device)]
            per_scale[scale_idx] = (e_by, plan["stamps"])
        timer.stop(h)
        self.multiscale_embeddings_and_timestamps = per_scale
        self.embs_and_timestamps = su.get_embs_and_timestamps(per_scale, self.multiscale_args_dict)
        # host work that only needs the window times runs here, while the GPU is still busy with the embeddings queued above
        from .clustering import get_argmin_mat

        for uniq_id, e in self.embs_and_timestamps.items():
            split = [int(x) for x in e["multiscale_segment_counts"].tolist()]
            e["scale_mapping"] = get_argmin_mat(list(torch.split(e["timestamps"], split, dim=0)))
        h = timer.start("cluster")
        todo = [u for u in self.AUDIO_RTTM_MAP.keys() if u in self.embs_and_timestamps]
        n_streams = min(len(todo), max(1, int(os.environ.get("B200D_RECORDING_STREAMS", "4"))))
        pending = {}
        if n_streams <= 1 or self.shard_windows:
            for uniq_id in todo:
                pending[uniq_id] = self._cluster_one(uniq_id, self.embs_and_timestamps[uniq_id])
        else:
            # recordings are independent and one recording's clustering is a chain of small kernels and host round trips:
            # several recordings in flight on separate streams (multi-stream region: CTA-pair GEMM off, see _cabi)
            from concurrent.futures import ThreadPoolExecutor

            main = torch.cuda.current_stream()
            if len(self._cluster_streams) < n_streams:
                self._cluster_streams.extend(torch.cuda.Stream() for _ in range(n_streams - len(self._cluster_streams)))
            free = list(self._cluster_streams[:n_streams])
            dev_index = self.device.index
            main.synchronize()

            def work(uniq_id):
                torch.cuda.set_device(dev_index)
                st = free.pop()
                try:
                    with torch.cuda.stream(st), torch.no_grad(), _cabi.single_cta_gemms():
                        labels, sc = self._cluster_one(uniq_id, self.embs_and_timestamps[uniq_id], chunk_streams=1)
                        labels.record_stream(main)
                        st.synchronize()
                        return uniq_id, (labels, sc)
                finally:
                    free.append(st)

            with _cabi.short_gil_switch(), ThreadPoolExecutor(max_workers=n_streams) as pool:
                for uniq_id, res in [f.result() for f in [pool.submit(work, u) for u in todo]]:
                    pending[uniq_id] = res
        timer.stop(h)
        labels_host = {}
        self.results = {}
        for uniq_id, (labels, sc) in pending.items():
            lab = labels.cpu().numpy()
            labels_host[uniq_id] = lab
            base_scale_idx = int(self.embs_and_timestamps[uniq_id]["multiscale_segment_counts"].shape[0]) - 1
            self._last_clusterers[uniq_id] = sc
            self.results[uniq_id] = {"labels": lab, "timestamps": sc.timestamps_in_scales[base_scale_idx], "base_scale_idx": base_scale_idx,
                                     "debug": dict(sc.speaker_clustering.debug), "fused_affinity": sc.speaker_clustering.fused_affinity}
        if timers:
            self.stage_ms = timer.totals_ms()
        return labels_host

    # ------------------------------------------------------------------ outputs
    def _write_outputs(self):
        out_rttm_dir = os.path.join(self._out_dir, "pred_rttms")
        lines_cluster_labels, base_scale_idx = [], 0
        for uniq_id, r in self.results.items():
            turns, lines = su.generate_cluster_labels(r["timestamps"], r["labels"])
            su.labels_to_rttmfile(turns, uniq_id, out_rttm_dir)
            lines_cluster_labels.extend(f"{uniq_id} {line}\n" for line in lines)
            base_scale_idx = r["base_scale_idx"]
            r["rttm_labels"] = turns
        su.write_cluster_labels(base_scale_idx, lines_cluster_labels, out_rttm_dir)
        if _get(self.cfg, "diarizer.speaker_embeddings.parameters.save_embeddings", False):
            emb_dir = os.path.join(self._speaker_dir, "embeddings")
            os.makedirs(emb_dir, exist_ok=True)
            for scale_idx, (e_by, _) in self.multiscale_embeddings_and_timestamps.items():
                with open(os.path.join(emb_dir, f"subsegments_scale{scale_idx}_embeddings.pkl"), "wb") as f:
                    pkl.dump({u: e.cpu() for u, e in e_by.items()}, f)

    def diarize(self, paths2audio_files: List[str] = None, batch_size: int = 0):
        """Diarize every recording of the manifest; writes `<out_dir>/pred_rttms/<stem>.rttm`.
        Returns None (upstream returns DER scores only when reference RTTMs are being scored)."""
        if batch_size:
            self.batch_size = int(batch_size)
        t0 = time.perf_counter()
        self._prepare(paths2audio_files)
        t1 = time.perf_counter()
        self.run_device()
        t2 = time.perf_counter()
        self._write_outputs()
        self.host_seconds = {"prepare": t1 - t0, "device_path": t2 - t1, "write": time.perf_counter() - t2}
        return None

    def diarize_async(self, paths2audio_files: List[str] = None, batch_size: int = 0):
        """`diarize()` on a worker thread and a CUDA stream of its own; returns a concurrent.futures.Future.

        In-process counterpart of the reference's parallel mode: diarize_parallel.py:117-120
        starts `python nemo_process.py` as a SUBPROCESS next to the Whisper transcription and joins it at :191-196
        (`nemo_process.wait()`), paying a second CUDA context, a second copy of the models and the interpreter start-up for
        a stage that takes a fraction of a second here.  The caller keeps transcribing on its own stream(s) and calls
        `.result()` where the reference calls `.wait()`.  The other streams of the process are busy while this runs, so the
        worker's GEMMs carry B200D_GEMM_NO_PAIR (include/b200d.h: the cluster-launched kernel never shares the device with
        another stream's kernels); the embeddings then differ from a plain `diarize()` in the last bits (SqueezeExcite means by
        a separate pass instead of the pair kernel's epilogue sums), within the same tolerances."""
        import threading
        from concurrent.futures import Future

        fut: Future = Future()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        if getattr(self, "_async_stream", None) is None:
            self._async_stream = torch.cuda.Stream(device=dev_index)
        stream = self._async_stream
        stream.wait_stream(torch.cuda.current_stream(dev_index))

        def work():
            try:
                torch.cuda.set_device(dev_index)
                with torch.cuda.stream(stream), torch.no_grad(), _cabi.single_cta_gemms():
                    res = self.diarize(paths2audio_files=paths2audio_files, batch_size=batch_size)
                    stream.synchronize()
                fut.set_result(res)
            except BaseException as exc:  # noqa: BLE001 -- delivered through the future
                fut.set_exception(exc)

        threading.Thread(target=work, name="b200d-diarize", daemon=True).start()
        return fut

    # ------------------------------------------------------------------ in-memory fast path
    def diarize_waveform(self, waveform, speech_regions, uniq_id: str = "mono_file", write_rttm_dir: Optional[str] = None):
        """Diarize one recording that is already in memory -- no WAV / manifest / RTTM round trip through the disk.

        The reference holds the 16 kHz mono waveform as a tensor before it writes `mono_file.wav` for NeMo
        (diarize.py:187-196) and parses the RTTM back into (start_ms, end_ms, speaker) triples right after
        (diarize.py:209-216); this entry point takes the tensor and returns those triples directly.
          waveform        float32 [n] torch tensor (host or CUDA) or numpy array, 16 kHz mono
          speech_regions  [(start_s, end_s), ...] from any VAD
        Returns [[start_ms, end_ms, speaker_index], ...] exactly as diarize.py builds `speaker_ts` from the RTTM text
        (same %.3f rounding).  `write_rttm_dir`: also write `<dir>/<uniq_id>.rttm`."""
        wav = torch.as_tensor(waveform)
        if wav.dim() != 1:
            raise ValueError("waveform must be mono: shape [n]")
        wav = wav.to(torch.float32)
        n = int(wav.numel())
        self.AUDIO_RTTM_MAP = {uniq_id: {"audio_filepath": uniq_id + ".wav", "rttm_filepath": None, "offset": 0, "duration": None,
                                         "text": "-", "num_speakers": None, "uem_filepath": None, "ctm_filepath": None}}
        self._wav_offset = {uniq_id: (0, n)}
        total_s = n / self.sample_rate
        regions = []
        for a, b in su._merge_on_grid([[float(a), float(b)] for a, b in speech_regions], 5):
            a, b = max(a, 0.0), min(b, total_s)
            if b > a:
                regions.append({"audio_filepath": uniq_id + ".wav", "offset": round(a, 5), "duration": round(b - a, 5), "label": "UNK", "uniq_id": uniq_id})
        self._plan_windows(regions, write_manifests=False)
        if wav.is_cuda:
            wav_dev = wav.contiguous()
        else:
            self._wav_host = _pinned_buffer(n)
            self._wav_host.copy_(wav)
            wav_dev = None
        self.run_device(wav_dev=wav_dev)
        r = self.results[uniq_id]
        turns, _ = su.generate_cluster_labels(r["timestamps"], r["labels"])
        r["rttm_labels"] = turns
        if write_rttm_dir:
            os.makedirs(write_rttm_dir, exist_ok=True)
            su.labels_to_rttmfile(turns, uniq_id, write_rttm_dir)
        speaker_ts = []
        for line in turns:
            start, end, speaker = line.split()
            start_s = float("{:.3f}".format(float(start)))
            dur_s = float("{:.3f}".format(float(end) - float(start)))
            s_ms = int(start_s * 1000)
            speaker_ts.append([s_ms, s_ms + int(dur_s * 1000), int(speaker.split("_")[-1])])
        return speaker_ts



class NeuralDiarizer:
    """The class the reference actually instantiates (diarize.py:19,200-201; nemo_process.py:6,31-32):
    `NeuralDiarizer(cfg=create_config(dir)).to(device).diarize()`.  NeMo's NeuralDiarizer builds a `ClusteringDiarizer`,
    runs it, and then refines the clustering result with the MSDD decoder (`diar_msdd_telephonic`).

    This wrapper keeps the constructor / `.to()` / `.diarize()` surface and runs the CLUSTERING STAGE on the B200 path.  The
    MSDD refinement is NOT performed (out of scope of the accelerated path): the RTTMs in
    `out_dir/pred_rttms/` are the clustering stage's -- what upstream writes before MSDD overwrites them -- and
    everything MSDD reads is left in `out_dir/speaker_outputs/` (`subsegments_scale<k>.json`, `..._cluster.label`,
    `embeddings/subsegments_scale<k>_embeddings.pkl`), so NeMo's own MSDD stage can be run on top of this output.
    `msdd_refinement_skipped` is True after `diarize()` so callers can tell.  Speech regions: `oracle_vad`, an
    `external_vad_manifest`, or a `vad_fn` callable (see ClusteringDiarizer._run_vad); MarbleNet itself is not bundled."""

    def __init__(self, cfg, speaker_model=None, vad_fn=None):
        self._cfg = as_config(cfg)
        self.clustering_embedding = self  # upstream attribute chain: neural_diarizer.clustering_embedding.clus_diar_model
        self.clus_diar_model = ClusteringDiarizer(self._cfg, speaker_model=speaker_model)
        if vad_fn is not None:
            self.clus_diar_model.vad_fn = vad_fn
        self.msdd_refinement_skipped = False

    def to(self, device):
        self.clus_diar_model.to(device)
        return self

    def diarize(self, paths2audio_files: List[str] = None, batch_size: int = 0):
        import warnings

        self.clus_diar_model._cfg.diarizer.speaker_embeddings.parameters.save_embeddings = True  # MSDD reads the pickles
        self.clus_diar_model.diarize(paths2audio_files=paths2audio_files, batch_size=batch_size)
        self.msdd_refinement_skipped = True
        warnings.warn("whisper_nemo_b200.NeuralDiarizer ran the clustering stage only; the MSDD refinement of NeMo's NeuralDiarizer was "
                      "skipped (its inputs are in out_dir/speaker_outputs/)", RuntimeWarning, stacklevel=2)
        return None
