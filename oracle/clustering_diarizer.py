"""Oracle: end-to-end `ClusteringDiarizer.diarize()` on CPU.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates upstream `nemo/collections/asr/models/clustering_diarizer.py`
(`ClusteringDiarizer.{_perform_speech_activity_detection (oracle_vad branch),
_run_segmentation, _extract_embeddings, diarize}`), the speaker-label dataset
collation of `nemo/collections/asr/data/audio_to_label.py`
(`_fixed_seq_collate_fn` / `_speech_collate_fn`) and
`speaker_utils.perform_clustering` -- the code behind the reference's call
`NeuralDiarizer(cfg=create_config(temp_path)).to(device).diarize()`
(diarize.py:200-201, nemo_process.py:31-32) up to, and excluding, the MSDD decoder.
This is the "CPU restatement of NeMo (NeMo unavailable)" timed as the CPU baseline.
"""
import json
import os
import shutil
import time
from typing import Dict, List

import numpy as np
import torch

from . import speaker_utils as su
from . import switches
from .longform_clustering import LongFormSpeakerClustering


def _get(cfg, dotted, default=None):
    cur = cfg
    for k in dotted.split("."):
        if cur is None:
            return default
        cur = cur.get(k, None) if hasattr(cur, "get") else getattr(cur, k, None)
    return default if cur is None else cur


def read_wav(path: str) -> np.ndarray:
    """soundfile-equivalent read of a mono WAV to float32 (int16 PCM scaled by 1/32768)."""
    from scipy.io import wavfile

    sr, data = wavfile.read(path)
    if sr != 16000:
        raise ValueError(f"expected 16 kHz audio, got {sr}")
    if data.ndim > 1:
        data = data.mean(axis=1)
    if data.dtype == np.int16:
        data = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        data = data.astype(np.float32) / 2147483648.0
    return np.ascontiguousarray(data, dtype=np.float32)


def collate(batch: List[torch.Tensor]):
    """audio_to_label collate for the speaker model's test dataloader."""
    lengths = [int(s.shape[0]) for s in batch]
    fixed_length = max(lengths)
    if switches.COLLATE == "fixed_seq":
        out = []
        for sig, sig_len in zip(batch, lengths):
            if sig_len < fixed_length:
                repeat = fixed_length // sig_len
                rem = fixed_length % sig_len
                sub = sig[-rem:] if rem > 0 else torch.tensor([])
                rep_sig = torch.cat(repeat * [sig])
                sig = torch.cat((rep_sig, sub))
            out.append(sig)
        return torch.stack(out), torch.full((len(batch),), fixed_length, dtype=torch.long)
    out = [torch.nn.functional.pad(s, (0, fixed_length - l)) for s, l in zip(batch, lengths)]
    return torch.stack(out), torch.tensor(lengths, dtype=torch.long)


class OracleClusteringDiarizer:
    def __init__(self, cfg, speaker_model):
        self.cfg = cfg
        self.model = speaker_model
        self.sample_rate = int(_get(cfg, "sample_rate", 16000))
        self.batch_size = int(_get(cfg, "batch_size", 64))
        p = _get(cfg, "diarizer.speaker_embeddings.parameters")
        self.multiscale_args_dict = su.parse_scale_configs(
            _get(p, "window_length_in_sec"), _get(p, "shift_length_in_sec"), _get(p, "multiscale_weights")
        )
        self.clus = _get(cfg, "diarizer.clustering.parameters")
        self.stage_seconds: Dict[str, float] = {}
        self.results: Dict[str, dict] = {}
        self.clusterers: Dict[str, LongFormSpeakerClustering] = {}

    # -- untimed preparation: manifests, VAD, WAV decode ------------------------------------
    def prepare(self):
        cfg = self.cfg
        self.out_dir = _get(cfg, "diarizer.out_dir")
        self.speaker_dir = os.path.join(self.out_dir, "speaker_outputs")
        if os.path.exists(self.speaker_dir):
            shutil.rmtree(self.speaker_dir, ignore_errors=True)
        os.makedirs(self.speaker_dir)
        os.makedirs(os.path.join(self.out_dir, "pred_rttms"), exist_ok=True)
        self.AUDIO_RTTM_MAP = su.audio_rttm_map(_get(cfg, "diarizer.manifest_filepath"))
        self.wavs = {u: read_wav(m["audio_filepath"]) for u, m in self.AUDIO_RTTM_MAP.items()}
        durations = {u: len(w) / self.sample_rate for u, w in self.wavs.items()}
        ext = _get(cfg, "diarizer.vad.external_vad_manifest")
        if _get(cfg, "diarizer.oracle_vad", False):
            self.speech_manifest = su.write_rttm2manifest(
                self.AUDIO_RTTM_MAP, os.path.join(self.speaker_dir, "oracle_vad_manifest.json"), durations
            )
        elif ext:
            self.speech_manifest = ext
        else:
            raise NotImplementedError("oracle covers oracle_vad / external_vad_manifest only (MarbleNet VAD is out of scope, SURVEY D8)")
        self.subseg_manifests = {}
        for scale_idx, (window, shift) in self.multiscale_args_dict["scale_dict"].items():
            path = os.path.join(self.speaker_dir, f"subsegments_scale{scale_idx}.json")
            su.segments_manifest_to_subsegments_manifest(self.speech_manifest, path, window, shift)
            self.subseg_manifests[scale_idx] = path

    # -- timed: waveform in RAM -> labels in RAM ----------------------------------------------
    def plan_batches(self):
        """Dataloader batches of the whole job, in upstream's order (scale by scale, `batch_size` windows each):
        [(scale_idx, first window, one past last window)].  `embed()` is `embed_batch` over this list + `finish_embed()`;
        bench.py's reference arm walks the same list in slices to time a bounded part of the pass per step."""
        self._entries, self._sigs, self._embs = {}, {}, {}
        plan = []
        for scale_idx in self.multiscale_args_dict["scale_dict"]:
            entries = [json.loads(l) for l in open(self.subseg_manifests[scale_idx]) if l.strip()]
            sigs = []
            for dic in entries:
                uniq = su.get_uniqname_from_filepath(dic["audio_filepath"])
                start = int(dic["offset"] * self.sample_rate)
                n = int(dic["duration"] * self.sample_rate)
                sigs.append(torch.from_numpy(self.wavs[uniq][start : start + n]))
            self._entries[scale_idx], self._sigs[scale_idx], self._embs[scale_idx] = entries, sigs, {}
            plan.extend((scale_idx, b0, min(b0 + self.batch_size, len(sigs))) for b0 in range(0, len(sigs), self.batch_size))
        return plan

    def embed_batch(self, item):
        scale_idx, b0, b1 = item
        audio, lens = collate(self._sigs[scale_idx][b0:b1])
        _, embs = self.model(audio, lens)
        self._embs[scale_idx][b0] = embs

    def finish_embed(self):
        self.multiscale_embeddings_and_timestamps = {}
        for scale_idx, entries in self._entries.items():
            parts = [self._embs[scale_idx][b0] for b0 in sorted(self._embs[scale_idx])]
            all_embs = torch.cat(parts) if parts else torch.empty([0])
            embeddings, time_stamps = {}, {}
            for i, dic in enumerate(entries):
                uniq = su.get_uniqname_from_filepath(dic["audio_filepath"])
                embeddings.setdefault(uniq, []).append(all_embs[i].view(1, -1))
                start = dic["offset"]
                time_stamps.setdefault(uniq, []).append([start, start + dic["duration"]])
            self.multiscale_embeddings_and_timestamps[scale_idx] = ({u: torch.cat(v) for u, v in embeddings.items()}, time_stamps)
        self.embs_and_timestamps = su.get_embs_and_timestamps(self.multiscale_embeddings_and_timestamps, self.multiscale_args_dict)
        self._sigs, self._embs = {}, {}
        if _get(self.cfg, "diarizer.speaker_embeddings.parameters.save_embeddings", False):
            # upstream _extract_embeddings: {uniq_id: [n, 192]} per scale, read back by the MSDD stage
            import pickle as pkl

            emb_dir = os.path.join(self.speaker_dir, "embeddings")
            os.makedirs(emb_dir, exist_ok=True)
            for scale_idx, (embeddings, _) in self.multiscale_embeddings_and_timestamps.items():
                with open(os.path.join(emb_dir, f"subsegments_scale{scale_idx}_embeddings.pkl"), "wb") as f:
                    pkl.dump(embeddings, f)

    def embed(self):
        t0 = time.perf_counter()
        for item in self.plan_batches():
            self.embed_batch(item)
        self.finish_embed()
        self.stage_seconds["embed"] = time.perf_counter() - t0

    def cluster(self):
        """speaker_utils.perform_clustering (labels only; RTTM writing is `write_outputs`)."""
        t0 = time.perf_counter()
        clus = self.clus
        for uniq_id, meta in self.AUDIO_RTTM_MAP.items():
            if uniq_id not in self.embs_and_timestamps:
                continue
            e = self.embs_and_timestamps[uniq_id]
            if _get(clus, "oracle_num_speakers", False):
                num_speakers = meta.get("num_speakers", None)
                if num_speakers is None:
                    raise ValueError("Provided option as oracle num of speakers but num_speakers in manifest is null")
            else:
                num_speakers = -1
            sc = LongFormSpeakerClustering()
            self.clusterers[uniq_id] = sc
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
            )
            base_scale_idx = e["multiscale_segment_counts"].shape[0] - 1
            self.results[uniq_id] = {
                "labels": labels.cpu().numpy(),
                "timestamps": sc.timestamps_in_scales[base_scale_idx],
                "base_scale_idx": base_scale_idx,
                "debug": dict(sc.speaker_clustering.debug),
                "fused_affinity": getattr(sc.speaker_clustering, "fused_affinity", None),
            }
        self.stage_seconds["cluster"] = time.perf_counter() - t0

    def write_outputs(self):
        out_rttm_dir = os.path.join(self.out_dir, "pred_rttms")
        lines_cluster_labels = []
        base_scale_idx = 0
        for uniq_id, r in self.results.items():
            labels, lines = su.generate_cluster_labels(r["timestamps"], r["labels"])
            su.labels_to_rttmfile(labels, uniq_id, out_rttm_dir)
            lines_cluster_labels.extend([f"{uniq_id} {seg_line}\n" for seg_line in lines])
            base_scale_idx = r["base_scale_idx"]
            r["rttm_labels"] = labels
        su.write_cluster_labels(base_scale_idx, lines_cluster_labels, out_rttm_dir)

    def diarize(self):
        self.prepare()
        self.embed()
        self.cluster()
        self.write_outputs()
        return self.results
