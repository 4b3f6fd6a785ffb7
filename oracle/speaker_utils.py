"""Oracle: host-side diarization utilities.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates the functions of upstream `nemo/collections/asr/parts/utils/speaker_utils.py`
that `ClusteringDiarizer.diarize()` executes:
manifest parsing, oracle-VAD manifest, multi-scale sub-segmentation, the
embeddings+timestamps packing, and the cluster-label -> RTTM writers whose exact
format strings the reference parses at diarize.py:213-216.
"""
import json
import math
import os
from copy import deepcopy
from typing import Dict, List

import numpy as np
import torch

from . import switches


def get_uniqname_from_filepath(filepath):
    return os.path.splitext(os.path.basename(filepath))[0]


def audio_rttm_map(manifest):
    """speaker_utils.audio_rttm_map: one entry per manifest line, key = file stem."""
    AUDIO_RTTM_MAP = {}
    with open(manifest, "r") as inp_file:
        for line in inp_file.readlines():
            line = line.strip()
            if not line:
                continue
            dic = json.loads(line)
            meta = {
                "audio_filepath": dic["audio_filepath"],
                "rttm_filepath": dic.get("rttm_filepath", None),
                "offset": dic.get("offset", None),
                "duration": dic.get("duration", None),
                "text": dic.get("text", None),
                "num_speakers": dic.get("num_speakers", None),
                "uem_filepath": dic.get("uem_filepath", None),
                "ctm_filepath": dic.get("ctm_filepath", None),
            }
            uniqname = get_uniqname_from_filepath(meta["audio_filepath"])
            if uniqname in AUDIO_RTTM_MAP:
                raise KeyError(f"file {meta['audio_filepath']} is already part of AUDIO_RTTM_MAP")
            AUDIO_RTTM_MAP[uniqname] = meta
    return AUDIO_RTTM_MAP


def parse_scale_configs(window_lengths_in_sec, shift_lengths_in_sec, multiscale_weights):
    """speaker_utils.parse_scale_configs -> {'use_single_scale_clustering', 'scale_dict', 'multiscale_weights'}."""
    check_float = lambda x: isinstance(x, (float, int)) and not isinstance(x, bool)
    check_list = lambda x: isinstance(x, (list, tuple)) or type(x).__name__ == "ListConfig"
    if check_float(window_lengths_in_sec) and check_float(shift_lengths_in_sec):
        return {
            "use_single_scale_clustering": True,
            "scale_dict": {0: (float(window_lengths_in_sec), float(shift_lengths_in_sec))},
            "multiscale_weights": [1.0],
        }
    if not (check_list(window_lengths_in_sec) and check_list(shift_lengths_in_sec)):
        raise ValueError("Multiscale parameters must be all floats or all lists")
    w, s = list(window_lengths_in_sec), list(shift_lengths_in_sec)
    if multiscale_weights is None or len(w) != len(s) or len(w) != len(multiscale_weights):
        raise ValueError("window/shift/multiscale_weights must be lists of equal length")
    if not all(a > b for a, b in zip(w, s)):
        raise ValueError("window length must be larger than shift length")
    if w != sorted(w, reverse=True) or len(set(w)) != len(w):
        raise ValueError("window lengths must be unique and in descending order")
    if len(w) == 1:
        return {"use_single_scale_clustering": True, "scale_dict": {0: (w[0], s[0])}, "multiscale_weights": [1.0]}
    return {
        "use_single_scale_clustering": False,
        "scale_dict": {k: (wl, sl) for k, (wl, sl) in enumerate(zip(w, s))},
        "multiscale_weights": list(multiscale_weights),
    }


def get_subsegments(offset: float, window: float, shift: float, duration: float) -> List[List[float]]:
    """speaker_utils.get_subsegments."""
    if switches.SUBSEGMENT_RULE == "v2":
        return _get_subsegments_v2(offset, window, shift, duration)
    subsegments: List[List[float]] = []
    start = offset
    slice_end = start + duration
    base = math.ceil((duration - window) / shift)
    slices = 1 if base < 0 else base + 1
    for slice_id in range(slices):
        end = start + window
        if end > slice_end:
            end = slice_end
        subsegments.append([start, end - start])
        start = offset + (slice_id + 1) * shift
    return subsegments


def _get_subsegments_v2(offset, window, shift, duration, min_subsegment_duration=0.01, decimals=2):
    """The torch.arange / round(decimals) variant present in NeMo 2.x (lower confidence recollection)."""
    subsegments: List[List[float]] = []
    start = offset
    slice_end = start + duration
    if min_subsegment_duration <= duration <= shift:
        slices = 1
    else:
        slices = int(np.ceil(1 + (duration - window) / shift))
    if slices == 1:
        if min(duration, window) >= min_subsegment_duration:
            subsegments.append([start, min(duration, window)])
    elif slices > 0:
        start_col = torch.arange(offset, slice_end, shift)[:slices]
        dur_col = window * torch.ones(slices)
        dur_col[-1] = min(slice_end - start_col[-1], window)
        dur_col = torch.round(dur_col, decimals=decimals)
        valid_mask = dur_col >= min_subsegment_duration
        subsegments = torch.stack([start_col[valid_mask], dur_col[valid_mask]], dim=1).tolist()
    return subsegments


def read_rttm_lines(rttm_file_path):
    with open(rttm_file_path, "r") as f:
        return [ln for ln in f.readlines() if ln.strip()]


def merge_float_intervals(ranges: List[List[float]], decimals: int = 5) -> List[List[float]]:
    """speaker_utils.combine_float_overlaps / merge_int_intervals: union of intervals,
    touching intervals are merged (start <= previous end), on a 10**decimals integer grid."""
    margin = 10 ** decimals
    ints = sorted([[int(round(a * margin)), int(round(b * margin))] for a, b in ranges])
    merged: List[List[int]] = []
    for st, en in ints:
        if merged and st <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], en)
        else:
            merged.append([st, en])
    return [[a / margin, b / margin] for a, b in merged]


def get_sub_range_list(target_range, source_range_list):
    out = []
    for st, en in source_range_list:
        if en > target_range[0] and st < target_range[1]:
            out.append([max(st, target_range[0]), min(en, target_range[1])])
    return out


def write_rttm2manifest(AUDIO_RTTM_MAP, manifest_file, audio_durations: Dict[str, float], decimals=5):
    """speaker_utils.write_rttm2manifest: the oracle-VAD speech-segments manifest."""
    with open(manifest_file, "w") as outfile:
        for uniq_id, meta in AUDIO_RTTM_MAP.items():
            rttm_lines = read_rttm_lines(meta["rttm_filepath"])
            offset = meta["offset"] if meta["offset"] is not None else 0.0
            duration = meta["duration"] if meta["duration"] is not None else audio_durations[uniq_id]
            offset, duration = round(float(offset), decimals), round(float(duration), decimals)
            raw = []
            for line in rttm_lines:
                fields = line.strip().split()
                start, dur = float(fields[3]), float(fields[4])
                raw.append([start, start + dur])
            vad = merge_float_intervals(raw, decimals)
            for st, en in get_sub_range_list([offset, offset + duration], vad):
                meta_out = {
                    "audio_filepath": meta["audio_filepath"],
                    "offset": round(st, decimals),
                    "duration": round(en - st, decimals),
                    "label": "UNK",
                    "uniq_id": uniq_id,
                }
                json.dump(meta_out, outfile)
                outfile.write("\n")
    return manifest_file


def segments_manifest_to_subsegments_manifest(segments_manifest_file, subsegments_manifest_file, window, shift):
    """speaker_utils.segments_manifest_to_subsegments_manifest."""
    min_subsegment_duration = switches.MIN_SUBSEGMENT_DURATION
    with open(segments_manifest_file, "r") as segments_manifest, open(subsegments_manifest_file, "w") as out:
        for segment in segments_manifest.readlines():
            segment = segment.strip()
            if not segment:
                continue
            dic = json.loads(segment)
            audio, offset, duration, label = dic["audio_filepath"], dic["offset"], dic["duration"], dic["label"]
            for start, dur in get_subsegments(offset=offset, window=window, shift=shift, duration=duration):
                if dur > min_subsegment_duration:
                    meta = {"audio_filepath": audio, "offset": start, "duration": dur, "label": label, "uniq_id": dic.get("uniq_id")}
                    json.dump(meta, out)
                    out.write("\n")
    return subsegments_manifest_file


def get_embs_and_timestamps(multiscale_embeddings_and_timestamps, multiscale_args_dict):
    """speaker_utils.get_embs_and_timestamps."""
    embs_and_timestamps = {uniq_id: {} for uniq_id in multiscale_embeddings_and_timestamps[0][0].keys()}
    if multiscale_args_dict["use_single_scale_clustering"]:
        _args = deepcopy(multiscale_args_dict)
        _args["scale_dict"] = {0: multiscale_args_dict["scale_dict"][0]}
        _args["multiscale_weights"] = multiscale_args_dict["multiscale_weights"][:1]
    else:
        _args = multiscale_args_dict
    embeddings, _ = multiscale_embeddings_and_timestamps[0]
    for uniq_id in embeddings.keys():
        embeddings_list, time_stamps_list, segment_index_list = [], [], []
        for scale_idx in sorted(_args["scale_dict"].keys()):
            embeddings_s, time_stamps = multiscale_embeddings_and_timestamps[scale_idx]
            if len(embeddings_s[uniq_id]) != len(time_stamps[uniq_id]):
                raise ValueError("Mismatch of counts between embedding vectors and timestamps")
            time_stamps_tensor = torch.tensor(time_stamps[uniq_id])
            embeddings_list.append(embeddings_s[uniq_id])
            segment_index_list.append(embeddings_s[uniq_id].shape[0])
            time_stamps_list.append(time_stamps_tensor)
        embs_and_timestamps[uniq_id]["multiscale_weights"] = torch.tensor(_args["multiscale_weights"]).unsqueeze(0).float()
        embs_and_timestamps[uniq_id]["embeddings"] = torch.cat(embeddings_list, dim=0)
        embs_and_timestamps[uniq_id]["timestamps"] = torch.cat(time_stamps_list, dim=0)
        embs_and_timestamps[uniq_id]["multiscale_segment_counts"] = torch.tensor(segment_index_list)
    return embs_and_timestamps


def get_contiguous_stamps(stamps):
    """speaker_utils.get_contiguous_stamps: split overlaps of adjacent windows at the midpoint."""
    lines = deepcopy(stamps)
    contiguous_stamps = []
    for i in range(len(lines) - 1):
        start, end, speaker = lines[i].split()
        next_start, next_end, next_speaker = lines[i + 1].split()
        if float(end) > float(next_start):
            avg = str((float(next_start) + float(end)) / 2.0)
            lines[i + 1] = " ".join([avg, next_end, next_speaker])
            contiguous_stamps.append(start + " " + avg + " " + speaker)
        else:
            contiguous_stamps.append(start + " " + end + " " + speaker)
    start, end, speaker = lines[-1].split()
    contiguous_stamps.append(start + " " + end + " " + speaker)
    return contiguous_stamps


def merge_stamps(lines):
    """speaker_utils.merge_stamps: merge adjacent same-speaker stamps."""
    stamps = deepcopy(lines)
    overlap_stamps = []
    for i in range(len(stamps) - 1):
        start, end, speaker = stamps[i].split()
        next_start, next_end, next_speaker = stamps[i + 1].split()
        if float(end) == float(next_start) and speaker == next_speaker:
            stamps[i + 1] = " ".join([start, next_end, next_speaker])
        else:
            overlap_stamps.append(start + " " + end + " " + speaker)
    start, end, speaker = stamps[-1].split()
    overlap_stamps.append(start + " " + end + " " + speaker)
    return overlap_stamps


def generate_cluster_labels(segment_ranges: torch.Tensor, cluster_labels):
    """speaker_utils.generate_cluster_labels; timestamps arrive as a float32 tensor, and
    f"{zero_dim_tensor}" formats `tensor.item()` (a python float holding the fp32 value)."""
    lines = []
    for idx, label in enumerate(cluster_labels):
        tag = "speaker_" + str(int(label))
        stt, end = segment_ranges[idx]
        lines.append(f"{stt.item()} {end.item()} {tag}")
    cont_lines = get_contiguous_stamps(lines)
    diar_hyp = merge_stamps(cont_lines)
    return diar_hyp, lines


def labels_to_rttmfile(labels, uniq_id, out_rttm_dir):
    """speaker_utils.labels_to_rttmfile -- the triple-space format diarize.py:213-216 depends on."""
    filename = os.path.join(out_rttm_dir, uniq_id + ".rttm")
    with open(filename, "w") as f:
        for line in labels:
            line = line.strip()
            start, end, speaker = line.split()
            duration = float(end) - float(start)
            start = float(start)
            f.write("SPEAKER {} 1   {:.3f}   {:.3f} <NA> <NA> {} <NA> <NA>\n".format(uniq_id, start, duration, speaker))
    return filename


def write_cluster_labels(base_scale_idx, lines_cluster_labels, out_rttm_dir):
    out_label_name = os.path.join(out_rttm_dir, "../speaker_outputs", f"subsegments_scale{base_scale_idx}_cluster.label")
    with open(out_label_name, "w") as f:
        for clus_label_line in lines_cluster_labels:
            f.write(clus_label_line)


def rttm_to_labels(rttm_filename):
    labels = []
    with open(rttm_filename, "r") as f:
        for line in f.readlines():
            rttm = line.strip().split()
            if not rttm:
                continue
            start, end, speaker = float(rttm[3]), float(rttm[4]) + float(rttm[3]), rttm[7]
            labels.append("{} {} {}".format(start, end, speaker))
    return labels
