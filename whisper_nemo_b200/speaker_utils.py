"""Host-side diarization utilities of the B200 path: manifests, oracle-VAD speech regions,
multi-scale sub-segmentation, embedding/timestamp packing and the cluster-label -> RTTM writers.

Mirrors the functions of upstream NeMo's `nemo/collections/asr/parts/utils/speaker_utils.py`
that `ClusteringDiarizer.diarize()` runs around the device work (same names, same argument
meaning, same files written), because the reference consumes their outputs byte-for-byte:
`diarize.py:209-216` / `diarize_parallel.py:202-208` split every RTTM line on a single space
and read fields [5], [8], [11], which only works for the triple-space format string of
`labels_to_rttmfile`; the manifest keys are the ones `helpers.py:267-275` writes.
This is index arithmetic and text; no kernel is involved.
"""
import json
import math
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

MIN_SUBSEGMENT_DURATION = 0.05


def get_uniqname_from_filepath(filepath: str) -> str:
    return os.path.splitext(os.path.basename(filepath))[0]


def audio_rttm_map(manifest: str) -> Dict[str, dict]:
    """One entry per manifest line, keyed by the audio file's stem (`mono_file` for the reference)."""
    out: Dict[str, dict] = {}
    with open(manifest, "r") as f:
        for raw in f:
            raw = raw.strip()
            if not raw:
                continue
            dic = json.loads(raw)
            uniq = get_uniqname_from_filepath(dic["audio_filepath"])
            if uniq in out:
                raise KeyError(f"file {dic['audio_filepath']} is already part of AUDIO_RTTM_MAP, it might be duplicated")
            out[uniq] = {key: dic.get(key, None) for key in
                         ("audio_filepath", "rttm_filepath", "offset", "duration", "text", "num_speakers", "uem_filepath", "ctm_filepath")}
    return out


def parse_scale_configs(window_lengths_in_sec, shift_lengths_in_sec, multiscale_weights) -> dict:
    """`diarizer.speaker_embeddings.parameters.{window_length_in_sec, shift_length_in_sec, multiscale_weights}`
    (diar_infer_*.yaml:42-44) -> {'use_single_scale_clustering', 'scale_dict': {idx: (window, shift)}, 'multiscale_weights'}."""
    is_num = lambda x: isinstance(x, (int, float)) and not isinstance(x, bool)
    is_seq = lambda x: isinstance(x, (list, tuple)) or type(x).__name__ == "ListConfig"
    if is_num(window_lengths_in_sec) and is_num(shift_lengths_in_sec):
        return {"use_single_scale_clustering": True, "scale_dict": {0: (float(window_lengths_in_sec), float(shift_lengths_in_sec))},
                "multiscale_weights": [1.0]}
    if not (is_seq(window_lengths_in_sec) and is_seq(shift_lengths_in_sec)):
        raise ValueError("Multiscale parameters are not properly setup: window and shift must both be floats or both be lists")
    windows, shifts = list(window_lengths_in_sec), list(shift_lengths_in_sec)
    if multiscale_weights is None or not (len(windows) == len(shifts) == len(multiscale_weights)):
        raise ValueError("Multiscale parameters are not properly setup: lists must have equal length")
    if any(w <= s for w, s in zip(windows, shifts)):
        raise ValueError("Multiscale parameters are not properly setup: window must be longer than shift")
    if windows != sorted(set(windows), reverse=True):
        raise ValueError("Multiscale parameters are not properly setup: windows must be unique and descending")
    if len(windows) == 1:
        return {"use_single_scale_clustering": True, "scale_dict": {0: (windows[0], shifts[0])}, "multiscale_weights": [1.0]}
    return {"use_single_scale_clustering": False, "scale_dict": dict(enumerate(zip(windows, shifts))),
            "multiscale_weights": list(multiscale_weights)}


# Which release of upstream's `get_subsegments` the window grid follows.  "classic" (NeMo 1.x / early 2.x): float64 python
# arithmetic, n = ceil((dur - w) / shift) + 1.  "v2" (later 2.x; the reference pins only nemo >= 2.dev): start times from
# torch.arange (float32), durations rounded to 2 decimals, a region shorter than the shift kept as one window.  The two differ in
# the last digits of window times and occasionally in the window count of a region; set B200D_SUBSEGMENT_RULE=v2 to follow a
# NeMo that has the later one (the oracle's oracle/switches.py SUBSEGMENT_RULE is the same switch on the CPU side).
SUBSEGMENT_RULE = os.environ.get("B200D_SUBSEGMENT_RULE", "classic")


def get_subsegments_v2(offset: float, window: float, shift: float, duration: float, min_subsegment_duration: float = 0.01,
                       decimals: int = 2) -> List[List[float]]:
    """The later form of upstream's get_subsegments (see SUBSEGMENT_RULE)."""
    slice_end = offset + duration
    slices = 1 if min_subsegment_duration <= duration <= shift else int(np.ceil(1 + (duration - window) / shift))
    if slices == 1:
        return [[offset, min(duration, window)]] if min(duration, window) >= min_subsegment_duration else []
    if slices <= 0:
        return []
    starts = torch.arange(offset, slice_end, shift)[:slices]
    durs = window * torch.ones(slices)
    durs[-1] = min(slice_end - starts[-1], window)
    durs = torch.round(durs, decimals=decimals)
    keep = durs >= min_subsegment_duration
    return torch.stack([starts[keep], durs[keep]], dim=1).tolist()


def get_subsegments(offset: float, window: float, shift: float, duration: float) -> List[List[float]]:
    """Sliding windows over one speech region: start_k = offset + k * shift, the last window is cut at
    the region's end.  n = ceil((duration - window) / shift) + 1 (1 if the region is shorter than a window)."""
    if SUBSEGMENT_RULE == "v2":
        return get_subsegments_v2(offset, window, shift, duration)
    base = math.ceil((duration - window) / shift)
    count = 1 if base < 0 else base + 1
    region_end = offset + duration
    out = []
    for k in range(count):
        start = offset if k == 0 else offset + k * shift
        stop = start + window
        if stop > region_end:
            stop = region_end
        out.append([start, stop - start])
    return out


def _merge_on_grid(ranges: Sequence[Sequence[float]], decimals: int) -> List[List[float]]:
    scale = 10 ** decimals
    grid = sorted([int(round(a * scale)), int(round(b * scale))] for a, b in ranges)
    merged: List[List[int]] = []
    for a, b in grid:
        if merged and a <= merged[-1][1]:
            if b > merged[-1][1]:
                merged[-1][1] = b
        else:
            merged.append([a, b])
    return [[a / scale, b / scale] for a, b in merged]


def write_rttm2manifest(AUDIO_RTTM_MAP: Dict[str, dict], manifest_file: str, audio_durations: Dict[str, float], decimals: int = 5) -> str:
    """Oracle VAD (`diarizer.oracle_vad: True`): speech regions = union of the RTTM turns of each
    recording, clipped to [offset, offset + duration], one manifest line per region."""
    with open(manifest_file, "w") as out:
        for uniq_id, meta in AUDIO_RTTM_MAP.items():
            if not meta.get("rttm_filepath"):
                raise ValueError(f"oracle_vad needs rttm_filepath in the manifest entry of {uniq_id}")
            turns = []
            with open(meta["rttm_filepath"], "r") as f:
                for line in f:
                    fields = line.split()
                    if fields:
                        turns.append([float(fields[3]), float(fields[3]) + float(fields[4])])
            begin = round(float(meta["offset"] if meta["offset"] is not None else 0.0), decimals)
            length = round(float(meta["duration"] if meta["duration"] is not None else audio_durations[uniq_id]), decimals)
            stop = begin + length
            for a, b in _merge_on_grid(turns, decimals):
                if b > begin and a < stop:
                    a, b = max(a, begin), min(b, stop)
                    json.dump({"audio_filepath": meta["audio_filepath"], "offset": round(a, decimals), "duration": round(b - a, decimals),
                               "label": "UNK", "uniq_id": uniq_id}, out)
                    out.write("\n")
    return manifest_file


def segments_manifest_to_subsegments_manifest(segments_manifest_file: str, subsegments_manifest_file: str, window: float, shift: float,
                                              min_subsegment_duration: float = MIN_SUBSEGMENT_DURATION) -> List[dict]:
    """Speech regions -> windows of one scale; writes `subsegments_scale<k>.json` (read later by MSDD) and
    returns the same entries as a list."""
    entries: List[dict] = []
    with open(segments_manifest_file, "r") as src, open(subsegments_manifest_file, "w") as out:
        for raw in src:
            raw = raw.strip()
            if not raw:
                continue
            dic = json.loads(raw)
            for start, dur in get_subsegments(offset=dic["offset"], window=window, shift=shift, duration=dic["duration"]):
                if dur > min_subsegment_duration:
                    meta = {"audio_filepath": dic["audio_filepath"], "offset": start, "duration": dur, "label": dic["label"],
                            "uniq_id": dic.get("uniq_id")}
                    json.dump(meta, out)
                    out.write("\n")
                    entries.append(meta)
    return entries


def read_speech_regions(segments_manifest_file: str):
    """Lines of a speech-region manifest (oracle VAD / external VAD / VAD output) -> list of dicts."""
    out = []
    with open(segments_manifest_file, "r") as src:
        for raw in src:
            raw = raw.strip()
            if raw:
                out.append(json.loads(raw))
    return out


def subsegment_arrays(offsets: np.ndarray, durations: np.ndarray, window: float, shift: float,
                      min_subsegment_duration: float = MIN_SUBSEGMENT_DURATION):
    """`get_subsegments` for all speech regions of a manifest at once (float64 numpy, the same IEEE operations in the same
    order as the per-region Python loop, so every start / duration is bit-identical to it).
    Returns (region index int64 [n], start float64 [n], duration float64 [n]) with windows <= min_subsegment_duration dropped."""
    offsets, durations = np.asarray(offsets, dtype=np.float64), np.asarray(durations, dtype=np.float64)
    if SUBSEGMENT_RULE == "v2":  # the later rule goes through torch.arange per region (float32 starts): no vectorised twin
        rows = [(r, st, du) for r, (o, d) in enumerate(zip(offsets.tolist(), durations.tolist()))
                for st, du in get_subsegments_v2(o, window, shift, d) if du > min_subsegment_duration]
        return (np.array([x[0] for x in rows], dtype=np.int64), np.array([x[1] for x in rows], dtype=np.float64),
                np.array([x[2] for x in rows], dtype=np.float64))
    base = np.ceil((durations - window) / shift)
    count = np.where(base < 0, 1, base + 1).astype(np.int64)
    region = np.repeat(np.arange(offsets.shape[0], dtype=np.int64), count)
    first = np.cumsum(count) - count
    k = np.arange(int(count.sum()), dtype=np.int64) - first[region]
    start = np.where(k == 0, offsets[region], offsets[region] + k * shift)
    stop = np.minimum(start + window, (offsets + durations)[region])
    dur = stop - start
    keep = dur > min_subsegment_duration
    return region[keep], start[keep], dur[keep]


def write_subsegments_manifest(path: str, regions: Sequence[dict], region_idx: np.ndarray, start: np.ndarray, dur: np.ndarray) -> None:
    """`subsegments_scale<k>.json`, byte-identical to the json.dump lines of segments_manifest_to_subsegments_manifest."""
    heads = [json.dumps(r["audio_filepath"]) for r in regions]
    tails = [f', "label": {json.dumps(r["label"])}, "uniq_id": {json.dumps(r.get("uniq_id"))}}}\n' for r in regions]
    with open(path, "w") as out:
        out.write("".join(f'{{"audio_filepath": {heads[r]}, "offset": {o!r}, "duration": {d!r}{tails[r]}'
                          for r, o, d in zip(region_idx.tolist(), start.tolist(), dur.tolist())))


def get_embs_and_timestamps(multiscale_embeddings_and_timestamps: dict, multiscale_args_dict: dict) -> Dict[str, dict]:
    """{scale: (embeddings{uniq}, time_stamps{uniq})} -> per recording: concatenated embeddings (device),
    float32 timestamps, per-scale counts and the weight row."""
    args = multiscale_args_dict
    scales = sorted(args["scale_dict"].keys())
    weights = args["multiscale_weights"]
    if args["use_single_scale_clustering"]:
        scales, weights = scales[:1], weights[:1]
    first_embs, _ = multiscale_embeddings_and_timestamps[scales[0]]
    out: Dict[str, dict] = {}
    for uniq_id in first_embs.keys():
        embs, stamps, counts = [], [], []
        for s in scales:
            e, t = multiscale_embeddings_and_timestamps[s]
            if len(e[uniq_id]) != len(t[uniq_id]):
                raise ValueError("Mismatch of counts between embedding vectors and timestamps")
            embs.append(e[uniq_id])
            stamps.append(t[uniq_id] if torch.is_tensor(t[uniq_id]) else torch.tensor(t[uniq_id]))
            counts.append(e[uniq_id].shape[0])
        out[uniq_id] = {"multiscale_weights": torch.tensor(weights).unsqueeze(0).float(), "embeddings": torch.cat(embs, dim=0),
                        "timestamps": torch.cat(stamps, dim=0), "multiscale_segment_counts": torch.tensor(counts)}
    return out


def generate_cluster_labels(segment_ranges: torch.Tensor, cluster_labels) -> Tuple[List[str], List[str]]:
    """Base-scale windows + labels -> (merged speaker turns, raw per-window lines), both as
    "start end speaker_<k>" strings.  Overlapping neighbours are cut at the midpoint of the overlap,
    then adjacent turns of one speaker that touch exactly are joined (upstream get_contiguous_stamps
    + merge_stamps, done here on floats: str(float) round-trips exactly in Python)."""
    ranges = segment_ranges.detach().cpu().tolist()  # python floats holding the fp32 values
    labels = [int(x) for x in cluster_labels]
    lines = [f"{st} {en} speaker_{lab}" for (st, en), lab in zip(ranges, labels)]
    n = len(labels)
    if n == 0:
        return [], lines
    starts = [r[0] for r in ranges]
    ends = [r[1] for r in ranges]
    for i in range(n - 1):  # cut overlaps (the moved start is seen by the next comparison, as upstream)
        if ends[i] > starts[i + 1]:
            mid = (starts[i + 1] + ends[i]) / 2.0
            ends[i] = mid
            starts[i + 1] = mid
    turns: List[str] = []
    run_start = starts[0]
    for i in range(n - 1):
        if ends[i] == starts[i + 1] and labels[i] == labels[i + 1]:
            continue
        turns.append(f"{run_start} {ends[i]} speaker_{labels[i]}")
        run_start = starts[i + 1]
    turns.append(f"{run_start} {ends[n - 1]} speaker_{labels[n - 1]}")
    return turns, lines


def labels_to_rttmfile(labels: List[str], uniq_id: str, out_rttm_dir: str) -> str:
    """`SPEAKER <id> 1   <start:.3f>   <dur:.3f> <NA> <NA> <spk> <NA> <NA>` -- the triple spaces are load-bearing."""
    filename = os.path.join(out_rttm_dir, uniq_id + ".rttm")
    with open(filename, "w") as f:
        for line in labels:
            start, end, speaker = line.strip().split()
            start, end = float(start), float(end)
            f.write("SPEAKER {} 1   {:.3f}   {:.3f} <NA> <NA> {} <NA> <NA>\n".format(uniq_id, start, end - start, speaker))
    return filename


def write_cluster_labels(base_scale_idx: int, lines_cluster_labels: List[str], out_rttm_dir: str) -> str:
    path = os.path.join(out_rttm_dir, "../speaker_outputs", f"subsegments_scale{base_scale_idx}_cluster.label")
    with open(path, "w") as f:
        f.writelines(lines_cluster_labels)
    return path


def rttm_to_turns(rttm_filename: str) -> List[Tuple[float, float, str]]:
    """Parse an RTTM the way the reference does (diarize.py:213-216 indices also resolve on this split())."""
    turns = []
    with open(rttm_filename, "r") as f:
        for line in f:
            fields = line.split()
            if fields:
                turns.append((float(fields[3]), float(fields[3]) + float(fields[4]), fields[7]))
    return turns


def read_wav(path: str, expected_sr: int = 16000) -> np.ndarray:
    """16 kHz mono WAV -> float32 in [-1, 1): float32 files as written by diarize.py:191-196
    (torchaudio.save) and int16 PCM as written by nemo_process.py:24-28 (pydub), like soundfile."""
    from scipy.io import wavfile

    try:
        sr, data = wavfile.read(path, mmap=True)  # float32 files are copied once, straight into the caller's pinned buffer
    except (ValueError, OSError):
        sr, data = wavfile.read(path)
    if sr != expected_sr:
        raise ValueError(f"{path}: expected {expected_sr} Hz audio, got {sr}")
    if data.ndim > 1:
        data = data.mean(axis=1)
    if data.dtype == np.int16:
        data = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        data = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        data = (data.astype(np.float32) - 128.0) / 128.0
    return data if data.dtype == np.float32 else np.ascontiguousarray(data, dtype=np.float32)
