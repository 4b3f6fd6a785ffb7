"""Benchmark / test WORKLOADS: seeded synthetic multi-speaker 16 kHz audio + ground-truth RTTM, the session
(manifest + config) builders around it and the label / RTTM comparison metrics.  Shared by tests/, bench.py (both arms)
and __graft_entry__.smoke(); it imports neither the product package nor the oracle, so the reference arm of the
benchmark can build its inputs without loading libb200d.so.

There is no network for datasets and the reference's only clip
(/root/reference/tests/assets/test.opus, 22.58 s stereo 48 kHz Opus) cannot be
decoded in this image, so every BASELINE.json config runs on audio generated
here (SURVEY.md section 8d "Synthetic audio").  Each speaker is a harmonic
source with its own F0, vocal-tract scale and spectral tilt, stepping through
vowel-like formant settings; turns are 2-10 s with 0.2-1 s silences and no
overlap.  The ground-truth RTTM doubles as oracle VAD (`oracle_vad: True`).
"""
import json
import os
from typing import List, Tuple

import numpy as np
import torch

SR = 16000
_VOWELS = np.array(
    [[730, 1090, 2440], [270, 2290, 3010], [300, 870, 2240], [530, 1840, 2480], [570, 840, 2410], [440, 1020, 2240]],
    dtype=np.float64,
)
_BW = np.array([90.0, 110.0, 170.0])


# "v2" speakers (the full-size benchmark recordings): the same source-filter model with speaker identity made robust for a
# RANDOM-INIT network -- every speaker owns a fixed subset of three vowels (which mel bands move together survives the
# per-feature normalisation), wider and jitter-free syllable / vibrato / vowel rates.  With "v1" voices an 8-speaker hour is an
# ill-posed clustering problem for untrained weights (the CPU oracle itself flips between 3, 7 and 8 speakers under an
# fp32-vs-fp64 eigh change); the small parity tests keep "v1".
STYLE = os.environ.get("B200D_SYNTH_STYLE", "v1")


def _speaker_params(n_speakers: int, rng: np.random.Generator):
    f0 = np.linspace(95.0, 245.0, n_speakers) if n_speakers > 1 else np.array([140.0])
    f0 = f0[rng.permutation(n_speakers)] * (1.0 + 0.03 * rng.standard_normal(n_speakers))
    tract = np.linspace(0.82, 1.22, n_speakers)[rng.permutation(n_speakers)]
    tilt = 0.6 + 1.0 * rng.random(n_speakers)
    # speaker-specific DYNAMICS: TitaNet's per-feature normalisation removes every static spectral envelope, so what
    # separates speakers for a (random-init) network is temporal texture -- syllable rate, vibrato, vowel pace
    spread = lambda lo, hi: np.geomspace(lo, hi, n_speakers)[rng.permutation(n_speakers)] if n_speakers > 1 else np.array([(lo * hi) ** 0.5])
    dyn = {"am_rate": spread(2.2, 9.0), "am_depth": 0.25 + 0.5 * rng.random(n_speakers), "vib_rate": spread(3.0, 11.0),
           "vib_depth": spread(0.01, 0.08), "vowel_s": spread(0.06, 0.32)}
    if STYLE == "v2":
        from itertools import combinations

        subsets = list(combinations(range(len(_VOWELS)), 3))
        order = rng.permutation(len(subsets))
        dyn.update({"am_rate": spread(1.8, 14.0), "vib_rate": spread(2.5, 15.0), "vowel_s": spread(0.045, 0.42),
                    "vowels": [np.asarray(subsets[order[i % len(subsets)]]) for i in range(n_speakers)]})
    return f0, tract, tilt, dyn


def make_turns(duration_s: float, n_speakers: int, rng: np.random.Generator, turn=(2.0, 10.0), gap=(0.2, 1.0)) -> List[Tuple[float, float, int]]:
    turns = []
    t = round(float(rng.uniform(0.1, 0.5)), 3)
    prev = -1
    while True:
        dur = round(float(rng.uniform(*turn)), 3)
        if t + dur > duration_s - 0.05:
            dur = round(duration_s - 0.05 - t, 3)
            if dur < 0.6:
                break
        spk = int(rng.integers(n_speakers))
        if n_speakers > 1 and spk == prev:
            spk = (spk + 1 + int(rng.integers(n_speakers - 1))) % n_speakers
        turns.append((t, round(t + dur, 3), spk))
        prev = spk
        t = round(t + dur + float(rng.uniform(*gap)), 3)
        if t >= duration_s - 0.7:
            break
    # guarantee every speaker appears at least once
    missing = [s for s in range(n_speakers) if s not in {x[2] for x in turns}]
    for i, s in enumerate(missing):
        if i < len(turns):
            a, b, _ = turns[-1 - i]
            turns[-1 - i] = (a, b, s)
    return turns


def synth_recording(duration_s: float, n_speakers: int, seed: int, n_harm: int = 24) -> Tuple[np.ndarray, List[Tuple[float, float, int]]]:
    """Returns (float32 waveform [duration_s * 16000], [(start_s, end_s, speaker)])."""
    rng = np.random.default_rng(seed)
    n_total = int(round(duration_s * SR))
    f0s, tracts, tilts, dyn = _speaker_params(n_speakers, rng)
    turns = make_turns(duration_s, n_speakers, rng)
    gen = torch.Generator().manual_seed(seed)
    wav = 0.0008 * torch.randn(n_total, generator=gen)
    harm = torch.arange(1, n_harm + 1, dtype=torch.float32).unsqueeze(1)
    for (st, en, spk) in turns:
        i0, i1 = int(round(st * SR)), int(round(en * SR))
        n = i1 - i0
        tt = torch.arange(n, dtype=torch.float32) / SR
        # slowly varying pitch
        f0 = f0s[spk] * (1.0 + 0.04 * torch.sin(2 * np.pi * float(rng.uniform(0.15, 0.5)) * tt + float(rng.uniform(0, 6.28)))
                         + float(dyn["vib_depth"][spk]) * torch.sin(2 * np.pi * float(dyn["vib_rate"][spk]) * tt + float(rng.uniform(0, 6.28))))
        phase = 2 * np.pi * torch.cumsum(f0.double(), 0).float() / SR
        # vowel-like segments of 90-280 ms
        seg_bounds = [0]
        while seg_bounds[-1] < n:
            seg_bounds.append(seg_bounds[-1] + max(320, int(rng.uniform(0.7, 1.3) * float(dyn["vowel_s"][spk]) * SR)))
        n_seg = len(seg_bounds) - 1
        vowel_idx = rng.integers(len(_VOWELS), size=n_seg)
        if "vowels" in dyn:
            vowel_idx = dyn["vowels"][spk][vowel_idx % 3]
        formants = _VOWELS[vowel_idx] * tracts[spk] * (1.0 + 0.02 * rng.standard_normal((n_seg, 3)))
        hf = (np.arange(1, n_harm + 1)[None, :] * f0s[spk])[:, :, None]  # [1,H,1]
        amp = (1.0 / (1.0 + ((hf - formants[:, None, :]) / _BW[None, None, :]) ** 2)).sum(-1)  # [n_seg,H]
        amp = amp / (np.arange(1, n_harm + 1)[None, :] ** tilts[spk])
        amp = amp / np.sqrt((amp ** 2).sum(1, keepdims=True))
        amp_t = torch.tensor(amp.T, dtype=torch.float32)  # [H,n_seg]
        seg_of_sample = torch.bucketize(torch.arange(n), torch.tensor(seg_bounds[1:-1], dtype=torch.long), right=True)
        sig = (amp_t[:, seg_of_sample] * torch.sin(harm * phase.unsqueeze(0))).sum(0)
        depth = float(dyn["am_depth"][spk])
        jitter = float(rng.uniform(0.93, 1.07))
        env = (1.0 - depth) + depth * torch.sin(2 * np.pi * float(dyn["am_rate"][spk]) * (1.0 if "vowels" in dyn else jitter) * tt + float(rng.uniform(0, 6.28)))
        fade = torch.clamp(torch.minimum(tt, tt.flip(0)) / 0.02, max=1.0)
        level = 0.08 * (0.8 + 0.4 * float(rng.random()))
        sig = level * sig * env * fade + 0.0025 * torch.randn(n, generator=gen)
        wav[i0:i1] += sig
    return wav.clamp_(-1.0, 1.0).numpy().astype(np.float32), turns


def write_wav(path: str, wav: np.ndarray, pcm16: bool = False) -> None:
    """16 kHz mono WAV: float32 (as diarize.py:191-196 writes through torchaudio.save) or
    int16 PCM (as nemo_process.py:24-28 writes through pydub)."""
    from scipy.io import wavfile

    if pcm16:
        wavfile.write(path, SR, np.round(np.clip(wav, -1, 1) * 32767.0).astype(np.int16))
    else:
        wavfile.write(path, SR, wav.astype(np.float32))


def write_rttm(path: str, uniq_id: str, turns) -> None:
    with open(path, "w") as f:
        for st, en, spk in turns:
            f.write(f"SPEAKER {uniq_id} 1 {st:.3f} {en - st:.3f} <NA> <NA> spk{spk} <NA> <NA>\n")


def write_manifest(path: str, entries) -> None:
    """One JSON object per line with the keys helpers.py:267-275 writes."""
    with open(path, "w") as f:
        for e in entries:
            meta = {
                "audio_filepath": e["audio_filepath"],
                "offset": 0,
                "duration": None,
                "label": "infer",
                "text": "-",
                "rttm_filepath": e.get("rttm_filepath"),
                "uem_filepath": None,
            }
            if "num_speakers" in e:
                meta["num_speakers"] = e["num_speakers"]
            json.dump(meta, f)
            f.write("\n")


def make_session(work_dir: str, name: str, duration_s: float, n_speakers: int, seed: int, pcm16: bool = False):
    """Write <work_dir>/<name>.wav + .rttm; return (wav_path, rttm_path, waveform, turns)."""
    os.makedirs(work_dir, exist_ok=True)
    wav, turns = synth_recording(duration_s, n_speakers, seed)
    wav_path = os.path.join(work_dir, name + ".wav")
    rttm_path = os.path.join(work_dir, name + ".rttm")
    write_wav(wav_path, wav, pcm16=pcm16)
    write_rttm(rttm_path, name, turns)
    return wav_path, rttm_path, wav, turns


# ------------------------------------------------------------------------------------------ sessions
class AttrDict(dict):
    """Minimal OmegaConf-DictConfig stand-in (attribute access on nested dicts); the product's `config.as_config` and the
    oracle's `_get` both accept it."""

    def __init__(self, data=None):
        super().__init__()
        for k, v in dict(data or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v)

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v


CONF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "whisper_nemo_b200", "conf")


def load_domain_config(domain: str) -> AttrDict:
    """diar_infer_<domain>.yaml (the reference's nemo_msdd_configs/*.yaml schema, bundled with the product as data)."""
    import yaml

    with open(os.path.join(CONF_DIR, f"diar_infer_{domain}.yaml")) as f:
        return AttrDict(yaml.safe_load(f))


def make_session_cfg(tmp_dir, domain, duration_s, n_speakers, seed, name="mono_file", pcm16=False, **overrides):
    """Synthetic recording + manifest + config (oracle VAD from the ground-truth RTTM).  Returns (cfg, waveform, turns)."""
    wav_path, rttm_path, wav, turns = make_session(str(tmp_dir), name, duration_s, n_speakers, seed, pcm16=pcm16)
    cfg = load_domain_config(domain)
    man = os.path.join(str(tmp_dir), "input_manifest.json")
    write_manifest(man, [{"audio_filepath": wav_path, "rttm_filepath": rttm_path}])
    cfg.diarizer.manifest_filepath = man
    cfg.diarizer.out_dir = str(tmp_dir)
    cfg.diarizer.oracle_vad = True
    for k, v in overrides.items():
        cfg.diarizer.clustering.parameters[k] = v
    return cfg, wav, turns


# ------------------------------------------------------------------------------------------ metrics
def best_permutation_agreement(a, b) -> float:
    """Fraction of items on which labelings a and b agree under the best one-to-one label mapping."""
    from scipy.optimize import linear_sum_assignment

    a, b = np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64)
    assert a.shape == b.shape
    k = max(int(a.max()) + 1, int(b.max()) + 1)
    cont = np.zeros((k, k), dtype=np.int64)
    np.add.at(cont, (a, b), 1)
    r, c = linear_sum_assignment(-cont)
    return cont[r, c].sum() / len(a)


def rttm_frames(path, step=0.01):
    """RTTM -> {speaker: set of 10 ms frame indices} for a permutation-invariant comparison of two RTTMs."""
    out = {}
    with open(path) as f:
        for line in f:
            fld = line.split()
            if not fld:
                continue
            st, du, spk = float(fld[3]), float(fld[4]), fld[7]
            out.setdefault(spk, set()).update(range(int(round(st / step)), int(round((st + du) / step))))
    return out


def rttm_der_between(path_a, path_b) -> float:
    """Speaker-confusion + miss + false-alarm time between two RTTMs over their union, best label mapping
    (a small permutation-invariant DER; 0.0 means identical turns up to speaker renaming)."""
    from scipy.optimize import linear_sum_assignment

    fa, fb = rttm_frames(path_a), rttm_frames(path_b)
    ka, kb = sorted(fa), sorted(fb)
    k = max(len(ka), len(kb))
    overlap = np.zeros((k, k))
    for i, sa in enumerate(ka):
        for j, sb in enumerate(kb):
            overlap[i, j] = len(fa[sa] & fb[sb])
    r, c = linear_sum_assignment(-overlap)
    matched = overlap[r, c].sum()
    total = max(sum(len(v) for v in fa.values()), sum(len(v) for v in fb.values()))
    return 1.0 - matched / max(total, 1)
