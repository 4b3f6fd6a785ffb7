"""Shared helpers of the test-suite (not product code)."""
import os

import numpy as np
import torch


def best_permutation_agreement(a, b) -> float:
    """Fraction of items on which labelings a and b agree under the best one-to-one label mapping."""
    from scipy.optimize import linear_sum_assignment

    a, b = np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64)
    assert a.shape == b.shape
    ka, kb = int(a.max()) + 1, int(b.max()) + 1
    k = max(ka, kb)
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


def make_session_cfg(tmp_dir, domain, duration_s, n_speakers, seed, name="mono_file", pcm16=False, **overrides):
    """Synthetic recording + manifest + config (oracle VAD from the ground-truth RTTM)."""
    from whisper_nemo_b200 import config, synth

    wav_path, rttm_path, wav, turns = synth.make_session(str(tmp_dir), name, duration_s, n_speakers, seed, pcm16=pcm16)
    cfg = config.load_config(domain)
    man = os.path.join(str(tmp_dir), "input_manifest.json")
    synth.write_manifest(man, [{"audio_filepath": wav_path, "rttm_filepath": rttm_path}])
    cfg.diarizer.manifest_filepath = man
    cfg.diarizer.out_dir = str(tmp_dir)
    cfg.diarizer.oracle_vad = True
    for k, v in overrides.items():
        cfg.diarizer.clustering.parameters[k] = v
    return cfg, wav, turns


def synthetic_multiscale_embeddings(duration_s, scales, n_speakers, seed, turn_s=12.0, dim=192, noise=1.0, sep=2.0):
    """Clustered random embeddings on a multi-scale window grid: (embeddings [sum N_s, dim], timestamps [sum N_s, 2] fp32,
    counts [S], true base labels)."""
    import math

    g = torch.Generator().manual_seed(seed)
    centers = torch.randn(n_speakers, dim, generator=g) * sep
    order = torch.randint(0, n_speakers, (int(duration_s / turn_s) + 2,), generator=g)
    embs, stamps, counts, base_lab = [], [], [], None
    for w, s in scales:
        n = int(math.ceil((duration_s - w) / s)) + 1
        t0 = torch.arange(n, dtype=torch.float64) * s
        ts = torch.stack([t0, torch.clamp(t0 + w, max=duration_s)], 1)
        spk = order[((t0 + w / 2) / turn_s).long()]
        embs.append(centers[spk] + noise * torch.randn(n, dim, generator=g))
        stamps.append(ts.float())
        counts.append(n)
        base_lab = spk
    return torch.cat(embs), torch.cat(stamps), torch.tensor(counts), base_lab.numpy()
