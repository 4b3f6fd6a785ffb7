"""Shared helpers of the test-suite (not product code)."""
import os

import numpy as np
import torch


from tools.workload import best_permutation_agreement, make_session_cfg, rttm_der_between, rttm_frames  # noqa: F401,E402


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
