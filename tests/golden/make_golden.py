"""Generates the regression vectors under tests/golden/ from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  They pin the ORACLE against accidental edits -- they are this repo's own
restatement of upstream NeMo, not outputs of NeMo (which cannot be run here: parity is unpinned, see oracle/__init__.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def featurizer_case():
    from oracle.features import FilterbankFeatures

    g = torch.Generator().manual_seed(11)
    t = torch.arange(24000) / 16000.0
    sig = 0.3 * torch.sin(2 * np.pi * 180.0 * t) + 0.1 * torch.sin(2 * np.pi * 1230.0 * t + 1.0) + 0.02 * torch.randn(24000, generator=g)
    audio = torch.stack([sig, sig.flip(0)])
    feats, lens = FilterbankFeatures()(audio, torch.tensor([24000, 24000]))
    return {"audio": audio.numpy(), "feats_sub": feats[:, ::8, ::10].numpy(), "feat_len": lens.numpy(),
            "feat_mean": feats.mean(dim=(1, 2)).numpy(), "feat_abs_sum": feats.abs().sum(dim=(1, 2)).numpy()}


def clustering_case():
    from oracle import offline_clustering as oc
    from tests.util import synthetic_multiscale_embeddings

    scales = [(1.5, 0.75), (1.0, 0.5), (0.5, 0.25)]
    embs, stamps, counts, truth = synthetic_multiscale_embeddings(120.0, scales, 3, seed=21, turn_s=9.0, dim=32)
    sc = oc.SpeakerClustering()
    labels = sc.forward_infer(embs, stamps, counts, torch.ones(1, 3), max_num_speakers=8, max_rp_threshold=0.25, sparse_search_volume=30)
    fused = sc.fused_affinity
    return {"embs": embs.numpy(), "stamps": stamps.numpy(), "counts": counts.numpy(), "labels": labels.numpy(),
            "est": np.int64(sc.debug["est_num_of_spk"]), "p_hat": np.int64(sc.debug["p_hat"]), "g_p": sc.debug["g_p"].numpy(),
            "fused_sub": fused[::7, ::7].numpy(), "fused_sum": np.float64(fused.double().sum().item())}


def rttm_case():
    from oracle import speaker_utils as su

    ts = torch.tensor([[0.0, 1.5], [0.75, 2.25], [1.5, 3.0], [2.25, 3.4], [5.0, 6.5], [5.75, 7.25], [6.5, 8.0], [8.0, 9.5]])
    labels = [0, 0, 1, 1, 1, 0, 0, 0]
    turns, lines = su.generate_cluster_labels(ts, labels)
    return {"turns": np.array(turns), "lines": np.array(lines)}


if __name__ == "__main__":
    torch.set_num_threads(4)
    np.savez_compressed(os.path.join(HERE, "featurizer.npz"), **featurizer_case())
    np.savez_compressed(os.path.join(HERE, "clustering.npz"), **clustering_case())
    np.savez_compressed(os.path.join(HERE, "rttm.npz"), **rttm_case())
    print("written", os.listdir(HERE))
