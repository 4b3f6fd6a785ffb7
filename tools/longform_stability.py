"""Development aid (GPU box): how well-posed is a 1-hour benchmark recording for the long-form clustering path?

    python tools/longform_stability.py meeting 3600 8 v1:100 v2:100 v2:101

For every style:seed the recording is embedded on the B200 path, then the CPU ORACLE clusters those embeddings three times --
upstream's fp32 eigh, float64 eigh (oracle.switches.SPECTRAL_EIGH_FP64) and fp32 on embeddings perturbed by 3e-3 relative --
and the speaker counts / label agreements between the three, the device labels and the ground truth are printed.  A
recording is a usable parity instance when all of them agree."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import switches  # noqa: E402
from oracle.longform_clustering import LongFormSpeakerClustering as OracleLF  # noqa: E402
from tools import workload  # noqa: E402
from whisper_nemo_b200 import ClusteringDiarizer, checkpoint  # noqa: E402


def main():
    domain, seconds, speakers = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
    torch.set_num_threads(os.cpu_count() or 8)
    weights = checkpoint.seeded()
    for case in sys.argv[4:]:
        style, seed = case.split(":")
        workload.STYLE = style
        with tempfile.TemporaryDirectory() as tmp:
            cfg, _, turns = workload.make_session_cfg(tmp, domain, seconds, speakers, int(seed))
            diar = ClusteringDiarizer(cfg=cfg, speaker_model=weights)
            diar.diarize()
        e = diar.embs_and_timestamps["mono_file"]
        r = diar.results["mono_file"]
        clus = cfg.diarizer.clustering.parameters
        kw = dict(max_num_speakers=int(clus.max_num_speakers), max_rp_threshold=float(clus.max_rp_threshold), sparse_search_volume=int(clus.sparse_search_volume),
                  chunk_cluster_count=clus.chunk_cluster_count, embeddings_per_chunk=clus.embeddings_per_chunk)
        emb = e["embeddings"].cpu()
        mid = r["timestamps"].numpy().mean(1)
        truth = np.full(len(mid), -1)
        for a, b, k in turns:
            truth[(mid >= a) & (mid <= b)] = k
        ok = truth >= 0

        def oracle(x, fp64=False):
            switches.SPECTRAL_EIGH_FP64 = fp64
            try:
                return OracleLF().forward_infer(x, e["timestamps"], e["multiscale_segment_counts"], e["multiscale_weights"], **kw).numpy()
            finally:
                switches.SPECTRAL_EIGH_FP64 = False

        t0 = time.time()
        gen = torch.Generator().manual_seed(0)
        runs = {"b200": r["labels"], "oracle32": oracle(emb)}
        if os.environ.get("STABILITY_FP64", "0") == "1":
            runs["oracle64"] = oracle(emb, True)
        for i in range(int(os.environ.get("STABILITY_PERTURBATIONS", "2"))):
            runs[f"oracle32_pert{i}"] = oracle(emb * (1.0 + 3e-3 * torch.randn(emb.shape, generator=gen)))
        ag = workload.best_permutation_agreement
        print(f"{case}: N={len(mid)} " + " ".join(f"{k}: k={len(set(v.tolist()))} purity={ag(v[ok], truth[ok]):.4f}" for k, v in runs.items()) +
              " | agreement with oracle32: " + " ".join(f"{k}={ag(v, runs['oracle32']):.5f}" for k, v in runs.items() if k != "oracle32") +
              f" ({time.time() - t0:.0f} s of oracle)", flush=True)


if __name__ == "__main__":
    main()
