"""Development aid (GPU box): diarize one bench workload on the B200 path and save what an offline (CPU-only) analysis of
label differences needs -- all-scale embeddings, timestamps, counts, final labels, per-chunk over-clustering labels.

    python tools/dump_recording.py meeting_1h gpurun_out/meeting_1h_gpu.npz
"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import SEED, WORKLOADS  # noqa: E402
from tools.workload import make_session_cfg  # noqa: E402
from whisper_nemo_b200 import ClusteringDiarizer, checkpoint  # noqa: E402


def main():
    workload, out = sys.argv[1], sys.argv[2]
    domain, seconds, speakers, _ = WORKLOADS[workload]
    with tempfile.TemporaryDirectory() as tmp:
        cfg, _, _ = make_session_cfg(tmp, domain, seconds, speakers, SEED)
        diar = ClusteringDiarizer(cfg=cfg, speaker_model=checkpoint.seeded())
        diar.diarize()
        e = diar.embs_and_timestamps["mono_file"]
        r = diar.results["mono_file"]
        sc = diar._last_clusterers["mono_file"]
        chunks = {f"chunk{w}_offset": np.asarray(off) for w, (off, _) in sc.chunk_labels.items()}
        chunks.update({f"chunk{w}_labels": y.numpy() for w, (_, y) in sc.chunk_labels.items()})
        np.savez_compressed(out, embeddings=e["embeddings"].cpu().numpy(), timestamps=e["timestamps"].numpy(),
                            counts=e["multiscale_segment_counts"].numpy(), weights=e["multiscale_weights"].numpy(), labels=r["labels"], **chunks)
        print(f"{out}: {e['embeddings'].shape[0]} windows, {len(r['labels'])} base, {len(set(r['labels'].tolist()))} speakers, "
              f"{len(sc.chunk_labels)} chunks, stages {diar.stage_ms}")


if __name__ == "__main__":
    main()
