"""Multi-GPU check (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py

    ... tools/multi_gpu_check.py 14400 telephonic fullmatrix

One long recording with its windows sharded across the ranks (embeddings all-gathered over NCCL; long-form chunks dealt to
the ranks, or -- `fullmatrix`: embeddings_per_chunk raised above the recording's length -- the N x N affinity, its graph and
the eigensolver's products row-sharded with peer stores over NVLink, rowshard.py) must give exactly the labels of the
single-GPU run; prints the device time and the stage split of both."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from tools.workload import make_session_cfg
from whisper_nemo_b200 import ClusteringDiarizer, checkpoint


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 1200.0
    domain = sys.argv[2] if len(sys.argv) > 2 else "telephonic"
    fullmatrix = len(sys.argv) > 3 and sys.argv[3] == "fullmatrix"
    weights = checkpoint.seeded()
    work = os.path.join(tempfile.gettempdir(), f"b200d_mgpu_r{rank}")
    cfg, _, _ = make_session_cfg(work, domain, seconds, 4 if seconds < 3000 else 8, seed=7)
    if fullmatrix:
        cfg.diarizer.clustering.parameters.embeddings_per_chunk = 10 ** 7
    results = {}
    for mode in ("single", "sharded"):
        diar = ClusteringDiarizer(cfg=cfg, speaker_model=weights, shard_windows=(mode == "sharded"))
        diar._prepare()
        wav = diar._wav_host.to(dev)
        for _ in range(2):
            diar.run_device(wav_dev=wav, timers=False)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        labels = diar.run_device(wav_dev=wav, timers=False)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        emb = diar.embs_and_timestamps["mono_file"]["embeddings"].clone()
        diar.run_device(wav_dev=wav, timers=True)
        results[mode] = (labels["mono_file"], t.item(), emb, dict(diar.stage_ms), dict(diar.results["mono_file"]["debug"]))
    same = np.array_equal(results["single"][0], results["sharded"][0])
    emb_diff = (results["single"][2] - results["sharded"][2]).abs().max().item()
    gathered = [None] * world
    dist.all_gather_object(gathered, results["sharded"][0].tolist())
    all_equal = all(g == gathered[0] for g in gathered)
    if rank == 0:
        dbg = {k: v for k, v in results["sharded"][4].items() if k in ("n_clusters", "p_hat", "row_sharded")}
        print(f"world {world}: {seconds:.0f} s {domain}{' full-matrix path' if fullmatrix else ''}: N={len(results['single'][0])} single {results['single'][1]:.1f} ms "
              f"{({k: round(v, 1) for k, v in results['single'][3].items()})}, sharded {results['sharded'][1]:.1f} ms (max over ranks) "
              f"{({k: round(v, 1) for k, v in results['sharded'][3].items()})} {dbg}; labels identical to single-GPU: {same}; "
              f"identical on all ranks: {all_equal}; max |emb diff| {emb_diff:.2e}")
        assert same and all_equal and emb_diff == 0.0
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
