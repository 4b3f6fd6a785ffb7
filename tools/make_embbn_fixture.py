"""Generate whisper_nemo_b200/data/titanet_large_random_seed<seed>_embbn.pt: the running statistics of the BatchNorm in
front of TitaNet-L's 6144 -> 192 embedding projection for the fixed-seed random-init weights (checkpoint.py).

    python tools/make_embbn_fixture.py [--seed 1234]

Seeded synthetic speech (tools/workload.py) is cut into 1.5 s windows inside speaker turns and pushed through the CPU
oracle's fp32 preprocessor + encoder + attentive pooling; the per-channel mean / biased variance of the pooled
[mu | sigma] vectors become running_mean / running_var.  CPU only (about a minute): the weights of the benchmark are
reproducible without a GPU and without the product's kernels."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def calibration_windows(seed: int, n_speakers: int = 8, duration_s: float = 192.0, window_s: float = 1.5):
    from tools import workload

    wav, turns = workload.synth_recording(duration_s, n_speakers, seed=seed + 77)
    n = int(window_s * workload.SR)
    starts = []
    for a, b, _ in turns:
        t = a
        while t + window_s <= b:
            starts.append(int(t * workload.SR))
            t += window_s
    return wav, starts, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1234)
    args = ap.parse_args()
    from oracle.titanet import TitaNetL
    from whisper_nemo_b200 import checkpoint

    sd = checkpoint.random_init_titanet_large(args.seed)
    model = TitaNetL(compute_logits=False)
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    model.eval()
    wav, starts, n = calibration_windows(args.seed)
    audio = torch.stack([torch.from_numpy(wav[s : s + n]) for s in starts])
    lens = torch.full((len(starts),), n, dtype=torch.long)
    with torch.no_grad():
        feats, flens = model.preprocessor(audio, lens)
        enc, elens = model.encoder(feats, flens)
        pool = model.decoder._pooling(enc, elens).squeeze(-1)  # [n, 6144] = [mu | sigma]
    out = {"running_mean": pool.mean(dim=0).float(), "running_var": pool.var(dim=0, unbiased=False).float(),
           "windows": torch.tensor(len(starts)), "seed": torch.tensor(args.seed)}
    path = checkpoint.embbn_fixture_path(args.seed)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(out, path)
    print(f"{path}: {len(starts)} windows, mean of means {out['running_mean'].mean():.4f}, mean var {out['running_var'].mean():.3e}")


if __name__ == "__main__":
    main()
