"""TitaNet-L weights for the B200 path.

The reference asks NeMo for the pretrained `titanet_large` checkpoint
(helpers.py:281,290: `speaker_embeddings.model_path = "titanet_large"`, an NGC download).
There is no network here and BASELINE.json fixes "random-init, fixed-seed TitaNet weights", so
`resolve("titanet_large")` returns a deterministic random initialisation with upstream's exact
parameter names and shapes (`encoder.encoder.{i}.mconv.{j}...`, `decoder._pooling...`,
`decoder.emb_layers.0...`), which is also what a converted NeMo checkpoint (`torch.save` of the
model's state_dict) looks like; a path to such a file is loaded instead.

Convolutions are variance-preserving Gaussians and every BatchNorm gets random affine
parameters plus running statistics drawn around the moments such a network produces on
per-feature-normalised log-mel input, so activations stay O(1) through all 14 sub-blocks and
the running statistics are not an identity (BN folding is exercised).  One BatchNorm cannot be
drawn blindly: the one in front of the 6144 -> 192 embedding projection sees pooled statistics
whose per-channel offsets are ~10x larger than their variation between speakers, and a random
guess of its running mean would bury the speaker information under a constant vector.
`calibrated()` therefore sets that layer's running statistics from one pass of seeded synthetic
speech through the B200 encoder itself (this package's own kernels; a trained checkpoint carries
the equivalent numbers from training).  Parity tests load the resulting state_dict into the CPU
oracle, so neither the init distribution nor the calibration enters any parity claim.
"""
import os
from typing import Dict

import torch

# titanet-large.yaml `jasper:` table: (filters, repeat, kernel, residual)
TITANET_L_BLOCKS = [(1024, 1, 3, False), (1024, 3, 7, True), (1024, 3, 11, True), (1024, 3, 15, True), (3072, 1, 1, False)]
FEAT_IN, EMB_SIZE, ATTN_CHANNELS, SE_REDUCTION = 80, 192, 128, 8
DEFAULT_SEED = 1234

# second moment of the activations feeding a pointwise conv (1.0 for the normalised features, ~0.6 after ReLU)
_M2_RELU = 0.6
# second moment of what the attention TDNN sees ([x | mean | std] of the 3072-channel encoder output)
_M2_TDNN = 0.04
# placeholder pooled statistics for the embedding BatchNorm (replaced by calibrated())
_POOL_MEAN, _POOL_MEAN_SPREAD, _POOL_STD, _POOL_STD_SPREAD = 0.13, 0.03, 0.07, 0.02


def _bn(sd, prefix, c, gen, mean_center=0.0, mean_spread=0.1, var_lo=0.6, var_hi=1.4):
    sd[prefix + ".weight"] = 0.75 + 0.5 * torch.rand(c, generator=gen)
    sd[prefix + ".bias"] = 0.2 * torch.randn(c, generator=gen)
    sd[prefix + ".running_mean"] = mean_center + mean_spread * torch.randn(c, generator=gen)
    sd[prefix + ".running_var"] = var_lo + (var_hi - var_lo) * torch.rand(c, generator=gen)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(1)


def random_init_titanet_large(seed: int = DEFAULT_SEED) -> Dict[str, torch.Tensor]:
    gen = torch.Generator().manual_seed(seed)
    randn = lambda *shape: torch.randn(*shape, generator=gen)
    sd: Dict[str, torch.Tensor] = {}
    cin = FEAT_IN
    m2 = 1.0
    for i, (filters, repeat, k, residual) in enumerate(TITANET_L_BLOCKS):
        pre = f"encoder.encoder.{i}."
        block_in, block_m2 = cin, m2
        j = 0
        for r in range(repeat):
            sd[f"{pre}mconv.{j}.conv.weight"] = randn(cin, 1, k) * (1.0 / k) ** 0.5  # depthwise
            sd[f"{pre}mconv.{j + 1}.conv.weight"] = randn(filters, cin, 1) * (1.0 / (cin * m2)) ** 0.5  # pointwise
            _bn(sd, f"{pre}mconv.{j + 2}", filters, gen)
            j += 3
            if r < repeat - 1:
                j += 1  # the ReLU(+dropout) slot of upstream's nn.ModuleList numbering
            cin, m2 = filters, _M2_RELU
        sd[f"{pre}mconv.{j}.fc.0.weight"] = randn(filters // SE_REDUCTION, filters) * (1.0 / filters) ** 0.5
        sd[f"{pre}mconv.{j}.fc.2.weight"] = randn(filters, filters // SE_REDUCTION) * (2.0 / (filters // SE_REDUCTION)) ** 0.5
        if residual:
            sd[f"{pre}res.0.0.conv.weight"] = randn(filters, block_in, 1) * (1.0 / (block_in * block_m2)) ** 0.5
            _bn(sd, f"{pre}res.0.1", filters, gen)
        m2 = _M2_RELU
    c_enc = cin
    ap = "decoder._pooling.attention_layer."
    sd[ap + "0.conv_layer.weight"] = randn(ATTN_CHANNELS, 3 * c_enc, 1) * (1.0 / (3 * c_enc * _M2_TDNN)) ** 0.5
    sd[ap + "0.conv_layer.bias"] = 0.1 * randn(ATTN_CHANNELS)
    _bn(sd, ap + "0.bn", ATTN_CHANNELS, gen, mean_center=0.35, var_lo=0.2, var_hi=0.5)
    sd[ap + "2.weight"] = randn(c_enc, ATTN_CHANNELS, 1) * (2.0 / ATTN_CHANNELS) ** 0.5
    sd[ap + "2.bias"] = 0.1 * randn(c_enc)
    # embedding BatchNorm over [mean | std] of the attentive pooling
    bn = "decoder.emb_layers.0.0"
    _bn(sd, bn, 2 * c_enc, gen)
    sd[bn + ".running_mean"] = torch.cat([_POOL_MEAN + _POOL_MEAN_SPREAD * randn(c_enc), _POOL_STD + _POOL_STD_SPREAD * randn(c_enc)])
    sd[bn + ".running_var"] = 1e-4 + 2e-4 * torch.rand(2 * c_enc, generator=gen)
    # embedding scale of a trained TitaNet: per-dimension std ~0.02; upstream's enhanced speaker
    # count (addAnchorEmb, sigma = 50 x the per-dimension std against unit-variance anchors) assumes that scale
    sd["decoder.emb_layers.0.1.weight"] = randn(EMB_SIZE, 2 * c_enc, 1) * (0.02 ** 2 / (2 * c_enc)) ** 0.5
    sd["decoder.emb_layers.0.1.bias"] = 0.005 * randn(EMB_SIZE)
    return sd


_CACHE: Dict[str, Dict[str, torch.Tensor]] = {}


def calibration_windows(seed: int = DEFAULT_SEED, n_speakers: int = 8, duration_s: float = 64.0, window_s: float = 1.5):
    """Seeded synthetic speech cut into 1.5 s windows (inside speaker turns): (waveform float32, starts, length)."""
    from . import synth

    wav, turns = synth.synth_recording(duration_s, n_speakers, seed=seed + 77)
    n = int(window_s * synth.SR)
    starts = []
    for a, b, _ in turns:
        t = a
        while t + window_s <= b:
            starts.append(int(t * synth.SR))
            t += window_s
    return wav, starts, n


def calibrate_embedding_bn(state_dict: Dict[str, torch.Tensor], device, seed: int = DEFAULT_SEED) -> Dict[str, torch.Tensor]:
    """Return a copy of `state_dict` whose `decoder.emb_layers.0.0.running_{mean,var}` are the statistics of the
    attentive-pooling output over seeded synthetic speech, measured by running the B200 encoder (TitaNetB200)."""
    from .titanet import TitaNetB200

    wav, starts, n = calibration_windows(seed)
    net = TitaNetB200(state_dict, device, max_frames=8192)
    wav_d = torch.from_numpy(wav).to(net.device)
    st = torch.tensor(starts, dtype=torch.int32, device=net.device)
    ln = torch.full((len(starts),), n, dtype=torch.int32, device=net.device)
    pools = []
    step = max(1, net.max_frames // (n // 160 + 1))  # an upper bound on the frames per window in either featurizer variant
    for c0 in range(0, len(starts), step):
        taps = {}
        net.embed_segments(wav_d, st[c0 : c0 + step], ln[c0 : c0 + step], n, taps=taps)
        pools.append(taps["pool"].float())
    pool = torch.cat(pools)
    out = dict(state_dict)
    out["decoder.emb_layers.0.0.running_mean"] = pool.mean(dim=0).cpu()
    out["decoder.emb_layers.0.0.running_var"] = pool.var(dim=0, unbiased=False).cpu()
    return out


def calibrated(device, seed: int = None) -> Dict[str, torch.Tensor]:
    """The fixed-seed random-init TitaNet-L the diarizer uses for `model_path: titanet_large`."""
    if seed is None:
        seed = int(os.environ.get("B200D_TITANET_SEED", DEFAULT_SEED))
    key = f"cal{seed}"
    if key not in _CACHE:
        _CACHE[key] = calibrate_embedding_bn(random_init_titanet_large(seed), device, seed)
    return _CACHE[key]


def resolve(model_path, device=None) -> Dict[str, torch.Tensor]:
    """`diarizer.speaker_embeddings.model_path` -> TitaNet-L state_dict.  A local file is loaded
    (`torch.load`, a state_dict or {'state_dict': ...}); `titanet_large` / None gives the fixed-seed
    random init (seed overridable through B200D_TITANET_SEED), calibrated on `device`."""
    if model_path and os.path.exists(str(model_path)):
        obj = torch.load(str(model_path), map_location="cpu", weights_only=True)
        return obj.get("state_dict", obj) if isinstance(obj, dict) and "state_dict" in obj else obj
    if model_path in (None, "titanet_large", "titanet_large_random"):
        if device is None:
            raise ValueError("the random-init TitaNet-L is calibrated on the GPU: pass the device")
        return calibrated(device)
    raise FileNotFoundError(f"speaker_embeddings.model_path={model_path!r}: not a local checkpoint and there is no network for NGC models")
