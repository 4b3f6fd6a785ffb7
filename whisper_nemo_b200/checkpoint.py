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
Its running statistics are therefore a committed fixture (`data/titanet_large_random_seed<seed>_embbn.pt`,
two 6144-vectors) measured once by `tools/make_embbn_fixture.py`, which pushes seeded synthetic
speech through the CPU oracle's fp32 encoder (a trained checkpoint carries the equivalent numbers
from training).  The weights are thus reproducible without a GPU and identical for the product, the
oracle and the benchmark's reference arm; neither the init distribution nor the calibration enters
any parity claim.  This module is pure Python / CPU torch: importing it never loads libb200d.so.
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
_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def embbn_fixture_path(seed: int = DEFAULT_SEED) -> str:
    return os.path.join(_DATA_DIR, f"titanet_large_random_seed{seed}_embbn.pt")


def seeded(seed: int = None) -> Dict[str, torch.Tensor]:
    """The fixed-seed random-init TitaNet-L (`model_path: titanet_large_random`): `random_init_titanet_large(seed)` with
    the embedding BatchNorm's running statistics taken from the committed fixture of that seed."""
    if seed is None:
        seed = int(os.environ.get("B200D_TITANET_SEED", DEFAULT_SEED))
    key = f"seeded{seed}"
    if key not in _CACHE:
        path = embbn_fixture_path(seed)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not found: generate it with `python tools/make_embbn_fixture.py --seed {seed}` (CPU, ~1 min)")
        fix = torch.load(path, map_location="cpu", weights_only=True)
        sd = random_init_titanet_large(seed)
        sd["decoder.emb_layers.0.0.running_mean"] = fix["running_mean"].float().clone()
        sd["decoder.emb_layers.0.0.running_var"] = fix["running_var"].float().clone()
        _CACHE[key] = sd
    return _CACHE[key]


def calibrated(device=None, seed: int = None) -> Dict[str, torch.Tensor]:
    """Round-1 name of `seeded` (the calibration ran on the GPU then; `device` is ignored now)."""
    return seeded(seed)


def _load_nemo_archive(path: str) -> Dict[str, torch.Tensor]:
    """A `.nemo` file is a (possibly gzipped) tar holding `model_weights.ckpt` (torch.save of the state_dict) next to
    `model_config.yaml` -- NeMo's normal `model_path` format (SaveRestoreConnector)."""
    import io
    import tarfile

    with tarfile.open(path, "r:*") as tar:
        member = next((m for m in tar.getmembers() if os.path.basename(m.name) == "model_weights.ckpt"), None)
        if member is None:
            raise ValueError(f"{path}: no model_weights.ckpt inside the .nemo archive")
        data = tar.extractfile(member).read()
    return torch.load(io.BytesIO(data), map_location="cpu", weights_only=True)


def _find_local_titanet() -> str:
    """A TitaNet-L checkpoint already on this machine: $B200D_TITANET_CKPT, or what NeMo's `from_pretrained` left in its
    cache (~/.cache/torch/NeMo/**/titanet-l*.nemo)."""
    env = os.environ.get("B200D_TITANET_CKPT")
    if env:
        return env
    import glob

    hits = sorted(glob.glob(os.path.join(os.path.expanduser("~"), ".cache", "torch", "NeMo", "**", "titanet-l*.nemo"), recursive=True))
    return hits[-1] if hits else ""


def resolve(model_path, device=None) -> Dict[str, torch.Tensor]:
    """`diarizer.speaker_embeddings.model_path` -> TitaNet-L state_dict.
      * a local file: a `.nemo` archive, or a `torch.save`d state_dict / {'state_dict': ...};
      * `titanet_large_random`: the fixed-seed random init (BASELINE.json's benchmark weights; seed via B200D_TITANET_SEED);
      * `titanet_large` / None (what the reference's create_config sets, helpers.py:281,290 -- an NGC download upstream):
        $B200D_TITANET_CKPT or NeMo's local download cache when one exists; otherwise, because there is no network here,
        the random init WITH A WARNING (speaker labels from untrained weights are meaningless), or an error when
        B200D_STRICT_WEIGHTS=1."""
    if model_path and os.path.exists(str(model_path)):
        path = str(model_path)
        if path.endswith(".nemo"):
            return _load_nemo_archive(path)
        obj = torch.load(path, map_location="cpu", weights_only=True)
        return obj.get("state_dict", obj) if isinstance(obj, dict) and "state_dict" in obj else obj
    if model_path == "titanet_large_random":
        return seeded()
    if model_path in (None, "titanet_large"):
        local = _find_local_titanet()
        if local:
            return resolve(local, device)
        if os.environ.get("B200D_STRICT_WEIGHTS") == "1":
            raise FileNotFoundError("speaker_embeddings.model_path='titanet_large': no local checkpoint ($B200D_TITANET_CKPT / NeMo cache) and "
                                    "no network for the NGC download")
        import warnings

        warnings.warn("speaker_embeddings.model_path='titanet_large' could not be resolved to a checkpoint (no $B200D_TITANET_CKPT, nothing in "
                      "~/.cache/torch/NeMo, no network): using the RANDOM-INIT fixed-seed TitaNet-L of the benchmark.  Speaker labels from "
                      "untrained weights are NOT meaningful; point model_path at a .nemo / state_dict file for real diarization "
                      "(B200D_STRICT_WEIGHTS=1 turns this warning into an error).", RuntimeWarning, stacklevel=2)
        return seeded()
    raise FileNotFoundError(f"speaker_embeddings.model_path={model_path!r}: not a local checkpoint and there is no network for NGC models")
