"""The CPU oracle against its committed regression vectors, closed-form properties and independent torch / scipy /
sklearn / torchaudio computations.  (The reference holds no golden vector for this path -- parity is unpinned; these
pins are this repo's own, see tests/golden/make_golden.py.)"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import offline_clustering as oc
from oracle import speaker_utils as osu
from tests.util import best_permutation_agreement, synthetic_multiscale_embeddings

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_featurizer():
    from oracle.features import FilterbankFeatures

    g = np.load(os.path.join(GOLD, "featurizer.npz"))
    feats, lens = FilterbankFeatures()(torch.from_numpy(g["audio"]), torch.tensor([24000, 24000]))
    assert np.array_equal(lens.numpy(), g["feat_len"])
    assert np.abs(feats[:, ::8, ::10].numpy() - g["feats_sub"]).max() < 2e-4
    assert np.allclose(feats.abs().sum(dim=(1, 2)).numpy(), g["feat_abs_sum"], rtol=1e-5)


def test_golden_clustering():
    g = np.load(os.path.join(GOLD, "clustering.npz"))
    sc = oc.SpeakerClustering()
    labels = sc.forward_infer(torch.from_numpy(g["embs"]), torch.from_numpy(g["stamps"]), torch.from_numpy(g["counts"]), torch.ones(1, 3),
                              max_num_speakers=8, max_rp_threshold=0.25, sparse_search_volume=30)
    assert sc.debug["est_num_of_spk"] == int(g["est"]) and sc.debug["p_hat"] == int(g["p_hat"])
    assert np.abs(sc.fused_affinity[::7, ::7].numpy() - g["fused_sub"]).max() < 1e-5
    assert best_permutation_agreement(labels.numpy(), g["labels"]) == 1.0


def test_golden_rttm_lines():
    g = np.load(os.path.join(GOLD, "rttm.npz"))
    ts = torch.tensor([[0.0, 1.5], [0.75, 2.25], [1.5, 3.0], [2.25, 3.4], [5.0, 6.5], [5.75, 7.25], [6.5, 8.0], [8.0, 9.5]])
    turns, lines = osu.generate_cluster_labels(ts, [0, 0, 1, 1, 1, 0, 0, 0])
    assert turns == list(g["turns"]) and lines == list(g["lines"])


def test_mel_filterbank_matches_torchaudio():
    import torchaudio

    from oracle.features import librosa_mel

    ours = librosa_mel(16000, 512, 80)
    ta = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="slaney").t().numpy()
    assert np.abs(ours - ta).max() < 1e-6


def test_featurizer_matches_an_independent_stft_pipeline():
    from oracle.features import FilterbankFeatures, librosa_mel

    g = torch.Generator().manual_seed(3)
    x = 0.1 * torch.randn(2, 16000, generator=g)
    feats, lens = FilterbankFeatures()(x, torch.tensor([16000, 16000]))
    y = torch.cat([x[:, :1], x[:, 1:] - 0.97 * x[:, :-1]], dim=1)
    win = torch.hann_window(400, periodic=False)
    spec = torch.stft(y, 512, hop_length=160, win_length=400, window=win, center=True, pad_mode="reflect", return_complex=True)
    mel = torch.matmul(torch.from_numpy(librosa_mel()), spec.abs() ** 2)
    logmel = torch.log(mel + 2.0 ** -24)
    T = 16000 // 160 + 1
    logmel = logmel[:, :, :T]
    norm = (logmel - logmel.mean(2, keepdim=True)) / (logmel.std(2, keepdim=True) + 1e-5)
    assert int(lens[0]) == T
    assert (feats[:, :, :T] - norm).abs().max().item() < 1e-3
    assert feats.shape[2] % 16 == 0 and feats[:, :, T:].abs().max().item() == 0.0


def test_featurizer_matches_the_transformers_port_of_nemo(monkeypatch):
    """transformers' ParakeetFeatureExtractor is an independent port of NeMo's FilterbankFeatures (same 16 kHz /
    n_fft 512 / win 400 / hop 160 / 80 slaney mels / pre-emphasis 0.97 / 2^-24 log guard / per-feature normalisation)
    that ships in this image.  With the two version-dependent details switched to that port's choice (constant STFT
    padding, no "+ 1" in the frame count) the restatement reproduces its features; librosa is absent, so the port gets
    the mel filterbank that test_mel_filterbank_matches_torchaudio checks."""
    import types

    fe = pytest.importorskip("transformers.models.parakeet.feature_extraction_parakeet")
    from oracle import switches
    from oracle.features import FilterbankFeatures, librosa_mel

    monkeypatch.setattr(fe, "librosa", types.SimpleNamespace(filters=types.SimpleNamespace(
        mel=lambda sr, n_fft, n_mels, fmin, fmax, norm: librosa_mel(sr, n_fft, n_mels, fmin, fmax))), raising=False)
    try:
        extractor = fe.ParakeetFeatureExtractor()
    except Exception as exc:  # noqa: BLE001 -- the class refuses to build without its optional backends
        pytest.skip(f"ParakeetFeatureExtractor unavailable: {exc}")
    monkeypatch.setattr(switches, "STFT_PAD_MODE", "constant")
    monkeypatch.setattr(switches, "SEQ_LEN_PLUS_ONE", False)
    g = torch.Generator().manual_seed(9)
    for n in (24000, 8000, 12345):
        x = 0.1 * torch.randn(n, generator=g)
        theirs = extractor(x.numpy(), sampling_rate=16000, return_tensors="pt")
        want = theirs["input_features"][0]  # [T, 80]
        frames = int(theirs["attention_mask"][0].sum())
        ours, lens = FilterbankFeatures()(x[None], torch.tensor([n]))
        assert int(lens[0]) == frames
        err = (ours[0, :, :frames].t() - want[:frames]).abs().max().item()
        assert err < 1e-5, (n, err)  # measured: 0.0 (the same torch ops in the same order)
    # and the two details do matter: with the defaults the same input differs from the port
    monkeypatch.setattr(switches, "STFT_PAD_MODE", "reflect")
    monkeypatch.setattr(switches, "SEQ_LEN_PLUS_ONE", True)
    ours, lens = FilterbankFeatures()(x[None], torch.tensor([n]))
    assert int(lens[0]) == frames + 1


@pytest.mark.parametrize("dur,w,s", [(10.0, 1.5, 0.75), (600.0, 0.5, 0.25), (3.3, 3.0, 1.5), (0.4, 1.5, 0.75), (1.5, 1.5, 0.75)])
def test_subsegment_count_formula(dur, w, s):
    segs = osu.get_subsegments(2.0, w, s, dur)
    want = 1 if dur < w else math.ceil((dur - w) / s) + 1
    assert len(segs) == want
    assert abs(segs[-1][0] + segs[-1][1] - (2.0 + dur)) < 1e-9 or segs[-1][1] == w
    assert all(abs(b[0] - a[0] - s) < 1e-9 for a, b in zip(segs, segs[1:]))


@pytest.mark.parametrize("k,n_per", [(2, 60), (3, 50), (5, 40)])
def test_block_diagonal_affinity_gives_k_speakers(k, n_per):
    n = k * n_per
    g = torch.Generator().manual_seed(k)
    mat = 0.05 * torch.rand(n, n, generator=g)
    for b in range(k):
        mat[b * n_per : (b + 1) * n_per, b * n_per : (b + 1) * n_per] += 0.9
    mat = 0.5 * (mat + mat.t())
    mat.fill_diagonal_(1.0)
    sc = oc.SpeakerClustering()
    labels = sc.forward_unit_infer(mat, max_num_speakers=8, max_rp_threshold=0.25, sparse_search_volume=30)
    assert sc.debug["est_num_of_spk"] == k
    truth = np.repeat(np.arange(k), n_per)
    assert best_permutation_agreement(labels.numpy(), truth) == 1.0


def test_kmeans_agrees_with_sklearn_on_separated_blobs():
    from sklearn.cluster import KMeans

    g = torch.Generator().manual_seed(5)
    centers = torch.tensor([[0.0, 0.0], [6.0, 0.0], [0.0, 6.0], [6.0, 6.0]])
    lab = torch.randint(0, 4, (400,), generator=g)
    x = centers[lab] + 0.5 * torch.randn(400, 2, generator=g)
    ours = oc.kmeans_torch(x, 4).numpy()
    sk = KMeans(4, n_init=5, random_state=0).fit(x.numpy()).labels_
    assert best_permutation_agreement(ours, sk) == 1.0


def test_kmeans_plusplus_picks_sklearns_centers_on_the_same_draws():
    """kmeans_plusplus_torch is upstream's port of sklearn's k-means++ (greedy, 30 local trials).  Fed the same uniform
    draws and the same first index, the restatement must choose the centers sklearn's own implementation chooses."""
    from sklearn.cluster import _kmeans

    from oracle import offline_clustering as oc

    class Draws(np.random.RandomState):
        def __init__(self, first, uniforms):
            super().__init__(0)
            self.first, self.uniforms, self.pos = first, uniforms, 0

        def choice(self, n, p=None):
            return self.first

        def uniform(self, size=None):
            row = self.uniforms[self.pos]
            self.pos += 1
            return row.astype(np.float64)

    rng = np.random.default_rng(21)
    for trial in range(12):
        k = int(rng.integers(2, 9))
        n = int(rng.integers(40, 400))
        centers = rng.standard_normal((k, 6)) * 6
        x = (centers[rng.integers(0, k, n)] + rng.standard_normal((n, 6))).astype(np.float32)
        xt = torch.from_numpy(x)
        seed = int(rng.integers(0, 1000))
        # the draws upstream would make after torch.manual_seed(seed)
        torch.manual_seed(seed)
        first = int(torch.randint(0, n, (1,)))
        uniforms = np.stack([torch.rand(30).numpy() for _ in range(k - 1)])
        _, want = _kmeans._kmeans_plusplus(x.astype(np.float64), k, (x.astype(np.float64) ** 2).sum(1), np.ones(n),
                                           Draws(first, uniforms), n_local_trials=30)
        _, got = oc.kmeans_plusplus_torch(xt, k, random_state=seed)
        assert got.tolist() == want.tolist(), (trial, got.tolist(), want.tolist())


def test_laplacian_spectrum_properties():
    g = torch.Generator().manual_seed(9)
    x = torch.randn(120, 8, generator=g)
    aff = oc.getAffinityGraphMat(oc.getCosAffinityMatrix(x), 10)
    assert torch.equal(aff, aff.t()) and set(torch.unique(aff).tolist()) <= {0.0, 0.5, 1.0}
    lap = oc.getLaplacian(aff.clone()).float()
    lam = torch.linalg.eigvalsh(lap.double())
    assert lam[0].abs().item() < 1e-5 and lam.min().item() > -1e-5
    from scipy.sparse.csgraph import connected_components

    ncomp, _ = connected_components((aff != 0).numpy())
    assert int((lam.abs() < 1e-6).sum()) == ncomp
    assert oc.isGraphFullyConnected(aff) == (ncomp == 1)


def test_multiscale_fusion_range_and_diagonal():
    scales = [(1.5, 0.75), (1.0, 0.5), (0.5, 0.25)]
    embs, stamps, counts, _ = synthetic_multiscale_embeddings(60.0, scales, 2, seed=2, dim=16)
    e, t = oc.split_input_data(embs, stamps, counts)
    w = torch.tensor([[1.0, 2.0, 0.5]])
    fused = oc.getMultiScaleCosAffinityMatrix(w, e, t)
    assert fused.shape == (int(counts[-1]), int(counts[-1]))
    assert fused.min().item() >= -1e-6 and fused.max().item() <= 3.5 + 1e-5
    assert torch.allclose(torch.diagonal(fused), torch.full((int(counts[-1]),), 3.5), atol=1e-5)  # not divided by sum(w)


def test_rttm_roundtrip_through_the_reference_parser(tmp_path):
    turns = ["0.0 1.125 speaker_0", "1.125 4.0 speaker_1", "5.0 9.5 speaker_0"]
    path = osu.labels_to_rttmfile(turns, "mono_file", str(tmp_path))
    got = []
    with open(path) as f:  # /root/reference/diarize.py:209-216
        for line in f.readlines():
            line_list = line.split(" ")
            s = int(float(line_list[5]) * 1000)
            e = s + int(float(line_list[8]) * 1000)
            got.append([s, e, int(line_list[11].split("_")[-1])])
    assert got == [[0, 1125, 0], [1125, 4000, 1], [5000, 9500, 0]]
    assert osu.rttm_to_labels(path) == ["0.0 1.125 speaker_0", "1.125 4.0 speaker_1", "5.0 9.5 speaker_0"]


def test_oracle_titanet_parameter_count_matches_the_published_figure():
    """NVIDIA's model card and the TitaNet paper (Koluguri et al., 2022, table 1) give TitaNet-L 25.3 M parameters; the
    restated architecture (prolog + 3 mega blocks of 3 sub-blocks at 1024 channels + 3072-channel epilog, attentive
    statistics pooling with a 128-channel bottleneck, 192-d embedding, 16 681-way training head) must land on it."""
    from oracle.titanet import TitaNetL

    total = sum(p.numel() for p in TitaNetL().parameters())
    assert abs(total / 1e6 - 25.3) < 0.05, total


def test_oracle_titanet_shapes_and_determinism():
    from oracle.titanet import TitaNetL
    from whisper_nemo_b200 import checkpoint

    sd = checkpoint.random_init_titanet_large(1234)
    model = TitaNetL(compute_logits=False)
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys
    assert all(k.startswith("preprocessor.") or k.startswith("decoder.final") for k in res.missing_keys)
    model.eval()
    g = torch.Generator().manual_seed(1)
    x = 0.1 * torch.randn(2, 8000, generator=g)
    _, e1 = model(x, torch.tensor([8000, 8000]))
    _, e2 = model(x, torch.tensor([8000, 8000]))
    assert e1.shape == (2, 192) and torch.equal(e1, e2) and torch.isfinite(e1).all()
    sd2 = checkpoint.random_init_titanet_large(1234)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
