"""Host-side logic of the product (manifests, windows, RTTM text, scale mapping, long-form bookkeeping, config,
sharding) against the oracle's restatement, on CPU.  The device kernels are exercised by the `-m gpu` tests."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import longform_clustering as olf
from oracle import offline_clustering as oc
from oracle import speaker_utils as osu
from tests.util import synthetic_multiscale_embeddings
from whisper_nemo_b200 import clustering as cl
from tools import workload as synth
from whisper_nemo_b200 import config, longform as lf, sharding
from whisper_nemo_b200 import speaker_utils as su

SCALES = [(1.5, .75), (1.25, .625), (1.0, .5), (.75, .375), (.5, .25), (3.0, 1.5), (1.9, .95), (1.2, .6)]


def test_subsegments_equal_oracle():
    rng = np.random.default_rng(0)
    for _ in range(3000):
        off, dur = round(float(rng.uniform(0, 100)), 3), round(float(rng.uniform(0.01, 30)), 3)
        w, s = SCALES[rng.integers(len(SCALES))]
        assert osu.get_subsegments(off, w, s, dur) == su.get_subsegments(off, w, s, dur)


@pytest.mark.parametrize("pcm16", [False, True])
def test_manifests_and_wav_equal_oracle(tmp_path, pcm16):
    from oracle.clustering_diarizer import read_wav

    wav_path, rttm_path, wav, turns = synth.make_session(str(tmp_path), "mono_file", 45.0, 3, 3, pcm16=pcm16)
    man = os.path.join(tmp_path, "m.json")
    synth.write_manifest(man, [{"audio_filepath": wav_path, "rttm_filepath": rttm_path}])
    A, B = osu.audio_rttm_map(man), su.audio_rttm_map(man)
    assert A == B and list(A) == ["mono_file"]
    assert np.array_equal(read_wav(wav_path), su.read_wav(wav_path))
    osu.write_rttm2manifest(A, str(tmp_path / "o_vad.json"), {"mono_file": 45.0})
    su.write_rttm2manifest(B, str(tmp_path / "g_vad.json"), {"mono_file": 45.0})
    assert open(tmp_path / "o_vad.json").read() == open(tmp_path / "g_vad.json").read()
    for w, s in SCALES[:5]:
        osu.segments_manifest_to_subsegments_manifest(str(tmp_path / "o_vad.json"), str(tmp_path / "o_s.json"), w, s)
        entries = su.segments_manifest_to_subsegments_manifest(str(tmp_path / "g_vad.json"), str(tmp_path / "g_s.json"), w, s)
        assert open(tmp_path / "o_s.json").read() == open(tmp_path / "g_s.json").read()
        assert len(entries) == sum(1 for _ in open(tmp_path / "g_s.json"))


def test_cluster_labels_and_rttm_text_equal_oracle(tmp_path):
    g = torch.Generator().manual_seed(1)
    t0 = torch.cumsum(torch.rand(300, generator=g) * 0.4 + 0.05, 0)
    ts = torch.stack([t0, t0 + 0.5], 1).float()
    for trial in range(50):
        lab = torch.randint(0, 4, (300,), generator=g)
        a, b = osu.generate_cluster_labels(ts, lab), su.generate_cluster_labels(ts, lab)
        assert a[0] == b[0] and a[1] == b[1]
    pa = osu.labels_to_rttmfile(a[0], "x", str(tmp_path))
    text_o = open(pa).read()
    pb = su.labels_to_rttmfile(b[0], "x", str(tmp_path))
    assert open(pb).read() == text_o and "SPEAKER x 1   " in text_o
    assert [(round(s, 3), round(e, 3), k) for s, e, k in su.rttm_to_turns(pb)] == \
           [(round(float(l.split()[0]), 3), round(float(l.split()[1]), 3), l.split()[2]) for l in osu.rttm_to_labels(pa)]


def test_scale_mapping_equals_dense_argmin():
    scales = [(3.0, 1.5), (2.5, 1.25), (2.0, 1.0), (1.5, .75), (1.0, .5), (.5, .25)]
    _, stamps, counts, _ = synthetic_multiscale_embeddings(1800.0, scales, 4, 1, dim=4)
    st = list(torch.split(stamps, counts.tolist()))
    for a, b in zip(oc.get_argmin_mat(st), cl.get_argmin_mat(st)):
        assert np.array_equal(a.numpy(), b)
    # gaps between speech regions and duplicated centres
    odd = [torch.tensor([[0.0, 1.0], [0.5, 1.5], [10.0, 11.0], [10.0, 11.0], [30.0, 30.2]]), torch.tensor([[0.0, 0.5], [0.25, 0.75], [0.5, 1.0], [5.2, 5.7], [10.2, 10.7], [20.0, 20.5], [30.0, 30.2]])]
    for a, b in zip(oc.get_argmin_mat(odd), cl.get_argmin_mat(odd)):
        assert np.array_equal(a.numpy(), b)


def test_longform_bookkeeping_equals_oracle():
    rng = np.random.default_rng(0)
    for trial in range(400):
        k = int(rng.integers(2, 50))
        n = int(rng.integers(k * 3, 2000))
        labels = torch.from_numpy(rng.integers(0, k, n))
        labels[:k] = torch.arange(k)
        target = int(rng.integers(k, min(n - 1, 4 * k) + 1)) if n - 1 > k else k
        mc = int(np.ceil(target / k))
        res = []
        for mod in (olf, lf):
            try:
                res.append(mod.get_merge_quantity(n - target, labels.clone(), mc))
            except ValueError as e:
                res.append(str(e))
        if isinstance(res[0], str) or isinstance(res[1], str):
            assert res[0] == res[1]
        else:
            assert torch.equal(res[0], res[1])


def test_merge_plan_equals_oracle_run_reducer():
    """plan_cluster_merges (the host half of the long-form reducer) against the oracle's run_reducer cluster by cluster:
    same kept / merged window lists, and gathering [rows ; means] in the planned order reproduces its merged vectors."""
    rng = np.random.default_rng(11)
    for trial in range(25):
        k = int(rng.integers(2, 12))
        n = int(rng.integers(6 * k, 400))
        d = 16
        emb = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
        labels = torch.from_numpy(rng.integers(0, k, n))
        labels[:k] = torch.arange(k)
        target = int(rng.integers(k + 1, max(k + 2, n // 2)))
        vol = olf.get_merge_quantity(n - target, labels.clone(), max(1, target // k))
        mat = torch.from_numpy(rng.random((n, n)).astype(np.float32))
        mat = 0.5 * (mat + mat.t())
        if trial % 5 == 0:  # ties in the masses
            mat = torch.round(mat * 2) / 2
        mass = torch.zeros(n)
        for c in range(k):
            idx = torch.where(labels == c)[0]
            mass[idx] = mat[:, idx][idx, :].sum(0)
        offset = int(rng.integers(0, 1000))
        mapping, sizes, sel_idx, seg_off, order, n_avg = lf.plan_cluster_merges(labels.numpy(), [int(v) for v in vol.tolist()],
                                                                                 mass.numpy(), n, offset)
        means = [emb[torch.from_numpy(ix)].mean(0) for ix in sel_idx]
        src = torch.cat([emb] + [m[None] for m in means]) if means else emb
        merged_all = src[torch.from_numpy(np.concatenate(order))]
        got = torch.split(merged_all, sizes)
        assert n_avg == len(sel_idx) == len(seg_off) - 1
        for c in range(k):
            want_emb, _, (kept, merged) = olf.run_reducer(emb, c, int(vol[c]), labels, total_affinity_mat=mat)
            assert torch.equal(mapping[c][0], kept + offset) and torch.equal(mapping[c][1], merged + offset), (trial, c)
            assert got[c].shape == want_emb.shape and (got[c] - want_emb).abs().max().item() < 1e-6


def test_nme_ratios_equal_the_per_p_loop():
    """The vectorised eigengap analysis == upstream's loop over p (NMESC.getEigRatio), bit for bit, ties included."""
    from whisper_nemo_b200.clustering import nme_ratios

    rng = np.random.default_rng(5)
    for trial in range(200):
        n = int(rng.integers(20, 600))
        max_spk = int(rng.integers(2, 12))
        n_low = min(max_spk, n - 1) + 1
        np_ = int(rng.integers(1, 40))
        p_list = sorted(int(v) for v in rng.integers(2, max(3, n // 4), np_))
        lam = np.sort(rng.random((np_, n_low)) * rng.choice([1e-3, 1.0, 50.0]), axis=1)
        if trial % 3 == 0:  # exact ties between eigengaps
            lam = np.round(lam * 4) / 4
        evals = torch.from_numpy(np.concatenate([lam, lam[:, -1:] + rng.random((np_, 1)) * 10], axis=1).astype(np.float32))
        spk, g = nme_ratios(evals, p_list, n, max_spk, 1e-10)
        for i, p_neighbors in enumerate(p_list):
            lambdas, lam_max = evals[i, :n_low], evals[i, n_low]
            gap = lambdas[1:] - lambdas[:-1]
            want_spk = torch.argmax(gap[: min(max_spk, gap.shape[0])]) + 1
            key = torch.argsort(gap[:max_spk], descending=True)[0]
            max_eig_gap = gap[key] / (lam_max.item() + 1e-10)
            want_g = (p_neighbors / n) / (max_eig_gap + 1e-10)
            assert float(spk[i]) == float(want_spk)
            assert torch.equal(g[i], want_g.float()) or (torch.isinf(g[i]) and torch.isinf(want_g)), (trial, i, g[i], want_g)


@pytest.mark.parametrize("n", [7, 100, 10000, 51261])
def test_batched_cpu_rng_draws_equal_sequential_draws(n):
    """kmeans_torch draws its k-means++ uniforms and fallback indices in two batched calls; upstream draws them one
    center at a time -- the CPU generator must give the same numbers either way."""
    g = torch.Generator().manual_seed(3)
    torch.randint(0, n, (1,), generator=g)
    seq_r = torch.stack([torch.rand(30, generator=g) for _ in range(49)])
    seq_f = torch.cat([torch.randint(n, (1,), generator=g) for _ in range(200)])
    g = torch.Generator().manual_seed(3)
    torch.randint(0, n, (1,), generator=g)
    assert torch.equal(torch.rand(49, 30, generator=g), seq_r)
    assert torch.equal(torch.randint(n, (200,), generator=g), seq_f)


def test_nmesc_p_value_list_equals_oracle():
    for n in (7, 37, 110, 515, 600, 1023):
        for thr, vol in ((0.25, 30), (0.15, 10), (0.05, 30)):
            o = oc.NMESC(torch.zeros(n, n), max_rp_threshold=thr, sparse_search_volume=vol)
            want = o.getPvalueList().tolist()
            g = cl.NMESC(torch.zeros(n, n), max_rp_threshold=thr, sparse_search_volume=vol)
            assert g.getPvalueList(n) == want and g.max_N == int(o.max_N)


def test_config_schema_and_create_config(tmp_path):
    for dom, n_scales in (("telephonic", 5), ("meeting", 6), ("general", 3)):
        cfg = config.load_config(dom)
        p = cfg.diarizer.speaker_embeddings.parameters
        assert len(p.window_length_in_sec) == len(p.shift_length_in_sec) == len(p.multiscale_weights) == n_scales
        c = cfg.diarizer.clustering.parameters
        assert c.max_num_speakers == 8 and c.chunk_cluster_count == 50 and c.embeddings_per_chunk == 10000
        with pytest.raises(config.MissingMandatoryValue):
            cfg.diarizer.manifest_filepath
    cfg = config.create_config(str(tmp_path))  # helpers.py:252-303
    meta = json.loads(open(cfg.diarizer.manifest_filepath).read())
    assert meta["audio_filepath"].endswith("mono_file.wav") and set(meta) == {"audio_filepath", "offset", "duration", "label", "text", "rttm_filepath", "uem_filepath"}
    assert cfg.diarizer.speaker_embeddings.model_path == "titanet_large" and cfg.diarizer.oracle_vad is False and cfg.num_workers == 0
    sd = su.parse_scale_configs([1.5, 1.0, 0.5], [0.75, 0.5, 0.25], [1, 1, 1])
    assert sd["scale_dict"] == {0: (1.5, 0.75), 1: (1.0, 0.5), 2: (0.5, 0.25)} and not sd["use_single_scale_clustering"]
    with pytest.raises(ValueError):
        su.parse_scale_configs([1.0, 1.5], [0.5, 0.75], [1, 1])


def test_reference_yaml_matches_bundled_yaml():
    ref = "/root/reference/nemo_msdd_configs"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present on this box")
    import yaml

    for dom in ("telephonic", "meeting", "general"):
        a = yaml.safe_load(open(os.path.join(ref, f"diar_infer_{dom}.yaml")))
        b = config.load_config(dom).to_dict()
        for key in ("speaker_embeddings", "clustering"):
            assert a["diarizer"][key]["parameters"] == b["diarizer"][key]["parameters"], (dom, key)
        assert a["batch_size"] == b["batch_size"] and a["sample_rate"] == b["sample_rate"]


def test_recording_assignment_and_ranges():
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 8):
        dur = rng.uniform(30, 4000, size=37).tolist()
        parts = sharding.assign_recordings(dur, world)
        assert sorted(i for p in parts for i in p) == list(range(37))
        loads = [sum(dur[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(dur)
        for n in (0, 1, 7, 100):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def _gloo_worker(rank, world, port, tmp):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d = 1001, 192
        full = torch.arange(n * d, dtype=torch.float32).view(n, d)
        lo, hi = sharding.shard_range(n, rank, world)
        got = sharding.all_gather_rows(full[lo:hi].clone(), n)
        assert torch.equal(got, full)
        # batch-of-recordings sharding: every rank labels its own recordings, results are gathered as objects
        dur = [600.0, 120.0, 900.0, 30.0, 450.0]
        mine = sharding.assign_recordings(dur, world)[rank]
        out = [None] * world
        dist.all_gather_object(out, {i: f"labels-of-{i}" for i in mine})
        merged = {k: v for part in out for k, v in part.items()}
        assert sorted(merged) == list(range(len(dur)))
        # long-form chunks dealt round-robin; every rank ends with every chunk's (labels, mass), bit for bit
        n_chunks, m_chunk = 5, 1000
        lens = [1000, 1000, 1000, 1000, 337]
        g = torch.Generator().manual_seed(7)
        truth = {w: (torch.randint(0, 50, (lens[w],), generator=g, dtype=torch.int32), torch.rand(lens[w], generator=g)) for w in range(n_chunks)}
        mine_chunks = sharding.chunks_of_rank(n_chunks, rank, world)
        assert sorted(sum((sharding.chunks_of_rank(n_chunks, r, world) for r in range(world)), [])) == list(range(n_chunks))
        got = sharding.all_gather_chunk_results({w: truth[w] for w in mine_chunks}, n_chunks, m_chunk, lens)
        for w in range(n_chunks):
            assert torch.equal(got[w][0], truth[w][0]) and torch.equal(got[w][1], truth[w][1])
        # row-sharded clustering: the small collectives of rowshard.DistComm (per-row thresholds / degrees / NME rows
        # gathered in rank order, per-scale (min, max) reduced) and the host-side planning of the NME subsample
        from whisper_nemo_b200 import rowshard

        comm = rowshard.DistComm()
        n = 4099
        shards = rowshard.row_shards(n, world)
        lo, hi = shards[rank]
        counts = [h - l for l, h in shards]
        full = torch.arange(n * 2, dtype=torch.int32).view(n, 2)
        assert torch.equal(comm.all_gather_rows(full[lo:hi].clone(), counts), full)
        mm = torch.tensor([[-0.25 - rank, 1.0], [0.1 * (rank + 1), 0.5 + rank]])
        red = comm.all_reduce_minmax(mm)
        assert torch.equal(red, torch.tensor([[-0.25 - (world - 1), 1.0], [0.1, 0.5 + (world - 1)]]))
        ratio = max(1, int(n / 512))
        sub_idx = np.arange(0, n, ratio)
        cnt = [int(((sub_idx >= l) & (sub_idx < h)).sum()) for l, h in shards]
        mat = torch.arange(n, dtype=torch.float32)[:, None] * 3 + torch.arange(n, dtype=torch.float32)[None, :]
        mine = torch.from_numpy(sub_idx[(sub_idx >= lo) & (sub_idx < hi)] - lo)
        sub = comm.all_gather_rows(mat[lo:hi].index_select(0, mine)[:, ::ratio].contiguous(), cnt)
        assert torch.equal(sub, mat[::ratio, ::ratio])
        # launch-group sharding of the embedding extraction (ClusteringDiarizer._extract_embeddings): every window computed once,
        # by a launch group of the single-GPU run, and returned in manifest order on every rank
        from whisper_nemo_b200.diarizer import ClusteringDiarizer

        class _StubModel:
            max_frames = 1000
            calls = []

            def group_windows(self, fixed_len):
                return 7

            def embed_segments(self, wav, st, ln, fl, logmel=None, seg_row0=None, n_on_stream=None):
                self.calls.append((int(fl), int(n_on_stream), st.tolist()))
                return (st.float()[:, None] * 1000 + float(fl)).expand(-1, 192).contiguous()

        rng = np.random.default_rng(3)
        n_win = 61
        plan = {"n": n_win, "fixed": rng.choice([8000, 24000], n_win), "start": np.arange(n_win) * 10, "len": np.full(n_win, 8000),
                "row0": np.where(rng.random(n_win) < 0.7, np.arange(n_win), -1)}
        diar = object.__new__(ClusteringDiarizer)
        diar.device, diar.shard_windows, diar._speaker_model, diar._shard_loads = torch.device("cpu"), True, _StubModel(), []
        got = diar._extract_embeddings(plan, torch.zeros(1), logmel=torch.zeros(1))
        want = torch.from_numpy(plan["start"] * 1000.0 + plan["fixed"]).float()[:, None].expand(-1, 192)
        assert torch.equal(got, want)
        seen = [None] * world
        dist.all_gather_object(seen, _StubModel.calls)
        starts = sorted(s for calls in seen for _, _, sts in calls for s in sts)
        assert starts == plan["start"].tolist()                      # every window exactly once over the ranks
        assert all(len(sts) <= 7 for calls in seen for _, _, sts in calls)
        for calls in seen:                                           # on-stream windows lead every launch group
            for fl, n_fast, sts in calls:
                on = [plan["row0"][s // 10] >= 0 for s in sts]
                assert on == sorted(on, reverse=True) and sum(on) == n_fast
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_window_sharding_all_gather_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    mp.spawn(_gloo_worker, args=(2, 29517, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def _random_regions(rng, n=300, path="/x/rec a.wav", uniq="rec a"):
    regs, t = [], 0.0
    for _ in range(n):
        d = float(round(rng.uniform(0.03, 9), 5))
        t = float(round(t + rng.uniform(0, 1), 5))
        regs.append({"audio_filepath": path, "offset": t, "duration": d, "label": "UNK", "uniq_id": uniq})
        t += d
    return regs


def test_vectorised_subsegments_equal_the_loop(tmp_path):
    """subsegment_arrays / write_subsegments_manifest (the product's prepare path) against the per-region loop that mirrors
    upstream: identical float64 starts and durations, byte-identical subsegments_scale<k>.json."""
    regs = _random_regions(np.random.default_rng(1))
    man = tmp_path / "seg.json"
    man.write_text("".join(json.dumps(r) + "\n" for r in regs))
    for w, s in SCALES:
        ref = su.segments_manifest_to_subsegments_manifest(str(man), str(tmp_path / "a.json"), w, s)
        region, start, dur = su.subsegment_arrays([r["offset"] for r in regs], [r["duration"] for r in regs], w, s)
        su.write_subsegments_manifest(str(tmp_path / "b.json"), regs, region, start, dur)
        assert (tmp_path / "a.json").read_text() == (tmp_path / "b.json").read_text()
        assert [e["offset"] for e in ref] == start.tolist() and [e["duration"] for e in ref] == dur.tolist()


def test_window_plan_and_mel_streams(tmp_path):
    """ClusteringDiarizer._plan_windows (host only): sample ranges, fixed_seq batch lengths and float32 stamps equal the
    entry-by-entry rule of the dataset / collate; every full-length window starts on a frame of its log-mel stream and
    the stream holds all of its interior frames."""
    from whisper_nemo_b200 import titanet as tn
    from whisper_nemo_b200.diarizer import ClusteringDiarizer

    sr, total = 16000, 16000 * 1400
    regs = _random_regions(np.random.default_rng(5), n=260)
    regs = [r for r in regs if r["offset"] + r["duration"] < total / sr - 1]
    diar = object.__new__(ClusteringDiarizer)  # host-side planning needs no device
    diar.sample_rate, diar.batch_size = sr, 64
    diar.AUDIO_RTTM_MAP = {"rec a": {}}
    diar._wav_offset = {"rec a": (1000, total)}
    diar.multiscale_args_dict = su.parse_scale_configs([1.5, 1.25, 1.0, 0.75, 0.5], [0.75, 0.625, 0.5, 0.375, 0.25], [1, 1, 1, 1, 1])
    diar._plan_windows(regs, write_manifests=False)
    man = tmp_path / "seg.json"
    man.write_text("".join(json.dumps(r) + "\n" for r in regs))
    for scale_idx, (w, s) in diar.multiscale_args_dict["scale_dict"].items():
        entries = su.segments_manifest_to_subsegments_manifest(str(man), str(tmp_path / "s.json"), w, s)
        plan = diar._scales[scale_idx]
        start = [1000 + int(e["offset"] * sr) for e in entries]
        length = [min(int(e["duration"] * sr), total - int(e["offset"] * sr)) for e in entries]
        fixed = [max(length[b0 : b0 + 64]) for b0 in range(0, len(length), 64) for _ in length[b0 : b0 + 64]]
        assert plan["start"].tolist() == start and plan["len"].tolist() == length and plan["fixed"].tolist() == fixed
        stamps = torch.tensor([[e["offset"], e["offset"] + e["duration"]] for e in entries])
        assert torch.equal(plan["stamps"]["rec a"], stamps)
    stream_start, stream_off = diar._streams
    assert np.all(stream_off % tn.STREAM_PAD == 0) and np.all(np.diff(stream_off) > 0)
    n_full = n_rows = 0
    for plan in diar._scales.values():
        full = plan["len"] == plan["fixed"]
        assert np.array_equal(plan["row0"] >= 0, full)  # exactly the windows that fill their tiled-up length sit on a stream
        r0, st, fx = plan["row0"][full], plan["start"][full], plan["fixed"][full]
        sid = np.searchsorted(stream_off, r0, side="right") - 1
        assert np.array_equal(stream_start[sid] + (r0 - stream_off[sid]) * tn.HOP, st)  # frame row0 is centred on the first sample
        assert np.all(r0 + (fx - 200) // tn.HOP < stream_off[sid + 1])                   # last interior frame inside the stream
        n_full += int(full.sum())
        n_rows += int((fx // tn.HOP + 1).sum())
    assert n_full > 0 and stream_off[-1] < 0.45 * n_rows  # telephonic scales: two grid phases per region instead of ~10 recomputations


def test_row_shards_cover_every_row_once():
    from whisper_nemo_b200 import rowshard

    for n, world in [(4096, 2), (10000, 3), (57599, 8), (130, 2), (4097, 8)]:
        sh = rowshard.row_shards(n, world)
        assert sh[0][0] == 0 and sh[-1][1] == n and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        assert all(hi > lo for lo, hi in sh)
        if n >= 16 * world * rowshard.ROW_ALIGN:
            assert all(lo % rowshard.ROW_ALIGN == 0 for lo, _ in sh)


def test_later_subsegment_rule_matches_the_oracle_switch(monkeypatch):
    """B200D_SUBSEGMENT_RULE=v2 (the get_subsegments of later NeMo 2.x releases) in the PRODUCT path: per region and through the
    vectorised planner, equal to the oracle under oracle.switches.SUBSEGMENT_RULE = "v2"."""
    from oracle import speaker_utils as osu
    from oracle import switches

    monkeypatch.setattr(switches, "SUBSEGMENT_RULE", "v2")
    monkeypatch.setattr(su, "SUBSEGMENT_RULE", "v2")
    rng = np.random.default_rng(11)
    offs = np.round(rng.uniform(0, 500, 200), 5)
    durs = np.round(np.concatenate([rng.uniform(0.005, 0.6, 60), rng.uniform(0.5, 30, 140)]), 5)
    for w, s in SCALES:
        want = []
        for r, (o, d) in enumerate(zip(offs.tolist(), durs.tolist())):
            segs = osu.get_subsegments(o, w, s, d)
            assert segs == su.get_subsegments(o, w, s, d)
            want += [(r, st, du) for st, du in segs if du > su.MIN_SUBSEGMENT_DURATION]
        region, start, dur = su.subsegment_arrays(offs, durs, w, s)
        assert region.tolist() == [x[0] for x in want] and start.tolist() == [x[1] for x in want] and dur.tolist() == [x[2] for x in want]


def test_round_robin_gather_order_restores_the_sweep():
    """The NME sweep's p values are dealt round-robin to the ranks and their eigenvalues gathered in rank order: the inverse
    permutation puts every row back where the replicated sweep has it (uneven counts and empty ranks included)."""
    for n_items, world in [(30, 8), (30, 2), (5, 8), (1, 4), (64, 3)]:
        counts, order = sharding.round_robin_counts_and_order(n_items, world)
        assert sum(counts) == n_items and sorted(order) == list(range(n_items))
        per_rank = [list(range(n_items))[r::world] for r in range(world)]
        assert counts == [len(x) for x in per_rank]
        gathered = torch.tensor([i for part in per_rank for i in part], dtype=torch.float32)
        out = torch.empty_like(gathered)
        out[torch.tensor(order)] = gathered
        assert out.tolist() == [float(i) for i in range(n_items)]
