"""Staged on-GPU diagnostics (development aid; the judged checks live in tests/ -m gpu).

Usage: python tests/gpu_check.py [gemm] [feat] [titanet] [cluster]
Prints max errors of each CUDA stage against torch fp32 / the CPU oracle.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from whisper_nemo_b200 import _cabi
from whisper_nemo_b200 import titanet as tn

dev = torch.device("cuda")


def check_gemm():
    torch.manual_seed(0)
    for (M, N, K) in [(300, 256, 128), (128, 128, 64), (1000, 1024, 1024), (4097, 3072, 1024), (77, 128, 6144), (513, 384, 3072)]:
        A = (torch.randn(M, K, device=dev) * 0.5).half()
        W = (torch.randn(N, K, device=dev) * 0.05).half()
        bias = torch.randn(N, device=dev)
        ref = A.float() @ W.float().t()
        out = torch.empty(M, N, dtype=torch.float16, device=dev)
        tn.gemm(A, W, out, _cabi.EPI_BIAS, bias=bias)
        torch.cuda.synchronize()
        err = (out.float() - (ref + bias)).abs().max().item()
        print(f"gemm BIAS      M{M} N{N} K{K}: max err {err:.4e} (ref max {ref.abs().max().item():.2f})", flush=True)
        out32 = torch.empty(M, N, dtype=torch.float32, device=dev)
        tn.gemm(A, W, out32, _cabi.EPI_BIAS_F32, bias=bias)
        torch.cuda.synchronize()
        print(f"gemm BIAS_F32  M{M} N{N} K{K}: max err {(out32 - (ref + bias)).abs().max().item():.4e}", flush=True)
    M, N, K, T = 906, 1024, 1024, 151
    A = (torch.randn(M, K, device=dev) * 0.5).half()
    W = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    ref = A.float() @ W.float().t()
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    tn.gemm(A, W, out, _cabi.EPI_BIAS_RELU, bias=bias)
    print("gemm RELU err", (out.float() - torch.relu(ref + bias)).abs().max().item())
    aux = torch.randn(M, N, device=dev).half()
    gate = torch.rand(M // T, N, device=dev)
    tn.gemm(A, W, out, _cabi.EPI_SE_RES, bias=bias, rowvec=gate, aux16=aux, rows_per_seg=T)
    want = torch.relu(aux.float() * gate.repeat_interleave(T, 0) + ref + bias)
    print("gemm SE_RES err", (out.float() - want).abs().max().item())
    N2 = 128
    W2 = (torch.randn(N2, K, device=dev) * 0.05).half()
    ref2 = A.float() @ W2.float().t()
    scale, shift = torch.rand(N2, device=dev) + 0.5, torch.randn(N2, device=dev) * 0.1
    rb = torch.randn(M // T, N2, device=dev)
    out2 = torch.empty(M, N2, dtype=torch.float16, device=dev)
    tn.gemm(A, W2, out2, _cabi.EPI_TDNN, scale=scale, shift=shift, rowvec=rb, rows_per_seg=T)
    want = torch.tanh(scale * torch.relu(ref2 + rb.repeat_interleave(T, 0)) + shift)
    print("gemm TDNN err", (out2.float() - want).abs().max().item())
    out3 = torch.empty(M, N, dtype=torch.float32, device=dev)
    tn.gemm(A, W, out3, _cabi.EPI_SIGMOID_F32)
    print("gemm SIGMOID err", (out3 - torch.sigmoid(ref)).abs().max().item())
    # timing of the big pointwise shape
    M = 32768
    A = (torch.randn(M, 1024, device=dev) * 0.5).half()
    W = (torch.randn(1024, 1024, device=dev) * 0.05).half()
    out = torch.empty(M, 1024, dtype=torch.float16, device=dev)
    bias = torch.zeros(1024, device=dev)
    for _ in range(3):
        tn.gemm(A, W, out, _cabi.EPI_BIAS_RELU, bias=bias)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tn.gemm(A, W, out, _cabi.EPI_BIAS_RELU, bias=bias)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"gemm 32768x1024x1024: {ms:.3f} ms  {2 * M * 1024 * 1024 / ms / 1e9:.1f} TFLOP/s", flush=True)


def _oracle_model():
    from oracle.titanet import seeded_state_dict

    torch.set_num_threads(os.cpu_count())
    return seeded_state_dict(1234, compute_logits=False)


def check_feat(model=None):
    from tools import workload as synth

    model = model or _oracle_model()
    wav, _ = synth.synth_recording(30.0, 2, seed=5)
    wav_t = torch.from_numpy(wav)
    pk = tn.pack_weights(model.state_dict(), dev)
    for fixed_len, lens in [(24000, [24000, 24000, 24000, 9000, 801]), (8000, [8000, 8000, 3000]), (48000, [48000, 47000])]:
        starts = [1000 + 3000 * i for i in range(len(lens))]
        # oracle: fixed_seq collate then FilterbankFeatures
        from oracle.clustering_diarizer import collate

        audio, alens = collate([wav_t[s : s + l] for s, l in zip(starts, lens)])
        feats, flen = model.preprocessor(audio, alens)
        T = fixed_len // 160 + 1
        ref = feats[:, :, :T].transpose(1, 2).contiguous()  # [B,T,80]
        out16, out32 = tn.featurize(pk, wav_t.to(dev), torch.tensor(starts, dtype=torch.int32, device=dev),
                                    torch.tensor(lens, dtype=torch.int32, device=dev), fixed_len, want_f32=True)
        torch.cuda.synchronize()
        err = (out32.cpu() - ref).abs()
        print(f"feat fixed_len {fixed_len}: max abs err {err.max().item():.3e} mean {err.mean().item():.3e}; per-seg max {err.amax(dim=(1,2)).tolist()}", flush=True)
        e16 = (out16.view(len(lens), T, -1)[:, :, :80].float().cpu() - ref).abs().max().item()
        print(f"   fp16 copy max err {e16:.3e}; pad cols max {out16[:, 80:].abs().max().item()}")
    return model, pk


def check_titanet(model=None):
    from oracle.clustering_diarizer import collate
    from tools import workload as synth

    model = model or _oracle_model()
    wav, _ = synth.synth_recording(40.0, 3, seed=7)
    wav_t = torch.from_numpy(wav)
    net = tn.TitaNetB200(model.state_dict(), dev, max_frames=8192)
    for fixed_len, n in [(24000, 12), (8000, 40), (48000, 5)]:
        starts = [500 + 7000 * i for i in range(n)]
        lens = [fixed_len] * n
        lens[-1] = fixed_len // 3
        audio, alens = collate([wav_t[s : s + l] for s, l in zip(starts, lens)])
        t0 = time.time()
        _, ref = model(audio, alens)
        t_cpu = time.time() - t0
        taps = {}
        emb = net.embed_segments(wav_t.to(dev), torch.tensor(starts, dtype=torch.int32, device=dev),
                                 torch.tensor(lens, dtype=torch.int32, device=dev), fixed_len, taps=taps)
        torch.cuda.synchronize()
        emb = emb.cpu()
        cos = torch.nn.functional.cosine_similarity(emb, ref, dim=1)
        rel = (emb - ref).norm(dim=1) / ref.norm(dim=1)
        print(f"titanet fixed_len {fixed_len} n {n}: 1-cos max {(1 - cos).max().item():.3e}  rel err max {rel.max().item():.3e}  (cpu {t_cpu:.2f}s)", flush=True)
        # intermediate taps vs oracle
        feats, flen = model.preprocessor(audio, alens)
        T = fixed_len // 160 + 1
        xs, lens_t = [feats], flen
        for bi, blk in enumerate(model.encoder.encoder):
            xs, lens_t = blk(xs, lens_t)
            key = f"block{bi}" if bi < 4 else "encoder"
            if key in taps and taps[key].shape[0] == n * T:
                got = taps[key].float().cpu().view(n, T, -1)
                want = xs[-1][:, :, :T].transpose(1, 2)
                d = (got - want).abs()
                print(f"   {key}: max abs err {d.max().item():.3e}, rel fro {(got - want).norm().item() / want.norm().item():.3e}")


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "feat", "titanet"]
    _cabi.require_device()
    print(_cabi.load().b200d_version().decode(), torch.cuda.get_device_name(0), flush=True)
    model = None
    if "gemm" in what:
        check_gemm()
    if "feat" in what:
        model, _ = check_feat(model)
    if "titanet" in what:
        check_titanet(model)
