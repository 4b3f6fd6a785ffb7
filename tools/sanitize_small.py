"""Small invocations of the hand-written kernels for compute-sanitizer (memcheck / racecheck are 10-100x slower than a
plain run, so every shape is tiny):

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py > profiles/r02_sanitizer_memcheck.log 2>&1
    compute-sanitizer --tool racecheck python tools/sanitize_small.py > profiles/r02_sanitizer_racecheck.log 2>&1
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py --rowshard   (adds the row-sharded path on two ranks)

Covers: both tcgen05 GEMM kernels (1-CTA, CTA-pair with the SE column-sum epilogue), the split Chebyshev GEMM, featurizer
(stream + window kernels), depthwise / statistics / pooling kernels through b200d_titanet_forward, k-means, top-p
binarisation, the batched eigenvalue kernels, b200d_eig_bottomk (split-K products + fix-up kernel) and the row-sharded path on
two LocalComm ranks (row kernels, peer stores of the fix-up kernel, device-side barriers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tools import workload  # noqa: E402
from whisper_nemo_b200 import _cabi, checkpoint  # noqa: E402
from whisper_nemo_b200 import clustering as cl  # noqa: E402
from whisper_nemo_b200 import titanet as tn  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    _cabi.require_device()
    wav, _ = workload.synth_recording(12.0, 2, seed=3)
    wav_d = torch.from_numpy(wav).to(dev)
    net = tn.TitaNetB200(checkpoint.seeded(), dev, max_frames=40960)
    # 260 windows of 1.5 s -> M = 39 260 frames: the large GEMMs take the CTA-pair kernel (with column sums), the small the 1-CTA
    start = (np.arange(260) * 400 + 160).astype(np.int64)
    length = np.full(260, 24000, dtype=np.int64)
    length[-1] = 9000
    s_start, s_off, row0 = tn.plan_mel_streams(start, length, np.full(260, 24000, dtype=np.int64))
    logmel = net.mel_stream(wav_d, torch.from_numpy(s_start).to(dev), torch.from_numpy(s_off).to(dev), int(s_off[-1]))
    to32 = lambda a: torch.from_numpy(a.astype(np.int32)).to(dev)
    emb = net.embed_segments(wav_d, to32(start), to32(length), 24000, logmel=logmel, seg_row0=to32(row0))
    torch.cuda.synchronize()
    print("titanet_forward", tuple(emb.shape), float(emb.abs().mean()))
    g = torch.Generator().manual_seed(0)
    x = torch.randn(700, 192, generator=g) + 3 * torch.randn(4, 192, generator=g)[torch.randint(0, 4, (700,), generator=g)]
    mat = cl.getCosAffinityMatrix(x.to(dev))
    nmesc = cl.NMESC(mat, max_num_speakers=8, max_rp_threshold=0.25, sparse_search_volume=10)
    k, p = nmesc.forward()
    graph = cl.getAffinityGraphMat(mat, p)
    vec = cl.bottom_eigvecs(graph[0], graph[1], k, p=p)            # CSR or dense products by density
    vec2 = cl.bottom_eigvecs(graph[0], graph[1], 30, p=None)       # 64-vector block, dense: the split Chebyshev GEMM
    lab = cl.kmeans_torch(vec, k)
    torch.cuda.synchronize()
    print("clustering", k, p, tuple(vec.shape), tuple(vec2.shape), int(lab.max()))
    if "--rowshard" not in sys.argv:
        return
    # row-sharded forms on two ranks (host threads, one stream and one peer buffer each).  A separate invocation: the device-side
    # barriers need both ranks' kernels in flight at once, which a sanitizer that serialises kernels turns into a barrier time-out
    import threading

    from whisper_nemo_b200 import rowshard

    n = mat.shape[0]
    comms = rowshard.LocalComm.make(2)
    shards = rowshard.row_shards(n, 2)
    streams = [torch.cuda.Stream() for _ in range(2)]
    out = [None, None]

    def work(r):
        torch.cuda.set_device(0)
        with torch.cuda.stream(streams[r]), torch.no_grad():
            lo, hi = shards[r]
            rows = rowshard.fused_affinity_rows(comms[r], [1.0], [x.to(dev)], [np.arange(n, dtype=np.int32)], lo, hi)
            a_rows, deg = rowshard.graph_rows(comms[r], rows, p, lo, hi, shards)
            out[r] = rowshard.bottom_eigvecs_sharded(comms[r], a_rows, deg, k, lo, hi)
            streams[r].synchronize()

    torch.cuda.synchronize()
    threads = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    print("row-sharded", tuple(out[0].shape), bool(torch.equal(out[0], out[1])))


if __name__ == "__main__":
    main()
