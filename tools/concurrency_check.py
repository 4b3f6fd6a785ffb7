"""Development aid: do two host threads launching kernels on two streams make progress together?
usage: python tools/concurrency_check.py <variantA> <variantB> [seconds]   variants: big2cta small1cta mid1cta cheb2cta featurize depthwise timestats jacobi kmeans eigvals torchcopy"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from whisper_nemo_b200 import _cabi, checkpoint
from whisper_nemo_b200 import titanet as tn
from whisper_nemo_b200 import clustering as cl

dev = torch.device("cuda")
_cabi.require_device()


def make(variant):
    if variant == "big2cta":
        A = (torch.randn(49152, 1024, device=dev) * 0.5).half(); W = (torch.randn(1024, 1024, device=dev) * 0.05).half()
        out = torch.empty(49152, 1024, dtype=torch.float16, device=dev); bias = torch.zeros(1024, device=dev)
        return lambda: tn.gemm(A, W, out, _cabi.EPI_BIAS_RELU, bias=bias)
    if variant == "small1cta":
        A = (torch.randn(325, 128, device=dev) * 0.5).half(); W = (torch.randn(1024, 128, device=dev) * 0.05).half()
        out = torch.empty(325, 1024, dtype=torch.float32, device=dev)
        return lambda: tn.gemm(A, W, out, _cabi.EPI_SIGMOID_F32)
    if variant == "mid1cta":
        A = (torch.randn(2000, 1024, device=dev) * 0.5).half(); W = (torch.randn(1024, 1024, device=dev) * 0.05).half()
        out = torch.empty(2000, 1024, dtype=torch.float16, device=dev); bias = torch.zeros(1024, device=dev)
        return lambda: tn.gemm(A, W, out, _cabi.EPI_BIAS, bias=bias)
    if variant == "cheb2cta":
        n, b = 10000, 64
        a16 = (torch.rand(n, n, device=dev) < 0.02).to(torch.bfloat16); deg = a16.float().sum(1)
        x = torch.randn(n, b, device=dev); out = torch.empty(n, b, device=dev)
        nw = cl.cheb_operand_rows(b)
        vt = torch.zeros(nw, n, dtype=torch.bfloat16, device=dev); vo = torch.zeros(nw, n, dtype=torch.bfloat16, device=dev)
        return lambda: cl._gemm_cheb(a16, n, vt, n, n, nw, out, deg, x, None, 1.0, 0.0, 0.0, vo)
    if variant == "depthwise":
        x = torch.randn(325 * 151, 1024, device=dev).half(); y = torch.empty_like(x); w = torch.randn(15, 1024, device=dev)
        return lambda: _cabi.call("b200d_depthwise_conv", _cabi.ptr(x), _cabi.ptr(y), _cabi.ptr(w), 325, 151, 1024, 15, _cabi._stream())
    if variant == "timestats":
        x = torch.randn(325 * 151, 1024, device=dev).half(); o = torch.empty(325, 1024, dtype=torch.float16, device=dev)
        return lambda: _cabi.call("b200d_time_stats", _cabi.ptr(x), 325, 151, 1024, 0, _cabi.ptr(o), _cabi._stream())
    if variant == "featurize":
        pk = tn.pack_weights(checkpoint.random_init_titanet_large(1), dev)
        wav = torch.randn(16000 * 600, device=dev) * 0.1
        st = (torch.arange(600, device=dev, dtype=torch.int32) * 12000).contiguous(); ln = torch.full((600,), 30400, dtype=torch.int32, device=dev)
        out16 = torch.empty(600 * 191, 128, dtype=torch.float16, device=dev)
        return lambda: tn.featurize(pk, wav, st, ln, 30400, out16=out16)
    if variant == "jacobi":
        g = torch.randn(64, 64, device=dev); g = g @ g.t()
        ev = torch.empty(64, device=dev); V = torch.empty(64, 64, device=dev)
        return lambda: _cabi.call("b200d_small_eig", _cabi.ptr(g), 64, _cabi.ptr(ev), _cabi.ptr(V), 0, _cabi._stream())
    if variant == "kmeans":
        X = torch.randn(10000, 50, device=dev)
        return lambda: cl.kmeans_torch(X, 50)
    if variant == "eigvals":
        n, batch = 515, 30
        a = torch.randn(batch, n, n, device=dev); a = a + a.transpose(1, 2)
        work = torch.empty_like(a)
        ws_bytes = _cabi.load().b200d_eigvals_workspace_bytes(batch, n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev); ev = torch.empty(batch, 10, device=dev)

        def run():
            work.copy_(a)
            _cabi.call("b200d_eigvals_batched", _cabi.ptr(work), batch, n, 9, _cabi.ptr(ev), _cabi.ptr(ws), ws_bytes, _cabi._stream())
        return run
    if variant == "torchcopy":
        a = torch.randn(4096, 1024, device=dev); b = torch.empty_like(a)
        return lambda: b.copy_(a)
    raise SystemExit(f"unknown variant {variant}")


def main():
    va, vb = sys.argv[1], sys.argv[2]
    secs = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
    counts = [0, 0]
    stop = time.time() + secs

    def run(i, variant):
        try:
            _run(i, variant)
        except Exception as e:  # noqa: BLE001
            print(f"EXC in {variant}: {e!r}", flush=True)

    def _run(i, variant):
        torch.cuda.set_device(0)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fn = make(variant)
            torch.cuda.synchronize()
            while time.time() < stop:
                for _ in range(20):
                    fn()
                st.synchronize()
                counts[i] += 20
                if counts[i] % 2000 == 0:
                    print(f"  thread {i} ({variant}) {counts[i]} launches", flush=True)

    import faulthandler

    faulthandler.enable()
    faulthandler.dump_traceback_later(secs + 15, exit=True)
    ts = [threading.Thread(target=run, args=(0, va)), threading.Thread(target=run, args=(1, vb))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    torch.cuda.synchronize()
    print(f"OK {va} x {vb}: launches {counts}", flush=True)


if __name__ == "__main__":
    main()
