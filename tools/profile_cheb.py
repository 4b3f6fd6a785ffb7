"""The spectral solver alone on a long-form-sized graph (10 000 nodes, k = 50 -> 64-vector block), for ncu:

    ncu --set full --clock-control none -k regex:"gemm_tcgen05_kernel|cheb_fixup" -s 20 -c 6 -o gpurun_out/cheb python tools/profile_cheb.py

Prints the solver's own statistics and the event-timed duration of one call (not under ncu: a bench value)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from whisper_nemo_b200 import _cabi  # noqa: E402
from whisper_nemo_b200 import clustering as cl  # noqa: E402


def main():
    n, k, p = int(os.environ.get("N", 10000)), int(os.environ.get("K", 50)), int(os.environ.get("P", 171))
    dev = torch.device("cuda", 0)
    _cabi.require_device()
    g = torch.Generator().manual_seed(0)
    centers = torch.randn(8, 192, generator=g) * 2
    x = centers[torch.randint(0, 8, (n,), generator=g)] + torch.randn(n, 192, generator=g)
    mat = cl.getCosAffinityMatrix(x.to(dev))
    a16, deg = cl.getAffinityGraphMat(mat, p)
    del mat
    cl.bottom_eigvecs(a16, deg, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cl.bottom_eigvecs(a16, deg, k)
    e1.record()
    torch.cuda.synchronize()
    st = cl.last_spectral_stats
    print(f"n={n} k={k} p={p}: {st.method} outer {st.outer} products {st.gemms} resid {st.max_resid:.2e}; {e0.elapsed_time(e1):.2f} ms per solve, "
          f"{1e3 * e0.elapsed_time(e1) / max(st.gemms, 1):.1f} us per product incl. the small dense steps")


if __name__ == "__main__":
    main()
