"""Exhaustive device-vs-host sweep of the restated glibc sincosf (SURVEY Appendix B acceptance test): EVERY float
bit pattern, both signs (4 294 967 296 values: |y| < 120, the reduce_large branch above it, Inf, every NaN).  UQS_SKIP_EXHAUSTIVE=1 skips it; the strided +
boundary version in test_gpu_parity.py runs always.  Last full run: profiles/r1_sincosf_exhaustive.log."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.slow
@pytest.mark.skipif(os.environ.get("UQS_SKIP_EXHAUSTIVE") == "1", reason="UQS_SKIP_EXHAUSTIVE=1")
def test_device_sincosf_exhaustive(gpu, oracle):
    hi = 0x80000000                      # every non-negative bit pattern (120.0f = 0x42F00000, Inf = 0x7F800000)
    chunk = 1 << 25
    threads = min(os.cpu_count() or 1, 32)
    bad = 0
    checked = 0
    first = None
    with ThreadPoolExecutor(threads) as pool:
        for lo in range(0, hi, chunk):
            bits = np.arange(lo, min(lo + chunk, hi), dtype=np.uint32)
            for sign in (0, 0x80000000):
                a = (bits | np.uint32(sign)).view(np.float32)
                ds, dc = gpu.sincosf_batch(a)
                parts = np.array_split(np.arange(a.size), threads)
                host = list(pool.map(lambda idx: oracle.libm_sincosf(a[idx[0]:idx[-1] + 1]), parts))
                hs = np.concatenate([h[0] for h in host])
                hc = np.concatenate([h[1] for h in host])
                ne = (ds.view(np.uint32) != hs.view(np.uint32)) | (dc.view(np.uint32) != hc.view(np.uint32))
                n = int(ne.sum())
                if n and first is None:
                    first = float(a[np.flatnonzero(ne)[0]])
                bad += n
                checked += a.size
    print(f"device sincosf vs host libm: {checked} floats checked, {bad} mismatches")
    assert bad == 0, f"{bad} mismatches, first at y={first!r}"
    assert checked == 2 * hi
