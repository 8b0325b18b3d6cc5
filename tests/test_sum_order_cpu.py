"""The resampler must reproduce torch.sum's CPU reduction order bit for bit (models/mip.py:253 feeds it into the
CDF, and a 1-ulp difference flips searchsorted indices).  This test pins the order csrc/render.cu implements -
ATen's vectorized_inner_sum: 4 interleaved accumulators of 8-float vectors, leftovers into accumulator 0, scalar
tail first in the final lane sum; 4 interleaved scalars below 8 elements - against torch.sum itself, using the same
lane mapping as the kernel (warp lane L = accumulator L // 8, vector lane L % 8, element i = L + 32 k)."""
import numpy as np
import pytest
import torch

f32 = np.float32


def kernel_order_sum(x):
    n = len(x)
    if n < 8:
        p = [f32(0)] * 4
        q4 = n // 4
        for i in range(q4):
            for k in range(4):
                p[k] = f32(p[k] + x[i * 4 + k])
        for i in range(q4 * 4, n):
            p[0] = f32(p[0] + x[i])
        return f32(f32(f32(p[0] + p[1]) + p[2]) + p[3])
    vec, ilp = n // 8, n // 32
    local = np.zeros(32, dtype=f32)
    for k in range(ilp):                       # every lane: local += blur[k], element lane + 32 k
        local = local + x[32 * k:32 * k + 32]
    for j in range(ilp * 4, vec):              # lanes 0..7: leftover vectors
        local[:8] = local[:8] + x[j * 8:j * 8 + 8]
    p = ((local[:8] + local[8:16]) + local[16:24]) + local[24:32]
    acc = f32(0)
    for e in range(vec * 8, n):
        acc = f32(acc + x[e])
    for l in range(8):
        acc = f32(acc + p[l])
    return acc


@pytest.mark.parametrize("n", [3, 7, 8, 10, 16, 24, 33, 40, 64, 65, 100, 128, 200, 256])
def test_kernel_sum_order_equals_torch_cpu_sum(n):
    g = torch.Generator().manual_seed(n)
    w = torch.rand(1500, n, generator=g) * torch.rand(1500, 1, generator=g)
    ref = w.sum(-1).numpy()
    got = np.array([kernel_order_sum(w[i].numpy()) for i in range(w.shape[0])])
    assert np.array_equal(got, ref)
