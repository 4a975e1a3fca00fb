"""Host model of the INT8 tensor-path square (csrc/gemm_i8.cu), bit for bit.

The device kernel is integer-exact, so the host can reproduce its result exactly: slice X
into balanced base-128 digits, form P_c = sum_{s+t=c} D_s D_t with dgemm on integer-valued
doubles (exact below 2^53), and fold sum_c 2^(2e-12-7c) P_c from the smallest weight up in
the kernel's order.  Test infrastructure only (used by tests/ and tools/i8_check.py).
"""
import numpy as np


def slices_of(X: np.ndarray, S: int, bits: int = 7):
    """Balanced digits of `bits` bits (7 or 8); the leading digit is at most 64 in magnitude."""
    vmax = float(np.max(np.abs(X)))
    e = int(np.frexp(vmax)[1])          # vmax = m * 2^e, m in [0.5, 1)  (== ilogb(vmax) + 1)
    q = np.rint(X * 2.0 ** (bits * (S - 1) + 6 - e)).astype(np.int64)
    half = 1 << (bits - 1)
    D = [None] * S
    for s in range(S - 1, 0, -1):
        d = ((q + half) & (2 * half - 1)) - half
        q = (q - d) >> bits
        D[s] = d.astype(np.float64)
    assert np.max(np.abs(q)) <= 64
    D[0] = q.astype(np.float64)
    return D, e


def exact_square(X: np.ndarray, S: int, bits: int = 7, kseg: int = 0) -> np.ndarray:
    """kseg: length of the K segments (a multiple of 128; 0 = the device default, 16384 for 8-bit
    digits and 32768 for 7-bit digits); every (accumulator pair, segment) is folded separately."""
    D, e = slices_of(X, S, bits)
    n = X.shape[0]
    kseg = kseg or (16384 if bits == 8 else 32768)
    out = np.zeros_like(X)
    c0 = S - 2
    while c0 >= -1:
        for k0 in range(0, max(n, 1), kseg):
            sl = slice(k0, min(n, k0 + kseg))

            def group(c):
                acc = np.zeros_like(X)
                for s in range(c + 1):
                    acc += D[s][:, sl] @ D[c - s][sl, :]
                return acc
            out = out + 2.0 ** (2 * e - 12 - bits * (c0 + 1)) * group(c0 + 1)
            if c0 >= 0:
                out = out + 2.0 ** (2 * e - 12 - bits * c0) * group(c0)
        c0 -= 2
    return out
