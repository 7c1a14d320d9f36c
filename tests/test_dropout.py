"""The dropout mask function (csrc/dropout.cuh) restated in numpy: statistical quality on CPU.  The GPU tests
(test_dropout_gpu.py) check that the kernels produce exactly this function and use it consistently in the
forward and backward passes."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def fmix32(h):
    h = h.astype(np.uint64) & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & M32
    h ^= h >> np.uint64(16)
    return h


def launch_key(seed, offset):
    a = fmix32(np.array([(seed & 0xFFFFFFFF) ^ 0x9E3779B9], dtype=np.uint64))
    a = fmix32(a ^ np.uint64(seed >> 32))
    a = fmix32(a ^ np.uint64(offset & 0xFFFFFFFF))
    b = fmix32(a ^ np.uint64((offset >> 32) ^ 0x7F4A7C15))
    return a[0], b[0]


def keep_mask(seed, offset, rows, cols, p):
    """bool [rows, cols]: the mask of make_dropout_args(p, seed, offset) — row/col hashes, 32-bit multiply, threshold."""
    ka, kb = launch_key(seed, offset)
    r = np.arange(rows, dtype=np.uint64)
    a = fmix32(fmix32(ka ^ (r & M32)) ^ (r >> np.uint64(32)) ^ kb) | np.uint64(1)
    c = np.arange(cols, dtype=np.uint64)
    b = fmix32(fmix32(np.uint64((seed & 0xFFFFFFFF) ^ 0x632BE5AB) ^ c) + np.uint64(seed >> 32)) | np.uint64(1)
    prod = (a[:, None] * b[None, :]) & M32                      # low 32 bits of the product
    return prod >= np.uint64(min(int(p * 4294967296.0 + 0.5), 4294967295))


def _corr(a, b):
    a, b = a - a.mean(), b - b.mean()
    return (a * b).mean() / np.sqrt((a * a).mean() * (b * b).mean())


def test_keep_rate_and_independence_of_the_mask_function():
    p = 0.1
    rows, cols = 2048, 2048
    m = keep_mask(1234, 7, rows, cols, p)
    d = 1.0 - m.astype(np.float64)
    n = d.size
    sigma = np.sqrt(p * (1 - p) / n)
    assert abs(d.mean() - p) < 5 * sigma
    # per-row and per-column drop rates spread like independent Bernoulli draws
    assert 0.8 < d.mean(1).std() / np.sqrt(p * (1 - p) / cols) < 1.2
    assert 0.8 < d.mean(0).std() / np.sqrt(p * (1 - p) / rows) < 1.2
    # neighbours along keys, along queries, on the diagonal, and at a few lags are uncorrelated
    noise = 5.0 / np.sqrt(n)
    assert abs(_corr(d[:, :-1], d[:, 1:])) < noise and abs(_corr(d[:-1], d[1:])) < noise
    assert abs(_corr(d[:-1, :-1], d[1:, 1:])) < noise and abs(_corr(d[:, :-7], d[:, 7:])) < noise
    # different call offsets / seeds give unrelated masks; the same pair gives the same mask
    assert abs(_corr(d, 1.0 - keep_mask(1234, 8, rows, cols, p))) < noise
    assert abs(_corr(d, 1.0 - keep_mask(1235, 7, rows, cols, p))) < noise
    assert np.array_equal(m, keep_mask(1234, 7, rows, cols, p))
    # no pair of rows (or of columns) is noticeably aligned
    sub = d[:256] - d[:256].mean(1, keepdims=True)
    cm = sub @ sub.T / cols / (p * (1 - p))
    np.fill_diagonal(cm, 0.0)
    assert np.abs(cm).max() < 6.0 / np.sqrt(cols)
    subc = d[:, :256] - d[:, :256].mean(0, keepdims=True)
    cc = subc.T @ subc / rows / (p * (1 - p))
    np.fill_diagonal(cc, 0.0)
    assert np.abs(cc).max() < 6.0 / np.sqrt(rows)
    # gaps between dropped keys of one query are geometric (mean 1/p, variance (1-p)/p^2), not periodic
    gaps = np.concatenate([np.diff(np.flatnonzero(d[i])) for i in range(64)])
    assert abs(gaps.mean() - 1 / p) < 0.5 and 0.8 < gaps.var() / ((1 - p) / p ** 2) < 1.2


def test_thresholds_cover_the_probability_range():
    for p in (0.01, 0.1, 0.5, 0.9):
        m = keep_mask(99, 3, 512, 1024, p)
        assert abs((1.0 - m.mean()) - p) < 5 * np.sqrt(p * (1 - p) / m.size)
