"""numpy mirror of csrc/pack.cu:bn_fill_kernel (test infrastructure): same counter-based bits, same thresholds."""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
    return z ^ (z >> np.uint64(31))


def bn_fill_numpy(thresholds, pop, n_samples, seed=0, first_variant=0):
    """-> int8 dosage [M, N] with -1 = missing."""
    M = thresholds.shape[0]
    n_words = (n_samples + 15) // 16
    with np.errstate(over="ignore"):
        v = (np.arange(M, dtype=np.uint64) + np.uint64(first_variant))[:, None]
        w = np.arange(n_words, dtype=np.uint64)[None, :]
        base = np.uint64(seed) ^ (v * np.uint64(0xD1B54A32D192ED03)) ^ (w * np.uint64(0x8CB92BA72F3D8DD7))
        out = np.empty((M, n_words * 16), dtype=np.int8)
        pop_pad = np.zeros(n_words * 16, dtype=np.int64)
        pop_pad[:n_samples] = pop
        for q in range(4):
            bits = mix64(base + np.uint64(q))
            for e in range(4):
                j = q * 4 + e
                u = ((bits >> np.uint64(16 * e)) & np.uint64(0xFFFF)).astype(np.int64)  # [M, n_words]
                p = pop_pad[j::16][None, :]  # population of sample 16w + j
                t = thresholds[np.arange(M)[:, None], p]  # [M, n_words, 3]
                code = np.where(u < t[..., 0], -1, np.where(u < t[..., 1], 0, np.where(u < t[..., 2], 1, 2)))
                out[:, j::16] = code
    return out[:, :n_samples]
