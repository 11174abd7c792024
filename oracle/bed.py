"""PLINK .bed helpers for the oracle (numpy).  TEST INFRASTRUCTURE ONLY.

Format authority: hail/hail/src/is/hail/io/plink/LoadPlink.scala:37-38 (magic 0x6c 0x1b, mode 1),
:525 (sample i at byte i>>2, bits (i&3)<<1), :475-481 (a2_reference=True codes:
0 -> 2 alt alleles, 1 -> missing, 2 -> 1, 3 -> 0).
"""
import numpy as np

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])
_DECODE = np.array([2.0, np.nan, 1.0, 0.0])
_ENCODE = {0: 3, 1: 2, 2: 0}


def bed_body(raw, n_samples, n_variants):
    raw = np.asarray(raw, dtype=np.uint8)
    stride = (n_samples + 3) // 4
    assert bytes(raw[:3]) == BED_MAGIC, "not a SNP-major .bed"
    assert raw.size == 3 + n_variants * stride, "bed size mismatch"
    return raw[3:].reshape(n_variants, stride)


def decode_rows(rows, n_samples):
    """uint8 [M, stride] -> float64 [M, N] of n_alt_alleles with NaN for missing calls."""
    rows = np.asarray(rows, dtype=np.uint8)
    shifts = np.array([0, 2, 4, 6], dtype=np.uint8)
    codes = (rows[:, :, None] >> shifts[None, None, :]) & 3
    codes = codes.reshape(rows.shape[0], -1)[:, :n_samples]
    return _DECODE[codes]


def encode_rows(x):
    """float64/int [M, N] dosage (NaN or negative = missing) -> uint8 [M, ceil(N/4)] PLINK codes."""
    x = np.asarray(x)
    M, N = x.shape
    code = np.full((M, N), 1, dtype=np.uint8)
    with np.errstate(invalid="ignore"):
        code[x == 0] = 3
        code[x == 1] = 2
        code[x == 2] = 0
    pad = (-N) % 4
    if pad:
        code = np.concatenate([code, np.zeros((M, pad), dtype=np.uint8)], axis=1)  # pad code 0 like PLINK writers
    code = code.reshape(M, -1, 4)
    return (code[:, :, 0] | (code[:, :, 1] << 2) | (code[:, :, 2] << 4) | (code[:, :, 3] << 6)).astype(np.uint8)
