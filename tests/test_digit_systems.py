"""CPU checks of the two exact-integer number systems behind the tensor-core sweeps (no GPU).

tc4 (csrc/tc4_kernel.cu digit13 / e2m1_code / imax13): balanced base-13 digits drawn from the E2M1 value set;
tc  (csrc/tc_kernel.cu quantize_kernel): balanced base-256 int8 digits.  These mirror the device code line by line
and pin the claims DESIGN.md makes: complete residue system, contiguous representable range, exactness bounds.
"""
import numpy as np

E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


def digit13(I):
    r = ((I % 13) + 13) % 13
    d = r if r <= 4 else (-8 if r == 5 else 6 if r == 6 else -6 if r == 7 else 8 if r == 8 else r - 13)
    return d, (I - d) // 13


def e2m1_code(u):
    a = abs(u)
    m = a if a <= 4 else (5 if a == 6 else 6)
    return m | (8 if u < 0 else 0)


def test_digit_set_is_a_complete_residue_system_inside_e2m1():
    digits = sorted({digit13(r)[0] for r in range(13)})
    assert digits == [-8, -6, -4, -3, -2, -1, 0, 1, 2, 3, 4, 6, 8]
    assert sorted(d % 13 for d in digits) == list(range(13))
    for u in digits:                                   # every digit / 2 is an E2M1 value and the code decodes to it
        code = e2m1_code(u)
        assert (-1 if code & 8 else 1) * E2M1[code & 7] == u / 2


def test_every_integer_up_to_imax_has_nd_digits():
    rng = np.random.default_rng(0)
    for nd in (1, 2, 3, 4, 9, 13):
        imax = 13 ** nd // 3
        cand = range(-imax, imax + 1) if imax < 30000 else \
            [int(v) for v in rng.integers(-imax, imax + 1, size=20000)] + [imax, -imax]
        for I in cand:
            rest, ds = I, []
            for _ in range(nd):
                d, rest = digit13(rest)
                ds.append(d)
            assert rest == 0 and sum(d * 13 ** k for k, d in enumerate(ds)) == I
    assert np.log2(13 ** 13 // 3) > 46.5 and np.log2(13 ** 9 // 3) > 31.7


def test_exactness_bounds_quoted_in_design():
    # tc4: f32 holds sum c u / 4 exactly while |sum c u| <= 2^24; worst case per sample is c = 3, |u| = 8
    assert (1 << 24) // (3 * 8) == 699050
    # epilogue recombination: two int64 halves of at most 7 digits each never overflow
    assert 7 * (1 << 24) * 13 ** 6 < 2 ** 63 and (1 << 24) * 13 ** 6 * 1.1 < 2 ** 53 * 8
    # tc: int8 digits, A bytes <= 12 (fields left in place as 4c): INT32 cannot overflow below 1.39 M samples
    assert (2 ** 31 - 1) // (12 * 128) > 1_390_000


def test_balanced_base256_digits():
    rng = np.random.default_rng(1)
    for nd in (4, 6):
        imax = 127 * (256 ** nd - 1) // 255
        for I in [imax, -imax, 0] + [int(v) for v in rng.integers(-imax, imax + 1, size=5000)]:
            rest, ds = I, []
            for _ in range(nd):
                d = ((rest + 128) & 255) - 128
                ds.append(d)
                rest = (rest - d) >> 8
            assert rest == 0 and all(-128 <= d <= 127 for d in ds) and sum(d * 256 ** k for k, d in enumerate(ds)) == I
