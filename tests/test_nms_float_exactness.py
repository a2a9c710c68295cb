"""sobel_nms_kernel (csrc/preprocess.cu) does OpenCV's integer Canny arithmetic in fp32.  This test proves, exhaustively over
every Sobel gradient pair an 8-bit image can produce, that the fp32 formulation takes the same decisions as the integer one
(cv::Canny with L2gradient, src/SolveDVO.cpp:1705,1767):
    horizontal  <=>  |dy| * 2^15 <  |dx| * 13573            (fp32:  |dy| < |dx| * (13573 / 2^15))
    vertical    <=>  |dy| * 2^15 >  |dx| * (13573 + 2^16)   (fp32:  |dy| - 2 |dx| > |dx| * (13573 / 2^15))
    magnitude   dx^2 + dy^2 as an exact fp32 integer, "m >= n" as "m > n - 0.5", sign test dx * dy < 0 where it is read.
CPU only: numpy float32 arithmetic is IEEE, like the kernel's (every intermediate is an exactly representable number, so FMA
contraction cannot change a result either)."""
import numpy as np


def test_direction_classes_and_magnitudes_match_the_integer_formulation():
    v = np.arange(-1020, 1021, dtype=np.int64)
    K = np.float32(13573.0 / 32768.0)
    assert float(K) * 32768.0 == 13573.0                                     # dyadic: exactly representable
    for dx in v[::1]:
        dy = v
        ax, ay = abs(int(dx)), np.abs(dy)
        # integer formulation (OpenCV)
        tg22x = ax * 13573
        horiz_i = (ay << 15) < tg22x
        vert_i = (ay << 15) > tg22x + (ax << 16)
        neg_i = (int(dx) ^ dy) < 0
        mag_i = int(dx) * int(dx) + dy * dy
        # fp32 formulation (the kernel)
        fdx, fdy = np.float32(dx), dy.astype(np.float32)
        adx, ady = np.abs(fdx), np.abs(fdy)
        t = np.float32(adx * K)
        assert float(t) * 32768.0 == float(ax * 13573)                        # the product is exact
        u = (np.float32(-2.0) * adx + ady).astype(np.float32)
        horiz_f = ady < t
        vert_f = u > t
        neg_f = (fdx * fdy) < np.float32(0.0)
        mag_f = (fdx * fdx + fdy * fdy).astype(np.float32)
        assert np.array_equal(horiz_i, horiz_f), dx
        assert np.array_equal(vert_i, vert_f), dx
        assert not np.any(horiz_i & vert_i)
        diag = ~horiz_i & ~vert_i & (mag_i > 0)
        assert np.array_equal(neg_i[diag], neg_f[diag]), dx                  # the sign is only read in the diagonal case
        assert np.array_equal(mag_f.astype(np.int64), mag_i), dx             # < 2^24: exact


def test_half_offset_comparison_is_greater_or_equal_on_integers():
    rng = np.random.default_rng(5)
    m = rng.integers(0, 2_080_801, 200_000).astype(np.float32)
    n = np.concatenate([m[:100_000] + rng.integers(-2, 3, 100_000).astype(np.float32), rng.integers(0, 2_080_801, 100_000).astype(np.float32)])
    n = np.clip(n, 0, 2_080_800).astype(np.float32)
    assert np.array_equal(m > (n - np.float32(0.5)).astype(np.float32), m >= n)
    big = np.float32(2_080_800.0)
    assert float(big - np.float32(0.5)) == 2_080_799.5                        # representable: ulp 0.25 at 2^21


def test_byte_to_float_trick_is_exact():
    b = np.arange(256, dtype=np.uint32)
    f = (b | np.uint32(0x4B000000)).view(np.float32) - np.float32(8388608.0)
    assert np.array_equal(f, b.astype(np.float32))
