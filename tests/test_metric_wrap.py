"""The FP32 form of the angular wrap used by the k-NN / radius kernels (csrc/knn_common.cuh::wrapped_abs) against the
reference's own arithmetic (src/primitives.h:277-292: float difference, ONE +-2*pi in double, narrowed to float).

numpy float32 arithmetic is IEEE round-to-nearest, the same as the kernels' __fsub_rn / __fadd_rn, so the identity can be
checked here on the CPU for EVERY float of the range the fast path is used for (|d| < 12.5; verified up to 14)."""
import numpy as np

HI = np.float32(6.28318548202514648)        # float(2*pi)
LO = np.float32(1.74845553146951715e-07)    # float(2*pi) - 2*pi, nearest float
PI_F = np.float32(np.pi)


def fast_wrap(a):
    """a = |d| >= 0, float32 array"""
    t = (a - HI) + LO
    return np.minimum(a, np.abs(t))


def reference_wrap(a):
    """|NormalizeAngle<float>(d)| for d = +-a: d >= M_PI (double compare) -> float(double(d) - 2*M_PI)"""
    wrapped = np.abs((a.astype(np.float64) - 2 * np.pi).astype(np.float32))
    return np.where(a.astype(np.float64) >= np.pi, wrapped, a)


def floats_between(lo, hi):
    """every float32 in [lo, hi), lo and hi positive"""
    b0, b1 = np.float32(lo).view(np.uint32), np.float32(hi).view(np.uint32)
    return np.arange(int(b0), int(b1), dtype=np.uint32).view(np.float32)


def test_fast_wrap_is_exact_for_every_float_from_pi_to_14():
    a = floats_between(PI_F, 14.0)
    assert len(a) > 18_000_000
    assert a[0].astype(np.float64) >= np.pi > np.nextafter(a[0], np.float32(0)).astype(np.float64)
    np.testing.assert_array_equal(fast_wrap(a).view(np.uint32), reference_wrap(a).view(np.uint32))


def test_fast_wrap_below_pi_returns_the_difference_itself():
    r = np.random.RandomState(0)
    a = np.concatenate([floats_between(3.0, PI_F), floats_between(1e-3, 1.0001e-3), np.float32([0.0, 1e-30, 1e-45]),
                        r.uniform(0, np.pi, 2_000_000).astype(np.float32)])
    a = a[a.astype(np.float64) < np.pi]
    np.testing.assert_array_equal(fast_wrap(a).view(np.uint32), a.view(np.uint32))
    np.testing.assert_array_equal(reference_wrap(a).view(np.uint32), a.view(np.uint32))


def test_fast_wrap_is_not_exact_far_out_which_is_why_the_wide_path_exists():
    """beyond 2*float(2*pi) the first subtraction rounds (no Sterbenz) and the double rounding shows: the kernels switch to the
    reference's float -> double -> float route from 12.5 on (wide_needed / kWideFrom)"""
    a = floats_between(14.0, 200.0)[::7]
    assert np.any(fast_wrap(a).view(np.uint32) != reference_wrap(a).view(np.uint32))


def test_oracle_metric_uses_the_reference_wrap():
    import oracle as O
    O.build(ref=False)
    r = np.random.RandomState(1)
    for _ in range(200):
        p, q = np.zeros(6, np.float32), np.zeros(6, np.float32)
        p[3:], q[3:] = r.uniform(-60, 60, 3), r.uniform(-60, 60, 3)
        d = (q[3:] - p[3:]).astype(np.float32)
        w = np.where(d.astype(np.float64) < -np.pi, (d.astype(np.float64) + 2 * np.pi).astype(np.float32),
                     np.where(d.astype(np.float64) >= np.pi, (d.astype(np.float64) - 2 * np.pi).astype(np.float32), d))
        want = np.float32(0)
        for x in w:
            want = np.float32(want + np.float32(x * x))
        assert np.float32(O.d6_float(p, q)).view(np.uint32) == want.view(np.uint32)
