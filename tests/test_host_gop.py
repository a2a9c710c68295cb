"""GOP<float> / GOP<double> host class (rgbd_odometry_b200/host/GOP.h) against the oracle restatement of src/GOP.cpp."""
import numpy as np

import host_lib as Hh
import oracle_lib as O


def test_gop_class_matches_oracle():
    rng = np.random.default_rng(11)
    n = 17
    kind = np.zeros(n, np.int32); kind[0] = 1; kind[[5, 10]] = 2; kind[13] = 1
    reason = np.full(n, 5, np.int32); reason[0] = 1
    rel = np.zeros((n, 12))
    for i in range(n):
        R, t = O.se3_exp(rng.normal(size=6) * 0.2)
        rel[i, :9], rel[i, 9:] = R.reshape(9), t
    o, ok, orr = O.gop_replay(kind, reason, rel)
    d, dk, dr = Hh.gop_replay(kind, reason, rel, use_float=False)
    assert np.allclose(d, o, atol=1e-13) and np.array_equal(dk, ok) and np.array_equal(dr, orr)
    f, fk, fr = Hh.gop_replay(kind, reason, rel, use_float=True)
    assert np.allclose(f, o, atol=5e-5) and np.array_equal(fk, ok) and np.array_equal(fr, orr)
