"""The exact pass ranks its candidates on 64-bit integer keys (csrc/exact.cuh: f64_orderable / f64_from_orderable) instead of
comparing doubles pair by pair.  The mapping is restated here in numpy and checked for the three properties the ranking relies
on: unsigned order == numeric order, the two zeros share one key, the inverse returns the score bit for bit; and for the
near-tie rule of the batched path: key distance == distance in representable doubles."""
import numpy as np

SIGN = np.uint64(1) << np.uint64(63)


def f64_orderable(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64) + 0.0                      # -0.0 -> +0.0, as the device function does
    b = x.view(np.uint64)
    return np.where(b >> np.uint64(63) != 0, ~b, b | SIGN)


def f64_from_orderable(o: np.ndarray) -> np.ndarray:
    o = np.asarray(o, dtype=np.uint64)
    b = np.where(o >> np.uint64(63) != 0, o & ~SIGN, ~o)
    return b.view(np.float64)


def _samples():
    rng = np.random.default_rng(0)
    tiny = np.float64(5e-324)
    special = np.array([0.0, -0.0, tiny, -tiny, 1.0, -1.0, np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0), np.inf, -np.inf,
                        np.finfo(np.float64).max, -np.finfo(np.float64).max, np.finfo(np.float64).tiny, 1e-300, -1e-300])
    cos = rng.uniform(-1.0, 1.0, size=20000)                         # what the scores are
    wide = rng.standard_normal(20000) * np.exp2(rng.integers(-1000, 1000, size=20000).astype(np.float64))
    return np.concatenate([special, cos, wide])


def test_unsigned_order_is_numeric_order_and_zeros_share_a_key():
    x = _samples()
    k = f64_orderable(x)
    order_num = np.argsort(x, kind="stable")
    xs, ks = x[order_num], k[order_num]
    assert np.all(ks[1:] >= ks[:-1])                                 # sorted by value => sorted by key
    assert np.array_equal(ks[1:] == ks[:-1], xs[1:] == xs[:-1])      # equal keys exactly where the values compare equal
    assert f64_orderable(np.array([0.0]))[0] == f64_orderable(np.array([-0.0]))[0]
    assert np.all(k != 0)                                            # key 0 is reserved for an empty candidate slot


def test_inverse_returns_the_bits():
    x = _samples()
    back = f64_from_orderable(f64_orderable(x))
    want = x + 0.0                                                   # -0.0 comes back as +0.0, everything else unchanged
    assert np.array_equal(back.view(np.uint64), want.view(np.uint64))


def test_key_distance_counts_representable_doubles():
    x = np.array([1.0, -1.0, 0.3, -0.3, 1e-300, 123456.789])
    up = x.copy()
    for _ in range(4):
        up = np.nextafter(up, np.inf)
    d = f64_orderable(up).astype(np.int64) - f64_orderable(x).astype(np.int64)     # wraps like uint64 subtraction
    assert np.all(d == 4)
    # across zero the key that -0.0 would have had stays unused: -tiny -> [unused] -> 0 -> +tiny is a distance of 3, i.e. the
    # near-tie rule (distance 1..4) is one step stricter there and exact everywhere else
    tiny = np.float64(5e-324)
    assert int(f64_orderable(np.array([tiny]))[0]) - int(f64_orderable(np.array([-tiny]))[0]) == 3
