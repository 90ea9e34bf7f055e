"""The C restatement of the NumPy legacy stream (oracle/mt19937_legacy.c) against NumPy itself."""
import numpy as np
import pytest


@pytest.mark.parametrize("seed", [0, 1, 5, 1791095845, 2147483646, 4294967295])
def test_raw_and_doubles_bit_exact(oracle, seed):
    rs = np.random.RandomState(seed)
    assert np.array_equal(oracle.random_sample_f64(seed, 5000), rs.random_sample(5000))
    st = np.random.RandomState(seed)
    raw = st.randint(0, 2 ** 32, size=2000, dtype=np.uint64).astype(np.uint32)   # one 32-bit word per draw
    assert np.array_equal(oracle.raw_u32(seed, 2000), raw)


@pytest.mark.parametrize("seed,n", [(1791095845, 512 * 1000), (7, 3), (2135392491, 100_001)])
def test_standard_normal_bit_exact(oracle, seed, n):
    ref = np.random.RandomState(seed).standard_normal(n)
    assert np.array_equal(oracle.standard_normal_f64(seed, n), ref)
    assert np.array_equal(oracle.standard_normal_f32(seed, n), ref.astype(np.float32))


@pytest.mark.parametrize("seed0", [1, 3, 7])
def test_seed_sequence_matches_global_state(oracle, seed0):
    np.random.seed(seed0)
    ref = [np.random.randint(np.iinfo(np.int32).max) for _ in range(200)]
    assert oracle.seed_sequence(seed0, 200).tolist() == ref


def test_golden_stream_heads(oracle, golden):
    g = golden("mapping_known_answers.npz")
    for s, head in zip(g["head_seeds"], g["heads"]):
        assert np.array_equal(oracle.standard_normal_f32(int(s), 64), head)
    assert oracle.seed_sequence(1, 16).tolist() == g["seeds_after_seed1"].tolist()
    assert oracle.seed_sequence(3, 16).tolist() == g["seeds_after_seed3"].tolist()
