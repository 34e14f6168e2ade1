"""oracle/synth.py -- the CPU restatement of the device's counter-based generator (Philox4x32-10 + Box-Muller).
The Philox rounds are pinned against the published known-answer vectors of Random123 (Salmon et al., SC'11: kat_vectors,
philox4x32 with 10 rounds); the device generator is compared with this restatement on the GPU
(tests/test_gpu_dense.py::test_generator_matches_cpu_restatement), so it is pinned to the same vectors transitively."""
import numpy as np

from oracle import synth

KAT = [  # (counter words, key words, expected output words)
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        got = synth._philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_normal_stream_is_a_function_of_the_element_index_only():
    """any slab / chunking regenerates the same values (what lets every rank and the CPU build the same tensor)"""
    full = synth.normal_range(7, 0, 5000, chunk=1 << 20)
    assert np.array_equal(synth.normal_range(7, 1234, 777, chunk=100), full[1234:2011])
    assert np.array_equal(synth.normal_range(7, 1, 11), full[1:12])             # odd start: second half of a Box-Muller pair first
    T = synth.generate_tensor((10, 9, 8), seed=7, elem_offset=0)
    assert np.array_equal(T.reshape(-1, order="F"), full[:720])
    slab = synth.generate_tensor((10, 9, 3), seed=7, elem_offset=10 * 9 * 4)
    assert np.array_equal(slab, T[:, :, 4:7])
    x = synth.normal_range(3, 0, 200000)
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01 and abs(np.mean(x ** 4) - 3.0) < 0.1
    assert not np.array_equal(synth.normal_range(4, 0, 100), x[:100])


def test_bench_initial_factors_are_unit_columns_from_one_stream():
    f = synth.init_factors((30, 20, 10), 7, seed=1)
    assert [a.shape for a in f] == [(30, 7), (20, 7), (10, 7)]
    for a in f:
        assert np.allclose(np.sum(a * a, axis=0), 1.0, atol=1e-14) and a.flags.f_contiguous
    import bench
    g = bench.init_factors((30, 20, 10), 7, seed=1)
    assert all(np.array_equal(a, b) for a, b in zip(f, g))
