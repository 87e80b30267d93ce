"""Counter-RNG restatements agree with each other and with the Random123 known answers."""
import numpy as np

from marl_ctf_development_b200 import draws
from oracle import ctf_oracle

# Random123 kat_vectors, philox4x32-10
KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    (
        [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344],
        [0xA4093822, 0x299F31D0],
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1],
    ),
]


def test_philox_known_answers_numpy():
    for ctr, key, want in KAT:
        got = draws.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert got.tolist() == want


def test_philox_known_answers_oracle_c():
    L = ctf_oracle.lib()
    for ctr, key, want in KAT:
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        L.ctf_oracle_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        assert out.tolist() == want


def test_move_order_is_permutation_and_roughly_uniform():
    counts = np.zeros((8, 8), dtype=np.int64)
    for step in range(1, 2001):
        w = draws.step_words(seed=5, env_id=9, episode=0, step=step)
        order = draws.move_order(w, 8)
        assert sorted(order) == list(range(8))
        for slot, a in enumerate(order):
            counts[slot, a] += 1
    assert counts.min() > 150 and counts.max() < 350  # expectation 250


def test_tag_roll_is_exact_threshold():
    w = draws.step_words(1, 2, 3, 4)
    for a in range(8):
        for j in range(4):
            u = draws.tag_roll_uniform(w, a, j)
            assert (u < 0.75) == (int(w[4 * a + j, 0]) < 0xC0000000)


def test_respawn_pick_in_range():
    w = draws.step_words(1, 2, 3, 4)
    for k in range(1, 10):
        assert 0 <= draws.respawn_pick(w, 3, 1, k) < k
