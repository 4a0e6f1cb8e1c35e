"""Regenerates tests/golden/seeders.npz from the reference's own seeders (oracle/_ref/libpu_ref.so:
CreateParticleSeeder<T> with T = Particle and T = LWParticle, src/Sim/IParticleSeeder.hpp:29-50).
Run where /root/reference exists:

    python tests/golden/make_golden_seeders.py

Keys: <seeder>_<record>_n<N>_s<seed>_x<scale>[_c] -> raw bytes, one row per record (`_c` = with the
colour ranges CASES_COLOURS below set through Set{Red,Green,Blue}Dist).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = {0: "random", 1: "galaxy", 2: "starsystem"}
CASES_COLOURS = ((0.2, 0.5), (0.3, 1.7), (-1.0, 0.4))   # out-of-range ends exercise the clamp
CASES = [  # kind, n, seed, scale, colours?, lw?
    (0, 64, 0, 1.0, False, False), (0, 64, 5, 4.0, False, True),
    (1, 300, 9, 0.1, False, True), (1, 300, 9, 1.0, True, False), (1, 300, (3 << 21) + 17, 0.1, True, True),
    (2, 64, 0, 1.0, False, False), (2, 64, 0, 1.0, False, True), (2, 1, 0, 1.0, False, False),
]


def key(kind, n, seed, scale, col, lw):
    return f"{NAMES[kind]}_{'lw' if lw else 'p'}_n{n}_s{seed}_x{scale}{'_c' if col else ''}"


def main():
    out = {}
    for kind, n, seed, scale, col, lw in CASES:
        p = ref.seed_ex(n, kind, seed, scale, CASES_COLOURS if col else None, lw)
        out[key(kind, n, seed, scale, col, lw)] = p.view(np.uint8).reshape(n, p.dtype.itemsize)
    path = os.path.join(HERE, "seeders.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
