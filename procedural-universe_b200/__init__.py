"""b200-nbody: B200-native N-body engine behind Procedural-Universe's INBodySim interface.

The product is the C-ABI library ``lib/libnbody_b200.so`` (hand-written sm_100a CUDA, sources in
``csrc/``) and the C++ adapter in ``host/``.  This Python package is only the ctypes binding that
tests and bench.py use to call through that C ABI; it contains no compute and no fallback: if the
library is missing or no B200 is present, calls raise.
"""
from .binding import (  # noqa: F401
    LIB_PATH,
    NBodyError,
    PARTICLE_DTYPE,
    MODE_ALLPAIRS,
    MODE_BARNESHUT,
    Config,
    Sim,
    load,
    seed_galaxy_host,
    seed_host,
    seed_device,
    SeedOptions,
    LWPARTICLE_DTYPE,
    SEEDER_RANDOM,
    SEEDER_GALAXY,
    SEEDER_STARSYSTEM,
    seed_collision_host,
    declared_symbols,
    save_nbody,
    load_nbody,
    recentre,
)
