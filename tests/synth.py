"""Seeded synthetic inputs of the test-suite: the generators live in ``meshflow_b200.workloads`` (bench.py uses the
same ones); this module keeps the tests' historical import path."""
from meshflow_b200.workloads import *  # noqa: F401,F403
from meshflow_b200.workloads import (random_homography, synthetic_paths, synthetic_tracks, synthetic_warp_inputs,  # noqa: F401
                                     textured_video)
