"""The bench's noise_sources point on its own: python scripts/rng_bench.py [--small]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dfd_starter_b200.device import get_context  # noqa: E402

print(json.dumps(bench.rng_noise_source_point(get_context(0), full="--small" not in sys.argv,
                                               rows_only="--rows-only" in sys.argv), indent=1))
