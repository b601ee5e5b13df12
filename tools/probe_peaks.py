#!/usr/bin/env python
"""Prints the measured FP64 (DFMA) peak and the shared-memory wavefront cost of 128-bit loads by lane pattern."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from degnorm_b200 import probes      # noqa: E402

print(json.dumps(dict(fp64=probes.fp64_peak("cuda:0"), lds128_cycles_per_request=probes.lds_wavefronts("cuda:0"))))
