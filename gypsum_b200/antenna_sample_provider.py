"""Boundary types of the sample stream, mirroring reference gypsum/antenna_sample_provider.py:24-35.

Any object with the same attribute names (e.g. the reference's own dataclasses) is accepted wherever these are.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SampleProviderAttributes:  # antenna_sample_provider.py:24-28
    samples_per_second: int
    samples_per_prn_transmission: int


@dataclass
class AntennaSampleChunk:  # antenna_sample_provider.py:31-35
    start_time: float
    end_time: float
    samples: np.ndarray
