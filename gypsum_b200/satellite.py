"""GpsSatellite: the replica holder the detector / tracker are constructed with (reference gypsum/satellite.py:8-31)."""
from __future__ import annotations

from dataclasses import dataclass
from functools import cached_property

import numpy as np

from gypsum_b200.gps_ca_prn_codes import GpsReplicaPrnSignal, GpsSatelliteId

ALL_SATELLITE_IDS = [GpsSatelliteId(i + 1) for i in range(32)]  # satellite.py:8


@dataclass
class GpsSatellite:
    satellite_id: GpsSatelliteId
    prn_code: GpsReplicaPrnSignal
    scale_factor: int

    def __hash__(self) -> int:
        return hash(self.satellite_id)

    @cached_property
    def prn_as_complex(self) -> np.ndarray:
        """satellite.py:20-31: each chip repeated scale_factor times, {0,1} -> {-1,+1}, complex128."""
        return (2.0 * np.repeat(np.asarray(self.prn_code.inner), self.scale_factor) - 1.0).astype(complex)
