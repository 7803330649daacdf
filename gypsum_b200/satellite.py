"""GpsSatellite: the replica holder the detector / tracker are constructed with (reference gypsum/satellite.py:8-31).

Same attributes and semantics as the reference's dataclass -- `satellite_id`, `prn_code`, `scale_factor`, hashed by the
satellite id, compared field by field -- written as a plain class; `prn_as_complex` is the sampled +-1 replica."""
from __future__ import annotations

import numpy as np

from gypsum_b200.gps_ca_prn_codes import GpsReplicaPrnSignal, GpsSatelliteId

ALL_SATELLITE_IDS = [GpsSatelliteId(prn) for prn in range(1, 33)]  # satellite.py:8


class GpsSatellite:
    _FIELDS = ("satellite_id", "prn_code", "scale_factor")

    def __init__(self, satellite_id: GpsSatelliteId, prn_code: GpsReplicaPrnSignal, scale_factor: int) -> None:
        self.satellite_id = satellite_id
        self.prn_code = prn_code
        self.scale_factor = scale_factor
        self._replica = None

    def _key(self):
        return tuple(getattr(self, f) for f in self._FIELDS)

    def __eq__(self, other) -> bool:
        return other.__class__ is self.__class__ and self._key() == other._key()

    def __hash__(self) -> int:
        return hash(self.satellite_id)

    def __repr__(self) -> str:
        return "GpsSatellite(" + ", ".join(f"{f}={getattr(self, f)!r}" for f in self._FIELDS) + ")"

    @property
    def prn_as_complex(self) -> np.ndarray:
        """satellite.py:20-31: each chip held for scale_factor samples, {0, 1} -> {-1, +1}, complex128 (computed once)."""
        if self._replica is None:
            chips = np.asarray(self.prn_code.inner)
            self._replica = (2.0 * np.repeat(chips, self.scale_factor) - 1.0).astype(complex)
        return self._replica
