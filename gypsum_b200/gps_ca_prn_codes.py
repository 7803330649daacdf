"""GPS L1 C/A Gold codes for SV 1..32 -- host-side input preparation for the replica table.

Same public surface as reference gypsum/gps_ca_prn_codes.py (GpsSatelliteId :30-45, GpsReplicaPrnSignal :48-52,
generate_replica_prn_signals :134-250), re-implemented with integer shift registers.  The IS-GPS-200 first-ten-chip
octal check of the reference (:192-247) is kept and raises ValueError on mismatch, as there.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np

from gypsum_b200.constants import PRN_CHIP_COUNT

# IS-GPS-200 table 3-Ia: the two G2 stages whose xor forms G2i (reference :145-178).
_G2_TAP_PAIRS = (
    (2, 6), (3, 7), (4, 8), (5, 9), (1, 9), (2, 10), (1, 8), (2, 9), (3, 10), (2, 3), (3, 4), (5, 6), (6, 7), (7, 8),
    (8, 9), (9, 10), (1, 4), (2, 5), (3, 6), (4, 7), (5, 8), (6, 9), (1, 3), (4, 6), (5, 7), (6, 8), (7, 9), (8, 10),
    (1, 6), (2, 7), (3, 8), (4, 9),
)
# IS-GPS-200 "first 10 chips, octal" column (reference :192-225).
_FIRST_TEN_CHIPS_OCTAL = (
    0o1440, 0o1620, 0o1710, 0o1744, 0o1133, 0o1455, 0o1131, 0o1454, 0o1626, 0o1504, 0o1642, 0o1750, 0o1764, 0o1772,
    0o1775, 0o1776, 0o1156, 0o1467, 0o1633, 0o1715, 0o1746, 0o1763, 0o1063, 0o1706, 0o1743, 0o1761, 0o1770, 0o1774,
    0o1127, 0o1453, 0o1625, 0o1712,
)


@dataclass
class GpsSatelliteId:
    id: int

    def __hash__(self) -> int:
        return hash(self.id)

    def __eq__(self, other: Any) -> bool:
        return getattr(other, "id", None) == self.id and hasattr(other, "id")


@dataclass
class GpsReplicaPrnSignal:
    inner: np.ndarray


def _bit(reg: int, stage: int) -> int:
    """Stage numbering follows IS-GPS-200: stage 1 is the input end, stage 10 the output end."""
    return (reg >> (stage - 1)) & 1


def ca_code_chips(satellite_number: int) -> np.ndarray:
    """int64[1023] in {0,1}.  G1: x^10+x^3+1, G2: x^10+x^9+x^8+x^6+x^3+x^2+1, both start all-ones."""
    t1, t2 = _G2_TAP_PAIRS[satellite_number - 1]
    g1 = g2 = 0x3FF
    chips = np.empty(PRN_CHIP_COUNT, dtype=np.int64)
    for n in range(PRN_CHIP_COUNT):
        chips[n] = _bit(g1, 10) ^ _bit(g2, t1) ^ _bit(g2, t2)
        fb1 = _bit(g1, 3) ^ _bit(g1, 10)
        fb2 = _bit(g2, 2) ^ _bit(g2, 3) ^ _bit(g2, 6) ^ _bit(g2, 8) ^ _bit(g2, 9) ^ _bit(g2, 10)
        g1 = ((g1 << 1) & 0x3FF) | fb1
        g2 = ((g2 << 1) & 0x3FF) | fb2
    return chips


def generate_replica_prn_signals() -> dict[GpsSatelliteId, GpsReplicaPrnSignal]:
    out = {}
    for sv in range(1, 33):
        chips = ca_code_chips(sv)
        head = int("".join(str(int(c)) for c in chips[:10]), 2)
        if head != _FIRST_TEN_CHIPS_OCTAL[sv - 1]:
            raise ValueError(f"SV {sv}: generated PRN starts {head:o}, IS-GPS-200 says {_FIRST_TEN_CHIPS_OCTAL[sv - 1]:o}")
        out[GpsSatelliteId(sv)] = GpsReplicaPrnSignal(chips)
    return out
