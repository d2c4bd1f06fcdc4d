"""Frame-sharding arithmetic of the multi-GPU keyframe round (host logic).

Round R, rank r of P ranks, b frames per rank: the rank owns the contiguous block of global keyframe indices
``[(R*P + r)*b, (R*P + r)*b + b)``.  One all-gather of [b,512] per rank lands rank-major in every rank's bank, so
bank row == global keyframe index (+ any pre-filled rows) and the sequential reference semantics
(keyframe.cpp:274-282: search rows [0, index-50]) hold exactly for any round size when the gather precedes the search.
"""
from __future__ import annotations

import numpy as np

EXCLUDE_RECENT = 50


def round_frame_ids(R: int, rank: int, world: int, b: int) -> np.ndarray:
    return np.arange(b, dtype=np.int64) + (R * world + rank) * b


def owner_of(frame_id: int, world: int, b: int) -> int:
    return int((frame_id // b) % world)


def nb_limit(index: int, exclude: int = EXCLUDE_RECENT) -> int:
    """Rows searched for the keyframe whose bank row is `index` (keyframe.cpp:274-282)."""
    return index - exclude + 1 if index >= exclude else index + 1


def previous_round_ids(ids: np.ndarray, world: int, b: int) -> np.ndarray:
    """The same rank's frames of the previous round (resident in its feature store)."""
    return ids - world * b
