"""CPU restatement of the keyframe retrieval (TEST INFRASTRUCTURE).

Follows keyframe.cpp:262-346 (KeyFrame::sort_vec_faiss): faiss::IndexFlatIP(512) rebuilt per keyframe over
bank rows [0, index-50] when index >= 50 (else [0, index]); search(nq=1, k=3); results by inner product
descending; faiss pads with (D=-inf, I=-1) when nb < k.  faiss 1.7.2 itself is un-vendored (README.md:20):
exact inner-product top-k is restated with numpy.  Tie rule: lowest index first.  Parity unpinned.
"""
from __future__ import annotations

import numpy as np

K = 3
EXCLUDE_RECENT = 50


def nb_limit(index: int, exclude: int = EXCLUDE_RECENT) -> int:
    """keyframe.cpp:274-282: searched prefix length for the keyframe with global index `index`."""
    return index - exclude + 1 if index >= exclude else index + 1


def knn_ip(bank: np.ndarray, q: np.ndarray, nb: int, k: int = K):
    """Exact max-inner-product top-k over bank[:nb].  Returns (D [k] f32, I [k] int64)."""
    D = np.full((k,), -np.inf, dtype=np.float32)
    I = np.full((k,), -1, dtype=np.int64)
    nb = int(max(0, min(nb, bank.shape[0])))
    if nb == 0:
        return D, I
    ip = (bank[:nb].astype(np.float32) @ q.astype(np.float32)).astype(np.float32)
    order = np.lexsort((np.arange(nb), -ip.astype(np.float64)))[:k]
    D[:len(order)] = ip[order]
    I[:len(order)] = order
    return D, I


def knn_reference_style(bank: np.ndarray, q: np.ndarray, nb: int, k: int = K):
    """Same result, but mirroring the reference's cost: copies the prefix then 'adds' it (second copy) before the
    search (keyframe.cpp:278-306) - used only as the CPU timing baseline."""
    prefix = np.array(bank[:nb], copy=True)
    index = np.array(prefix, copy=True)
    ip = index @ q
    if nb <= k:
        order = np.argsort(-ip, kind="stable")
    else:
        part = np.argpartition(-ip, k)[:k]
        order = part[np.argsort(-ip[part], kind="stable")]
    D = np.full((k,), -np.inf, np.float32); I = np.full((k,), -1, np.int64)
    D[:len(order)] = ip[order][:k]; I[:len(order)] = order[:k]
    return D, I
