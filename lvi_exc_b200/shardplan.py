"""Host-side bookkeeping of the sharded map path, restated from csrc/shard.cu so that it can be exercised without a GPU
(tests/test_multi_gloo.py runs it at world size 2 over gloo):

  balanced_splitters   ownership ranges of the voxel index from the all-reduced 4096-bin histogram
  owner_of             owner rank of a voxel index
  decimation_plan      which of every rank's associated points survive the global every-`step`-th decimation
                       (SurfelAssociation::averageTimeDownSmaple over ALL points in time order, L/src/core/surfel_association.cpp:240-244)
"""
from __future__ import annotations

import numpy as np

HIST_BINS = 4096


def key_histogram(keys: np.ndarray, ncell: int) -> np.ndarray:
    b = (keys.astype(np.uint64) * np.uint64(HIST_BINS)) // np.uint64(ncell)
    return np.bincount(b.astype(np.int64), minlength=HIST_BINS).astype(np.uint64)


def balanced_splitters(hist: np.ndarray, ncell: int, world: int) -> np.ndarray:
    """split[q] = first voxel index owned by rank q (split[0] = 0, split[world] = ncell): the boundary moves to the end of the first
    histogram bin at which the running count reaches q / world of all points"""
    total = int(hist.sum())
    split = np.zeros(world + 1, dtype=np.int64)
    acc, q = 0, 1
    for b in range(HIST_BINS):
        if q >= world:
            break
        acc += int(hist[b])
        while q < world and acc * world >= total * q:
            k = ((b + 1) * ncell + HIST_BINS - 1) // HIST_BINS
            split[q] = min(k, ncell)
            q += 1
    split[q:world] = ncell
    split[world] = ncell
    return split


def owner_of(keys: np.ndarray, split: np.ndarray) -> np.ndarray:
    world = len(split) - 1
    o = np.zeros(len(keys), dtype=np.int64)
    for q in range(1, world):
        o += keys >= split[q]
    return o


def decimation_plan(tots, step: int):
    """tots[q] = associated points of rank q (time order across ranks).  Returns (first[q], count[q]): rank q keeps its local hits
    first, first + step, ... (count of them); their global rank is a multiple of `step`."""
    first, count = [], []
    o = 0
    for t in tots:
        f = (step - o % step) % step
        first.append(f)
        count.append((t - f + step - 1) // step if t > f else 0)
        o += t
    return first, count
