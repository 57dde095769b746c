"""Drop-in for the on-disk submap format of the reference's loading_pointclouds.py: a submap is a raw
little-endian float64 file of 4096 x 3 coordinates (reference :26-35); a batch of them becomes one [n, 4096, 3] array
(:38-47).  Host-side only (disk I/O, no arithmetic): the float64 -> float32 narrowing happens once, in `to_model_input`,
on the way into pinned memory for the embedding driver (`lpdnet_b200.evaluate.get_latent_vectors`).
"""
from __future__ import annotations

import os

import numpy as np

NUM_POINTS = 4096  # reference config.py / loading_pointclouds.py:30


def load_pc_file(filename, dataset_folder: str = "", num_points: int = NUM_POINTS):
    """reference :26-35: -> [num_points, 3] float64, or an empty array (and a message) when the file has another size."""
    pc = np.fromfile(os.path.join(dataset_folder, filename), dtype=np.float64)
    if pc.shape[0] != num_points * 3:
        print("Error in pointcloud shape")
        return np.array([])
    return np.reshape(pc, (pc.shape[0] // 3, 3))


def load_pc_files(filenames, dataset_folder: str = "", num_points: int = NUM_POINTS):
    """reference :38-47: files of the wrong size are skipped, the rest stacked -> [n, num_points, 3] float64."""
    pcs = []
    for filename in filenames:
        pc = load_pc_file(filename, dataset_folder, num_points)
        if pc.shape[0] != num_points:
            continue
        pcs.append(pc)
    return np.array(pcs)


def to_model_input(pcs):
    """[n, N, 3] float64 (load_pc_files) -> [n, 1, N, 3] float32 contiguous, the layout every caller feeds the model
    (reference train_pointnetvlad.py:204-207, evaluate.py:112-117)."""
    a = np.ascontiguousarray(np.asarray(pcs), dtype=np.float32)
    return a.reshape(a.shape[0], 1, a.shape[1], 3)
