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


# ---------------------------------------------------------------------------------------------------------------------
# Augmentation and tuple sampling around the loaders (reference :50-142).  Host-side numpy / random, same RNG call order as
# the reference so that a seeded run draws the same angles, noise and tuple members (pinned by tests/golden/host_pipeline.npz,
# generated from the reference's own functions by oracle/gen_golden_host.py).
# ---------------------------------------------------------------------------------------------------------------------
import random  # noqa: E402


def rotate_point_cloud(batch_data):
    """reference :50-71: one random rotation about the up axis per cloud, angle uniform in [-90, 90) degrees;
    [B, N, 3] -> [B, N, 3] float32."""
    rotated_data = np.zeros(batch_data.shape, dtype=np.float32)
    for k in range(batch_data.shape[0]):
        rotation_angle = (np.random.uniform() * np.pi) - np.pi / 2.0
        cosval = np.cos(rotation_angle)
        sinval = np.sin(rotation_angle)
        rotation_matrix = np.array([[cosval, -sinval, 0],
                                    [sinval, cosval, 0],
                                    [0, 0, 1]])
        shape_pc = batch_data[k, ...]
        rotated_data[k, ...] = np.dot(shape_pc.reshape((-1, 3)), rotation_matrix)
    return rotated_data


def jitter_point_cloud(batch_data, sigma=0.005, clip=0.05):
    """reference :74-85: clipped Gaussian noise per coordinate."""
    B, N, C = batch_data.shape
    assert clip > 0
    jittered_data = np.clip(sigma * np.random.randn(B, N, C), -1 * clip, clip)
    jittered_data += batch_data
    return jittered_data


def get_query_tuple(dict_value, num_pos, num_neg, QUERY_DICT, hard_neg=[], other_neg=False, dataset_folder=""):
    """reference :88-142: [query, positives, negatives(, other negative)] for one training entry.  Positives / negatives are
    shuffled IN PLACE in `dict_value` (as the reference does); hard negatives come first, topped up from the shuffled list;
    the extra negative of the quadruplet loss is any entry that is neither a positive of the query nor of a chosen negative."""
    query = load_pc_file(dict_value["query"], dataset_folder)
    random.shuffle(dict_value["positives"])
    pos_files = []
    for i in range(num_pos):
        pos_files.append(QUERY_DICT[dict_value["positives"][i]]["query"])
    positives = load_pc_files(pos_files, dataset_folder)

    neg_files = []
    neg_indices = []
    if len(hard_neg) == 0:
        random.shuffle(dict_value["negatives"])
        for i in range(num_neg):
            neg_files.append(QUERY_DICT[dict_value["negatives"][i]]["query"])
            neg_indices.append(dict_value["negatives"][i])
    else:
        random.shuffle(dict_value["negatives"])
        for i in hard_neg:
            neg_files.append(QUERY_DICT[i]["query"])
            neg_indices.append(i)
        j = 0
        while len(neg_files) < num_neg:
            if not dict_value["negatives"][j] in hard_neg:
                neg_files.append(QUERY_DICT[dict_value["negatives"][j]]["query"])
                neg_indices.append(dict_value["negatives"][j])
            j += 1
    negatives = load_pc_files(neg_files, dataset_folder)

    if other_neg is False:
        return [query, positives, negatives]
    neighbors = []
    for pos in dict_value["positives"]:
        neighbors.append(pos)
    for neg in neg_indices:
        for pos in QUERY_DICT[neg]["positives"]:
            neighbors.append(pos)
    possible_negs = list(set(QUERY_DICT.keys()) - set(neighbors))
    random.shuffle(possible_negs)
    if len(possible_negs) == 0:
        return [query, positives, negatives, np.array([])]
    neg2 = load_pc_file(QUERY_DICT[possible_negs[0]]["query"], dataset_folder)
    return [query, positives, negatives, neg2]
