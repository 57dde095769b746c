"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Restatement of get_recall, reference evaluate.py:162-206.

The reference builds sklearn.neighbors.KDTree(database) (:168) and queries one vector at a time with k=25
(:186-187); KDTree promotes float32 to float64 and ranks by Euclidean distance.  Here the 25 neighbours come
from the fp64 brute-force search in knn_canonical.c (ties to the lower index); everything after the search —
first-hit histogram (:189-198), top-1% test with threshold = max(int(round(len(db)/100.0)), 1) (:174,:200-201),
skip of queries without ground truth (:181-182), cumulative recall (:203-205) — follows the reference line by line.
"""
import numpy as np

from . import retrieval_bruteforce

RECALL_NUM = 25  # evaluate.py:20


def get_recall(m, n, DATABASE_VECTORS, QUERY_VECTORS, QUERY_SETS, recall_num=RECALL_NUM):
    database_output = np.asarray(DATABASE_VECTORS[m], dtype=np.float32)
    queries_output = np.asarray(QUERY_VECTORS[n], dtype=np.float32)
    kq = min(recall_num, len(database_output))
    indices, _ = retrieval_bruteforce(database_output, queries_output, kq)
    recall = [0] * recall_num
    top1_similarity_score = []
    one_percent_retrieved = 0
    threshold = max(int(round(len(database_output) / 100.0)), 1)
    num_evaluated = 0
    for i in range(len(queries_output)):
        true_neighbors = QUERY_SETS[n][i][m]
        if len(true_neighbors) == 0:
            continue
        num_evaluated += 1
        row = indices[i]
        for j in range(len(row)):
            if row[j] in true_neighbors:
                if j == 0:
                    top1_similarity_score.append(np.dot(queries_output[i], database_output[row[j]]))
                recall[j] += 1
                break
        if len(set(row[0:threshold].tolist()).intersection(set(true_neighbors))) > 0:
            one_percent_retrieved += 1
    one_percent_recall = (one_percent_retrieved / float(num_evaluated)) * 100
    recall = (np.cumsum(recall) / float(num_evaluated)) * 100
    return recall, top1_similarity_score, one_percent_recall
