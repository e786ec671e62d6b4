"""Oracle for the embedding-search front end (test infrastructure only; see oracle/__init__.py).  PARITY UNPINNED: the
reference ships no tests or golden vectors for this path and the script cannot run as committed (see below), so this
restatement is pinned only by reading the source.

Reference lines restated (embedding_search/similarity_search.py):
  :25-29   generation embeddings + file names from embedding.pkl
  :34-36   gallery sub-folders, sorted
  :39-41   torch.chunk(queries, num_chunks)
  :46-48   per chunk: running best score = -1 (float64), running best key = 0
  :50-56   per folder (sorted): load embedding.pkl; on any exception print and skip
  :62-63   dist = features @ batch.T ; (score, row) = dist.max(dim=0)
  :66-67   key = folder + ':' + keys[row]
  :70-74   merge: argmax over vstack([previous, current]) -> previous wins ties
  :82-88   concatenate chunks -> {'scores', 'keys', 'gen_images'}
Bugs in the committed script that are NOT reproduced: `args.laion_embeddings_folders` (:34,
attribute does not exist), `os.path.join(laion_folder, 'embedding.pkl')` without the root folder (:52),
`pkl.dump(f, dump_dict)` with swapped arguments into `open(dump_dict, 'wb')` (:90-91).

Scores follow the oracle contract of oracle/similarity.py (dot products in float64 from the float32 inputs, rounded
to float32, lowest row wins ties), then widened to float64 as the reference's accumulator is.
"""
from __future__ import annotations

import os
import pickle as pkl

import numpy as np


def torch_chunk_bounds(n: int, chunks: int):
    """Row ranges of torch.chunk(x, chunks) along dim 0: ceil(n/chunks) rows each, possibly fewer chunks."""
    if n == 0:
        return []
    size = -(-n // chunks)
    return [(s, min(n, s + size)) for s in range(0, n, size)]


def similarity_search(laion_embedding_folder: str, generation_embedding_path: str, num_chunks: int = 100) -> dict:
    with open(generation_embedding_path, "rb") as f:                               # :25-29
        data = pkl.load(f)
    gen = np.asarray(data["features"], dtype=np.float32)
    gen_images = data["indexes"]
    folders = sorted(x for x in os.listdir(laion_embedding_folder)                  # :34-35
                     if os.path.isdir(os.path.join(laion_embedding_folder, x)))
    top_scores, top_keys = [], []
    for lo, hi in torch_chunk_bounds(gen.shape[0], num_chunks):                     # :39-45
        batch = gen[lo:hi]
        best_s = -np.ones(batch.shape[0])                                          # :47 (float64)
        best_k = np.zeros(batch.shape[0])                                          # :48
        for folder in folders:                                                     # :50
            try:
                with open(os.path.join(laion_embedding_folder, folder, "embedding.pkl"), "rb") as f:
                    d = pkl.load(f)
                feats = np.asarray(d["features"], dtype=np.float32)
                keys = d["indexes"]
            except Exception as e:                                                 # :54-56
                print(e)
                continue
            if feats.shape[0] == 0:
                continue
            dist = feats.astype(np.float64) @ batch.astype(np.float64).T           # :62 [G_f, Q_c]
            row = dist.argmax(axis=0)                                              # :63 (first maximum = lowest row)
            cur_s = dist[row, np.arange(batch.shape[0])].astype(np.float32).astype(np.float64)
            cur_k = np.array([folder + ":" + str(keys[r]) for r in row])           # :66-67
            temp_s = np.vstack([best_s, cur_s])                                    # :70
            temp_k = np.vstack([best_k, cur_k])                                    # :71 (numeric 0 -> '0.0')
            mx = temp_s.argmax(axis=0).reshape(1, -1)                              # :72
            best_s = np.take_along_axis(temp_s, mx, axis=0).reshape(-1)            # :73
            best_k = np.take_along_axis(temp_k, mx, axis=0).reshape(-1)            # :74
        top_scores.append(best_s)
        top_keys.append(np.array([str(k) for k in best_k]))
    scores = np.concatenate(top_scores) if top_scores else np.zeros((0,))
    keys = np.concatenate(top_keys) if top_keys else np.array([], dtype=str)
    return {"scores": scores, "keys": keys, "gen_images": gen_images}              # :84-86
