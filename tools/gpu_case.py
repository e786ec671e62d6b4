"""Run ONE sim_topk case in this process and print a verdict line (used under `timeout` by tools/gpu_probe.sh so a
hung kernel cannot take the whole gpurun call with it)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dcr_b200 import similarity, synthetic
from oracle import similarity as osim

nq, ng, d, k = map(int, sys.argv[1:5])
q, g = synthetic.descriptors(nq, ng, d, seed=nq + ng)
qc, gc = q.cuda(), g.cuda()
torch.cuda.synchronize()
t0 = time.time()
v, i = similarity.sim_topk(qc, gc, k)
torch.cuda.synchronize()
t1 = time.time()
st = similarity.sim_topk_stats()
rows = np.arange(nq) if nq <= 512 else np.sort(np.random.default_rng(0).choice(nq, 256, replace=False))
ov, oi = osim.sim_topk(q.numpy()[rows], g.numpy(), k)
vi, ii = v.cpu().numpy()[rows], i.cpu().numpy()[rows]
bad = int((ii != oi).any(axis=1).sum())
err = float(np.abs(vi - ov).max())
# timing (warm)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    similarity.sim_topk(qc, gc, k)
ev0.record()
for _ in range(5):
    similarity.sim_topk(qc, gc, k)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 5
print(f"CASE nq={nq} ng={ng} d={d} k={k} cg={st['cta_group']} bad_rows={bad}/{len(rows)} max_score_err={err:.3e} "
      f"flagged={st['n_flagged']} first_call_s={t1 - t0:.3f} ms_per_call={ms:.3f} "
      f"tflops={2.0 * nq * ng * d / ms / 1e9:.1f} stats={st}", flush=True)
