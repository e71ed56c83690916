#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_sharded.py: labels of the row-sharded DBSCAN (N ranks, NCCL) must equal the
single-GPU fused DBSCAN bit for bit on a clustered synthetic feature matrix (sorted by a 'parameter' row like
MCBB.solve's output, so the per-row cost varies along the index).  Prints the time of both on rank 0."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from __graft_entry__ import load_package
    m = load_package()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda")
    m.check(m.lib().mcbb_init(torch.cuda.current_device()))
    os.dup2(2, 1) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl")
    n, F = int(os.environ.get("N", 60000)), 40
    rng = np.random.default_rng(5)
    par = np.sort(rng.uniform(0, 8, n))
    centers = rng.normal(0, 3, size=(6, F))
    which = np.minimum((par / 8 * 6).astype(int), 5)
    spread = np.where(par < 2.5, 2.5, 0.4)[:, None]  # the low-parameter end is diffuse (noise / border points)
    pts = centers[which] + spread * rng.normal(0, 1, size=(n, F))
    X = torch.from_numpy(np.ascontiguousarray(np.concatenate([pts.T, par[None]], axis=0))).to(dev)
    gl = np.array([F, 1], dtype=np.int32)
    gw = np.array([1.0, 1.0])
    eps, minpts = 0.62 * F * 0.4 * 2, 15
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a1, s1, c1 = m.dist.dbscan_features_device(X, gl, gw, eps, minpts)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    a2, s2, c2 = m.dist.dbscan_features_sharded(X, gl, gw, eps, minpts, rank, world, dist if world > 1 else None)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ok = bool(torch.equal(a1, a2) and torch.equal(s1, s2) and torch.equal(c1, c2))
    print(f"rank {rank}/{world}: n={n} clusters={s1.numel()} noise={(a1 == 0).sum().item()} equal={ok} "
          f"fused {t1 - t0:.3f}s sharded {t2 - t1:.3f}s", file=sys.stderr, flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
