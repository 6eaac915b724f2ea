"""Host-side logic of the data-parallel path, exercised with 2 CPU processes over gloo (no GPU):
batch sharding covers disjoint full batches, and ONE all-reduce of the flat [dW | dbh | dbv] buffer + 1/P scaling equals
the oracle's "P batches, mean of gradients" step (mode A)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import random_csr, mask_csr, xavier, rel_err


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from dae_rnn_news_recommendation_b200.autoencoder import utils
    from oracle.dae_oracle import OracleDAE
    N, F, H, B = 200, 120, 8, 40
    x = random_csr(N, F, 8, seed=0)
    xc, _ = mask_csr(x, 0.3)
    labels = np.random.default_rng(1).integers(0, 3, N).astype(np.float32)
    np.random.seed(5)
    perm = np.random.permutation(N)                     # identical on every rank (same seed)
    starts = utils.shard_batch_starts(N, B, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, starts)
    model = OracleDAE(xavier(F, H, 2))
    flat_sizes = [F * H, H, F]
    for s in starts:                                     # each rank: local gradient of ITS batch, then one all-reduce
        idx = perm[s:s + B]
        outp = model.forward(x[idx], xc[idx], labels[idx])
        g = torch.cat([t.reshape(-1) for t in model.grads(outp)])
        dist.all_reduce(g)                               # the single collective of the step
        g = g / world
        model.apply_gradients([t.view_as(p) for t, p in zip(torch.split(g, flat_sizes), model.params())])
    if rank == 0:
        out.put((gathered, model.get_parameters()['enc_w']))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_dp_matches_mean_of_gradients():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, w_dp = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # sharding: disjoint, full batches only, same number of steps per rank
    assert len(gathered[0]) == len(gathered[1]) == 2 and not set(gathered[0]) & set(gathered[1])
    assert sorted(gathered[0] + gathered[1]) == [0, 40, 80, 120]
    # single-process reference of mode A: per global step, mean of the P local gradients
    from oracle.dae_oracle import OracleDAE
    N, F, H, B = 200, 120, 8, 40
    x = random_csr(N, F, 8, seed=0)
    xc, _ = mask_csr(x, 0.3)
    labels = np.random.default_rng(1).integers(0, 3, N).astype(np.float32)
    np.random.seed(5)
    perm = np.random.permutation(N)
    model = OracleDAE(xavier(F, H, 2))
    for g0 in range(2):
        grads = []
        for r in range(world):
            idx = perm[(g0 * world + r) * B:(g0 * world + r + 1) * B]
            grads.append(model.grads(model.forward(x[idx], xc[idx], labels[idx])))
        model.apply_gradients([sum(gs) / world for gs in zip(*grads)])
    assert rel_err(w_dp, model.get_parameters()['enc_w']) < 1e-6


def test_shard_batch_starts_single_rank_keeps_short_tail():
    from dae_rnn_news_recommendation_b200.autoencoder import utils
    assert utils.shard_batch_starts(10, 4) == [0, 4, 8]
    assert utils.shard_batch_starts(10, 4, 2, 0) == [0] and utils.shard_batch_starts(10, 4, 2, 1) == [4]
