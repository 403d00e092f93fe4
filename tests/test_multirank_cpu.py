"""world_size-2 gloo tests (CPU) of the N>1 path's host-side logic: token-balanced document shards,
global counter bases, the count exchange and the vocabulary-sliced Phi normaliser give the same state
as one rank.  The arithmetic is the oracle's (the CUDA kernels need a GPU); what is under test is the
sharding / exchange scheme of DESIGN.md section 6."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, make_corpus


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, scheme, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import ldagroupedgibbssampler_b200 as L
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V, K, beta, seed = 120, 24, 0.05, 11
    alpha = np.full(K, 0.4)
    off, tokens = make_corpus(90, V, 25, seed=4, empty_every=10)
    d0, d1 = L.shard_documents_by_tokens(off, world)[rank]
    loff, ltok, doc_base, token_base = L.take_shard(off, tokens, d0, d1)
    z_all = O.java_next_ints(seed, K, len(tokens))          # the sequential stream; a rank skips token_base draws
    z = z_all[token_base: token_base + len(ltok)].copy()

    def exchange(z):
        n_wk, n_k = O.rebuild_counts(ltok, z, V, K)
        t = torch.from_numpy(n_wk.astype(np.int32)); dist.all_reduce(t)      # reduce-scatter + all-gather
        k = torch.from_numpy(n_k.astype(np.int32)); dist.all_reduce(k)
        return t.numpy(), k.numpy()

    n_wk, n_k = exchange(z)
    phi = O.phi_contract(n_wk, beta, seed, 0)               # cell-indexed counters: every rank draws the same Phi
    for it in (1, 2):
        if scheme == O.GGS:
            th = O.theta_contract(loff, z, K, alpha, seed, it, doc_base)
            z = O.z_ggs_contract(loff, ltok, z, K, th, phi, seed, it, token_base)
        else:
            z = O.z_pcgs_contract(loff, ltok, z, K, alpha, phi, seed, it, token_base)
        n_wk, n_k = exchange(z)
        phi = O.phi_contract(n_wk, beta, seed, it)
    # log-likelihood: document part per shard, type part once
    zs = [None] * world
    dist.all_gather_object(zs, z)
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), z=np.concatenate(zs), n_wk=n_wk, n_k=n_k, phi=phi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scheme", [0, 1])
def test_two_ranks_match_one(tmp_path, oracle, scheme):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), scheme, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(tmp_path, "sharded.npz"))
    V, K, beta, seed = 120, 24, 0.05, 11
    alpha = np.full(K, 0.4)
    off, tokens = make_corpus(90, V, 25, seed=4, empty_every=10)
    z0 = oracle.java_next_ints(seed, K, len(tokens))
    n_wk0, _ = oracle.rebuild_counts(tokens, z0, V, K)
    st = oracle.sweeps("contract", scheme, off, tokens, z0, V, K, alpha, beta, seed, 1, 2,
                       oracle.phi_contract(n_wk0, beta, seed, 0))
    assert np.array_equal(got["z"], st["z"])
    assert np.array_equal(got["n_wk"], st["n_wk"]) and np.array_equal(got["n_k"], st["n_k"])
    assert np.array_equal(got["phi"], st["phiT"])


def test_phi_normaliser_tree_is_rank_count_independent(oracle):
    """The Phi normaliser's fixed tree (8 vocabulary segments) must give the same sums whether 1, 2, 4
    or 8 ranks own the segments: emulate rank-local segment sums and the all-gather."""
    rng = np.random.default_rng(0)
    V, K = 700, 9
    Vp = (V + 63) // 64 * 64
    g = np.zeros((Vp, K), np.float32)
    g[:V] = rng.gamma(0.3, size=(V, K)).astype(np.float32)
    seg_rows = Vp // 8

    def seg_sum(s):
        acc = np.zeros(K)
        for b in range(seg_rows // 8):
            part = np.zeros(K)
            for w in range(s * seg_rows + b * 8, s * seg_rows + b * 8 + 8):
                part = part + g[w].astype(np.float64)
            acc = acc + part
        return acc

    def tree(s):
        return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]))

    ref = tree([seg_sum(s) for s in range(8)])
    for world in (2, 4, 8):
        gathered = []
        for r in range(world):                    # rank r owns segments [8r/G, 8(r+1)/G)
            gathered += [seg_sum(s) for s in range(r * 8 // world, (r + 1) * 8 // world)]
        assert np.array_equal(tree(gathered), ref)
    # and it is what the oracle's contract Phi draw divides by: columns of the result sum to 1
    n_wk = rng.integers(0, 5, size=(V, K)).astype(np.int32)
    phi = oracle.phi_contract(n_wk, 0.1, 3, 1)
    assert np.all(np.abs(phi.astype(np.float64).sum(axis=0) - 1) < 1e-5)
