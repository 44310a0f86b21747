"""world_size-2 gloo checks of the multi-rank host logic of bench.py (no GPU): utterance shards
are disjoint per rank, the step time is the max over ranks, the reference arm runs on rank 0 only."""
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    x, m, labels, lm = bench.synthetic_batch(4, 64, 40, 8, 32, seed=bench.shard_seed(rank))
    local_ms = 10.0 + 5.0 * rank                       # rank 1 is the slow one
    step_ms = bench.max_over_ranks(local_ms, world, None, dist)
    sums = torch.tensor([float(x.sum())], dtype=torch.float64)
    gathered = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    out.put((rank, step_ms, [float(g) for g in gathered], int(lm.sum()), labels.max()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_time():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, t0, g0, n0, v0), (r1, t1, g1, n1, v1) = res
    assert (r0, r1) == (0, 1)
    assert t0 == t1 == 15.0                       # max over ranks, identical on every rank
    assert g0 == g1 and g0[0] != g0[1]            # shards differ between ranks
    assert v0 <= 31 and v1 <= 31 and n0 > 0 and n1 > 0


def test_reference_arm_only_rank0_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_synthetic_batch_contract():
    import bench
    W = bench.WORKLOAD
    x, m, labels, lm = bench.synthetic_batch(8, 200, W["F"], 25, W["V"], seed=3)
    assert x.shape == (200, 8, 40) and m.shape == (200, 8) and labels.shape == (25, 8)
    assert m.sum(axis=0).max() == 200 and m.sum(axis=0).min() >= 120          # lengths in [0.6T, T]
    assert np.all(x[m == 0] == 0)
    last = (lm.sum(axis=0) - 1).astype(int)
    assert np.all(labels[last, np.arange(8)] == W["V"] - 1)                   # eos closes every label sequence
    assert bench.attention_step_bytes(64, 250, 512, 512) == 65792000          # 65.8 MB / step


def _grad_worker(rank, world, port, out):
    """Data-parallel training step, host logic only: per-rank gradient SUMS of an utterance shard (float64 oracle
    on a tiny model), the single all-reduce of [grads | batch | cost] (algorithms.allreduce_step_buffer), then the
    oracle's step rules with 1 / global batch -- must equal the single-process step on the concatenated batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as graft
    from oracle import lvsr_oracle as O
    from oracle import lvsr_oracle_grad as G
    pkg = graft.load_package()
    cfg = O.make_config(num_features=5, dims_bidir=[4], subsample=[1], dim_dec=6, dim_matcher=8, conv_n=3,
                        conv_num_filters=2, num_phonemes=5, post_merge_dims=[6], maxout_pieces=2)
    params = O.init_params(cfg, seed=4, weights_std=0.3)
    x, m, labels, lm = O.synthetic_batch(cfg, B=6, T=12, seed=5, label_div=4)
    sl = slice(rank * 3, rank * 3 + 3)
    local = (x[:, sl], m[:, sl], labels[:, sl], lm[:, sl])
    cost_mean, grads_mean = G.cost_and_grads(cfg, params, *local)        # per-shard mean; the CUDA call returns sums
    names = list(params)
    sizes = [params[k].size for k in names]
    n = sum(sizes)
    buf = torch.zeros(n + 64, dtype=torch.float32)
    buf[:n] = torch.as_tensor(np.concatenate([(grads_mean[k] * 3).ravel() for k in names]), dtype=torch.float32)
    bg, cost_sum = pkg.algorithms.allreduce_step_buffer(buf, n, 3, torch.tensor(cost_mean * 3, dtype=torch.float32), dist)
    out.put((rank, int(round(float(bg))), float(cost_sum), buf[:n].numpy().copy(), names, sizes))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_process_step():
    from oracle import lvsr_oracle as O
    from oracle import lvsr_oracle_grad as G
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + ((os.getpid() + 7) % 1000)
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((out.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = O.make_config(num_features=5, dims_bidir=[4], subsample=[1], dim_dec=6, dim_matcher=8, conv_n=3,
                        conv_num_filters=2, num_phonemes=5, post_merge_dims=[6], maxout_pieces=2)
    params = O.init_params(cfg, seed=4, weights_std=0.3)
    batch = O.synthetic_batch(cfg, B=6, T=12, seed=5, label_div=4)
    cost, grads = G.cost_and_grads(cfg, params, *batch)                   # sum / 6 on the concatenated batch
    for rank, bg, cost_sum, flat, names, sizes in res:
        assert bg == 6
        assert abs(cost_sum / bg - cost) < 1e-5 * abs(cost)
        off = 0
        for k, sz in zip(names, sizes):
            got = flat[off:off + sz].reshape(params[k].shape) / bg
            assert np.abs(got - grads[k]).max() <= 1e-5 * max(1e-6, np.abs(grads[k]).max()), k
            off += sz
    assert np.array_equal(res[0][3], res[1][3])                           # every replica holds the same reduced buffer
