"""torchrun worker: N-GPU data-parallel training step == 1-GPU step on the concatenated batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/dist_train_worker.py

Every rank builds the same model, takes its contiguous utterance shard of ONE global batch, runs two
GradientDescent.process_batch calls (NCCL all-reduce of the flat gradient buffer inside); rank 0 also runs the
same two steps on the whole batch with a second, single-GPU model and compares the updated parameters."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from helpers import O, PYRAMID, make_recognizer, package
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = package()
    cfg = O.make_config(**PYRAMID)
    params = O.init_params(cfg, seed=5, scale=10.0)
    tc = dict(gradient_threshold=2.0, rules=("momentum", "adadelta"), scale=0.5, momentum=0.3, decay_rate=0.95,
              epsilon=1e-6)
    Bg = 4 * world

    def make():
        rec = make_recognizer(cfg, params)
        algo = pkg.GradientDescent(recognizer=rec, step_rule=pkg.step_rule_from_config(tc, dict(max_norm=1.0)))
        algo.initialize()
        return rec, algo
    rec, algo = make()
    batches = [O.synthetic_batch(cfg, B=Bg, T=48, seed=40 + s) for s in range(2)]
    for x, m, labels, lm in batches:
        sl = slice(rank * 4, rank * 4 + 4)
        algo.process_batch(dict(recordings=x[:, sl], recordings_mask=m[:, sl], labels=labels[:, sl], labels_mask=lm[:, sl]))
    got = rec.get_parameter_values()
    cost_dp = float(algo.last_cost.item())
    ok = True
    if rank == 0:
        # single-GPU reference on the concatenated batch: temporarily hide the process group from GradientDescent
        rec1, algo1 = make()
        algo1._world = lambda: (None, 1)
        for x, m, labels, lm in batches:
            algo1.process_batch(dict(recordings=x, recordings_mask=m, labels=labels, labels_mask=lm))
        want = rec1.get_parameter_values()
        worst = 0.0
        for k, v in want.items():
            err = float(np.abs(got[k] - v).max() / max(1e-12, np.abs(v).max()))
            worst = max(worst, err)
        cost1 = float(algo1.last_cost.item())
        print("world %d: worst relative parameter difference after 2 steps %.3e, cost %.6f vs %.6f" % (world, worst, cost_dp, cost1))
        ok = worst <= 1e-5 and abs(cost_dp - cost1) <= 1e-5 * abs(cost1)
    # every replica must hold identical parameters
    flat = torch.cat([torch.as_tensor(v).reshape(-1) for v in got.values()]).cuda()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = bool((ref == flat).all().item())
    flag = torch.tensor([1.0 if (ok and same) else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if flag.item() != 1.0:
        print("rank %d: FAILED (ok=%s identical_replicas=%s)" % (rank, ok, same))
        sys.exit(1)
    if rank == 0:
        print("DIST_TRAIN_OK")


if __name__ == "__main__":
    main()
