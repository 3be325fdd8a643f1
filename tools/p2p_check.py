"""torchrun --nproc-per-node N tools/p2p_check.py — doc-id-range shards, one per GPU: the
fused peer-memory exchange (`ShardedReranker(p2p=True)`: top-k lists stored by the scoring
kernel's epilogue straight into the owner rank's buffer) against the NCCL all-to-all path and
against one unsharded index, bit for bit; then both paths timed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200"))
from fast_forward import _ffx  # noqa: E402
from fast_forward.sharded import ShardedReranker, plan_doc_shards  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(0)  # same data on every rank
n_docs, D = 60_000, 768
nq, C, k = int(os.environ.get("P2P_NQ", "2000")), 1500, 100
cnt = rng.integers(1, 12, n_docs)
off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
vec = rng.standard_normal((int(off[-1]), D), dtype=np.float32)
bounds = plan_doc_shards(cnt, world)
lo, hi = int(bounds[rank]), int(bounds[rank + 1])


def shard_index(d0, d1):
    ix = _ffx.DeviceIndex(D, capacity=int(off[d1] - off[d0]), device=local)
    ix.stage(0, vec[off[d0]:off[d1]])
    ix.set_docs(off[d0:d1 + 1] - off[d0])
    return ix


qv = torch.from_numpy(rng.standard_normal((nq, D), dtype=np.float32)).to(dev)
cand_np = np.concatenate([rng.choice(n_docs, C, replace=False) for _ in range(nq)]).astype(np.int32)
cand = torch.from_numpy(cand_np).to(dev)
lex = torch.from_numpy((rng.integers(0, 40, nq * C) / 2).astype(np.float32)).to(dev)
q_off = (torch.arange(nq + 1, device=dev, dtype=torch.int64) * C).contiguous()

results = {}
for name, p2p in (("all_to_all", False), ("p2p", True)):
    sh = ShardedReranker(shard_index(lo, hi), lo, n_docs, int(off[lo]), int(off[-1]), p2p=p2p)
    for _ in range(3):  # alternate the buffer sets, warm up
        s, p = sh.rerank(2, qv, q_off, cand, lex, 0.1, k, C, gather_result=False)
    torch.cuda.synchronize()
    dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        s, p = sh.rerank(2, qv, q_off, cand, lex, 0.1, k, C, gather_result=False)
    ev[1].record()
    torch.cuda.synchronize()
    full_s, full_p = sh.rerank(2, qv, q_off, cand, lex, 0.1, k, C, gather_result=True)
    results[name] = (s.clone(), p.clone(), full_s.clone(), full_p.clone(), ev[0].elapsed_time(ev[1]) / 10)
    sh.index.close()

a, b = results["all_to_all"], results["p2p"]
ok = all(torch.equal(x, y) for x, y in zip(a[:4], b[:4]))
if rank == 0:
    whole = shard_index(0, n_docs)
    want = whole.rerank_host(2, qv.cpu().numpy(), q_off.cpu().numpy(), cand_np, lex.cpu().numpy(), 0.1, k, want_ff=False)
    ok = ok and (b[3].cpu().numpy() == want["topk_pos"]).all() and \
        (b[2].cpu().numpy().view(np.uint32) == want["topk_score"].view(np.uint32)).all()
    whole.close()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"p2p_check world={world} nq={nq} identical={bool(flag.item())} "
          f"all_to_all_ms={a[4]:.3f} p2p_ms={b[4]:.3f}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
