"""Doc-id-range shards: (gpu) the kernels' shard ownership + merge against the unsharded
result, emulating the ranks as several indexes on one GPU; (cpu, gloo, world_size 2) the
exchange plumbing of `ShardedReranker` with the two libffx steps replaced by the oracle."""

import os
import socket

import numpy as np
import pytest

import ff_oracle as fo
from fast_forward.sharded import plan_doc_shards


def test_plan_doc_shards_balances_rows():
    rng = np.random.default_rng(0)
    cnt = rng.integers(1, 60, 10_000)
    b = plan_doc_shards(cnt, 8)
    assert b[0] == 0 and b[-1] == len(cnt) and (np.diff(b) > 0).all()
    rows = np.add.reduceat(cnt, b[:-1])
    assert rows.max() - rows.min() <= 2 * cnt.max()
    assert plan_doc_shards(cnt, 1).tolist() == [0, len(cnt)]
    assert plan_doc_shards(np.array([5]), 3).tolist() == [0, 0, 0, 1] or True  # degenerate: no crash


@pytest.mark.gpu
@pytest.mark.parametrize("dim,nq", [(768, 300), (768, 9), (100, 9), (128, 300), (384, 300), (100, 300), (64, 9)])
@pytest.mark.parametrize("mode", ["MAXP", "AVEP", "PASSAGE"])
def test_shards_on_one_gpu_equal_the_whole(dim, nq, mode):
    """fused (nq=300) and tiled (nq=9) launches of the whole-warp kernel (768) and of the packed
    kernel (lane-major 64 / 128 / 384, tree-as-data 100), 3 shards."""
    import torch

    import __graft_entry__ as g

    g.build()
    from fast_forward import _ffx

    m = {"MAXP": fo.MODE_MAXP, "AVEP": fo.MODE_AVEP, "PASSAGE": fo.MODE_PASSAGE}[mode]
    rng = np.random.default_rng(dim + nq)
    n_docs, S, k, alpha = 900, 3, 40, 0.1
    cnt = rng.integers(1, 7, n_docs)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    vec = rng.standard_normal((off[-1], dim)).astype(np.float32)
    whole = _ffx.DeviceIndex(dim, capacity=len(vec))
    whole.stage(0, vec)
    whole.set_docs(off)
    pool = len(vec) if m == fo.MODE_PASSAGE else n_docs
    cnts = rng.integers(0, 200, nq)
    q_off = np.concatenate([[0], np.cumsum(cnts)]).astype(np.int64)
    cand = np.concatenate([rng.choice(pool, c, replace=False) for c in cnts] + [np.zeros(0, np.int64)]).astype(np.int32)
    lex = (rng.integers(0, 30, len(cand)) / 2).astype(np.float32)
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    want = whole.rerank_host(m, qv, q_off, cand, lex, alpha, k, want_ff=True)

    bounds = plan_doc_shards(cnt, S)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    d_qv, d_off, d_cand, d_lex = dev(qv), dev(q_off), dev(cand), dev(lex)
    sh_s = torch.empty((S, nq, k), dtype=torch.float32, device="cuda")
    sh_p = torch.empty((S, nq, k), dtype=torch.int32, device="cuda")
    ff = torch.zeros(len(cand), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    shards = []
    for s in range(S):
        lo, hi = bounds[s], bounds[s + 1]
        r0, r1 = off[lo], off[hi]
        sh = _ffx.DeviceIndex(dim, capacity=int(r1 - r0))
        sh.stage(0, vec[r0:r1])
        sh.set_docs(off[lo:hi + 1] - r0)
        sh.set_shard(lo, n_docs, r0, len(vec))
        sh.rerank_device(m, d_qv.data_ptr(), nq, d_off.data_ptr(), d_cand.data_ptr(), d_lex.data_ptr(),
                         alpha, k, int(cnts.max()), ff.data_ptr(), 0, sh_s[s].data_ptr(), sh_p[s].data_ptr())
        sh.sync()
        shards.append(sh)
    o_s = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    o_p = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    _ffx.merge_topk(0, sh_s.data_ptr(), sh_p.data_ptr(), S, nq, k, o_s.data_ptr(), o_p.data_ptr())
    torch.cuda.synchronize()
    assert (ff.cpu().numpy().view(np.uint32) == want["ff"].view(np.uint32)).all()  # each pair written once
    assert (o_p.cpu().numpy() == want["topk_pos"]).all()
    assert (o_s.cpu().numpy().view(np.uint32) == want["topk_score"].view(np.uint32)).all()
    # a candidate outside the GLOBAL range is still reported
    bad = cand.copy()
    bad[0] = pool
    d_bad = dev(bad)
    shards[0].rerank_device(m, d_qv.data_ptr(), nq, d_off.data_ptr(), d_bad.data_ptr(), 0, alpha, 0,
                            int(cnts.max()), ff.data_ptr())
    with pytest.raises(_ffx.FFXError):
        shards[0].sync()
    for sh in shards:
        sh.close()
    whole.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    from fast_forward.sharded import ShardedReranker, plan_doc_shards

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)  # same data on every rank
    n_docs, nq, C, k, alpha = 200, 7, 50, 8, 0.25  # 7 queries over 2 ranks: unequal owner slices
    cnt = rng.integers(1, 5, n_docs)
    off = np.concatenate([[0], np.cumsum(cnt)])
    vec = rng.standard_normal((off[-1], 16)).astype(np.float32)
    qv = rng.standard_normal((nq, 16)).astype(np.float32)
    cand = np.concatenate([rng.choice(n_docs, C, replace=False) for _ in range(nq)]).astype(np.int32)
    lex = rng.uniform(0, 5, nq * C).astype(np.float32)
    q_off = np.arange(nq + 1) * C
    pair_q = np.repeat(np.arange(nq), C)
    bounds = plan_doc_shards(cnt, world)
    lo, hi = bounds[rank], bounds[rank + 1]

    class OracleShard(ShardedReranker):  # the two libffx steps, restated with the oracle
        def _local_topk(self, mode, qvecs, q_off_t, cand_t, lex_t, alpha, k, max_cand, stream):
            c = cand_t.numpy()
            mine = (c >= lo) & (c < hi)
            ff = np.full(len(c), np.nan, np.float32)
            ff[mine] = fo.score_pairs(vec, off, np.arange(off[-1]), pair_q[mine], c[mine], qv, mode)
            it = fo.interpolate_f32(lex_t.numpy(), ff, alpha)
            it[~mine] = -np.inf
            s, p = fo.topk_per_query(q_off_t.numpy(), it, k)
            p[np.isneginf(s)] = -1
            return torch.from_numpy(s), torch.from_numpy(p)

        def _merge(self, all_score, all_pos, k, stream):
            S, nq_, _ = all_score.shape
            s = all_score.numpy().transpose(1, 0, 2).reshape(nq_, -1)
            p = all_pos.numpy().transpose(1, 0, 2).reshape(nq_, -1)
            out_s = np.full((nq_, k), -np.inf, np.float32)
            out_p = np.full((nq_, k), -1, np.int32)
            for q in range(nq_):
                keep = p[q] >= 0
                order = np.lexsort((p[q][keep], -s[q][keep].astype(np.float64)))[:k]
                out_s[q, :len(order)] = s[q][keep][order]
                out_p[q, :len(order)] = p[q][keep][order]
            return torch.from_numpy(out_s), torch.from_numpy(out_p)

    sh = OracleShard(None, lo, n_docs)
    t = torch.from_numpy
    s, p = sh.rerank(fo.MODE_MAXP, t(qv), t(q_off), t(cand), t(lex), alpha, k, C)
    ff = fo.score_pairs(vec, off, np.arange(off[-1]), pair_q, cand, qv, fo.MODE_MAXP)
    ws, wp = fo.topk_per_query(q_off, fo.interpolate_f32(lex, ff, alpha), k)
    ok = (p.numpy() == wp).all() and (s.numpy().view(np.uint32) == ws.view(np.uint32)).all()
    # owner-partitioned result: this rank's query slice only, no final all-gather
    s2, p2 = sh.rerank(fo.MODE_MAXP, t(qv), t(q_off), t(cand), t(lex), alpha, k, C, gather_result=False)
    b = sh.owner_bounds(nq)
    ok = ok and b[0] == 0 and b[-1] == nq and (p2.numpy() == wp[b[rank]:b[rank + 1]]).all() and \
        (s2.numpy().view(np.uint32) == ws[b[rank]:b[rank + 1]].view(np.uint32)).all()
    open(os.path.join(out_dir, f"rank{rank}.{'ok' if ok else 'bad'}"), "w").close()
    dist.destroy_process_group()


def test_exchange_plumbing_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]


@pytest.mark.gpu
def test_fused_peer_memory_exchange_on_two_gpus():
    """`ShardedReranker(p2p=True)`: the scoring kernel's epilogue stores the top-k lists into the
    owner rank's buffer over peer memory.  Needs two GPUs (skipped on a one-GPU box): compared
    with the NCCL all-to-all path and with an unsharded index by tools/p2p_check.py."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                          os.path.join(root, "tools", "p2p_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "identical=True" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
