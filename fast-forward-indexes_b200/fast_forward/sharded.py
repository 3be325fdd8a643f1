"""Doc-id-range sharded re-ranking across the GPUs of one box (BASELINE config 5).

The reference has no counterpart: its index must fit one process (index/memory.py,
index/disk.py).  Here a corpus larger than one GPU's HBM is cut into contiguous document
ranges, one per rank (one process per GPU, `torch.distributed`, NCCL over NVLink).  Every rank
receives the FULL integer-coded candidate lists (global document ordinals) and

  1. scores the pairs whose document it owns — the kernel skips foreign pairs, so there is no
     host-side bucketing and positions stay positions in the full candidate block,
  2. interpolates and takes a LOCAL top-k per query (same fused kernel),
  3. exchanges the `[nq, k]` (score, position) lists with ONE all-to-all: every query has an
     owner rank (contiguous query ranges), which receives that query's list from every shard
     and merges them with `ffx_merge_topk`.  The merged lists stay with their owners
     (`gather_result=False`, what a serving system wants) or are all-gathered so that every
     rank holds the full result.

With `p2p=True` step 3 disappears into step 2: the ranks map each other's receive buffers
(symmetric memory over NVLink / NVSwitch), and the epilogue of the fused kernel stores every
query's list directly into the owner's buffer (`ffx_index_set_topk_scatter`) — compute and
exchange are one kernel; what is left is a device-side barrier and the owner's merge.

Exact: interpolation is per pair and the top-k of a union is the top-k of the per-shard
top-ks.  Each rank sends and receives `nq * k * 8` bytes (C5, k=1000, 8 GPUs: 0.8 GB per rank
against 1.2 TB of HBM traffic per rank) and merges `nq / world` queries — an all-gather would
move and merge `world` times as much.  When the semantic score of every pair
is wanted instead, each rank writes only its own pairs into a zeroed buffer and one
all-reduce(SUM) assembles the vector (`x + 0 == x` exactly).

torch is plumbing here (device buffers, the process group); all compute is libffx.
"""

from __future__ import annotations

import numpy as np

from fast_forward import _ffx


def plan_doc_shards(rows_per_doc: np.ndarray, world: int) -> np.ndarray:
    """Contiguous document ranges with (nearly) equal row counts.

    :return: int64 `[world + 1]` document boundaries; rank r owns docs [b[r], b[r+1])."""
    rows_per_doc = np.asarray(rows_per_doc, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(rows_per_doc)])
    targets = cum[-1] * np.arange(1, world) / world
    inner = np.searchsorted(cum, targets, side="left")
    bounds = np.concatenate([[0], inner, [len(rows_per_doc)]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


class ShardedReranker:
    """One rank's shard (`index`, a `_ffx.DeviceIndex` holding documents
    [doc_base, doc_base + n_local)) plus the exchange with the other ranks of `group`."""

    def __init__(self, index, doc_base: int, global_docs: int, row_base: int = 0, global_rows: int = 0,
                 group=None, p2p: bool = False) -> None:
        import torch.distributed as dist

        self.index = index
        self.group = group
        self.p2p = p2p
        self._p2p_sets = None  # (nq, k) -> two alternating sets of symmetric receive buffers
        self._p2p_step = 0
        # bench.py sets this to a list: every rerank() then appends the CUDA events
        # (start, local kernel done, exchange / barrier done, merge done) of its three phases
        self.trace = None
        self._io = None  # copy streams of rerank_host
        self._host_out = None  # its pinned result tensors
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if index is not None:
            index.set_shard(doc_base, global_docs, row_base, global_rows or row_base + len(index))

    # -- the two compute steps (libffx); tests on CPU substitute them ------------------------
    def _local_topk(self, mode, qvecs, q_off, cand, lex, alpha, k, max_cand, stream):
        import torch

        nq = qvecs.shape[0]
        score = torch.empty((nq, k), dtype=torch.float32, device=qvecs.device)
        pos = torch.empty((nq, k), dtype=torch.int32, device=qvecs.device)
        self.index.rerank_device(mode, qvecs.data_ptr(), nq, q_off.data_ptr(), cand.data_ptr(),
                                 lex.data_ptr() if lex is not None else 0, alpha, k, max_cand,
                                 0, 0, score.data_ptr(), pos.data_ptr(), stream)
        return score, pos

    def _merge(self, all_score, all_pos, k, stream):
        import torch

        world, nq, _ = all_score.shape
        score = torch.empty((nq, k), dtype=torch.float32, device=all_score.device)
        pos = torch.empty((nq, k), dtype=torch.int32, device=all_score.device)
        _ffx.merge_topk(all_score.device.index or 0, all_score.data_ptr(), all_pos.data_ptr(), world, nq, k,
                        score.data_ptr(), pos.data_ptr(), stream)
        return score, pos

    def _mark(self, marks):
        if self.trace is not None:
            import torch

            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append(ev)

    # -- the exchange ---------------------------------------------------------------------------
    def owner_bounds(self, nq: int) -> list[int]:
        """Query q is merged on the rank r with bounds[r] <= q < bounds[r+1]."""
        return [nq * r // self.world for r in range(self.world + 1)]

    def _to_owners(self, local, bounds):
        """`local` [nq, k] on every rank -> [world, mine, k]: slice [bounds[rank], bounds[rank+1])
        of every rank's `local`.  One all-to-all on NCCL; gloo (CPU tests) has none, so one
        gather per owner."""
        import torch
        import torch.distributed as dist

        mine = bounds[self.rank + 1] - bounds[self.rank]
        recv = torch.empty((self.world, mine) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        if dist.get_backend(self.group) == "nccl":
            dist.all_to_all_single(recv.view((self.world * mine,) + tuple(local.shape[1:])), local.contiguous(),
                                   output_split_sizes=[mine] * self.world,
                                   input_split_sizes=[bounds[r + 1] - bounds[r] for r in range(self.world)],
                                   group=self.group)
        else:
            for r in range(self.world):
                part = local[bounds[r]:bounds[r + 1]].contiguous()
                dist.gather(part, list(recv.unbind(0)) if r == self.rank else None,
                            dst=dist.get_global_rank(self.group, r) if self.group is not None else r,
                            group=self.group)
        return recv

    # -- fused exchange over peer memory --------------------------------------------------------
    def _p2p_buffers(self, nq: int, k: int, device):
        """Two alternating sets of receive buffers [world, cap, k] in symmetric memory (every rank
        maps every other rank's).  Collective; cached per (nq, k)."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        if self._p2p_sets is None:
            self._p2p_sets = {}
        if (nq, k) in self._p2p_sets:
            return self._p2p_sets[(nq, k)]
        bounds = self.owner_bounds(nq)
        cap = max(bounds[r + 1] - bounds[r] for r in range(self.world))
        group = self.group if self.group is not None else dist.group.WORLD
        sets = []
        for _ in range(2):
            score = symm.empty((self.world, cap, k), dtype=torch.float32, device=device)
            pos = symm.empty((self.world, cap, k), dtype=torch.int32, device=device)
            hs, hp = symm.rendezvous(score, group), symm.rendezvous(pos, group)
            sets.append({"score": score, "pos": pos, "hs": hs, "hp": hp, "cap": cap, "bounds": bounds})
        self._p2p_sets[(nq, k)] = sets
        return sets

    def _rerank_p2p(self, mode, qvecs, q_off, cand, lex, alpha, k, max_cand, stream):
        """Local scoring whose epilogue stores each query's list into the owner's buffer, a
        device-side barrier, then the merge of the `world` blocks this rank received."""
        nq = qvecs.shape[0]
        buf = self._p2p_buffers(nq, k, qvecs.device)[self._p2p_step & 1]
        self._p2p_step += 1
        self.index.set_topk_scatter(self.world, self.rank, buf["cap"], buf["bounds"], buf["hs"].buffer_ptrs,
                                    buf["hp"].buffer_ptrs)
        marks = []
        self._mark(marks)
        try:
            self.index.rerank_device(mode, qvecs.data_ptr(), nq, q_off.data_ptr(), cand.data_ptr(),
                                     lex.data_ptr() if lex is not None else 0, alpha, k, max_cand, 0, 0, 0, 0, stream)
        finally:
            self.index.set_topk_scatter(0)
        self._mark(marks)
        buf["hs"].barrier()  # every rank's stores have landed (stream-ordered, on the device)
        self._mark(marks)
        mine = buf["bounds"][self.rank + 1] - buf["bounds"][self.rank]
        score, pos = self._merge(buf["score"], buf["pos"], k, stream)  # [cap, k]
        self._mark(marks)
        if self.trace is not None:
            self.trace.append(marks)
        return score[:mine], pos[:mine]

    # -- public ---------------------------------------------------------------------------------
    def rerank(self, mode: int, qvecs, q_off, cand, lex, alpha: float, k: int, max_cand: int,
               gather_result: bool = True):
        """Global per-query top-k `(score [nq,k], position [nq,k])`.  All inputs are identical on
        all ranks: qvecs f32 [nq,D], q_off i64 [nq+1], cand i32 [n] (GLOBAL ordinals), lex f32 [n]
        or None.  With `gather_result=False` every rank returns only the lists of the queries it
        owns (`owner_bounds(nq)`), skipping the final all-gather."""
        import torch
        import torch.distributed as dist

        stream = torch.cuda.current_stream().cuda_stream if qvecs.is_cuda else 0
        nq = qvecs.shape[0]
        bounds = self.owner_bounds(nq)
        if self.world > 1 and self.p2p and (self._p2p_sets is None or (nq, k) not in self._p2p_sets):
            try:  # collective set-up: it fails on every rank or on none
                self._p2p_buffers(nq, k, qvecs.device)
            except Exception as e:  # no peer access between the GPUs: say so, use NCCL
                import warnings

                warnings.warn(f"symmetric memory unavailable ({e}); exchanging the top-k lists with NCCL all-to-all")
                self.p2p = False
        if self.world > 1 and self.p2p:
            mine_s, mine_p = self._rerank_p2p(mode, qvecs, q_off, cand, lex, alpha, k, max_cand, stream)
        else:
            marks = []
            self._mark(marks)
            score, pos = self._local_topk(mode, qvecs, q_off, cand, lex, alpha, k, max_cand, stream)
            self._mark(marks)
            if self.world == 1:
                return score, pos
            recv_s, recv_p = self._to_owners(score, bounds), self._to_owners(pos, bounds)
            self._mark(marks)
            mine_s, mine_p = self._merge(recv_s, recv_p, k, stream)
            self._mark(marks)
            if self.trace is not None:
                self.trace.append(marks)
        if not gather_result:
            return mine_s, mine_p
        # owners hold unequal slices when world does not divide nq: gather padded slices
        cap = max(bounds[r + 1] - bounds[r] for r in range(self.world))
        out = []
        for part, fill in ((mine_s, float("-inf")), (mine_p, -1)):
            padded = torch.full((cap, k), fill, dtype=part.dtype, device=part.device)
            padded[:part.shape[0]] = part
            everyone = torch.empty((self.world, cap, k), dtype=part.dtype, device=part.device)
            dist.all_gather(list(everyone.unbind(0)), padded, group=self.group)
            out.append(torch.cat([everyone[r, :bounds[r + 1] - bounds[r]] for r in range(self.world)]))
        return out[0], out[1]

    def rerank_host(self, mode: int, qvecs, q_off, cand, lex, alpha: float, k: int, max_cand: int,
                    chunk_queries: int = 0):
        """`rerank(..., gather_result=False)` from HOST inputs (torch CPU tensors, ideally pinned;
        identical on all ranks), pipelined: the job is cut into chunks of `chunk_queries` queries
        (default: 8 chunks); while chunk i is scored and exchanged, chunk i+1 crosses PCIe on a copy
        stream — with NCCL only this rank's 1/world piece of it, the pieces being all-gathered over
        NVLink — and the merged lists of chunk i-1 go back on another.  Returns host tensors
        `(query index [m], score [m, k], position [m, k])`: the queries this rank owns (every
        chunk is split over the ranks by `owner_bounds`) and their merged lists; the two pinned
        list tensors are reused by the next call of the same shape."""
        import torch

        dev = torch.device("cuda", self.index.device)
        nq = qvecs.shape[0]
        step = int(chunk_queries) if chunk_queries else -(-nq // 8)
        step = max(1, min(step, nq))
        chunks = [(lo, min(nq, lo + step)) for lo in range(0, nq, step)]
        if len(chunks) > 1 and chunks[-1][1] - chunks[-1][0] < step // 2:  # a short tail rides with the last chunk
            chunks[-2:] = [(chunks[-2][0], nq)]
        main = torch.cuda.current_stream(dev)
        if self._io is None:
            self._io = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        h2d, d2h = self._io
        q_off_cpu = q_off if not q_off.is_cuda else q_off.cpu()
        starts = q_off_cpu[[lo for lo, _ in chunks] + [nq]].tolist()
        widest_rows = max(starts[i + 1] - starts[i] for i in range(len(chunks)))
        widest_q = max(hi - lo for lo, hi in chunks)
        # Every shard needs every query's full candidate list, but not over PCIe: with NCCL each rank
        # uploads 1/world of a chunk (the inputs are identical on all ranks) and the pieces are
        # all-gathered over NVLink — 4.3 GB per rank and step at C5 become 0.54 GB of PCIe + an
        # NVLink all-gather that hides behind the previous chunk's scoring.
        import torch.distributed as dist

        W = self.world
        gather = W > 1 and dist.is_initialized() and dist.get_backend(self.group) == "nccl"

        def piece(n):  # elements per rank (equal pieces: all_gather_into_tensor)
            return -(-n // W) if gather else n

        cap_rows, cap_q = piece(widest_rows) * (W if gather else 1), piece(widest_q) * (W if gather else 1)
        # two sets of device buffers alternate between chunks
        bufs = [{"qv": torch.empty((cap_q, qvecs.shape[1]), dtype=torch.float32, device=dev),
                 "off": torch.empty(widest_q + 1, dtype=torch.int64, device=dev),
                 "cand": torch.empty(cap_rows, dtype=torch.int32, device=dev),
                 "lex": torch.empty(cap_rows, dtype=torch.float32, device=dev) if lex is not None else None,
                 "free": torch.cuda.Event(), "ready": torch.cuda.Event()} for _ in range(2)]
        if gather:
            for b in bufs:  # landing buffers of this rank's own pieces
                b["qv_in"] = torch.empty((piece(widest_q), qvecs.shape[1]), dtype=torch.float32, device=dev)
                b["cand_in"] = torch.empty(piece(widest_rows), dtype=torch.int32, device=dev)
                b["lex_in"] = torch.empty(piece(widest_rows), dtype=torch.float32, device=dev) if lex is not None else None

        def upload(dst, landing, src, lo_, n):
            """rows [lo_, lo_ + n) of the host tensor `src` into dst[:n] on the copy stream"""
            if not gather:
                dst[:n].copy_(src[lo_:lo_ + n], non_blocking=True)
                return
            per = piece(n)
            mine_lo = min(n, self.rank * per)
            mine_n = min(n, mine_lo + per) - mine_lo
            if mine_n:
                landing[:mine_n].copy_(src[lo_ + mine_lo:lo_ + mine_lo + mine_n], non_blocking=True)
            flat_out, flat_in = dst[:per * W], landing[:per]
            dist.all_gather_into_tensor(flat_out.view(-1), flat_in.reshape(-1), group=self.group, async_op=True).wait()

        # the merged lists of the queries this rank owns land in ONE pair of pinned tensors
        owned = [self.owner_bounds(hi - lo) for lo, hi in chunks]
        counts = [ob[self.rank + 1] - ob[self.rank] for ob in owned]
        total_mine = sum(counts)
        cache = self._host_out
        if cache is None or cache[0].shape != (total_mine, k):
            cache = (torch.empty((total_mine, k), dtype=torch.float32, pin_memory=True),
                     torch.empty((total_mine, k), dtype=torch.int32, pin_memory=True))
            self._host_out = cache
        out_s, out_p = cache
        index_parts, at = [], 0
        for i, (lo, hi) in enumerate(chunks):
            b = bufs[i & 1]
            r0, r1 = starts[i], starts[i + 1]
            with torch.cuda.stream(h2d):
                if i >= 2:
                    h2d.wait_event(b["free"])  # the chunk that used this buffer set has been scored
                upload(b["qv"], b.get("qv_in"), qvecs, lo, hi - lo)
                b["off"][:hi - lo + 1].copy_(q_off_cpu[lo:hi + 1] - r0, non_blocking=True)
                upload(b["cand"], b.get("cand_in"), cand, r0, r1 - r0)
                if lex is not None:
                    upload(b["lex"], b.get("lex_in"), lex, r0, r1 - r0)
                b["ready"].record(h2d)
            main.wait_event(b["ready"])
            s, p = self.rerank(mode, b["qv"][:hi - lo], b["off"][:hi - lo + 1], b["cand"][:r1 - r0],
                               b["lex"][:r1 - r0] if lex is not None else None, alpha, k, max_cand, gather_result=False)
            b["free"].record(main)
            done = torch.cuda.Event()
            done.record(main)
            bounds = owned[i]
            index_parts.append(torch.arange(lo + bounds[self.rank], lo + bounds[self.rank + 1]))
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                out_s[at:at + counts[i]].copy_(s, non_blocking=True)
                out_p[at:at + counts[i]].copy_(p, non_blocking=True)
                s.record_stream(d2h)
                p.record_stream(d2h)
            at += counts[i]
        d2h.synchronize()
        main.synchronize()
        return torch.cat(index_parts), out_s, out_p  # (the two tensors are reused by the next call)

    def scores(self, mode: int, qvecs, q_off, cand, max_cand: int):
        """Semantic score of EVERY pair on every rank (plain `Index.__call__` over a sharded
        corpus): own pairs into a zeroed buffer, then all-reduce(SUM)."""
        import torch
        import torch.distributed as dist

        stream = torch.cuda.current_stream().cuda_stream if qvecs.is_cuda else 0
        ff = torch.zeros(cand.shape[0], dtype=torch.float32, device=qvecs.device)
        self.index.rerank_device(mode, qvecs.data_ptr(), qvecs.shape[0], q_off.data_ptr(), cand.data_ptr(),
                                 0, 0.0, 0, max_cand, ff.data_ptr(), 0, 0, 0, stream)
        if self.world > 1:
            dist.all_reduce(ff, op=dist.ReduceOp.SUM, group=self.group)
        return ff
