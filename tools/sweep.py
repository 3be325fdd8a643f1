#!/usr/bin/env python
"""Tuning sweep over the scoring-kernel knobs (ffx_set_option) on one resident synthetic index.

    python tools/sweep.py [--workload c3_msmarco_doc_maxp] [--scale 1] [--steps 5] CONFIG...

CONFIG = kernel:warps:stages:batch, e.g. 1:0:0:0 (register kernel) 2:8:6:16 2:16:3:16.
Prints one line per config: ms/step, pairs/s, GB/s of algorithmic bytes."""

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200"))

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3_msmarco_doc_maxp")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--nq-list", default="", help="comma-separated query counts: time every config at each (wave-tail study)")
    ap.add_argument("configs", nargs="+")
    args = ap.parse_args()
    import torch

    from fast_forward import _ffx

    wl = dict(bench.WORKLOADS[args.workload])
    if args.k:
        wl["k"] = args.k
    dev = torch.device("cuda", 0)
    mode = bench.MODES[wl["mode"]]
    cands, k = wl["cands"], wl["k"]
    nq = wl["nq"] if args.scale == 1 else max(296, int(wl["nq"] * args.scale))
    nq_list = [int(x) for x in args.nq_list.split(",") if x]
    nq = max([nq] + nq_list)
    n_docs, cnt = bench.build_corpus(wl, args.scale, 0, 1)
    if n_docs is not None:
        off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        n_rows, pool = int(off[-1]), n_docs
    else:
        n_rows = pool = max(cands * 2, int(wl["n_rows"] * args.scale))
        off = None
    idx = _ffx.DeviceIndex(bench.DIM, capacity=n_rows)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    for r0 in range(0, n_rows, 1 << 20):
        nr = min(1 << 20, n_rows - r0)
        t = torch.randn((nr, bench.DIM), device=dev, generator=gen)
        torch.cuda.synchronize()
        idx.stage_device(r0, nr, t.data_ptr())
        del t
    if off is not None:
        idx.set_docs(off)
    gen.manual_seed(99)
    qv = torch.randn((nq, bench.DIM), device=dev, generator=gen)
    bucket = pool // cands
    perm = torch.rand((nq, cands), device=dev, generator=gen).argsort(dim=1)
    within = torch.randint(0, bucket, (nq, cands), device=dev, generator=gen)
    cand = (perm * bucket + within).to(torch.int32).contiguous().view(-1)
    del perm, within
    lex = (torch.rand((nq * cands,), device=dev, generator=gen) * 20).contiguous()
    q_off = (torch.arange(nq + 1, device=dev, dtype=torch.int64) * cands).contiguous()
    ts = torch.empty((nq, k), device=dev, dtype=torch.float32)
    tp = torch.empty((nq, k), device=dev, dtype=torch.int32)
    if cnt is not None:
        rows_touched = int(torch.from_numpy(cnt).to(dev)[cand.long()].sum().item())
    else:
        rows_touched = nq * cands
    algo = rows_touched * bench.DIM * 4 + nq * cands * 16 + nq * (bench.DIM * 4 + k * 8)
    stream = torch.cuda.current_stream()
    ref_pos = None
    for cfg, n_run in [(c, n) for c in args.configs for n in (nq_list or [nq])]:
        kern, warps, stages, batch = (int(x) for x in cfg.split(":"))
        _ffx.set_option("kernel", kern)
        _ffx.set_option("tma_warps", warps)
        _ffx.set_option("tma_stages", stages)
        _ffx.set_option("batch", batch)

        def step():
            idx.rerank_device(mode, qv.data_ptr(), n_run, q_off.data_ptr(), cand.data_ptr(), lex.data_ptr(), 0.1, k,
                              cands, 0, 0, ts.data_ptr(), tp.data_ptr(), stream.cuda_stream)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        idx.sync(stream.cuda_stream)
        ms = e0.elapsed_time(e1) / args.steps
        pos = tp[:64].cpu().numpy()
        same = "" if ref_pos is None else (" same-output" if (pos == ref_pos).all() else " OUTPUT-DIFFERS")
        ref_pos = pos if ref_pos is None else ref_pos
        print(f"{cfg:>12s}  nq {n_run:6d}  {ms:9.3f} ms  {n_run * cands / ms / 1e3:9.1f} Mpairs/s  "
              f"{algo * n_run / nq / ms / 1e6:8.1f} GB/s{same}", flush=True)


if __name__ == "__main__":
    main()
