#!/usr/bin/env python
"""bench.py — re-ranked (query, doc) pairs/s of the Fast-Forward hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ffx|reference]

One "step" = one pass of the hot path (gather passage vectors by id -> q.p dots -> per-doc
MAXP -> interpolate with the lexical score -> per-query top-k) over one batch of synthetic
input: BASELINE.json configs[2], "MS MARCO-doc scale Mode.MAXP: 3.2M docs / ~20M passages
ragged, 5193 queries x 5000 candidates" (61 GB of fp32 vectors resident in HBM; 499 GB of
algorithmic traffic per step, so nothing is L2-resident between steps).

  value     whole-job pairs/s with the integer-coded inputs already in HBM (CUDA events on the
            launching stream, max over ranks)
  e2e       the same metric through the host-buffer C-ABI call (ffx_rerank_host): pinned host
            inputs -> H2D -> kernel -> D2H of the ranked lists, all inside the timed region
  roofline  algorithmic bytes per launch / mean launch time of the fused kernel vs the measured
            HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the numpy restatement of the reference's algorithm (oracle/, "port") on a bounded
            sample of the same workload, one process per host core

`--impl reference` times that CPU port alone (rank 0 only) and prints the same line shape.
Multi-GPU (torchrun): queries shard across ranks, the index is replicated, no data-path
collective; scaling is weak (every rank re-ranks its own 5193 x 5000 pairs).
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

# stdout carries exactly one JSON line: everything libraries print there (NCCL's version
# banner, ...) is sent to stderr at the file-descriptor level; emit() writes to the real stdout
_REAL_STDOUT = None


def _claim_stdout() -> None:
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    if _REAL_STDOUT is None:
        print(json.dumps(line), flush=True)
    else:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "fast-forward-indexes_b200")
REF_DIR = os.path.join(ROOT, "baseline", "_ref")  # the unmodified reference, pip-installed offline

METRIC = "re-ranked (query,doc) pairs/sec, MAXP k=5000"
DIM = 768
MODE_MAXP = 2

MODES = {"PASSAGE": 1, "MAXP": 2, "FIRSTP": 3, "AVEP": 4}

# BASELINE.json configs; the default (and the driver's) workload is configs[2] = C3, the one
# the metric is quoted on.  The others are selectable for profiling / DESIGN.md numbers.
WORKLOADS = {
    "c1_passage_10k": dict(mode="PASSAGE", kind="f32", n_rows=10_000, nq=100, cands=1000, k=1000),
    "c2_msmarco_passage": dict(mode="PASSAGE", kind="f32", n_rows=8_800_000, nq=6980, cands=1000, k=1000),
    "c3_msmarco_doc_maxp": dict(mode="MAXP", kind="f32", n_docs=3_200_000, mean_psg=6.25, nq=5193,
                                cands=5000, k=5000),
    "c4_opq_avep": dict(mode="AVEP", kind="opq", n_docs=3_200_000, mean_psg=6.25, nq=5193, cands=5000,
                        k=5000, M=96, Ks=256),
    # SURVEY 8d's skewed variant (reported, not the headline): Zipf-like document popularity inside
    # every candidate stratum, so that popular documents recur across queries and hit in L2
    "c3_zipf_doc_maxp": dict(mode="MAXP", kind="f32", n_docs=3_200_000, mean_psg=6.25, nq=5193,
                             cands=5000, k=5000, popularity="zipf"),
    # per-GPU shard of 1.2M docs / 7.5M passages (60M passages at 8 GPUs), 12 500 queries per GPU
    "c5_sharded_maxp": dict(mode="MAXP", kind="f32", n_docs=1_200_000, mean_psg=6.25, nq=12_500,
                            cands=5000, k=1000, sharded=True),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ffx", choices=["ffx", "reference"])
    ap.add_argument("--workload", default="c3_msmarco_doc_maxp", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=float(os.environ.get("FFX_BENCH_SCALE", "1")),
                    help="shrink docs and queries (debug only; the line says so)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--alpha", type=float, default=0.1)
    ap.add_argument("--cpu-budget", default="short", choices=["short", "long"],
                    help="--impl reference: queries per core and step (short: a few seconds per step)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="sharded workloads: fused peer-memory exchange in the kernel epilogue, or an NCCL all-to-all")
    ap.add_argument("--emulate-shards", type=int, default=0,
                    help="single GPU: hold shard 0 of this many doc-id-range shards and time the local step only "
                         "(profiling the sharded kernel without the other GPUs; the line says so)")
    return ap.parse_args()


def doc_lengths(n_docs, mean_psg, seed=0):
    """Clipped geometric passages/doc, mean ~6.25, min 1, max 64."""
    rng = np.random.default_rng(seed)
    return np.clip(rng.geometric(1.0 / mean_psg, n_docs), 1, 64).astype(np.int64)


# ------------------------------------------------------------------------------------------
# CPU baseline: the numpy port of the reference algorithm, one process per core
# ------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker(args):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ff_oracle as fo

    q_lo, q_hi, cands, k, alpha = args
    vec, off, qv, cand, lex = _CPU["vec"], _CPU["off"], _CPU["qv"], _CPU["cand"], _CPU["lex"]
    rows = _CPU["rows"]
    sl = slice(q_lo * cands, q_hi * cands)
    pair_q = np.repeat(np.arange(q_lo, q_hi), cands)
    # index/base.py:445-459 batches queries; 2 queries/batch keeps temporaries ~200 MB
    ff = fo.score_pairs(vec, off, rows, pair_q, cand[sl], qv, fo.MODE_MAXP, chunk_pairs=2 * cands)
    it = fo.interpolate_f32(lex[sl], ff, alpha)
    q_off = np.arange(q_hi - q_lo + 1) * cands
    s, p = fo.topk_per_query(q_off, it, k)
    return float(s[0, 0])


def cpu_port_run(wl, alpha, q_per_core=4, cores=None):
    """Times the numpy port on `cores` processes x `q_per_core` queries of the workload's
    shape over a scaled-down index (cost per pair does not depend on the index size)."""
    import multiprocessing as mp

    cores = cores or os.cpu_count() or 1
    cores = min(cores, 64)
    n_docs = 40_000
    cnt = doc_lengths(n_docs, wl["mean_psg"], seed=1)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    rng = np.random.default_rng(2)
    _CPU["vec"] = rng.standard_normal((int(off[-1]), DIM), dtype=np.float32)
    _CPU["off"] = off
    _CPU["rows"] = np.arange(off[-1], dtype=np.int64)
    nq = cores * q_per_core
    cands = wl["cands"]
    _CPU["qv"] = rng.standard_normal((nq, DIM), dtype=np.float32)
    _CPU["cand"] = np.concatenate([rng.choice(n_docs, cands, replace=False) for _ in range(nq)])
    _CPU["lex"] = rng.uniform(0, 20, nq * cands).astype(np.float32)
    jobs = [(c * q_per_core, (c + 1) * q_per_core, cands, wl["k"], alpha) for c in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(0, 1, cands, wl["k"], alpha)] * cores)  # warm the workers
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    pairs = nq * cands
    sample = (f"{nq} queries x {cands} candidates MAXP over a {n_docs}-doc / {int(off[-1])}-passage "
              f"768-d fp32 host index, numpy port of index/base.py:279-314 + ranking.py:319,279-291, "
              f"{cores} processes x {q_per_core} queries")
    _CPU.clear()
    return pairs / dt, dt, cores, sample, pairs


# ------------------------------------------------------------------------------------------
def clocks_start():
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    try:
        p = subprocess.Popen(
            ["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,"
             "clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
            stdout=f, stderr=subprocess.DEVNULL)
    except OSError:
        return None, f.name
    return p, f.name


def clocks_stop(p, path, device_index):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    if p is not None:
        p.terminate()
        try:
            p.wait(timeout=5)
        except Exception:
            p.kill()
    try:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9 or not c[0].isdigit() or int(c[0]) != device_index:
                continue
            sm.append(float(c[1]))
            mx.append(float(c[2]))
            for nme, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
    except OSError:
        pass
    finally:
        try:
            os.unlink(path)
        except OSError:
            pass
    return out


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# CPU baseline, preferred form: the UNMODIFIED reference package from baseline/_ref
# ------------------------------------------------------------------------------------------
def reference_installed() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "fast_forward"))


def _import_reference():
    """`import fast_forward` from baseline/_ref.  Its two dependencies that are not in the image
    (h5py: only OnDiskIndex needs it; nanopq: only the quantizers) are stubbed with empty modules —
    the path timed here (InMemoryIndex fp32 -> Index.__call__ -> Ranking.interpolate -> cut)
    touches neither."""
    import types

    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    for name in ("h5py", "nanopq"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                stub = types.ModuleType(name)
                stub.File = None  # index/disk.py:138 names h5py.File in an annotation at import time
                sys.modules[name] = stub
    import fast_forward

    assert os.path.abspath(fast_forward.__file__).startswith(os.path.abspath(REF_DIR)), fast_forward.__file__
    return fast_forward


def _ref_worker(args):
    """One process: the reference's own three calls over its share of the queries."""
    import pandas as pd

    q_lo, q_hi, cands, k, alpha, batch = args
    ff = _import_reference()
    from fast_forward.encoder import LambdaEncoder
    from fast_forward.index import InMemoryIndex, Mode

    qv, cand, lex = _CPU["qv"], _CPU["cand"], _CPU["lex"]
    if "index" not in _CPU:  # built once per worker process (outside the timed region: see the warm-up)
        table = {f"text {i}": qv[i] for i in range(len(qv))}
        index = InMemoryIndex(LambdaEncoder(lambda q: table[q]), mode=Mode.MAXP, init_size=len(_CPU["vec"]))
        index.add(_CPU["vec"], doc_ids=_CPU["doc_ids"])
        _CPU["index"] = index
    if q_hi <= q_lo:
        return 0.0
    index = _CPU["index"]
    rows = slice(q_lo * cands, q_hi * cands)
    frame = pd.DataFrame({"q_id": np.repeat([f"q{i}" for i in range(q_lo, q_hi)], cands),
                          "id": [f"D{d}" for d in cand[rows]], "score": lex[rows]})
    first = ff.Ranking(frame, queries={f"q{i}": f"text {i}" for i in range(q_lo, q_hi)})
    t0 = time.perf_counter()
    out = first.interpolate(index(first, batch_size=batch), alpha).cut(k)
    dt = time.perf_counter() - t0
    assert len(out._df) == (q_hi - q_lo) * min(k, cands)
    return dt


def cpu_reference_run(wl, alpha, q_per_core=5, cores=None):
    """Times the unmodified reference (`Index.__call__` + `Ranking.interpolate` + `Ranking.cut`,
    index/base.py:389-469, ranking.py:293-326,279-291) on `cores` processes x `q_per_core`
    queries of the workload's shape over a scaled-down in-memory index.  `batch_size=2` keeps
    its three [pairs*passages, 768] temporaries near 300 MB per process (index/base.py:445-459)."""
    import multiprocessing as mp

    cores = min(cores or os.cpu_count() or 1, 64)
    n_docs = 10_000
    cnt = doc_lengths(n_docs, wl["mean_psg"], seed=1)
    rng = np.random.default_rng(2)
    n_rows = int(cnt.sum())
    _CPU["vec"] = rng.standard_normal((n_rows, DIM), dtype=np.float32)
    _CPU["doc_ids"] = np.repeat([f"D{d}" for d in range(n_docs)], cnt).tolist()
    nq = cores * q_per_core
    cands = wl["cands"]
    _CPU["qv"] = rng.standard_normal((nq, DIM), dtype=np.float32)
    _CPU["cand"] = np.concatenate([rng.choice(n_docs, cands, replace=False) for _ in range(nq)])
    _CPU["lex"] = rng.uniform(0, 20, nq * cands).astype(np.float32)
    if q_per_core % 2 == 0:
        q_per_core_batch = 3  # the reference crashes on an empty trailing batch (batch_size | #queries)
    else:
        q_per_core_batch = 2
    jobs = [(c * q_per_core, (c + 1) * q_per_core, cands, wl["k"], alpha, q_per_core_batch) for c in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_ref_worker, [(0, 0, cands, wl["k"], alpha, 2)] * cores, chunksize=1)  # import + build the index
        # every process times its own three calls (the first-stage Ranking is built before the
        # clock starts); the processes run side by side, the slowest one is the step
        dt = max(pool.map(_ref_worker, jobs, chunksize=1))
    pairs = nq * cands
    sample = (f"{nq} queries x {cands} candidates MAXP over a {n_docs}-doc / {n_rows}-passage 768-d fp32 "
              f"InMemoryIndex, the UNMODIFIED reference package (baseline/_ref): Index.__call__(batch_size="
              f"{q_per_core_batch}) + Ranking.interpolate + Ranking.cut, {cores} processes x {q_per_core} queries")
    _CPU.clear()
    return pairs / dt, dt, cores, sample, pairs


def cpu_baseline_run(wl, alpha, budget="long"):
    """(value, seconds, cores, sample, kind): the reference itself when baseline/_ref travelled
    with the repo, else the numpy port of oracle/."""
    force = int(os.environ.get("FFX_CPU_Q_PER_CORE", "0"))  # tests shrink the sample
    if reference_installed() and os.environ.get("FFX_CPU_BASELINE", "") != "port":
        v, dt, cores, sample, _ = cpu_reference_run(wl, alpha, q_per_core=force or (9 if budget == "long" else 5))
        return v, dt, cores, sample, "reference"
    v, dt, cores, sample, _ = cpu_port_run(wl, alpha, q_per_core=force or (64 if budget == "long" else 24))
    return v, dt, cores, sample, "port"


# ------------------------------------------------------------------------------------------
def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, times = [], []
    for step in range(args.warmup + args.steps):
        v, dt, cores, sample, kind = cpu_baseline_run(wl, args.alpha, budget=args.cpu_budget)
        if step >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = float(np.mean(vals)) if vals else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)) if times else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "mode": "MAXP", "dim": DIM, "candidates_per_query": wl["cands"],
                   "cut_k": wl["k"], "alpha": args.alpha,
                   "note": "each step = a bounded sample of the workload on the host cores"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline_subprocess(args):
    """The CPU baseline runs in its own interpreter: the reference package has the same import
    name (`fast_forward`) as the drop-in this process has loaded."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--alpha", str(args.alpha), "--cpu-budget", "long"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    for ln in out.stdout.splitlines():
        if ln.startswith("{"):
            base = json.loads(ln)["cpu_baseline"]
            base["seconds"] = round(json.loads(ln)["ms_per_step"] / 1e3, 2)
            return base
    raise RuntimeError("cpu baseline failed: " + out.stderr[-2000:])


def build_corpus(wl, scale, rank, world):
    """Row counts per document of the (rank's part of the) synthetic corpus."""
    if "n_docs" in wl:
        n_docs = max(wl["cands"] * 2, int(wl["n_docs"] * scale))
        if wl.get("sharded"):
            n_docs *= world
        cnt = doc_lengths(n_docs, wl["mean_psg"], seed=0)
        return n_docs, cnt
    return None, None


def run_ffx(args, wl):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, PKG)
    from fast_forward import _ffx
    from fast_forward.sharded import ShardedReranker, plan_doc_shards

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    emulate = args.emulate_shards if world == 1 and wl.get("sharded") else 0
    if emulate:
        world = emulate  # sizes and shard plan of an `emulate`-GPU job; only rank 0's local work runs
    if _ffx.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: libffx has no CPU path")
    if os.environ.get("FFX_CHUNK_WAVES"):
        _ffx.set_option("chunk_waves", int(os.environ["FFX_CHUNK_WAVES"]))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not emulate:
        dist.init_process_group("nccl", device_id=dev)

    mode, sharded, pq = MODES[wl["mode"]], bool(wl.get("sharded")), wl["kind"] == "opq"
    cands, k = wl["cands"], wl["k"]
    nq = wl["nq"] if args.scale == 1 else max(296, int(wl["nq"] * args.scale))
    if sharded:
        nq *= world  # weak scaling: every rank scores ~nq*cands/world pairs of a world-times larger job
    n_docs, cnt = build_corpus(wl, args.scale, rank, world)
    if n_docs is not None:
        off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        total_rows = int(off[-1])
        pool = n_docs
    else:
        total_rows = max(cands * 2, int(wl["n_rows"] * args.scale))
        off = None
        pool = total_rows
    doc_lo, doc_hi, row_lo, row_hi = 0, n_docs, 0, total_rows
    if sharded:
        bounds = plan_doc_shards(cnt, world)
        doc_lo, doc_hi = int(bounds[rank]), int(bounds[rank + 1])
        row_lo, row_hi = int(off[doc_lo]), int(off[doc_hi])
    n_rows = row_hi - row_lo
    width = wl["M"] if pq else DIM

    # ---- index: synthetic rows generated on the device and staged into the store
    t_stage = time.perf_counter()
    idx = _ffx.DeviceIndex(width, capacity=n_rows, row_kind=_ffx.ROWS_PQ_U8 if pq else _ffx.ROWS_F32,
                           device=local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + (rank if sharded else 0))  # replicas are identical on every rank
    chunk = 1 << 20
    for r0 in range(0, n_rows, chunk):
        nr = min(chunk, n_rows - r0)
        if pq:
            t = torch.randint(0, wl["Ks"], (nr, width), device=dev, dtype=torch.uint8, generator=gen)
        else:
            t = torch.randn((nr, DIM), device=dev, dtype=torch.float32, generator=gen)
        torch.cuda.synchronize()
        idx.stage_device(r0, nr, t.data_ptr())
        del t
    if off is not None:
        idx.set_docs(off[doc_lo:doc_hi + 1] - row_lo)
    if pq:
        g_cpu = torch.Generator().manual_seed(7)
        cw = torch.randn((wl["M"], wl["Ks"], DIM // wl["M"]), generator=g_cpu)
        R = torch.linalg.qr(torch.randn((DIM, DIM), generator=g_cpu))[0]
        idx.set_pq(cw.numpy(), R.numpy())
    torch.cuda.empty_cache()
    t_stage = time.perf_counter() - t_stage

    # ---- queries: stratified distinct candidates, uniform over the corpus.  Query-DP ranks
    # draw their own queries; shards all see the same ones.
    gen.manual_seed(99 + (0 if sharded else rank))
    qv = torch.randn((nq, DIM), device=dev, dtype=torch.float32, generator=gen)
    bucket = pool // cands
    if bucket < 1:
        raise RuntimeError("corpus smaller than the candidate list")
    cand_parts = []
    for q0 in range(0, nq, 4096):  # bounded temporaries
        qn = min(4096, nq - q0)
        perm = torch.rand((qn, cands), device=dev, generator=gen).argsort(dim=1)
        if wl.get("popularity") == "zipf":  # P(rank r inside the stratum) ~ 1/r: log-uniform draw
            u = torch.rand((qn, cands), device=dev, generator=gen)
            within = (torch.exp(u * float(np.log(bucket))) - 1.0).long().clamp_(0, bucket - 1)
        else:
            within = torch.randint(0, bucket, (qn, cands), device=dev, generator=gen)
        cand_parts.append((perm * bucket + within).to(torch.int32))
        del perm, within
    cand = torch.cat(cand_parts).contiguous().view(-1)
    del cand_parts
    lex = (torch.rand((nq * cands,), device=dev, generator=gen) * 20).contiguous()
    q_off = (torch.arange(nq + 1, device=dev, dtype=torch.int64) * cands).contiguous()
    topk_s = torch.empty((nq, k), device=dev, dtype=torch.float32)
    topk_p = torch.empty((nq, k), device=dev, dtype=torch.int32)
    n_pairs = nq * cands

    # algorithmic bytes: row bytes of every pair this rank scores + 16 B/pair
    # (candidate, lexical score, span) + per query (query vector + top-k out)
    row_bytes = width * (1 if pq else 4)
    if cnt is not None and mode != MODES["FIRSTP"]:
        d_cnt = torch.from_numpy(cnt).to(dev)
        c64 = cand.long()
        mine = (c64 >= doc_lo) & (c64 < doc_hi)
        rows_touched = int(d_cnt[c64[mine]].sum().item())
        my_pairs = int(mine.sum().item())
        del d_cnt, c64, mine
    else:
        rows_touched = my_pairs = n_pairs
    algo_bytes = rows_touched * row_bytes + n_pairs * 8 + my_pairs * 8 + nq * (DIM * 4 + k * 8)

    stream = torch.cuda.current_stream()
    p2p = sharded and world > 1 and not emulate and args.exchange == "p2p"
    reranker = ShardedReranker(idx, doc_lo, n_docs, row_lo, total_rows, p2p=p2p) if sharded else None

    def step():
        if sharded:
            return reranker.rerank(mode, qv, q_off, cand, lex, args.alpha, k, cands, gather_result=False)
        idx.rerank_device(mode, qv.data_ptr(), nq, q_off.data_ptr(), cand.data_ptr(), lex.data_ptr(),
                          args.alpha, k, cands, 0, 0, topk_s.data_ptr(), topk_p.data_ptr(),
                          stream.cuda_stream)
        return topk_s, topk_p

    def barrier():
        torch.cuda.synchronize()
        if world > 1 and not emulate:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        out_s, out_p = step()
    barrier()
    idx.sync(stream.cuda_stream)
    clk_p, clk_path = clocks_start() if rank == 0 else (None, None)
    launches0 = _ffx.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for i in range(args.steps):
        out_s, out_p = step()
        ev[i + 1].record(stream)
    barrier()
    launches = _ffx.launch_count() - launches0
    per_step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = clocks_stop(clk_p, clk_path, local) if rank == 0 else None

    # sanity on the timed output: ranked lists are sorted, positions distinct and in range
    s_host = out_s[:4].cpu().numpy()
    p_host = out_p[:4].cpu().numpy()
    if not emulate:  # a lone shard's lists are padded with (-inf, -1)
        assert (np.diff(s_host, axis=1) <= 0).all(), "top-k not sorted"
        assert all(len(set(r.tolist())) == k for r in p_host) and p_host.min() >= 0 and p_host.max() < cands

    # ---- e2e: host buffers through ffx_rerank_host (H2D + kernel + D2H inside the timed region)
    e2e_s = h2d = d2h = None
    if not sharded:
        h_q = _ffx.PinnedBuffer((nq, DIM), np.float32)
        h_off = _ffx.PinnedBuffer((nq + 1,), np.int64)
        h_cand = _ffx.PinnedBuffer((n_pairs,), np.int32)
        h_lex = _ffx.PinnedBuffer((n_pairs,), np.float32)
        h_ts = _ffx.PinnedBuffer((nq, k), np.float32)
        h_tp = _ffx.PinnedBuffer((nq, k), np.int32)
        h_q.array[:] = qv.cpu().numpy()
        h_off.array[:] = q_off.cpu().numpy()
        h_cand.array[:] = cand.cpu().numpy()
        h_lex.array[:] = lex.cpu().numpy()
        out = {"topk_score": h_ts.array, "topk_pos": h_tp.array}

        def e2e_step():
            idx.rerank_host(mode, h_q.array, h_off.array, h_cand.array, h_lex.array, args.alpha, k,
                            want_ff=False, want_int=False, out=out)

        e2e_step()
        if not (h_tp.array[:4] == p_host).all():  # say where before failing
            for q in range(4):
                d = np.flatnonzero(h_tp.array[q] != p_host[q])
                if len(d):
                    r = int(d[0])
                    print(f"query {q}: {len(d)} ranks differ, first at {r}: host pos/score "
                          f"{h_tp.array[q, r:r + 4].tolist()} {h_ts.array[q, r:r + 4].tolist()} device "
                          f"{p_host[q, r:r + 4].tolist()} {s_host[q, r:r + 4].tolist()}", file=sys.stderr)
                    for pos in (int(h_tp.array[q, r]), int(p_host[q, r])):
                        rh, rd = np.flatnonzero(h_tp.array[q] == pos), np.flatnonzero(p_host[q] == pos)
                        print(f"   pos {pos}: host rank {rh.tolist()} score {h_ts.array[q, rh].tolist()}; "
                              f"device rank {rd.tolist()} score {s_host[q, rd].tolist()}", file=sys.stderr)
        assert (h_tp.array[:4] == p_host).all(), "host-buffer path disagrees with the device path"
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        h2d = h_q.array.nbytes + h_off.array.nbytes + h_cand.array.nbytes + h_lex.array.nbytes
        d2h = h_ts.array.nbytes + h_tp.array.nbytes

    times = torch.tensor([total_ms, (e2e_s or 0.0) * 1e3], device=dev, dtype=torch.float64)
    if world > 1 and not emulate:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = times.tolist()

    if rank == 0:
        job_pairs = n_pairs if sharded else world * n_pairs  # shards split ONE job's pairs
        value = job_pairs * args.steps / (total_ms * 1e-3)
        kern_ms = float(np.mean(per_step_ms))
        achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
        peak, peak_src = measured_peak()
        traffic = ncu_traffic()
        kernel = ("ffx_adc_xor_kernel<3,true> (TMA-staged codes, XOR-swizzled conflict-free LUT look-ups, "
                  "AVEP-interpolate-topk fused; bound by shared-memory look-ups, not HBM)" if pq else
                  "ffx_score_tma_kernel<2,12,%s> (TMA-staged gather-dot-%s-interpolate%s)" % (
                      "true" if nq >= 296 else "false", wl["mode"], "-topk fused" if nq >= 296 else " ; ffx_topk_kernel"))
        line = {
            "metric": METRIC if args.workload == "c3_msmarco_doc_maxp" else
            f"re-ranked (query,doc) pairs/sec, {wl['mode']} k={cands}",
            "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 codes + f32 LUT" if pq else "f32",
            "data": "synthetic",
            "config": {
                "workload": (args.workload if args.scale == 1 else f"{args.workload} (SCALED x{args.scale}: debug)") +
                (f" (EMULATED: local step of shard 0 of {emulate} on one GPU, no exchange; profiling only)"
                 if emulate else ""),
                "mode": wl["mode"], "dim": DIM, "docs": n_docs, "passages": total_rows,
                "index_gb_per_gpu": n_rows * row_bytes / 1e9,
                "queries": nq if sharded else f"{nq} per GPU", "candidates_per_query": cands, "cut_k": k,
                "alpha": args.alpha,
                "parallelism": ((f"doc-id-range shards x{world}: fused score+top-k kernel whose epilogue stores each "
                                 f"query's [k] list into its owner rank's buffer over NVLink peer memory, device "
                                 f"barrier, ffx_merge_topk at the owner" if p2p else
                                 f"doc-id-range shards x{world}: local fused top-k, one NCCL all-to-all of the "
                                 f"[nq,k] lists to the queries' owner ranks, ffx_merge_topk there") if sharded else
                                f"query-dp{world} (replicated index, no data-path collective)"),
                "l2": "inputs larger than L2 (index %.1f GB/GPU, %.1f GB touched per step per GPU)" % (
                    n_rows * row_bytes / 1e9, algo_bytes / 1e9),
                "index_stage_s": round(t_stage, 2),
            },
            "e2e": None if e2e_s is None else {
                "value": world * n_pairs * args.steps / (e2e_ms * 1e-3), "unit": "pairs/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": traffic.get(args.workload) if traffic and args.scale == 1 else None,
                         "kernel": kernel, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kern_ms,
                         "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0},
            "clocks": clocks,
        }
        if pq:
            line["roofline"]["lut_lookups_per_s"] = rows_touched * wl["M"] / (kern_ms * 1e-3)
        if not args.no_cpu_baseline and world == 1 and args.workload == "c3_msmarco_doc_maxp":
            line["cpu_baseline"] = cpu_baseline_subprocess(args)
        emit(line)
    if world > 1 and not emulate:
        dist.barrier()
        dist.destroy_process_group()
    idx.close()


def main():
    _claim_stdout()
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, WORKLOADS["c3_msmarco_doc_maxp"])
    else:
        run_ffx(args, wl)


if __name__ == "__main__":
    main()
