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
  api_e2e   (1 GPU) the same job through the drop-in Python API: a first-stage `Ranking` of id
            STRINGS -> `index.rerank(ranking, alpha, k)` -> `Ranking`, wall clock
  roofline  algorithmic bytes per launch / mean launch time of the fused kernel vs the measured
            HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference package (baseline/_ref; else the numpy restatement of
            oracle/, "port") on a bounded sample of the same workload, one process per host core

Multi-GPU (torchrun, one process per GPU).  The headline stays query-DP weak scaling (every
rank re-ranks its own 5193 x 5000 pairs over a replica of the index; no data-path collective).
The same line carries two more records, measured in the same process group:
  strong    the SAME 5193 x 5000 job split by query over the N ranks (total work fixed)
  sharded   BASELINE configs[4] per GPU: a doc-id-range shard of 1.2 M docs / 7.5 M passages per
            rank, 12 500 x N queries x 5000 candidates, k = 1000 — the one path with a real
            exchange: the fused kernel's epilogue stores every query's top-k list into its owner
            rank's buffer over NVLink peer memory, or (checked bit for bit against it on the
            same inputs) one NCCL all-to-all; split into local kernel / exchange / merge

`--impl reference` times the CPU reference alone (rank 0 only) and prints the same line shape.
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
L2_BYTES = 126e6  # B200 L2

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
    ap.add_argument("--no-api", action="store_true", help="skip the api_e2e record (drop-in Python API, 1 GPU)")
    ap.add_argument("--no-extra", action="store_true",
                    help="N > 1: skip the strong-scaling and doc-id-sharded records")
    ap.add_argument("--alpha", type=float, default=0.1)
    ap.add_argument("--cpu-budget", default="short", choices=["short", "long"],
                    help="--impl reference: queries per core and step (short: a few seconds per step)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="sharded workloads: fused peer-memory exchange in the kernel epilogue, or an NCCL all-to-all")
    ap.add_argument("--emulate-shards", type=int, default=0,
                    help="single GPU: hold shard 0 of this many doc-id-range shards and time the local step only "
                         "(profiling the sharded kernel without the other GPUs; the line says so)")
    return ap.parse_args()


def build_corpus(wl, scale, rank, world):
    """Row counts per document of the (rank's part of the) synthetic corpus (tools/sweep.py)."""
    if "n_docs" in wl:
        n_docs = max(wl["cands"] * 2, int(wl["n_docs"] * scale)) * (world if wl.get("sharded") else 1)
        return n_docs, doc_lengths(n_docs, wl["mean_psg"], seed=0)
    return None, None


def doc_lengths(n_docs, mean_psg, seed=0):
    """Clipped geometric passages/doc, mean ~6.25, min 1, max 64."""
    rng = np.random.default_rng(seed)
    return np.clip(rng.geometric(1.0 / mean_psg, n_docs), 1, 64).astype(np.int64)


# ------------------------------------------------------------------------------------------
# CPU baseline: the unmodified reference (baseline/_ref) or the numpy port of oracle/, one
# process per core.  The pool, the host index and the sample are built ONCE; every step times
# only the reference's own three calls.
# ------------------------------------------------------------------------------------------
_CPU = {}
CPU_DOCS = 40_000  # 250 k passages x 3 KB = 768 MB: not L3-resident


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


def _port_worker(args):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ff_oracle as fo

    q_lo, q_hi, cands, k, alpha, _ = args
    if q_hi <= q_lo:
        return 0.0
    vec, off, qv, cand, lex = _CPU["vec"], _CPU["off"], _CPU["qv"], _CPU["cand"], _CPU["lex"]
    sl = slice(q_lo * cands, q_hi * cands)
    pair_q = np.repeat(np.arange(q_lo, q_hi), cands)
    t0 = time.perf_counter()
    # index/base.py:445-459 batches queries; 2 queries/batch keeps temporaries ~200 MB
    ff = fo.score_pairs(vec, off, _CPU["rows"], pair_q, cand[sl], qv, fo.MODE_MAXP, chunk_pairs=2 * cands)
    it = fo.interpolate_f32(lex[sl], ff, alpha)
    fo.topk_per_query(np.arange(q_hi - q_lo + 1) * cands, it, k)
    return time.perf_counter() - t0


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


class CpuBaseline:
    """`kind` "reference": the UNMODIFIED reference package — `Index.__call__(batch_size=2|3)` +
    `Ranking.interpolate` + `Ranking.cut` (index/base.py:389-469, ranking.py:293-326,279-291);
    `kind` "port": the numpy restatement of oracle/ (when baseline/_ref did not travel).  `cores`
    processes x `q_per_core` queries of the workload's shape over a scaled-down host index
    (cost per pair does not depend on the index size once it is out of cache)."""

    def __init__(self, wl, alpha, q_per_core, cores=None):
        import multiprocessing as mp

        self.kind = "reference" if reference_installed() and os.environ.get("FFX_CPU_BASELINE", "") != "port" else "port"
        self.cores = cores = min(cores or os.cpu_count() or 1, 64)
        self.q_per_core, self.alpha, self.wl = q_per_core, alpha, wl
        n_docs = int(os.environ.get("FFX_CPU_DOCS", CPU_DOCS))
        cnt = doc_lengths(n_docs, wl["mean_psg"], seed=1)
        rng = np.random.default_rng(2)
        n_rows = int(cnt.sum())
        _CPU["vec"] = rng.standard_normal((n_rows, DIM), dtype=np.float32)
        if self.kind == "reference":
            _CPU["doc_ids"] = np.repeat([f"D{d}" for d in range(n_docs)], cnt).tolist()
        else:
            _CPU["off"] = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
            _CPU["rows"] = np.arange(n_rows, dtype=np.int64)
        nq = cores * q_per_core
        cands = wl["cands"]
        _CPU["qv"] = rng.standard_normal((nq, DIM), dtype=np.float32)
        _CPU["cand"] = np.concatenate([rng.choice(n_docs, cands, replace=False) for _ in range(nq)])
        _CPU["lex"] = rng.uniform(0, 20, nq * cands).astype(np.float32)
        # the reference crashes on an empty trailing batch (batch_size | #queries, index/base.py:455)
        batch = 3 if q_per_core % 2 == 0 else 2
        self.jobs = [(c * q_per_core, (c + 1) * q_per_core, cands, wl["k"], alpha, batch) for c in range(cores)]
        self.worker = _ref_worker if self.kind == "reference" else _port_worker
        self.pool = mp.get_context("fork").Pool(cores)
        # import + build the host index in every worker (outside any timed step)
        self.pool.map(self.worker, [(0, 0, cands, wl["k"], alpha, 2)] * cores, chunksize=1)
        self.pairs = nq * cands
        what = (f"the UNMODIFIED reference package (baseline/_ref): Index.__call__(batch_size={batch}) + "
                f"Ranking.interpolate + Ranking.cut" if self.kind == "reference" else
                "numpy port of index/base.py:279-314 + ranking.py:319,279-291 (oracle/)")
        self.sample = (f"{nq} queries x {cands} candidates MAXP over a {n_docs}-doc / {n_rows}-passage 768-d fp32 host "
                       f"index ({n_rows * DIM * 4 / 1e6:.0f} MB), {what}, {cores} processes x {q_per_core} queries")

    def step(self):
        """Every process times its own calls (its first-stage Ranking is built before its clock
        starts); the processes run side by side, the slowest one is the step."""
        dt = max(self.pool.map(self.worker, self.jobs, chunksize=1))
        return self.pairs / dt, dt

    def close(self):
        self.pool.close()
        self.pool.join()
        _CPU.clear()


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    force = int(os.environ.get("FFX_CPU_Q_PER_CORE", "0"))  # tests shrink the sample
    base = CpuBaseline(wl, args.alpha, force or (9 if args.cpu_budget == "long" else 5))
    vals, times = [], []
    for step in range(args.warmup + args.steps):
        v, dt = base.step()
        if step >= args.warmup:
            vals.append(v)
            times.append(dt)
    base.close()
    value = float(np.mean(vals)) if vals else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)) if times else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "c3_msmarco_doc_maxp", "mode": "MAXP", "dim": DIM, "candidates_per_query": wl["cands"],
                   "cut_k": wl["k"], "alpha": args.alpha,
                   "note": "each step = a bounded sample of the workload on the host cores"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": base.cores, "kind": base.kind,
                         "sample": base.sample},
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
    """dram bytes per launch of the timed kernel from the committed ncu capture of this very
    command (profiles/ncu_traffic.json names the capture and the kernel symbol), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# synthetic workloads on the device
# ------------------------------------------------------------------------------------------
class Ctx:
    """Process-group facts of this rank."""

    def __init__(self, args):
        import torch

        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local)
        self.dist = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), device=self.dev, dtype=self.torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def all_true(self, flag: bool) -> bool:
        t = self.torch.tensor([1 if flag else 0], device=self.dev, dtype=self.torch.int32)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())


def stage_index(ctx, _ffx, wl, scale, shards, shard):
    """The (rank's part of the) synthetic corpus, generated on the device and staged into a
    libffx row store.  `shards` > 1: this index holds doc-id-range shard `shard`."""
    from fast_forward.sharded import plan_doc_shards

    torch, dev = ctx.torch, ctx.dev
    pq = wl["kind"] == "opq"
    cands = wl["cands"]
    m = {"pq": pq, "width": wl["M"] if pq else DIM}
    if "n_docs" in wl:
        n_docs = max(cands * 2, int(wl["n_docs"] * scale)) * (shards if wl.get("sharded") else 1)
        cnt = doc_lengths(n_docs, wl["mean_psg"], seed=0)
        off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        total_rows, pool = int(off[-1]), n_docs
    else:
        n_docs, cnt, off = None, None, None
        total_rows = pool = max(cands * 2, int(wl["n_rows"] * scale))
    doc_lo, doc_hi, row_lo, row_hi = 0, n_docs, 0, total_rows
    if shards > 1:
        bounds = plan_doc_shards(cnt, shards)
        doc_lo, doc_hi = int(bounds[shard]), int(bounds[shard + 1])
        row_lo, row_hi = int(off[doc_lo]), int(off[doc_hi])
    n_rows = row_hi - row_lo
    t0 = time.perf_counter()
    idx = _ffx.DeviceIndex(m["width"], capacity=n_rows, row_kind=_ffx.ROWS_PQ_U8 if pq else _ffx.ROWS_F32,
                           device=ctx.local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + (shard if shards > 1 else 0))  # replicas are identical on every rank
    chunk = 1 << 20
    for r0 in range(0, n_rows, chunk):
        nr = min(chunk, n_rows - r0)
        if pq:
            t = torch.randint(0, wl["Ks"], (nr, m["width"]), device=dev, dtype=torch.uint8, generator=gen)
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
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    m.update(idx=idx, n_docs=n_docs, cnt=cnt, off=off, total_rows=total_rows, pool=pool, doc_lo=doc_lo,
             doc_hi=doc_hi, row_lo=row_lo, row_hi=row_hi, n_rows=n_rows, stage_s=time.perf_counter() - t0,
             row_bytes=m["width"] * (1 if pq else 4))
    return m


def draw_queries(ctx, nq, cands, pool, seed, zipf=False):
    """Query vectors + stratified distinct candidates, uniform over the corpus (or Zipf-like
    inside every stratum) + first-stage scores, all on the device."""
    torch, dev = ctx.torch, ctx.dev
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    qv = torch.randn((nq, DIM), device=dev, dtype=torch.float32, generator=gen)
    bucket = pool // cands
    if bucket < 1:
        raise RuntimeError("corpus smaller than the candidate list")
    parts = []
    for q0 in range(0, nq, 4096):  # bounded temporaries
        qn = min(4096, nq - q0)
        perm = torch.rand((qn, cands), device=dev, generator=gen).argsort(dim=1)
        if zipf:  # P(rank r inside the stratum) ~ 1/r: log-uniform draw
            u = torch.rand((qn, cands), device=dev, generator=gen)
            within = (torch.exp(u * float(np.log(bucket))) - 1.0).long().clamp_(0, bucket - 1)
        else:
            within = torch.randint(0, bucket, (qn, cands), device=dev, generator=gen)
        parts.append((perm * bucket + within).to(torch.int32))
        del perm, within
    cand = torch.cat(parts).contiguous().view(-1)
    del parts
    lex = (torch.rand((nq * cands,), device=dev, generator=gen) * 20).contiguous()
    q_off = (torch.arange(nq + 1, device=dev, dtype=torch.int64) * cands).contiguous()
    return {"qv": qv, "cand": cand, "lex": lex, "q_off": q_off, "nq": nq, "cands": cands}


def algorithmic_bytes(ctx, m, q, mode, k):
    """row bytes of every pair this rank scores + 16 B/pair (candidate 4, lexical
    score 4, span 8 — the span only for the pairs this rank owns) + per query (query vector +
    top-k out).  Returns (bytes, rows touched, own pairs)."""
    torch = ctx.torch
    n_pairs = q["nq"] * q["cands"]
    if m["cnt"] is not None and mode != MODES["FIRSTP"] and mode != MODES["PASSAGE"]:
        d_cnt = torch.from_numpy(m["cnt"]).to(ctx.dev)
        c64 = q["cand"].long()
        mine = (c64 >= m["doc_lo"]) & (c64 < m["doc_hi"])
        rows_touched = int(d_cnt[c64[mine]].sum().item())
        my_pairs = int(mine.sum().item())
        del d_cnt, c64, mine
    else:
        rows_touched = my_pairs = n_pairs
    total = rows_touched * m["row_bytes"] + n_pairs * 8 + my_pairs * 8 + q["nq"] * (DIM * 4 + k * 8)
    return total, rows_touched, my_pairs


class Flusher:
    """Between timed steps of a workload whose index fits L2: overwrite L2 with a 512 MB store."""

    def __init__(self, ctx, needed):
        self.buf = ctx.torch.empty(512 << 20, dtype=ctx.torch.uint8, device=ctx.dev) if needed else None

    def __call__(self):
        if self.buf is not None:
            self.buf.add_(1)


def time_steps(ctx, step, steps, warmup, flush=None):
    """W untimed steps, then K timed ones bracketed by barrier + synchronize on both sides; CUDA
    events on the launching (current) stream.  Without an L2 flush the region is one event pair
    around all K steps; with one, every step has its own pair (the flush is not timed).
    Returns (total ms of this rank, per-step ms, last output)."""
    torch = ctx.torch
    stream = torch.cuda.current_stream()
    out = None
    for _ in range(warmup):
        out = step()
    ctx.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        if flush is not None:
            flush()
        a.record(stream)
        out = step()
        b.record(stream)
    ctx.barrier()
    per_step = [a.elapsed_time(b) for a, b in ev]
    total = sum(per_step) if (flush is not None and flush.buf is not None) else ev[0][0].elapsed_time(ev[-1][1])
    return total, per_step, out


def check_lists(out_s, out_p, k, cands, padded=False):
    """Sanity on a timed output: ranked lists sorted, positions distinct and in range."""
    s_host, p_host = out_s[:4].cpu().numpy(), out_p[:4].cpu().numpy()
    if not padded:  # a lone shard's lists are padded with (-inf, -1)
        assert (np.diff(s_host, axis=1) <= 0).all(), "top-k not sorted"
        assert all(len(set(r.tolist())) == k for r in p_host) and p_host.min() >= 0 and p_host.max() < cands
    return s_host, p_host


def e2e_host(ctx, _ffx, idx, q, mode, alpha, k, steps, check_pos):
    """The same job through ffx_rerank_host: pinned host inputs -> H2D -> kernel -> D2H of the
    ranked lists, wall clock around K calls (each call synchronises)."""
    torch = ctx.torch
    nq, n_pairs = q["nq"], q["nq"] * q["cands"]
    h_q = _ffx.PinnedBuffer((nq, DIM), np.float32)
    h_off = _ffx.PinnedBuffer((nq + 1,), np.int64)
    h_cand = _ffx.PinnedBuffer((n_pairs,), np.int32)
    h_lex = _ffx.PinnedBuffer((n_pairs,), np.float32)
    h_ts = _ffx.PinnedBuffer((nq, k), np.float32)
    h_tp = _ffx.PinnedBuffer((nq, k), np.int32)
    h_q.array[:] = q["qv"].cpu().numpy()
    h_off.array[:] = q["q_off"].cpu().numpy()
    h_cand.array[:] = q["cand"].cpu().numpy()
    h_lex.array[:] = q["lex"].cpu().numpy()
    out = {"topk_score": h_ts.array, "topk_pos": h_tp.array}

    def step():
        idx.rerank_host(mode, h_q.array, h_off.array, h_cand.array, h_lex.array, alpha, k,
                        want_ff=False, want_int=False, out=out)

    step()
    assert (h_tp.array[:4] == check_pos).all(), "host-buffer path disagrees with the device path"
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    h2d = h_q.array.nbytes + h_off.array.nbytes + h_cand.array.nbytes + h_lex.array.nbytes
    d2h = h_ts.array.nbytes + h_tp.array.nbytes
    host = {"qv": h_q, "q_off": h_off, "cand": h_cand, "lex": h_lex}
    return dt, h2d, d2h, host


def kernel_label(_ffx, wl):
    name = _ffx.last_kernel()  # the launched symbol: "void ffx::ffx_score_tma_kernel<2, 12, true, 32>(...)"
    fused = ", true" in name or "<true" in name
    what = ("TMA-staged codes, XOR-swizzled conflict-free LUT look-ups, %s-interpolate%s; bound by shared-memory "
            "look-ups, not HBM" if wl["kind"] == "opq" else "TMA-staged gather-dot-%s-interpolate%s")
    return f"{name} ({what % (wl['mode'], '-topk fused' if fused else ' ; ffx_topk_kernel')})"


# ------------------------------------------------------------------------------------------
# the records
# ------------------------------------------------------------------------------------------
def run_plain(ctx, _ffx, args, wl, m, warmup):
    """Query-DP (or single GPU): every rank re-ranks its own queries over its replica."""
    torch = ctx.torch
    mode, cands, k = MODES[wl["mode"]], wl["cands"], wl["k"]
    nq = wl["nq"] if args.scale == 1 else max(296, int(wl["nq"] * args.scale))
    idx = m["idx"]
    q = draw_queries(ctx, nq, cands, m["pool"], 99 + ctx.rank, zipf=wl.get("popularity") == "zipf")
    topk_s = torch.empty((nq, k), device=ctx.dev, dtype=torch.float32)
    topk_p = torch.empty((nq, k), device=ctx.dev, dtype=torch.int32)
    algo, rows_touched, _ = algorithmic_bytes(ctx, m, q, mode, k)
    stream = torch.cuda.current_stream()

    def step():
        idx.rerank_device(mode, q["qv"].data_ptr(), nq, q["q_off"].data_ptr(), q["cand"].data_ptr(),
                          q["lex"].data_ptr(), args.alpha, k, cands, 0, 0, topk_s.data_ptr(), topk_p.data_ptr(),
                          stream.cuda_stream)
        return topk_s, topk_p

    unique_bytes = m["n_rows"] * m["row_bytes"]
    flush = Flusher(ctx, unique_bytes < 4 * L2_BYTES)
    for _ in range(warmup):
        step()
    idx.sync(stream.cuda_stream)
    clk = clocks_start() if ctx.rank == 0 else (None, None)
    launches0 = _ffx.launch_count()
    total_ms, per_step, (out_s, out_p) = time_steps(ctx, step, args.steps, 0, flush)
    launches = _ffx.launch_count() - launches0
    clocks = clocks_stop(clk[0], clk[1], ctx.local) if ctx.rank == 0 else None
    kernel = kernel_label(_ffx, wl)
    _, p_host = check_lists(out_s, out_p, k, cands)
    e2e_s, h2d, d2h, host = e2e_host(ctx, _ffx, idx, q, mode, args.alpha, k, args.steps, p_host)
    total_ms, e2e_ms = ctx.max_over_ranks([total_ms, e2e_s * 1e3])
    l2 = (f"index {unique_bytes / 1e6:.1f} MB fits the 126 MB L2: L2 flushed between timed steps (512 MB store, untimed); "
          f"{algo / 1e9:.2f} GB touched per step" if flush.buf is not None else
          "inputs larger than L2 (index %.1f GB/GPU, %.1f GB touched per step per GPU vs 126 MB of L2)" % (
              unique_bytes / 1e9, algo / 1e9))
    return {"q": q, "host": host, "nq": nq, "total_ms": total_ms, "per_step": per_step, "e2e_ms": e2e_ms,
            "h2d": h2d, "d2h": d2h, "algo": algo, "rows_touched": rows_touched, "launches": launches,
            "clocks": clocks, "kernel": kernel, "l2": l2, "mode": mode, "k": k, "cands": cands}


def run_strong(ctx, _ffx, args, wl, m):
    """Strong scaling: ONE C3 job (same seed on every rank), queries split contiguously over the
    ranks, index replicated; time = max over ranks."""
    torch = ctx.torch
    mode, cands, k = MODES[wl["mode"]], wl["cands"], wl["k"]
    nq_all = wl["nq"] if args.scale == 1 else max(296, int(wl["nq"] * args.scale))
    lo, hi = nq_all * ctx.rank // ctx.world, nq_all * (ctx.rank + 1) // ctx.world
    full = draw_queries(ctx, nq_all, cands, m["pool"], 4242)
    nq = hi - lo
    qv = full["qv"][lo:hi].contiguous()
    cand = full["cand"][lo * cands:hi * cands].contiguous()
    lex = full["lex"][lo * cands:hi * cands].contiguous()
    q_off = full["q_off"][:nq + 1].contiguous()
    del full
    topk_s = torch.empty((nq, k), device=ctx.dev, dtype=torch.float32)
    topk_p = torch.empty((nq, k), device=ctx.dev, dtype=torch.int32)
    stream = torch.cuda.current_stream()
    idx = m["idx"]

    def step():
        idx.rerank_device(mode, qv.data_ptr(), nq, q_off.data_ptr(), cand.data_ptr(), lex.data_ptr(), args.alpha,
                          k, cands, 0, 0, topk_s.data_ptr(), topk_p.data_ptr(), stream.cuda_stream)
        return topk_s, topk_p

    launches0 = _ffx.launch_count()
    total_ms, per_step, (out_s, out_p) = time_steps(ctx, step, args.steps, 3)
    launches = (_ffx.launch_count() - launches0) // (args.steps + 3)
    idx.sync(stream.cuda_stream)
    check_lists(out_s, out_p, k, cands)
    (total_ms,) = ctx.max_over_ranks([total_ms])
    return {"value": nq_all * cands * args.steps / (total_ms * 1e-3), "unit": "pairs/s",
            "ms_per_step": total_ms / args.steps, "scaling": "strong",
            "queries_total": nq_all, "queries_per_rank": f"{nq_all // ctx.world}..{-(-nq_all // ctx.world)}",
            "launches_per_step": launches, "kernel": _ffx.last_kernel(),
            "note": "the 1-GPU job (5193 x 5000, MAXP, k = 5000) split by query over the ranks, replicated index, "
                    "no collective; device time, max over ranks"}


def run_sharded(ctx, _ffx, args, wl, m, emulate, warmup):
    """Doc-id-range shards: every rank scores its part of every query's candidates, local top-k,
    exchange to the query's owner, merge there."""
    from fast_forward.sharded import ShardedReranker

    torch = ctx.torch
    world = emulate or ctx.world
    mode, cands, k = MODES[wl["mode"]], wl["cands"], wl["k"]
    nq = (wl["nq"] if args.scale == 1 else max(296, int(wl["nq"] * args.scale))) * world
    q = draw_queries(ctx, nq, cands, m["pool"], 99)  # shards all see the same queries
    algo, rows_touched, my_pairs = algorithmic_bytes(ctx, m, q, mode, k)
    idx = m["idx"]
    use_p2p = ctx.world > 1 and args.exchange == "p2p"
    rr = ShardedReranker(idx, m["doc_lo"], m["n_docs"], m["row_lo"], m["total_rows"], p2p=use_p2p)

    def step_of(reranker):
        def step():
            return reranker.rerank(mode, q["qv"], q["q_off"], q["cand"], q["lex"], args.alpha, k, cands,
                                   gather_result=False)
        return step

    def phases(reranker, steps):
        reranker.trace = []
        total_ms, per_step, out = time_steps(ctx, step_of(reranker), steps, 0)
        marks, reranker.trace = reranker.trace, None
        split = [float(np.mean([mk[i].elapsed_time(mk[i + 1]) for mk in marks])) for i in range(3)] if marks and \
            len(marks[0]) == 4 else None
        return total_ms, per_step, out, split

    step = step_of(rr)
    for _ in range(warmup):
        step()
    ctx.barrier()
    idx.sync(torch.cuda.current_stream().cuda_stream)
    clk = clocks_start() if ctx.rank == 0 else (None, None)
    launches0 = _ffx.launch_count()
    total_ms, per_step, (out_s, out_p), split = phases(rr, args.steps)
    launches = _ffx.launch_count() - launches0
    clocks = clocks_stop(clk[0], clk[1], ctx.local) if ctx.rank == 0 else None
    kernel = _ffx.last_kernel()
    check_lists(out_s, out_p, k, cands, padded=bool(emulate) or ctx.world == 1)

    # the other exchange on the same inputs: must give the same lists, bit for bit
    other, verified = None, None
    if ctx.world > 1:
        rr2 = ShardedReranker(idx, m["doc_lo"], m["n_docs"], m["row_lo"], m["total_rows"], p2p=not use_p2p)
        step2 = step_of(rr2)
        for _ in range(2):
            step2()
        o_ms, _, (s2, p2), o_split = phases(rr2, max(2, min(args.steps, 5)))
        same = bool(torch.equal(out_s, s2)) and bool(torch.equal(out_p, p2))
        verified = ctx.all_true(same)
        (o_ms,) = ctx.max_over_ranks([o_ms / max(2, min(args.steps, 5))])
        other = {"exchange": "nccl all-to-all" if use_p2p else "p2p", "ms_per_step": o_ms, "split_ms": o_split}

    # e2e: the job starts in pinned host memory on every rank (every shard needs every query's
    # full candidate list), results = this rank's merged lists read back
    host = {kk: q[kk].cpu().pin_memory() for kk in ("qv", "q_off", "cand", "lex")}
    h2d = sum(t.numel() * t.element_size() for t in host.values())
    res = None

    def e2e_step():
        nonlocal res
        res = rr.rerank_host(mode, host["qv"], host["q_off"], host["cand"], host["lex"], args.alpha, k, cands,
                             chunk_queries=-(-nq // int(os.environ.get("FFX_SHARD_CHUNKS", "8"))))

    del q["cand"], q["lex"]
    torch.cuda.empty_cache()
    e2e_step()
    assert res[1].shape[0] > 0 and res[1].shape[1] == k
    ctx.barrier()
    e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
    res_s, res_p = res[1], res[2]
    d2h = res_s.numel() * 4 + res_p.numel() * 4
    total_ms, e2e_ms = ctx.max_over_ranks([total_ms, e2e_ms])
    kern_ms = float(np.mean(per_step))
    rec = {
        "value": nq * cands * args.steps / (total_ms * 1e-3), "unit": "pairs/s", "ms_per_step": total_ms / args.steps,
        "scaling": "weak", "workload": "c5_sharded_maxp" + (f" (SCALED x{args.scale})" if args.scale != 1 else ""),
        "docs": m["n_docs"], "passages": m["total_rows"], "index_gb_per_gpu": m["n_rows"] * m["row_bytes"] / 1e9,
        "queries": nq, "candidates_per_query": cands, "cut_k": k,
        "exchange": ("p2p: fused kernel epilogue stores each query's [k] list into its owner rank's buffer over "
                     "NVLink peer memory, device barrier, ffx_merge_topk at the owner") if use_p2p else
                    "nccl: local fused top-k, one all-to-all of the [nq,k] lists to the owners, ffx_merge_topk there",
        "split_ms": None if split is None else {"local_kernel": split[0], "exchange" if not use_p2p else "barrier": split[1],
                                                "merge": split[2]},
        "exchange_ms": None if split is None else split[1] + split[2],
        "verified": verified, "other_exchange": other, "kernel": kernel,
        "per_gpu_hbm_gbs": algo / (kern_ms * 1e-3) / 1e9 if split is None else algo / (split[0] * 1e-3) / 1e9,
        "algorithmic_bytes_per_gpu": algo, "own_pairs_per_gpu": my_pairs,
        "e2e": {"value": nq * cands / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "ShardedReranker.rerank_host, per rank: the full candidate lists H2D from pinned memory in 8 "
                        "query chunks pipelined against the sharded re-rank of the previous chunk, D2H of the "
                        "merged lists of the queries this rank owns"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    return rec, {"total_ms": total_ms, "per_step": per_step, "algo": algo, "rows_touched": rows_touched, "nq": nq}


def run_api(ctx, args, wl, m, q_host, alpha, k):
    """The drop-in Python API at the bench's shape: a first-stage `Ranking` of id STRINGS (built
    once, outside the clock, like the reference arm's) -> `index.rerank(ranking, alpha, k)` ->
    `Ranking`.  First call = ids hashed against the index's dictionaries; second call = the codes
    cached on the ranking.  Wall clock."""
    import pandas as pd
    import pyarrow as pa
    import pyarrow.compute as pc

    sys.path.insert(0, PKG)
    import fast_forward
    from fast_forward.encoder import TableEncoder
    from fast_forward.index import InMemoryIndex, Mode

    nq, cands = len(q_host["qv"].array), wl["cands"]
    qv = q_host["qv"].array
    t0 = time.perf_counter()
    enc = TableEncoder({f"text {i}": qv[i] for i in range(nq)})
    doc_ids = pc.binary_join_element_wise(pa.scalar("D", pa.large_string()), pa.array(np.repeat(np.arange(m["n_docs"]), m["cnt"])).cast(
        pa.large_string()), pa.scalar("", pa.large_string()))
    index = InMemoryIndex._adopt(m["idx"], doc_ids=doc_ids, query_encoder=enc, mode=Mode[wl["mode"]],
                                 encoder_batch_size=1024)
    index._device()  # document id -> rows table pushed to the device (once per index)
    t_index = time.perf_counter() - t0
    t0 = time.perf_counter()
    ids = pc.binary_join_element_wise(pa.scalar("D", pa.large_string()), pa.array(q_host["cand"].array).cast(pa.large_string()), pa.scalar("", pa.large_string()))
    q_ids = pc.binary_join_element_wise(pa.scalar("q", pa.large_string()), pa.array(np.repeat(np.arange(nq), cands)).cast(
        pa.large_string()), pa.scalar("", pa.large_string()))
    str_dtype = pd.StringDtype("pyarrow", na_value=np.nan)  # pandas' `str`
    frame = pd.DataFrame({"q_id": pd.Series(pd.arrays.ArrowStringArray(pa.chunked_array([q_ids]), dtype=str_dtype)),
                          "id": pd.Series(pd.arrays.ArrowStringArray(pa.chunked_array([ids]), dtype=str_dtype)),
                          "score": q_host["lex"].array})
    t_frame_in = time.perf_counter() - t0  # the bench's own id strings (pyarrow), not the API
    t0 = time.perf_counter()
    first = fast_forward.Ranking(frame, queries={f"q{i}": f"text {i}" for i in range(nq)})
    t_ranking = time.perf_counter() - t0
    times = []
    out = None
    for call in range(6):
        first_profile = call == 0 and os.environ.get("FFX_API_PROFILE") == "first"
        if first_profile:  # where the FIRST call (ids hashed, buffers pinned) spends its time (stderr)
            import cProfile
            import pstats

            prof = cProfile.Profile()
            prof.enable()
        t0 = time.perf_counter()
        out = index.rerank(first, alpha, k)
        n_out = out.num_rows
        times.append(time.perf_counter() - t0)
        if first_profile:
            prof.disable()
            pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(22)
    assert n_out == nq * min(k, cands)
    if os.environ.get("FFX_API_PROFILE"):  # where a steady-state call spends its time (stderr)
        import cProfile
        import pstats

        prof = cProfile.Profile()
        prof.enable()
        index.rerank(first, alpha, k)
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(18)
    t0 = time.perf_counter()
    df = out._df
    t_frame = time.perf_counter() - t0
    assert len(df) == n_out
    pairs = nq * cands
    return {"value": pairs / times[1], "unit": "pairs/s", "first_call_s": times[0], "second_call_s": times[1],
            "later_calls_s": times[2:], "result_frame_s": t_frame, "build_index_ids_s": t_index, "build_input_frame_s": t_frame_in,
            "build_first_stage_ranking_s": t_ranking,
            "call": "Ranking(q_id / id strings, float32 scores, queries attached) -> index.rerank(ranking, alpha, "
                    f"{k}) -> Ranking; query encoder = table look-up of the precomputed vectors",
            "note": "first call hashes every id string against the index's C++ dictionaries; later calls reuse the "
                    "integer codes cached on the ranking.  The result is a Ranking whose pandas frame is "
                    "materialised on first access (result_frame_s, not included in the call times)"}


def run_ffx(args, wl):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, PKG)
    from fast_forward import _ffx

    ctx = Ctx(args)
    sharded = bool(wl.get("sharded"))
    emulate = args.emulate_shards if ctx.world == 1 and sharded else 0
    if _ffx.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: libffx has no CPU path")
    if os.environ.get("FFX_CHUNK_WAVES"):
        _ffx.set_option("chunk_waves", int(os.environ["FFX_CHUNK_WAVES"]))
    torch.cuda.set_device(ctx.local)
    if ctx.world > 1:
        dist.init_process_group("nccl", device_id=ctx.dev)
        ctx.dist = dist
    warmup = max(args.warmup, 3)
    peak, peak_src = measured_peak()
    traffic = ncu_traffic()
    scaled = "" if args.scale == 1 else f" (SCALED x{args.scale}: debug)"

    if sharded:
        world = emulate or ctx.world
        m = stage_index(ctx, _ffx, wl, args.scale, world, 0 if emulate else ctx.rank)
        rec, raw = run_sharded(ctx, _ffx, args, wl, m, emulate, warmup)
        if ctx.rank == 0:
            achieved = rec["per_gpu_hbm_gbs"]
            line = {
                "metric": f"re-ranked (query,doc) pairs/sec, {wl['mode']} k={wl['k']}",
                "value": rec["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload + scaled + (
                    f" (EMULATED: local step of shard 0 of {emulate} on one GPU, no exchange; profiling only)"
                    if emulate else ""),
                    "mode": wl["mode"], "dim": DIM, "docs": m["n_docs"], "passages": m["total_rows"],
                    "index_gb_per_gpu": rec["index_gb_per_gpu"], "queries": rec["queries"],
                    "candidates_per_query": wl["cands"], "cut_k": wl["k"], "alpha": args.alpha,
                    "parallelism": f"doc-id-range shards x{world}: {rec['exchange']}",
                    "l2": "inputs larger than L2 (index %.1f GB/GPU, %.1f GB touched per step per GPU vs 126 MB of L2)" % (
                        rec["index_gb_per_gpu"], raw["algo"] / 1e9),
                    "index_stage_s": round(m["stage_s"], 2)},
                "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "sharded": rec,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "kernel": rec["kernel"], "algorithmic_bytes_per_launch": raw["algo"],
                             "kernel_ms": (rec["split_ms"] or {}).get("local_kernel", float(np.mean(raw["per_step"]))),
                             "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0},
                "clocks": rec["clocks"],
            }
            emit(line)
    else:
        m = stage_index(ctx, _ffx, wl, args.scale, 1, 0)
        r = run_plain(ctx, _ffx, args, wl, m, warmup)
        n_pairs = r["nq"] * r["cands"]
        extra = {}
        default_job = args.workload == "c3_msmarco_doc_maxp"
        if ctx.world > 1 and default_job and not args.no_extra:
            extra["strong"] = run_strong(ctx, _ffx, args, wl, m)
        api = None
        if ctx.world == 1 and default_job and not args.no_api:
            try:
                api = run_api(ctx, args, wl, m, r["host"], args.alpha, r["k"])
            except Exception as e:  # the headline must not die with the API record
                api = {"error": f"{type(e).__name__}: {e}"}
        r["host"] = None
        if ctx.world > 1 and default_job and not args.no_extra:
            m["idx"].close()
            del r["q"]
            torch.cuda.empty_cache()
            swl = WORKLOADS["c5_sharded_maxp"]
            sm = stage_index(ctx, _ffx, swl, args.scale, ctx.world, ctx.rank)
            extra["sharded"], _ = run_sharded(ctx, _ffx, args, swl, sm, 0, warmup)
            sm["idx"].close()
            m["idx"] = None
        if ctx.rank == 0:
            value = ctx.world * n_pairs * args.steps / (r["total_ms"] * 1e-3)
            kern_ms = float(np.mean(r["per_step"]))
            achieved = r["algo"] / (kern_ms * 1e-3) / 1e9
            pq = m["pq"]
            line = {
                "metric": METRIC if default_job else f"re-ranked (query,doc) pairs/sec, {wl['mode']} k={wl['k']}",
                "value": value, "unit": "pairs/s", "n_gpus": ctx.world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": r["total_ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8 codes + f32 LUT" if pq else "f32", "data": "synthetic",
                "config": {
                    "workload": args.workload + scaled, "mode": wl["mode"], "dim": DIM, "docs": m["n_docs"],
                    "passages": m["total_rows"], "index_gb_per_gpu": m["n_rows"] * m["row_bytes"] / 1e9,
                    "queries": f"{r['nq']} per GPU", "candidates_per_query": r["cands"], "cut_k": r["k"],
                    "alpha": args.alpha,
                    "parallelism": f"query-dp{ctx.world} (replicated index, no data-path collective)",
                    "l2": r["l2"], "index_stage_s": round(m["stage_s"], 2)},
                "e2e": {"value": ctx.world * n_pairs * args.steps / (r["e2e_ms"] * 1e-3), "unit": "pairs/s",
                        "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                        "ms_per_step": r["e2e_ms"] / args.steps},
                "gpu_launches": int(r["launches"]),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak,
                             "traffic": (traffic or {}).get(args.workload) if args.scale == 1 else None,
                             "traffic_source": (traffic or {}).get("_source") if args.scale == 1 else None,
                             "kernel": r["kernel"], "algorithmic_bytes_per_launch": r["algo"], "kernel_ms": kern_ms,
                             "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0},
                "clocks": r["clocks"],
            }
            if pq:
                lut = r["rows_touched"] * wl["M"] / (kern_ms * 1e-3)
                line["roofline"]["lut_lookups_per_s"] = lut
                smem_peak = (traffic or {}).get("smem_lookup_peak_per_s")
                if smem_peak:
                    line["roofline"].update({
                        "bound": "shared-memory look-ups", "achieved": lut / 1e9, "peak": smem_peak / 1e9,
                        "unit": "G look-ups/s", "frac": lut / smem_peak,
                        "peak_source": (traffic or {}).get("smem_lookup_peak_source"),
                        "hbm_gbs": achieved, "frac_of_nominal_8TBs": None})
            line.update(extra)
            if api is not None:
                line["api_e2e"] = api
            if not args.no_cpu_baseline and ctx.world == 1 and default_job:
                line["cpu_baseline"] = cpu_baseline_subprocess(args)
            emit(line)
        if m["idx"] is not None:
            m["idx"].close()
    if ctx.dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
