#!/usr/bin/env python
"""bench.py — headline benchmark of the SDF glyph rendering path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # the B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

One "step" = one pass of the hot path over one batch: the whole workload rendered once.  At N = 1 the workload is
BASELINE.json configs[1], the 20-file Noto Sans `merge` (6480 glyphs / 6445 bitmaps / 3.96 M segments / 2.67 G
pixel x segment pairs).  At N > 1 it is ONE job of N such fonts (the merge under N font names, i.e. a multi-font
`recurse`), cut into N shards by font x GlyphBlock with longest-processing-time-first over per-block cost estimates
(FontManager::shard_owners; reference task list src/font/manager.rs:88-97); rank r renders shard r on GPU r and
rank 0 checks that the gathered files equal an unsharded run.  Per-GPU work is constant: "scaling": "weak".

  value      glyphs/s with glyph requests, fonts (glyf tables) and bitmaps resident in HBM: K times
             [glyf_decode_kernel + sdf_tiles_strided_kernel], timed with CUDA events on the launching stream, L2
             flushed between steps.
  e2e        glyphs/s through the reference-facing host API (FontManager.render_glyphs: parsed fonts in host memory ->
             PBF bytes in host memory): cmap / hmtx / loca lookups, glyph requests written to pinned memory (read by the
             device across PCIe), both kernels, bitmaps + frames written back across PCIe, PBF encode.
  roofline   the dominant kernel (SDF): FP32-ALU bound (north star).  `achieved` counts the flops of the kernel's own
             formulation — 3 flop per pixel x vertex + 11 flop per pixel x long segment — over the kernel's duration
             measured in this run (event recorded between the two kernels); `bruteforce_equivalent` restates the launch
             as SURVEY.md 8(d)'s 11 flop x pixel x segment pairs.  peak = FFMA-chain microbenchmark of this run.
  strong     C4 (BASELINE.json configs[3]: synthetic 63 487-glyph font) sharded over the N ranks the same way: kernel
             and e2e milliseconds of the slowest rank, per-rank times and the cost balance.
  cpu_baseline  the oracle (C port of the reference's CPU algorithm) on the box's host cores, run in a child process
             so that the product process never maps anything under oracle/.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FLOP_PER_PAIR = 11  # pax 1, t 2, qx 2, qy 2, d2 3, min 1 (brute force / long segments)
FLOP_PER_VERTEX_PAIR = 3  # fma(pax, pax, pay^2) 2, min 1
LONG_L2 = 0.25  # sdf_kernel.cuh kLongL2: squared length above which a segment takes the clamped projection
METRIC = "SDF glyphs/sec (24px, buffer 3)"
TESTDATA = os.path.join(ROOT, "testdata")
FIRA = os.path.join(TESTDATA, "Fira Sans - Regular.ttf")
NOTO_DIR = os.path.join(TESTDATA, "Noto Sans")
L2_NOTE = "flushed between timed steps (256 MiB memset outside the event pair)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="noto", choices=["noto", "fira", "dense", "c4"])
    ap.add_argument("--kernel-only", action="store_true", help="value leg only (ncu runs)")
    ap.add_argument("--no-strong", action="store_true", help="skip the C4 strong-scaling leg")
    ap.add_argument("--diag", action="store_true", help="also time the host-planned (sorted, one CTA per tile job) SDF kernel on the same records")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (default: min(steps, 20))")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads per rank in the e2e leg (default: the box's cores / ranks)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample budget")
    ap.add_argument("--cpu-child", action="store_true", help=argparse.SUPPRESS)
    return ap.parse_args()


# ---- workloads ---------------------------------------------------------------------------------------
def noto_paths():
    """C2 input order: byte-wise lexicographic path order."""
    names = sorted(os.listdir(NOTO_DIR), key=lambda s: s.encode())
    return [os.path.join(NOTO_DIR, n) for n in names if n.endswith(".ttf")]


def c4_font_bytes():
    """The C4 font is generated once per box (20 s of Python) and shared by the ranks through a file."""
    import synth_font

    path = os.path.join(tempfile.gettempdir(), "vgb_bench_c4_b200.ttf")
    if not os.path.exists(path):
        blob = synth_font.full_bmp_font()
        tmp = f"{path}.{os.getpid()}"
        open(tmp, "wb").write(blob)
        os.replace(tmp, path)
    return open(path, "rb").read()


def workload_fonts(name, copies=1):
    """-> (label, [(font name, [font bytes])]): `copies` fonts with the same files under different names."""
    import synth_font

    if name == "noto":
        label, base, blobs = "C2: testdata/Noto Sans (20 files) merge", "Noto Sans Regular", [open(p, "rb").read() for p in noto_paths()]
    elif name == "fira":
        label, base, blobs = "C1: testdata/Fira Sans - Regular.ttf", "Fira Sans - Regular", [open(FIRA, "rb").read()]
    elif name == "dense":
        label, base, blobs = "C3: synthetic dense outlines, 4096 glyphs", "Synth Dense", [synth_font.dense_font(4096)]
    else:
        label, base, blobs = "C4: synthetic full-BMP font, 63487 glyphs", "Synth Full", [c4_font_bytes()]
    return label, [(base if k == 0 else f"{base} {k + 1}", blobs) for k in range(copies)]


def config_of(label, copies, glyphs):
    """Identical in both arms (the driver compares the dicts)."""
    return {"workload": label, "fonts": copies, "glyphs": int(glyphs),
            "sharding": "one job, font x GlyphBlock tasks, LPT by estimated cost; rank r renders shard r" if copies > 1 else "none",
            "l2": L2_NOTE}


class ClockSampler:
    """SM clock + throttle reasons of this rank's GPU, read in-process through NVML at moments the caller chooses (while
    the timed launches are queued and running).  A polling `nvidia-smi -lms` child per rank — what round 1 used — takes
    driver locks at random moments: with several ranks on a box it showed up as 20-40 ms outlier steps of the e2e leg."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.rows = []
        self.h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001 — no NVML: the line reports no clocks rather than failing
            self.h = None

    def sample(self):
        if self.h is None:
            return
        try:
            sm = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            why = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.rows.append((sm, int(why)))
        except Exception:  # noqa: BLE001
            pass

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted(n for n, bit in self.REASONS.items() if any(w & bit for _, w in self.rows))
        return {"sm_mhz": statistics.median(r[0] for r in self.rows), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(self.rows), "how": "NVML, in-process, sampled while the timed launches were running"}


# ---- the reference arm / CPU baseline (the only code here that touches oracle/) -------------------------
def cpu_fontset(font_name, font_blobs):
    import oracle_lib as O

    paths = []
    for i, blob in enumerate(font_blobs):
        p = os.path.join(tempfile.gettempdir(), f"_bench_font_{os.getpid()}_{i}.ttf")
        open(p, "wb").write(blob)
        paths.append(p)
    fs = O.FontSet(font_name, paths)
    for p in paths:
        os.unlink(p)
    return fs


def cpu_stride(fs, seconds, threads):
    """Block stride that brings one pass of the oracle over the workload under `seconds`."""
    import oracle_lib as O

    pop = fs.block_population()
    t0 = time.perf_counter()
    probe = fs.render_all(O.MODE_PRECISE, threads=threads, stride=16)
    dt = time.perf_counter() - t0
    est_full = dt * sum(pop) / max(1, probe["glyphs"])
    stride = 1
    while est_full / stride > seconds and stride < 64:
        stride *= 2
    return stride, sum(pop)


def run_cpu_child(args):
    """Child process of the b200 arm: the CPU baseline numbers as one JSON object on stdout."""
    import oracle_lib as O

    label, fonts = workload_fonts(args.workload)
    threads = len(os.sched_getaffinity(0)) or 1
    fs = cpu_fontset(*fonts[0])
    stride, _ = cpu_stride(fs, args.cpu_seconds, threads)
    t0 = time.perf_counter()
    st = fs.render_all(O.MODE_PRECISE, threads=threads, stride=stride)
    dt = time.perf_counter() - t0
    out = {"value": st["glyphs"] / dt, "unit": "glyphs/s", "cores": threads, "kind": "port",
           "sample": f"every {stride}th GlyphBlock of {label} ({st['glyphs']} glyphs, {st['pairs']} pairs, {dt:.2f} s)",
           "built": "oracle/vg_oracle.c, gcc -O3 -ffp-contract=off; STR bulk-loaded R-tree (fan-out 6) stands in for rstar's OMT"}
    # the reference's hidden --single-thread flag (src/commands/recurse.rs:52-53) on a bounded sample
    s1, _ = cpu_stride(fs, min(6.0, args.cpu_seconds), 1)
    t0 = time.perf_counter()
    st1 = fs.render_all(O.MODE_PRECISE, threads=1, stride=s1)
    dt1 = time.perf_counter() - t0
    out["single_thread"] = {"value": st1["glyphs"] / dt1, "unit": "glyphs/s", "cores": 1,
                            "sample": f"every {s1}th GlyphBlock ({st1['glyphs']} glyphs, {dt1:.2f} s)"}
    # BASELINE.json configs[0]: Fira Sans `recurse` on the CPU reference, whole font
    _, ffonts = workload_fonts("fira")
    ffs = cpu_fontset(*ffonts[0])
    ffs.render_all(O.MODE_PRECISE, threads=threads, stride=4)
    t0 = time.perf_counter()
    fst = ffs.render_all(O.MODE_PRECISE, threads=threads)
    fdt = time.perf_counter() - t0
    out["c1_fira_recurse"] = {"value": fst["glyphs"] / fdt, "unit": "glyphs/s", "cores": threads, "seconds": fdt,
                              "sample": f"all {fst['glyphs']} glyphs of Fira Sans Regular"}
    sys.stdout.write(json.dumps(out) + "\n")


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Rust and cannot
    be built here (no cargo/rustc), so this is the oracle port, all host threads, one task per GlyphBlock
    (reference src/font/manager.rs:117-118), on a bounded sample of the b200 arm's workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O

    copies = max(1, args.gpus)
    label, fonts = workload_fonts(args.workload, copies)
    threads = len(os.sched_getaffinity(0)) or 1
    budget = max(0.05, 150.0 / max(1, args.steps + args.warmup))
    fs = cpu_fontset(*fonts[0])  # the `copies` fonts of the job are identical: one of them is the sample
    stride, glyphs_per_font = cpu_stride(fs, budget, threads)
    glyphs = 0
    for _ in range(args.warmup):
        fs.render_all(O.MODE_PRECISE, threads=threads, stride=stride)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = fs.render_all(O.MODE_PRECISE, threads=threads, stride=stride)
        glyphs += st["glyphs"]
    dt = time.perf_counter() - t0
    value = glyphs / dt
    sample = (f"every {stride}th GlyphBlock of one of the job's {copies} identical font(s) "
              f"({st['glyphs']} glyphs, {st['pairs']} pairs per step)")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "glyphs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "testdata fonts" if args.workload in ("noto", "fira") else "synthetic",
        "config": config_of(label, copies, glyphs_per_font * copies),
        "path": "oracle port of reference recurse/merge (f64, per-row crossing sort, +-8px R-tree filter), gcc -O3",
        "cpu_baseline": {"value": value, "unit": "glyphs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "glyphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries print to stdout (NCCL: "NCCL version ..."): keep fd 1 for the ONE JSON line, send the rest to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(result):
    sys.stdout.flush()
    line = (json.dumps(result) + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, line)
    else:
        os.write(_REAL_STDOUT, line)


# ---- the B200 arm -------------------------------------------------------------------------------------------
class Job:
    """One rank's share of a font set: the FontManager (whole job) and this shard's glyphs as one device-resident batch."""

    def __init__(self, V, np, torch, dev, renderer, fonts, rank, world):
        self.V, self.np, self.torch = V, np, torch
        self.manager = V.FontManager(parallel=True)
        entries = {}
        for name, blobs in fonts:
            for blob in blobs:
                self.manager.add_font_bytes_with_name(name, blob)
            entries[V.name_to_id(name)] = [V.FontFileEntry(data=b) for b in blobs]
        self.entries = entries  # keeps the faces (and their device-resident glyf tables) alive
        owners, loads = self.manager.shard_owners(world)
        self.loads = loads
        self.ctx = V.SdfContext.of_renderer(renderer)
        batch = renderer.new_batch()  # glyph-level batch: requests only, the device decodes
        for fi, fid in enumerate(self.manager.font_ids()):
            owner = {}
            for f in entries[fid]:
                for cp in f.codepoints().tolist():
                    if cp <= 0xFFFF and cp not in owner:
                        owner[cp] = f  # first file wins (reference src/font/glyph_block.rs:34-36)
            for cp in sorted(owner):
                if world == 1 or owners[fi, cp >> 8] == rank:
                    batch.add_glyph(owner[cp], cp)
        self.batch = batch
        self.n_glyphs = len(batch)
        self.reqs, self.parts = batch.requests(), batch.parts()
        self.hcurves, self.hsegs = batch.curves(), batch.segments().copy()
        self.curve_slots, self.tile_cap, self.est_cost = batch.curve_slots, batch.tile_cap, batch.est_cost
        self.out_bytes = int((self.reqs["out_off"] + self.reqs["out_cap"]).max()) if len(self.reqs) else 0

        def to_dev(a):
            return torch.from_numpy(np.frombuffer(a.tobytes() or b"\0" * 16, dtype=np.uint8).copy()).to(dev)

        self.d_reqs, self.d_parts = to_dev(self.reqs), to_dev(self.parts)
        self.d_hcurves, self.d_hsegs = to_dev(self.hcurves), to_dev(self.hsegs)
        self.d_frames = torch.zeros(max(1, len(self.reqs)) * 24, dtype=torch.uint8, device=dev)
        self.d_out = torch.zeros(max(self.out_bytes, 1), dtype=torch.uint8, device=dev)
        self.in_bytes = self.reqs.nbytes + self.parts.nbytes + self.hcurves.nbytes + self.hsegs.nbytes

    def launch(self, stream, mid_event=0):
        self.ctx.render_glyphs_device(self.d_reqs.data_ptr(), len(self.reqs), self.d_parts.data_ptr(), len(self.parts),
                                      self.d_hcurves.data_ptr(), len(self.hcurves), self.d_hsegs.data_ptr(), len(self.hsegs),
                                      self.curve_slots, self.tile_cap, self.est_cost, self.d_frames.data_ptr(), self.d_out.data_ptr(),
                                      self.out_bytes, stream.cuda_stream, mid_event)

    def frames(self):
        from versatiles_glyphs_rs_b200.api import GLYPH_FRAME_DT

        return self.np.frombuffer(self.d_frames.cpu().numpy().tobytes(), dtype=GLYPH_FRAME_DT)[: len(self.reqs)]

    def time_device(self, stream, flush, steps, warmup, sampler=None):
        """-> per-step (total ms, decode ms, sdf ms) lists, CUDA events on `stream`."""
        torch = self.torch
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                flush.zero_()
                self.launch(stream)
        torch.cuda.synchronize()
        evs = []
        with torch.cuda.stream(stream):
            for _ in range(steps):
                flush.zero_()  # L2 flush between timed steps (not inside the event pair)
                e0, em, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                em.record(stream)  # creates the underlying CUDA event, so that its handle can be passed down
                e0.record(stream)
                self.launch(stream, em.cuda_event)
                e1.record(stream)
                evs.append((e0, em, e1))
                if sampler is not None and len(evs) % 4 == 2:
                    sampler.sample()  # (the launches are asynchronous: the GPU is busy with the queued steps)
        torch.cuda.synchronize()
        return [a.elapsed_time(c) for a, _, c in evs], [a.elapsed_time(b) for a, b, _ in evs], [b.elapsed_time(c) for _, b, c in evs]

    def file_hashes(self, renderer, rank, world, threads):
        w = self.V.Writer.new_memory()
        self.manager.render_glyphs(w, renderer, shard=rank, n_shards=world, threads=threads)
        return {n: hashlib.sha1(d).hexdigest() for n, is_dir, d in w.entries() if not is_dir}


def main():
    args = parse_args()
    if args.cpu_child:
        return run_cpu_child(args)
    _quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    torch.set_num_threads(1)  # torch is plumbing here: no intra-op pool competing with the pipeline's workers for cores

    import versatiles_glyphs_rs_b200 as V
    from versatiles_glyphs_rs_b200 import _native as N

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SDF path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        cpus_before = os.sched_getaffinity(0)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()  # NCCL communicator is built here
        # NCCL pins the calling thread to the GPU's NUMA-local cores while it initialises; threads created later
        # (the pipeline's workers) would inherit a narrowed mask.  Restore what the launcher gave this process.
        cpus_after = os.sched_getaffinity(0)
        if cpus_after != cpus_before:
            os.sched_setaffinity(0, cpus_before)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def gather(obj):
        if dist is None:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    host_threads = args.host_threads or max(1, len(os.sched_getaffinity(0)) // world)  # the ranks of one box share its host cores
    label, fonts = workload_fonts(args.workload, world)
    renderer = V.Renderer.new_precise(device=local_rank)
    job = Job(V, np, torch, dev, renderer, fonts, rank, world)
    ctx = job.ctx
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)
    fp32_peak_tflops, _ = ctx.measure_fp32_peak(5)
    fp32x2_peak_tflops, _ = ctx.measure_fp32_peak(5, packed=True)

    # ---- value: the device-resident job (decode + SDF kernels) ----
    sampler = ClockSampler(local_rank)
    warm = max(3, args.warmup)
    job.time_device(stream, flush, 0, warm)
    barrier()
    l0 = ctx.launch_count
    t_wall0 = time.perf_counter()
    total_ms, decode_ms, sdf_ms = job.time_device(stream, flush, args.steps, 0, sampler)
    barrier()
    t_wall1 = time.perf_counter()
    launches = ctx.launch_count - l0
    ms_per_step = max_over_ranks(sum(total_ms)) / args.steps
    frames = job.frames()
    if (frames["status"] >= N.GLYPH_NEEDS_HOST).any():
        raise SystemExit(f"rank {rank}: the device handed back / rejected {(frames['status'] >= N.GLYPH_NEEDS_HOST).sum()} glyph(s) of the "
                         "device-resident job: the timed work would be incomplete")
    okf = frames["status"] == N.GLYPH_OK
    n_bitmaps = int(okf.sum())
    bitmaps_all = int(sum_over_ranks(n_bitmaps))
    glyphs_all = int(sum_over_ranks(job.n_glyphs))
    pixels = int((frames["width"][okf].astype(np.int64) * frames["height"][okf]).sum())
    pairs = int((frames["width"][okf].astype(np.int64) * frames["height"][okf] * frames["seg_cnt"][okf]).sum())
    value = bitmaps_all / (ms_per_step * 1e-3)
    kernel_s = statistics.mean(sdf_ms) * 1e-3
    # checksum of what was just timed (bitmap bytes of every rendered glyph)
    out_host = job.d_out.cpu().numpy()
    checksum = 0
    for i in np.flatnonzero(okf):
        o = int(job.reqs["out_off"][i])
        checksum += int(out_host[o : o + int(frames["width"][i]) * int(frames["height"][i])].sum(dtype=np.int64))

    # executed flops of the SDF kernel's formulation (outside the timed region, from the device's own decoding and
    # flattening): per glyph W*H * (3 * vertices + 11 * long segments); SEGMENTS glyphs stage both end points
    dframes, djobs, dcurves, dev_tiles = ctx.decode_glyphs(job.reqs, job.parts, job.curve_slots, curves=job.hcurves, n_seg=len(job.hsegs),
                                                          est_cost=job.est_cost)
    assert np.array_equal(dframes["status"], frames["status"])
    is_curves = (djobs["kind"] == N.KIND_CURVES) & okf
    is_segs = (djobs["kind"] == N.KIND_SEGMENTS) & okf
    cj = djobs[is_curves]
    flat = ctx.flatten_outlines(dcurves, cj) if len(cj) else np.zeros((0, 4), np.float32)
    l2 = (flat[:, 2] - flat[:, 0]) ** 2 + (flat[:, 3] - flat[:, 1]) ** 2
    csum = np.concatenate([[0], np.cumsum((l2 > LONG_L2).astype(np.int64))])
    seg_cnt = djobs["seg_cnt"].astype(np.int64)
    px = djobs["width"].astype(np.int64) * djobs["height"].astype(np.int64)
    starts = np.concatenate([[0], np.cumsum(seg_cnt[is_curves])])[:-1]
    long_per_job = np.zeros(len(djobs), np.int64)
    long_per_job[is_curves] = csum[starts + seg_cnt[is_curves]] - csum[starts]
    if is_segs.any():
        sl2 = (job.hsegs[:, 2] - job.hsegs[:, 0]) ** 2 + (job.hsegs[:, 3] - job.hsegs[:, 1]) ** 2
        scs = np.concatenate([[0], np.cumsum((sl2 > LONG_L2).astype(np.int64))])
        so = djobs["src_off"].astype(np.int64)
        long_per_job[is_segs] = scs[(so + seg_cnt)[is_segs]] - scs[so[is_segs]]
    vertices = np.where(is_segs, 2 * seg_cnt, seg_cnt)
    vertex_pairs = int((px * vertices)[okf].sum())
    long_pairs = int((px * long_per_job)[okf].sum())
    executed_flop = FLOP_PER_VERTEX_PAIR * vertex_pairs + FLOP_PER_PAIR * long_pairs
    achieved_tflops = executed_flop / kernel_s / 1e12
    equivalent_tflops = FLOP_PER_PAIR * pairs / kernel_s / 1e12
    n_tiles = int(len(dev_tiles))
    curve_records = int(djobs["src_cnt"][is_curves].sum())
    # bytes the SDF kernel has to move per launch: curve records + job / tile records in, bitmaps out
    alg_bytes = curve_records * 32 + len(job.hsegs) * 16 + len(djobs) * 56 + n_tiles * 32 + pixels
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    traffic, traffic_source = None, None  # DRAM bytes per launch of the SDF kernel from this round's ncu --set full capture
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if t.get("workload") == args.workload and world == 1:
            traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
            traffic_source = "profiles/r02_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this command"
    except (OSError, ValueError, KeyError):
        pass

    t_e2e_end = t_wall1
    host_planned_ms = device_plan_ms = None
    if args.diag:
        # the same curve records through the other kernel form: tile jobs planned and sorted on the host, one CTA each
        okjobs = np.ascontiguousarray(djobs[okf])
        tiles_h, n_tiles_h, _ = ctx.plan_outline_tiles(okjobs, len(dcurves), len(job.hsegs), job.out_bytes)
        d_c = torch.from_numpy(np.frombuffer(dcurves.tobytes(), dtype=np.uint8).copy()).to(dev)
        d_j = torch.from_numpy(np.frombuffer(okjobs.tobytes(), dtype=np.uint8).copy()).to(dev)
        d_t = torch.from_numpy(tiles_h).to(dev)
        evs = []
        with torch.cuda.stream(stream):
            for i in range(23):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                ctx.render_outlines_device(d_c.data_ptr(), job.d_hsegs.data_ptr(), d_j.data_ptr(), d_t.data_ptr(), n_tiles_h,
                                           job.d_out.data_ptr(), stream.cuda_stream)
                b.record(stream)
                evs.append((a, b))
        torch.cuda.synchronize()
        host_planned_ms = statistics.mean(a.elapsed_time(b) for a, b in evs[3:])
        # ... and the device-planned tile list (claim order) through that one-CTA-per-job kernel: separates what the plan
        # costs from what the persistent claim loop costs
        d_ja = torch.from_numpy(np.frombuffer(np.ascontiguousarray(djobs).tobytes(), dtype=np.uint8).copy()).to(dev)
        d_ta = torch.from_numpy(np.frombuffer(np.ascontiguousarray(dev_tiles).tobytes(), dtype=np.uint8).copy()).to(dev)
        evs = []
        with torch.cuda.stream(stream):
            for i in range(23):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                ctx.render_outlines_device(d_c.data_ptr(), job.d_hsegs.data_ptr(), d_ja.data_ptr(), d_ta.data_ptr(), len(dev_tiles),
                                           job.d_out.data_ptr(), stream.cuda_stream)
                b.record(stream)
                evs.append((a, b))
        torch.cuda.synchronize()
        device_plan_ms = statistics.mean(a.elapsed_time(b) for a, b in evs[3:])
    config = config_of(label, world, glyphs_all)
    result = {
        "metric": METRIC, "value": value, "unit": "glyphs/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "testdata fonts" if args.workload in ("noto", "fira") else "synthetic",
        "config": config,
        "detail": {
            "glyphs_this_rank": job.n_glyphs, "bitmaps": n_bitmaps, "bitmaps_all_ranks": bitmaps_all, "pixels": pixels, "pairs": pairs,
            "segments": int(seg_cnt[okf].sum()), "curve_records": curve_records, "host_recorded_glyphs": int((job.reqs["kind"] != N.KIND_GLYF).sum()),
            "tile_jobs": n_tiles, "bitmap_checksum": checksum,
            "value_counts": "glyphs with a bitmap, all ranks; the timed step = glyf_decode_kernel + sdf_tiles_strided_kernel over this rank's shard",
            "decode_kernel_ms": statistics.mean(decode_ms), "sdf_kernel_ms": statistics.mean(sdf_ms), "step_ms_this_rank": statistics.mean(total_ms),
            "step_ms_best_median_this_rank": [min(total_ms), statistics.median(total_ms)],
            "pixels_per_s_this_rank": pixels / (statistics.mean(total_ms) * 1e-3),
            "resident_in_hbm": "glyf tables of the fonts (uploaded when first used, outside the timed region), glyph requests, bitmaps",
            "shard_cost_estimates": [int(x) for x in job.loads], "host_planned_sdf_kernel_ms": host_planned_ms,
            "device_plan_in_one_cta_per_job_kernel_ms": device_plan_ms,
        },
        "roofline": {
            "bound": "fp32", "kernel": "sdf_tiles_strided_kernel", "achieved": achieved_tflops, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
            "frac": achieved_tflops / fp32_peak_tflops, "traffic": traffic, "traffic_source": traffic_source,
            "peak_source": "FFMA-chain microbenchmark in this run (b200sdf_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
            "flop_model": "3 flop x pixel x vertex (FFMA + min) + 11 flop x pixel x long segment (|d| > 0.5 px); band "
                          "rasterisation and staging not credited",
            "vertex_pairs_per_launch": vertex_pairs, "long_segment_pairs_per_launch": long_pairs,
            "flop_per_launch": int(executed_flop),
            "bruteforce_equivalent": {"flop_per_pair": FLOP_PER_PAIR, "tflops": equivalent_tflops,
                                      "frac_of_peak": equivalent_tflops / fp32_peak_tflops,
                                      "note": "SURVEY.md 8(d) accounting: what a brute-force pixel x segment kernel would "
                                              "have to sustain to match this launch time"},
            "pairs_per_launch": int(pairs), "kernel_ms": kernel_s * 1e3,
            "peak_ffma2_tflops": fp32x2_peak_tflops,
            "pairs_per_s": pairs / kernel_s,
            "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / kernel_s / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs"), "peak_source": "MEASURED_PEAKS.json (measured copy)"},
        },
        "gpu_launches": int(launches),
    }

    if not args.kernel_only:
        import gc

        # ---- e2e: the reference-facing host API, host buffers in, PBF bytes out ----
        e2e_steps = args.e2e_steps or min(args.steps, 20)
        e2e_warmup = max(8, args.warmup)  # pooled pinned buffers reach their steady-state count and size during the first calls

        def e2e_loop(steps, threads, make_writer):
            ms = []
            st = None
            for k in range(steps):
                w = make_writer()
                if os.environ.get("VGB_ALLOC_TRACE"):
                    print(f"--- rank {rank} step {k} of {steps}", file=sys.stderr, flush=True)
                ts = time.perf_counter()
                st = job.manager.render_glyphs(w, renderer, shard=rank, n_shards=world, threads=threads)
                ms.append(1e3 * (time.perf_counter() - ts))
            return ms, st

        e2e_loop(e2e_warmup, host_threads, V.Writer.new_memory)
        gc.collect()
        gc.disable()  # a generational collection of the interpreter's heap (torch, numpy, ...) is a 50 ms pause between steps
        e2e_loop(2, host_threads, V.Writer.new_memory)  # (untimed: the collection above let the worker threads go to sleep)
        barrier()
        t0 = time.perf_counter()
        step_ms, st = e2e_loop(e2e_steps, host_threads, V.Writer.new_memory)
        t_loop = time.perf_counter() - t0
        gc.enable()
        barrier()
        t_e2e_end = time.perf_counter()
        e2e_s = max_over_ranks(t_loop)
        slow = [(i, round(x, 2)) for i, x in enumerate(step_ms) if x > 1.5 * statistics.median(step_ms)]
        print(f"[bench] rank {rank}: e2e loop {1e3 * t_loop:.1f} ms, steps min/median/max "
              f"{min(step_ms):.2f}/{statistics.median(step_ms):.2f}/{max(step_ms):.2f} ms, {st.workers} workers, {st.submits} submits"
              + (f", steps over 1.5 x median (index, ms): {slow}" if slow else ""), file=sys.stderr)
        per_rank = gather({"rank": rank, "median_ms": statistics.median(step_ms), "max_ms": max(step_ms), "glyphs": st.glyphs,
                           "cost": st.cost_shard})
        result["e2e"] = {
            "value": glyphs_all * e2e_steps / e2e_s, "unit": "glyphs/s",
            "h2d_bytes_per_step": int(st.h2d_bytes), "d2h_bytes_per_step": int(st.pixels + 24 * n_bitmaps),
            "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps, "warmup": e2e_warmup, "pbf_bytes_per_step": int(st.pbf_bytes),
            "api": "FontManager.render_glyphs(Writer.new_memory(), Renderer.new_precise(), shard=rank, n_shards=world): glyph requests in "
                   "pinned host memory -> glyf decode + SDF kernels (read them, write frames and bitmaps over PCIe) -> PBF",
            "step_ms_min_median_max": [min(step_ms), statistics.median(step_ms), max(step_ms)],
            "host_threads_per_rank": host_threads, "step_ms": [round(x, 3) for x in step_ms],
            "per_rank": per_rank,
            "rank_time_max_over_mean": max(p["median_ms"] for p in per_rank) / statistics.mean(p["median_ms"] for p in per_rank),
            "host_phases_ms_last_step": {
                "workers": st.workers, "submits": st.submits, "wall": st.wall_ns / 1e6, "handed_back": st.handed_back,
                "request_per_worker": st.outline_ns / 1e6 / max(1, st.workers), "submit_per_worker": st.submit_ns / 1e6 / max(1, st.workers),
                "wait_per_worker": st.wait_ns / 1e6 / max(1, st.workers), "encode_per_worker": st.encode_ns / 1e6 / max(1, st.workers),
                "write_per_worker": st.write_ns / 1e6 / max(1, st.workers),
            },
            "cached_across_calls": "per font set: the 256-block table and its cost estimates (pure functions of the files, built at first "
                                   "use like the parsed cmap); glyf tables resident in HBM; pooled pinned buffers.  No outline, frame or "
                                   "bitmap is kept from one call to the next",
        }
        # the other protocol lines of SURVEY.md 8(d): files written to a directory, and the reference's --single-thread
        few = max(3, min(8, e2e_steps))
        tmp = tempfile.mkdtemp(prefix="vgb_bench_")
        io_ms, _ = e2e_loop(few, host_threads, lambda: V.Writer.new_file(tmp))
        e2e_loop(3, 1, V.Writer.new_memory)  # (a different thread count means different batch sizes: let the pools settle first)
        one_ms, _ = e2e_loop(few, 1, V.Writer.new_memory)
        import shutil

        shutil.rmtree(tmp, ignore_errors=True)
        result["e2e"]["incl_file_io_ms_per_step"] = max_over_ranks(statistics.median(io_ms))
        result["e2e"]["single_thread_ms_per_step"] = max_over_ranks(statistics.median(one_ms))

        # ---- the shards together are the job: rank 0 compares the gathered files with an unsharded run ----
        mine = job.file_hashes(renderer, rank, world, host_threads)
        allh = gather(mine)
        if rank == 0:
            union = {}
            for h in allh:
                assert not set(h) & set(union), "shards overlap"
                union.update(h)
            whole = job.file_hashes(renderer, 0, 1, host_threads) if world > 1 else mine
            assert union == whole, "gathered shard output differs from the unsharded run"
            digest = hashlib.sha1("".join(f"{k}:{union[k]}" for k in sorted(union)).encode()).hexdigest()
            result["e2e"]["output_check"] = {"files": len(union), "sha1_of_file_sha1s": digest,
                                             "equals_unsharded_run": True if world > 1 else "n/a (one shard)"}

        # ---- strong scaling: C4 cut into `world` shards ----
        if not args.no_strong and args.workload == "noto":
            if rank == 0:
                c4_font_bytes()
            barrier()
            _, c4fonts = workload_fonts("c4")
            c4 = Job(V, np, torch, dev, renderer, c4fonts, rank, world)
            sk, _, ssdf = c4.time_device(stream, flush, 5, 2)
            c4f = c4.frames()
            assert not (c4f["status"] >= N.GLYPH_NEEDS_HOST).any()
            c4ok = c4f["status"] == N.GLYPH_OK
            c4pairs = int((c4f["width"][c4ok].astype(np.int64) * c4f["height"][c4ok] * c4f["seg_cnt"][c4ok]).sum())
            for _ in range(8):  # (pooled pinned buffers reach their steady-state size)
                c4.manager.render_glyphs(V.Writer.new_memory(), renderer, shard=rank, n_shards=world, threads=host_threads)
            barrier()
            c4ms = []
            for _ in range(5):
                barrier()
                ts = time.perf_counter()
                cst = c4.manager.render_glyphs(V.Writer.new_memory(), renderer, shard=rank, n_shards=world, threads=host_threads)
                c4ms.append(1e3 * (time.perf_counter() - ts))
            ranks = gather({"rank": rank, "glyphs": cst.glyphs, "pairs": c4pairs, "kernel_ms": statistics.median(sk), "e2e_ms": statistics.median(c4ms),
                            "cost": int(c4.loads[rank]) if world > 1 else int(c4.loads[0])})
            c4hash = gather(c4.file_hashes(renderer, rank, world, host_threads))
            strong = {
                "workload": "C4: synthetic full-BMP font, 63487 glyphs, one job cut into n_gpus shards (scaling: strong)",
                "glyphs": int(sum(r["glyphs"] for r in ranks)), "pairs": int(sum(r["pairs"] for r in ranks)),
                "kernel_ms": max(r["kernel_ms"] for r in ranks), "e2e_ms": max(r["e2e_ms"] for r in ranks),
                "kernel_glyphs_per_s": sum(r["glyphs"] for r in ranks) / (max(r["kernel_ms"] for r in ranks) * 1e-3),
                "e2e_glyphs_per_s": sum(r["glyphs"] for r in ranks) / (max(r["e2e_ms"] for r in ranks) * 1e-3),
                "kernel_time_max_over_mean": max(r["kernel_ms"] for r in ranks) / statistics.mean(r["kernel_ms"] for r in ranks),
                "e2e_time_max_over_mean": max(r["e2e_ms"] for r in ranks) / statistics.mean(r["e2e_ms"] for r in ranks),
                "cost_estimate_max_over_mean": max(r["cost"] for r in ranks) / statistics.mean(r["cost"] for r in ranks),
                "pairs_max_over_mean": max(r["pairs"] for r in ranks) / statistics.mean(r["pairs"] for r in ranks),
                "per_rank": ranks,
            }
            if rank == 0:
                union = {}
                for h in c4hash:
                    assert not set(h) & set(union), "C4 shards overlap"
                    union.update(h)
                strong["output_check"] = {"files": len(union),
                                          "sha1_of_file_sha1s": hashlib.sha1("".join(f"{k}:{union[k]}" for k in sorted(union)).encode()).hexdigest()}
                if world > 1:
                    whole = c4.file_hashes(renderer, 0, 1, host_threads)
                    assert union == whole, "gathered C4 shard output differs from the unsharded run"
                    strong["output_check"]["equals_unsharded_run"] = True
            result["strong"] = strong
            del c4

        # ---- CPU baseline: a child process on rank 0 (the product process never maps anything under oracle/) ----
        barrier()
        if rank == 0:
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-child", "--workload", args.workload,
                                    "--cpu-seconds", str(args.cpu_seconds)], capture_output=True, text=True, timeout=300)
                result["cpu_baseline"] = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
            except Exception as e:  # noqa: BLE001 — a baseline that cannot run must not hide the measured line
                result["cpu_baseline"] = {"value": None, "unit": "glyphs/s", "cores": 0, "kind": "port", "sample": f"failed: {e!r}"}
        barrier()
    result["clocks"] = sampler.summary()
    if rank == 0:
        emit(result)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
