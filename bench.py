#!/usr/bin/env python
"""bench.py — headline benchmark of the SDF glyph rendering path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # the B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

One "step" = one pass of the hot path over one batch: the whole workload (default: BASELINE.json
configs[1], the 20-file Noto Sans `merge`, 6480 glyphs / 6445 bitmaps / 3.96 M segments /
2.67 G pixel x segment pairs) rendered once.

  value      glyphs/s with segments, tile jobs and bitmaps resident in HBM: K kernel launches, each
             timed with CUDA events on the launching stream; L2 is flushed between launches.
  e2e        glyphs/s through the reference-facing host API (FontManager.render_glyphs: parsed fonts
             in host memory -> PBF bytes in host memory): outline recording, host->device transfer of
             the records, kernel, device->host transfer of the bitmaps, PBF encode.  The transfers are
             in the timed region; with pinned buffers they are zero-copy (the kernel reads the records
             and writes the bitmaps across PCIe itself), h2d/d2h_bytes_per_step are those bytes.
  roofline   FP32-ALU bound (north star).  The kernel computes min over segments as
             min(vertices) + band interiors of short segments + clamped projection of long segments
             (DESIGN.md §3.3), so `achieved` counts the flops of THAT formulation:
             3 flop per pixel x vertex (one FFMA + one min) + 11 flop per pixel x long segment, over kernel
             time; `bruteforce_equivalent` restates the same launch as SURVEY.md §8(d)'s 11 flop x
             pixel x segment pairs (what a brute-force kernel would have to sustain for this time).
             peak = FFMA-chain microbenchmark measured in this run (MEASURED_PEAKS.json has no FP32
             figure); HBM figures for the same launch are reported beside it.
  cpu_baseline  the oracle (C port of the reference's CPU algorithm) on the box's host cores.

Under torchrun each rank renders the same workload on its own GPU (font x block shards are
independent: weak scaling, no collective on the data path); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FLOP_PER_PAIR = 11  # pax 1, t 2, qx 2, qy 2, d2 3, min 1 (brute force / long segments)
FLOP_PER_VERTEX_PAIR = 3  # fma(pax, pax, pay^2) 2, min 1
LONG_L2 = 0.25  # sdf_kernel.cuh kLongL2: squared length above which a segment takes the clamped projection
METRIC = "SDF glyphs/sec (24px, buffer 3)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="noto", choices=["noto", "fira", "dense", "c4"])
    ap.add_argument("--kernel-only", action="store_true", help="skip e2e and cpu_baseline (ncu runs)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (default: min(steps, 20))")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline sample budget")
    return ap.parse_args()


# ---- workloads ---------------------------------------------------------------------------------------
def workload_fonts(name):
    """-> (display name, font name, [font bytes])"""
    import oracle_lib as O  # only for the fixture paths (no oracle code runs here)
    import synth_font

    if name == "noto":
        return "C2: testdata/Noto Sans (20 files) merge", "Noto Sans Regular", [open(p, "rb").read() for p in O.noto_paths()]
    if name == "fira":
        return "C1: testdata/Fira Sans - Regular.ttf", "Fira Sans - Regular", [open(O.FIRA, "rb").read()]
    if name == "dense":
        return "C3: synthetic dense outlines, 4096 glyphs", "Synth Dense", [synth_font.dense_font(4096)]
    return "C4: synthetic full-BMP font, 63487 glyphs", "Synth Full", [synth_font.full_bmp_font()]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        return [r for t, r in self.rows if t0 <= t <= t1]

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    @staticmethod
    def summary(rows):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(font_name, font_blobs, seconds, threads):
    """Oracle (port of the reference's CPU path) on a bounded, strided sample of the workload's blocks."""
    import oracle_lib as O

    paths = []
    for i, blob in enumerate(font_blobs):
        p = f"/tmp/_bench_font_{os.getpid()}_{i}.ttf"
        open(p, "wb").write(blob)
        paths.append(p)
    fs = O.FontSet(font_name, paths)
    for p in paths:
        os.unlink(p)
    pop = fs.block_population()
    t0 = time.perf_counter()
    probe = fs.render_all(O.MODE_PRECISE, threads=threads, stride=16)
    dt = time.perf_counter() - t0
    est_full = dt * sum(pop) / max(1, probe["glyphs"])
    stride = 1
    while est_full / stride > seconds and stride < 64:
        stride *= 2
    return fs, stride


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Rust and cannot
    be built here (no cargo/rustc), so this is the oracle port, all host threads, one task per GlyphBlock
    (reference src/font/manager.rs:117-118)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O

    label, font_name, blobs = workload_fonts(args.workload)
    threads = len(os.sched_getaffinity(0)) or 1
    budget = max(0.05, 150.0 / max(1, args.steps + args.warmup))
    fs, stride = cpu_sample(font_name, blobs, budget, threads)
    glyphs = 0
    for _ in range(args.warmup):
        fs.render_all(O.MODE_PRECISE, threads=threads, stride=stride)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = fs.render_all(O.MODE_PRECISE, threads=threads, stride=stride)
        glyphs += st["glyphs"]
    dt = time.perf_counter() - t0
    value = glyphs / dt
    sample = f"every {stride}th GlyphBlock of the workload ({st['glyphs']} glyphs, {st['pairs']} pairs per step)"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "glyphs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "testdata fonts" if args.workload in ("noto", "fira") else "synthetic",
        "config": {"workload": label, "path": "oracle port of reference recurse/merge (f64, per-row crossing sort, +-8px R-tree filter)"},
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


def main():
    args = parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    torch.set_num_threads(1)  # torch is plumbing here: no intra-op pool competing with the pipeline's workers for cores

    import versatiles_glyphs_rs_b200 as V

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
        print(f"[bench] rank {rank}: cpu affinity {len(cpus_before)} cpus at start, {len(cpus_after)} after NCCL init",
              file=sys.stderr)

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

    label, font_name, blobs = workload_fonts(args.workload)
    font_id = V.name_to_id(font_name)

    # ---- build the workload on the host (not timed): one flat batch of every glyph ----
    manager = V.FontManager(parallel=True)
    for blob in blobs:
        manager.add_font_bytes_with_name(font_name, blob)
    fonts = [V.FontFileEntry(data=b) for b in blobs]
    owner = {}
    for f in fonts:
        for cp in f.codepoints().tolist():
            if cp <= 0xFFFF and cp not in owner:
                owner[cp] = f  # first file wins (reference src/font/glyph_block.rs:34-36)
    renderer = V.Renderer.new_precise(device=local_rank)
    ctx = V.SdfContext(device=local_rank, n_slots=1)
    batch = renderer.new_batch()
    for cp in sorted(owner):
        batch.add_glyph(owner[cp], cp)
    n_glyphs = len(batch)
    segs = batch.segments().copy()  # only glyphs that cannot be flattened exactly on the device (scaled composites)
    curves = batch.curves()
    jobs = batch.jobs()
    n_bitmaps = len(jobs)
    n_segments = batch.total_segments
    out_bytes = int(jobs["out_off"][-1] + jobs["width"][-1].astype(np.uint64) * jobs["height"][-1]) if n_bitmaps else 0
    tiles, n_tiles, pairs = ctx.plan_outline_tiles(jobs, len(curves), len(segs), out_bytes)

    def to_dev(a):
        return torch.from_numpy(np.frombuffer(a.tobytes() or b"\0" * 16, dtype=np.uint8).copy()).to(dev)

    d_segs = to_dev(segs)
    d_curves = to_dev(curves)
    d_jobs = to_dev(jobs)
    d_tiles = torch.from_numpy(tiles).to(dev)
    d_out = torch.zeros(max(out_bytes, 1), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)

    fp32_peak_tflops, _ = ctx.measure_fp32_peak(5)
    fp32x2_peak_tflops, _ = ctx.measure_fp32_peak(5, packed=True)

    def launch():
        ctx.render_outlines_device(d_curves.data_ptr(), d_segs.data_ptr(), d_jobs.data_ptr(), d_tiles.data_ptr(), n_tiles,
                                   d_out.data_ptr(), stream.cuda_stream)

    # ---- value: device-resident kernel time ----
    sampler = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            flush.zero_()
            launch()
    barrier()
    l0 = ctx.launch_count
    evs = []
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(args.steps):
            flush.zero_()  # L2 flush between timed launches (not inside the event pair)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            launch()
            e1.record(stream)
            evs.append((e0, e1))
    barrier()
    t_wall1 = time.perf_counter()
    launches = ctx.launch_count - l0
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = max_over_ranks(sum(kernel_ms))
    ms_per_step = total_ms / args.steps

    # parity spot check of what was just timed (metrics are host-side; bitmap checksum must be stable)
    checksum = int(d_out[:out_bytes].to(torch.int64).sum().item())

    value = world * n_bitmaps / (ms_per_step * 1e-3)
    kernel_s = statistics.mean(kernel_ms) * 1e-3
    # executed flops of the kernel's formulation (computed outside the timed region, from the device's own
    # flattening): per glyph W*H * (3 * vertices + 11 * long segments); SEGMENTS glyphs stage both end points
    flat = ctx.flatten_outlines(curves, jobs) if len(curves) else np.zeros((0, 4), np.float32)
    l2 = (flat[:, 2] - flat[:, 0]) ** 2 + (flat[:, 3] - flat[:, 1]) ** 2
    is_curves = (jobs["kind"] == 0) if n_bitmaps else np.zeros(0, bool)  # B200SDF_KIND_CURVES
    seg_cnt = jobs["seg_cnt"].astype(np.int64)
    px = jobs["width"].astype(np.int64) * jobs["height"].astype(np.int64)
    starts = np.concatenate([[0], np.cumsum(np.where(is_curves, seg_cnt, 0))])[:-1]
    long_per_job = np.zeros(n_bitmaps, np.int64)
    long_mask = (l2 > LONG_L2).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(long_mask)])
    long_per_job[is_curves] = (csum[(starts + seg_cnt)[is_curves]] - csum[starts[is_curves]])
    if len(segs) and (~is_curves).any():  # host-flattened glyphs: classify their uploaded segments
        sl2 = (segs[:, 2] - segs[:, 0]) ** 2 + (segs[:, 3] - segs[:, 1]) ** 2
        scs = np.concatenate([[0], np.cumsum((sl2 > LONG_L2).astype(np.int64))])
        so = jobs["src_off"].astype(np.int64)
        long_per_job[~is_curves] = (scs[(so + seg_cnt)[~is_curves]] - scs[so[~is_curves]])
    vertices = np.where(is_curves, seg_cnt, 2 * seg_cnt)
    vertex_pairs = int((px * vertices).sum())
    long_pairs = int((px * long_per_job).sum())
    executed_flop = FLOP_PER_VERTEX_PAIR * vertex_pairs + FLOP_PER_PAIR * long_pairs
    achieved_tflops = executed_flop / kernel_s / 1e12
    equivalent_tflops = FLOP_PER_PAIR * pairs / kernel_s / 1e12
    in_bytes = len(segs) * 16 + len(curves) * 32 + len(jobs) * 56 + n_tiles * 32
    alg_bytes = in_bytes + out_bytes
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    traffic = None  # DRAM bytes per launch of the SDF kernel from the committed ncu --set full capture (same workload only)
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if t.get("workload") == args.workload:
            traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    except (OSError, ValueError, KeyError):
        pass

    result = {
        "metric": METRIC, "value": value, "unit": "glyphs/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "testdata fonts" if args.workload in ("noto", "fira") else "synthetic",
        "config": {
            "workload": label, "glyphs": n_glyphs, "bitmaps": n_bitmaps, "segments": int(n_segments), "curve_records": int(len(curves)), "host_flattened_segments": int(len(segs)), "pixels": out_bytes,
            "pairs": int(pairs), "ctas": int(n_tiles), "l2": "flushed between timed launches (256 MiB memset)",
            "value_counts": "glyphs with a bitmap; per rank the same workload (font x block shards are independent)",
            "bitmap_checksum": checksum,
        },
        "roofline": {
            "bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
            "frac": achieved_tflops / fp32_peak_tflops, "traffic": traffic,
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

    # ---- e2e: the reference-facing host API, host buffers in, PBF bytes out ----
    if not args.kernel_only:
        e2e_steps = args.e2e_steps or min(args.steps, 20)
        # the ranks of one box share its host cores: give each rank its share instead of oversubscribing
        # (threads = the rank's total, the calling thread included: with >= 14 the caller is the pipeline's dedicated CUDA
        # thread and the rest record / encode; with fewer every thread works and whoever is free talks to CUDA)
        host_threads = max(1, len(os.sched_getaffinity(0)) // world)
        # untimed calls: the pipeline's pooled pinned buffers reach their steady-state count and size during the first
        # calls (batch composition varies from call to call); a pinned allocation inside a timed step stalls the CUDA
        # context, for 200 ms when eight ranks allocate on one host
        e2e_warmup = max(8, args.warmup)
        for _ in range(e2e_warmup):
            manager.render_glyphs(V.Writer.new_memory(), renderer, threads=host_threads)
        import gc

        gc.collect()
        gc.disable()  # a generational collection of the interpreter's heap (torch, numpy, ...) is a 50 ms pause between steps
        barrier()
        t0 = time.perf_counter()
        step_ms = []
        for _ in range(e2e_steps):
            ts = time.perf_counter()
            st = manager.render_glyphs(V.Writer.new_memory(), renderer, threads=host_threads)
            step_ms.append(1e3 * (time.perf_counter() - ts))
        t_loop = time.perf_counter() - t0
        gc.enable()
        tb = time.perf_counter()
        barrier()
        print(f"[bench] rank {rank}: e2e loop {1e3 * t_loop:.1f} ms, closing barrier {1e3 * (time.perf_counter() - tb):.1f} ms, "
              f"steps min/median/max {min(step_ms):.2f}/{statistics.median(step_ms):.2f}/{max(step_ms):.2f} ms", file=sys.stderr)
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        result["e2e"] = {
            "value": world * st.glyphs * e2e_steps / e2e_s, "unit": "glyphs/s",
            "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(st.pixels),
            "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps, "warmup": e2e_warmup, "pbf_bytes_per_step": int(st.pbf_bytes),
            "api": "FontManager.render_glyphs(Writer.new_memory(), Renderer.new_precise()): outline records in pinned host memory -> "
                   "flatten+SDF kernel (reads them and writes the bitmaps over PCIe) -> PBF",
            "step_ms_min_median_max": [min(step_ms), statistics.median(step_ms), max(step_ms)],
            "host_threads_per_rank": host_threads, "step_ms": [round(x, 3) for x in step_ms],
            "host_phases_ms_last_step": {
                "workers": st.workers, "submits": st.submits, "wall": st.wall_ns / 1e6,
                "outline_per_worker": st.outline_ns / 1e6 / max(1, st.workers), "submit_per_worker": st.submit_ns / 1e6 / max(1, st.workers),
                "wait_per_worker": st.wait_ns / 1e6 / max(1, st.workers), "encode_per_worker": st.encode_ns / 1e6 / max(1, st.workers),
                "write_per_worker": st.write_ns / 1e6 / max(1, st.workers),
            },
        }
        if rank == 0 and world == 1:
            import oracle_lib as O

            threads = len(os.sched_getaffinity(0)) or 1
            fs, stride = cpu_sample(font_name, blobs, args.cpu_seconds, threads)
            t0 = time.perf_counter()
            cst = fs.render_all(O.MODE_PRECISE, threads=threads, stride=stride)
            dt = time.perf_counter() - t0
            result["cpu_baseline"] = {
                "value": cst["glyphs"] / dt, "unit": "glyphs/s", "cores": threads, "kind": "port",
                "sample": f"every {stride}th GlyphBlock ({cst['glyphs']} glyphs, {cst['pairs']} pairs, {dt:.2f} s)",
            }
    result["clocks"] = ClockSampler.summary(sampler.window(t_wall0, time.perf_counter()))
    sampler.stop()
    if rank == 0:
        emit(result)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
