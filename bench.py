#!/usr/bin/env python
"""bench.py — headline benchmark of the LZSS (+Huffman) hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload lzss|layered]

One "step" = one pass of the hot path over one batch: compress then decompress the whole
stream (the reference's own timed region, engine/engine.go:379-406).  The N=1 workload is
BASELINE.json configs[1]: lzss compress+decompress of a 64 MiB synthetic text stream with the
reference's window (4096) and min-match (-1, "smart") parameters.  For N>1 every rank handles
its own independent 64 MiB stream (north_star: batches of independent files partitioned across
GPUs; no data-path collective) — weak scaling.

Prints ONE JSON line on rank 0.  `value` = uncompressed bytes through compress+decompress per
second, inputs resident in HBM; `e2e` = the same through the C-ABI host-buffer calls with
pinned host buffers (H2D and D2H inside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BYTES = 64 << 20
WINDOW = 4096
METRIC = "lzss compress+decompress GB/s (uncompressed bytes / (encode+decode time)), bit-exact vs reference semantics"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def k2_traffic(n):
    """DRAM bytes of one K2 launch from the committed ncu --set full capture (per launch)."""
    p = os.path.join(ROOT, "profiles", "r1_k2_ncu_full.md")
    try:
        rd = wr = None
        for line in open(p):
            parts = line.split()
            if line.startswith("dram__bytes_read.sum"):
                rd = float(parts[1]) * (1e6 if parts[2].startswith("Mbyte") else 1e9 if parts[2].startswith("Gbyte") else 1e3 if parts[2].startswith("Kbyte") else 1)
            if line.startswith("dram__bytes_write.sum"):
                wr = float(parts[1]) * (1e6 if parts[2].startswith("Mbyte") else 1e9 if parts[2].startswith("Gbyte") else 1e3 if parts[2].startswith("Kbyte") else 1)
        if rd is None or wr is None:
            return None
        return int((rd + wr) * n / N_BYTES)
    except OSError:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval_ms=100):
        self.index = index
        self.interval_ms = interval_ms
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.interval_ms), "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ reference arm


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path, timed on this box's host cores.  The Go
    reference cannot be built here (no Go toolchain), so this is the oracle port in literal mode
    (per position, repeated leftmost-substring search over the window, all host threads), on a
    bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    from raisin_b200 import synth

    cores = os.cpu_count() or 1
    sample_n = 2 << 20
    data = synth.text(N_BYTES if args.full_reference else sample_n, 2)[:sample_n]
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        comp = po.lzss_compress_async(data, WINDOW, literal=True, threads=cores)
        back = po.lzss_decompress(comp)
        dt = time.perf_counter() - t0
        assert back == data
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    gbs = sample_n / (ms * 1e-3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "lzss compress+decompress, 64 MiB synthetic text stream, window 4096 (BASELINE configs[1])",
                   "window": WINDOW, "stream_bytes": N_BYTES},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": f"first {sample_n >> 20} MiB of the 64 MiB text stream per step; C port of the Go "
                                   "reference (no Go toolchain in this image), literal per-position search, "
                                   f"{cores} threads"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ batch workload


def run_batch(args, rank, local_rank, world):
    """BASELINE configs[3] shape: independent 256 KiB files (kind j mod 3, seed 1000 + j), layered
    lzss,huffman, files partitioned round-robin over ranks; every rank runs `--files` of them through
    rsn_batch_layers (host buffers in and out).  One step = compress all + decompress all."""
    import torch
    import torch.distributed as dist

    import raisin_b200 as rsn
    from raisin_b200 import parallel, synth

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = rsn._lib.lib()
    rsn._lib.check(lib.rsn_init(local_rank))
    mine = parallel.partition_files(args.files * world, world, rank)
    files = [synth.batch_file(j) for j in mine]
    total = sum(len(f) for f in files)
    n = len(files)
    keep = [rsn._lib._as_ptr(f) for f in files]
    ins = (C.c_void_p * n)(*[k[0] for k in keep])
    ns = (C.c_size_t * n)(*[k[1] for k in keep])

    def step(check=False):
        outs, out_ns, rcs = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        rsn._lib.check(lib.rsn_batch_layers(b"lzss,huffman", 1, n, ins, ns, outs, out_ns, rcs, args.workers, 0))
        b_outs, b_ns = (C.c_void_p * n)(), (C.c_size_t * n)()
        rsn._lib.check(lib.rsn_batch_layers(b"lzss,huffman", 0, n, outs, out_ns, b_outs, b_ns, rcs, args.workers, 0))
        csum = sum(out_ns)
        lossless = 0
        if check:
            lossless = sum(C.string_at(b_outs[i], b_ns[i]) == files[i] for i in range(n))
        for i in range(n):
            lib.rsn_free(outs[i])
            lib.rsn_free(b_outs[i])
        return csum, lossless

    csum, lossless = step(check=True)
    for _ in range(max(args.warmup, 3) - 1):
        step()
    lib.rsn_reset_kernel_launches()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # this workload is host-side launch heavy: poll the driver once a second, not ten times
    sampler = ClockSampler(local_rank, interval_ms=1000)
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop()
    launches = int(lib.rsn_kernel_launches())
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        gbs = world * total / (ms * 1e-3) / 1e9
        print(json.dumps({
            "metric": "lzss,huffman compress+decompress GB/s over a batch of independent files (host buffers)",
            "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"batch of {n} x 256 KiB files per GPU (BASELINE configs[3] shape), lzss,huffman",
                       "files_per_gpu": n, "bytes_per_gpu": total, "workers": args.workers},
            "compressed_bytes_per_gpu": csum, "lossless_files": lossless, "of_files": n,
            "gpu_launches": launches, "clocks": clocks,
            "timing": "host wall clock around the C-ABI calls (host buffers in and out), max over ranks",
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": total + csum, "d2h_bytes_per_step": total + csum},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ our arm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=N_BYTES)
    ap.add_argument("--full-reference", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="stream", choices=["stream", "batch"],
                    help="stream: BASELINE configs[1] (default, the headline); batch: configs[3] shape")
    ap.add_argument("--files", type=int, default=512, help="batch workload: files per GPU (256 KiB each)")
    ap.add_argument("--workers", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "batch":
        run_batch(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist

    import raisin_b200 as rsn
    from raisin_b200 import synth

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = rsn._lib.lib()
    rsn._lib.check(lib.rsn_init(local_rank))
    n = args.bytes
    data = synth.text(n, 2 + rank)  # rank r: its own independent stream
    stream = torch.cuda.Stream()  # a real (non-default) stream: the library launches on exactly this one
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0

    # ---- device-resident input
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def dev_step():
        out = C.c_void_p()
        out_n = C.c_size_t()
        rsn._lib.check(lib.rsn_dev_lzss_compress(d_in.data_ptr(), n, WINDOW, 0, C.byref(out), C.byref(out_n), sptr))
        back = C.c_void_p()
        back_n = C.c_size_t()
        rsn._lib.check(lib.rsn_dev_lzss_decompress(out, out_n.value, C.byref(back), C.byref(back_n), sptr))
        return out, out_n.value, back, back_n.value

    def free(*ptrs):
        for p in ptrs:
            lib.rsn_dev_free(p, sptr)

    # correctness of what is being timed: round trip on device
    out, c_bytes, back, back_n = dev_step()
    assert back_n == n
    h_back = (C.c_uint8 * n)()
    rsn._lib.check(lib.rsn_dev_download(back, n, h_back, sptr))
    assert bytes(h_back) == data, "device round trip differs"
    free(out, back)
    del h_back

    def timed(fn, steps, warmup, after_warmup=None):
        for _ in range(warmup):
            r = fn()
            if r:
                free(r[0], r[2])
        if after_warmup:
            after_warmup()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for a, b in ev:
            flush.fill_(1)  # L2 flush between timed iterations (outside the events)
            a.record(stream)
            r = fn()
            b.record(stream)
            if r:
                free(r[0], r[2])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return [a.elapsed_time(b) for a, b in ev]

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_list = timed(dev_step, args.steps, args.warmup, after_warmup=lib.rsn_reset_kernel_launches)
    launches = int(lib.rsn_kernel_launches())
    clocks = sampler.stop()
    ms = sum(ms_list) / len(ms_list)

    # ---- separate encode / decode timings (device resident)
    keep = {}

    def enc_only():
        o = C.c_void_p()
        on = C.c_size_t()
        rsn._lib.check(lib.rsn_dev_lzss_compress(d_in.data_ptr(), n, WINDOW, 0, C.byref(o), C.byref(on), sptr))
        if "c" in keep:
            lib.rsn_dev_free(keep["c"], sptr)
        keep["c"], keep["cn"] = o, on.value
        return None

    enc_ms = statistics.mean(timed(enc_only, max(3, args.steps // 2), 1))

    def dec_only():
        o = C.c_void_p()
        on = C.c_size_t()
        rsn._lib.check(lib.rsn_dev_lzss_decompress(keep["c"], keep["cn"], C.byref(o), C.byref(on), sptr))
        lib.rsn_dev_free(o, sptr)
        return None

    dec_ms = statistics.mean(timed(dec_only, max(3, args.steps // 2), 1))

    # ---- Huffman layer on the same stream (BASELINE configs[0]/[2] use it): encode/decode of the
    # LZSS output, device resident
    hk = {}

    def henc_only():
        o = C.c_void_p()
        on = C.c_size_t()
        rsn._lib.check(lib.rsn_dev_huff_compress(keep["c"], keep["cn"], C.byref(o), C.byref(on), sptr))
        if "h" in hk:
            lib.rsn_dev_free(hk["h"], sptr)
        hk["h"], hk["hn"] = o, on.value
        return None

    henc_ms = statistics.mean(timed(henc_only, max(3, args.steps // 2), 2))

    def hdec_only():
        o = C.c_void_p()
        on = C.c_size_t()
        rsn._lib.check(lib.rsn_dev_huff_decompress(hk["h"], hk["hn"], 0, C.byref(o), C.byref(on), sptr))
        lib.rsn_dev_free(o, sptr)
        return None

    hdec_ms = statistics.mean(timed(hdec_only, max(3, args.steps // 2), 2))

    # ---- dominant kernel (K2 match search) alone, CUDA events on its launch stream
    # the text stream has no '<', '\\' or 0xFF bytes, so the escaped buffer equals the raw one
    d_packed = torch.empty(n, dtype=torch.int32, device="cuda")

    def k2_only():
        rsn._lib.check(lib.rsn_dev_lzss_match(d_in.data_ptr(), n, WINDOW, d_packed.data_ptr(), sptr))
        return None

    k2_ms = statistics.mean(timed(k2_only, max(3, args.steps // 2), 1))

    # ---- e2e through the host-buffer C ABI, pinned host memory in and out
    h_in = lib.rsn_host_alloc(n)
    C.memmove(h_in, data, n)
    e2e_ms = []
    for it in range(2 + max(3, args.steps // 2)):
        o = C.POINTER(C.c_uint8)()
        on = C.c_size_t()
        b = C.POINTER(C.c_uint8)()
        bn = C.c_size_t()
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rsn._lib.check(lib.rsn_lzss_compress(h_in, n, WINDOW, 0, C.byref(o), C.byref(on)))
        rsn._lib.check(lib.rsn_lzss_decompress(o, on.value, C.byref(b), C.byref(bn)))
        dt = (time.perf_counter() - t0) * 1e3
        if it == 0:
            assert C.string_at(b, bn.value) == data, "host round trip differs"
        h2d, d2h = n + on.value, on.value + bn.value
        lib.rsn_free(o)
        lib.rsn_free(b)
        if it >= 2:
            e2e_ms.append(dt)
    lib.rsn_host_free(h_in)
    e2e = sum(e2e_ms) / len(e2e_ms)

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e, enc_ms, dec_ms, k2_ms, henc_ms, hdec_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e, enc_ms, dec_ms, k2_ms, henc_ms, hdec_ms = t.tolist()

    peak, peak_src = peaks()
    algo_bytes_k2 = n + keep["cn"]  # LZSS compress = n + c per stream; one K2 launch = one stream
    achieved = algo_bytes_k2 / (k2_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": world * n / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "lzss compress+decompress, 64 MiB synthetic text stream, window 4096 (BASELINE configs[1])",
                   "window": WINDOW, "stream_bytes": n, "streams_per_gpu": 1, "parallelism": f"independent streams x{world}",
                   "l2": "flushed between timed iterations (256 MiB write)"},
        "encode_GBps": world * n / (enc_ms * 1e-3) / 1e9, "decode_GBps": world * n / (dec_ms * 1e-3) / 1e9,
        "compressed_bytes": keep["cn"],
        "huffman_layer": {"input_bytes": keep["cn"], "output_bytes": hk["hn"],
                          "encode_GBps": world * keep["cn"] / (henc_ms * 1e-3) / 1e9,
                          "decode_GBps": world * keep["cn"] / (hdec_ms * 1e-3) / 1e9,
                          "layered_encode_GBps": world * n / ((enc_ms + henc_ms) * 1e-3) / 1e9,
                          "layered_decode_GBps": world * n / ((dec_ms + hdec_ms) * 1e-3) / 1e9},
        "e2e": {"value": world * n / (e2e * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "lzss match search (K2)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": k2_traffic(n), "peak_source": peak_src,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one k_match_tile launch on the "
                                       "64 MiB stream, profiles/r1_k2_ncu_full.md (scaled by n if --bytes differs)",
                     "kernel_ms": k2_ms, "algorithmic_bytes": algo_bytes_k2,
                     "note": "K2 is integer/shared-memory bound, not HBM bound; see DESIGN.md"},
    }
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import pyoracle as po

        cores = os.cpu_count() or 1
        sample_n = min(len(data), 64 << 20)  # ~11 s of CPU work on 16 cores
        sample = data[:sample_n]
        t0 = time.perf_counter()
        comp = po.lzss_compress_async(sample, WINDOW, literal=True, threads=cores)
        backb = po.lzss_decompress(comp)
        dt = time.perf_counter() - t0
        assert backb == sample
        line["cpu_baseline"] = {"value": sample_n / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                                "sample": f"first {sample_n >> 20} MiB of the stream, compress+decompress, C port of "
                                          f"the Go reference in literal mode, {cores} threads, {dt:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
