#!/usr/bin/env python
"""bench.py — headline benchmark of the LZSS + Huffman hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload batch|stream]

Default workload = BASELINE.json configs[3], the configuration the metric ("lzss+huffman
encode/decode GB/s (1/2/4/8 B200)") is quoted on: a batch of 4096 x 256 KiB synthetic files (kind =
j mod 3: text / logs / random bytes, seed 1000 + j), layered `lzss,huffman`, the files partitioned
round-robin over the N ranks (one process per GPU, no data-path collective; FIXED total, i.e.
strong scaling).  One "step" = one pass of the hot path over the batch: compress every file, then
decompress every result (the reference's own timed region, engine/engine.go:379-406, per file).

Prints ONE JSON line on rank 0:
  value     uncompressed bytes of the whole batch / max-over-ranks step time, files resident in HBM
            before the timed region and results left in HBM (rsn_batch_layers, device buffers)
  e2e       the same through rsn_batch_layers with pinned HOST buffers in and out
  roofline  the dominant kernel of the step (per-kernel CUDA-event times from the library's own
            launch timer, measured live in an extra pass), algorithmic bytes per SURVEY 8(d)
  parity    GPU output bytes compared with the CPU oracle on a sample of the same files
  cpu_baseline   the oracle port of the Go reference on a bounded sample (rank 0, N = 1 only)
  stream_c2 (N = 1) BASELINE configs[1]: lzss compress+decompress of one 64 MiB text stream

`--workload stream` makes configs[1] the headline instead (same keys).
`--impl reference` times the CPU port of the reference on the host cores (rank 0 only).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ProcessPoolExecutor, ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STREAM = 64 << 20
WINDOW = 4096
FILE_BYTES = 262144
N_FILES = 4096
ALGOS = b"lzss,huffman"
METRIC = ("lzss,huffman compress+decompress GB/s over a batch of independent files (uncompressed bytes / "
          "(encode+decode time)), bit-exact vs reference semantics")
METRIC_STREAM = "lzss compress+decompress GB/s (uncompressed bytes / (encode+decode time)), bit-exact vs reference semantics"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture, or None."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except (OSError, ValueError):
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
        self.first = 0
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

    def mark(self):
        """The timed region starts here: earlier samples are not reported."""
        self.first = len(self.rows)

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
        for r in self.rows[self.first:]:
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


def _gen_file(j):
    from raisin_b200 import synth

    return synth.batch_file(j, FILE_BYTES)


def make_files(indices):
    """Synthetic files of config 4 (deterministic per index), generated on a few processes."""
    procs = min(16, os.cpu_count() or 1, max(1, len(indices) // 64))
    if procs <= 1:
        return [_gen_file(j) for j in indices]
    with ProcessPoolExecutor(procs) as ex:
        return list(ex.map(_gen_file, indices, chunksize=32))


def batch_config(world, n_files):
    return {"workload": f"batch of {n_files} x 256 KiB synthetic files (text/logs/random by j mod 3), layered "
                        "lzss,huffman, compress + decompress (BASELINE configs[3])",
            "files": n_files, "file_bytes": FILE_BYTES, "window": WINDOW, "algorithms": "lzss,huffman",
            "parallelism": f"files partitioned round-robin over {world} rank(s), no data-path collective"}


# ------------------------------------------------------------------------------------------ CPU oracle legs


def oracle_batch(files, literal, threads):
    """lzss,huffman compress + decompress of every file with the CPU oracle; files run side by side
    on `threads` host threads (ctypes releases the GIL).  Returns (seconds, compressed, decoded)."""
    from oracle import pyoracle as po

    def one(f):
        lz = po.lzss_compress_async(f, WINDOW, literal=literal, threads=1)
        hf = po.huff_compress(lz)
        try:
            back = po.lzss_decompress(po.huff_decompress(hf))
        except po.OracleError:
            back = None
        return hf, back

    po.lib()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(one, files))
    dt = time.perf_counter() - t0
    return dt, [r[0] for r in res], [r[1] for r in res]


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path on this box's host cores.  The Go code cannot
    be built here (no Go toolchain in the image), so this is the oracle port in literal mode (per
    position, repeated leftmost-substring search over the window — the structure of lzss.go:156-184;
    Huffman with the reference's heap and linear-time packing), files side by side on all host
    threads, on a bounded sample of the same workload per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.workload == "stream":
        from oracle import pyoracle as po
        from raisin_b200 import synth

        sample_n = 2 << 20
        data = synth.text(sample_n, 2)
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
        metric, config = METRIC_STREAM, stream_config(1)
        sample = f"first {sample_n >> 20} MiB of the 64 MiB text stream per step"
    else:
        sample_files = max(3, min(args.ref_files, args.files))
        idx = list(range(sample_files))
        files = make_files(idx)
        total = sum(len(f) for f in files)
        times = []
        for it in range(args.warmup + args.steps):
            dt, _, _ = oracle_batch(files, True, cores)
            if it >= args.warmup:
                times.append(dt)
        ms = 1e3 * sum(times) / len(times)
        gbs = total / (ms * 1e-3) / 1e9
        metric, config = METRIC, batch_config(world, args.files)
        sample = f"files 0..{sample_files - 1} of the batch ({total >> 20} MiB) per step"
    line = {
        "impl": "reference", "metric": metric, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.workload == "batch" else "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": sample + "; C port of the Go reference (no Go toolchain in this image), literal "
                                            f"per-position search, {cores} threads"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ helpers (GPU)


def kernel_report(lib):
    n = lib.rsn_kernel_timing_report(None, 0)
    buf = C.create_string_buffer(n + 1)
    lib.rsn_kernel_timing_report(buf, n + 1)
    rows = []
    for ln in buf.value.decode().splitlines():
        name, cnt, ms = ln.rsplit(" ", 2)
        rows.append((name, int(cnt), float(ms)))
    return rows


STAGE_OF = [  # kernel-name prefix -> stage whose algorithmic bytes it is charged with
    ("kb_match", "lzss_compress"), ("k_match", "lzss_compress"), ("kb_escape", "lzss_compress"),
    ("k_escape", "lzss_compress"), ("kb_parse", "lzss_compress"), ("k_parse", "lzss_compress"),
    ("kb_emit", "lzss_compress"), ("k_emit", "lzss_compress"), ("kb_lz_", "lzss_compress"), ("k_lz_", "lzss_compress"),
    ("kb_tok", "lzss_decompress"), ("k_tok", "lzss_decompress"), ("k_resolve", "lzss_decompress"),
    ("kb_unesc", "lzss_decompress"), ("k_unesc", "lzss_decompress"),
    ("kb_huff_tree", "huffman_tree"), ("kb_rune", "huffman_compress"), ("k_rune", "huffman_compress"),
    ("kb_enc", "huffman_compress"), ("k_enc", "huffman_compress"), ("kb_hist", "huffman_compress"),
    ("kb_hdec", "huffman_decompress"), ("k_hdec", "huffman_decompress"),
]


def stage_of(kernel):
    for pre, st in STAGE_OF:
        if kernel.startswith(pre):
            return st
    return "other"


# ------------------------------------------------------------------------------------------ stream workload (c2)


def stream_config(world):
    return {"workload": "lzss compress+decompress, 64 MiB synthetic text stream, window 4096 (BASELINE configs[1])",
            "window": WINDOW, "stream_bytes": N_STREAM, "streams_per_gpu": 1,
            "parallelism": f"independent streams x{world}", "l2": "flushed between timed iterations (256 MiB write)"}


def measure_stream(args, torch, rsn, lib, rank, steps, warmup, with_e2e=True, with_oracle=True):
    """BASELINE configs[1] on this rank's GPU.  Returns a dict of numbers (ms, bytes, parity)."""
    from raisin_b200 import synth

    n = args.bytes
    data = synth.text(n, 2 + rank)
    stream = torch.cuda.Stream()
    sptr = C.c_void_p(stream.cuda_stream)
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {"n": n}

    def free(*ptrs):
        for p in ptrs:
            lib.rsn_dev_free(p, sptr)

    def dev_step():
        out, out_n, back, back_n = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        rsn._lib.check(lib.rsn_dev_lzss_compress(d_in.data_ptr(), n, WINDOW, 0, C.byref(out), C.byref(out_n), sptr))
        rsn._lib.check(lib.rsn_dev_lzss_decompress(out, out_n.value, C.byref(back), C.byref(back_n), sptr))
        return out, out_n.value, back, back_n.value

    with torch.cuda.stream(stream):
        # what is being timed, checked: the compressed stream and the round trip, downloaded once
        out, c_bytes, back, back_n = dev_step()
        assert back_n == n
        h_back = (C.c_uint8 * n)()
        rsn._lib.check(lib.rsn_dev_download(back, n, h_back, sptr))
        assert bytes(h_back) == data, "device round trip differs"
        h_comp = (C.c_uint8 * c_bytes)()
        rsn._lib.check(lib.rsn_dev_download(out, c_bytes, h_comp, sptr))
        gpu_comp = bytes(h_comp)
        free(out, back)
        del h_back, h_comp
        res["c"] = c_bytes

        def timed(fn, k, w):
            for _ in range(w):
                fn()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
            torch.cuda.synchronize()
            for a, b in ev:
                flush.fill_(1)  # L2 flush between timed iterations (outside the events)
                a.record(stream)
                fn()
                b.record(stream)
            torch.cuda.synchronize()
            return statistics.mean(a.elapsed_time(b) for a, b in ev)

        def both():
            o, _, b, _ = dev_step()
            free(o, b)

        lib.rsn_reset_kernel_launches()
        res["ms"] = timed(both, steps, warmup)
        res["launches_per_step"] = int(lib.rsn_kernel_launches()) // (steps + warmup)
        keep = {}

        def enc_only():
            o, on = C.c_void_p(), C.c_size_t()
            rsn._lib.check(lib.rsn_dev_lzss_compress(d_in.data_ptr(), n, WINDOW, 0, C.byref(o), C.byref(on), sptr))
            if "c" in keep:
                lib.rsn_dev_free(keep["c"], sptr)
            keep["c"], keep["cn"] = o, on.value

        def dec_only():
            o, on = C.c_void_p(), C.c_size_t()
            rsn._lib.check(lib.rsn_dev_lzss_decompress(keep["c"], keep["cn"], C.byref(o), C.byref(on), sptr))
            lib.rsn_dev_free(o, sptr)

        sub = max(3, steps // 2)
        res["enc_ms"] = timed(enc_only, sub, 1)
        res["dec_ms"] = timed(dec_only, sub, 1)
        hk = {}

        def henc_only():
            o, on = C.c_void_p(), C.c_size_t()
            rsn._lib.check(lib.rsn_dev_huff_compress(keep["c"], keep["cn"], C.byref(o), C.byref(on), sptr))
            if "h" in hk:
                lib.rsn_dev_free(hk["h"], sptr)
            hk["h"], hk["hn"] = o, on.value

        def hdec_only():
            o, on = C.c_void_p(), C.c_size_t()
            rsn._lib.check(lib.rsn_dev_huff_decompress(hk["h"], hk["hn"], 0, C.byref(o), C.byref(on), sptr))
            lib.rsn_dev_free(o, sptr)

        res["henc_ms"] = timed(henc_only, sub, 2)
        res["hdec_ms"] = timed(hdec_only, sub, 2)
        res["hn"] = hk["hn"]

        # per-kernel device times of one compress + decompress, from the library's launch timer
        lib.rsn_kernel_timing(1)
        flush.fill_(1)
        both()
        res["kernels"] = kernel_report(lib)
        lib.rsn_kernel_timing(0)
        free(keep["c"], hk["h"])

    if with_e2e:
        h_in = lib.rsn_host_alloc(n)
        C.memmove(h_in, data, n)
        e2e_ms = []
        for it in range(2 + max(3, steps // 2)):
            o, on = C.POINTER(C.c_uint8)(), C.c_size_t()
            b, bn = C.POINTER(C.c_uint8)(), C.c_size_t()
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rsn._lib.check(lib.rsn_lzss_compress(h_in, n, WINDOW, 0, C.byref(o), C.byref(on)))
            rsn._lib.check(lib.rsn_lzss_decompress(o, on.value, C.byref(b), C.byref(bn)))
            dt = (time.perf_counter() - t0) * 1e3
            if it == 0:
                assert C.string_at(b, bn.value) == data, "host round trip differs"
                assert C.string_at(o, on.value) == gpu_comp, "host-buffer call and device call disagree"
            res["h2d"], res["d2h"] = n + on.value, on.value + bn.value
            lib.rsn_free(o)
            lib.rsn_free(b)
            if it >= 2:
                e2e_ms.append(dt)
        lib.rsn_host_free(h_in)
        res["e2e_ms"] = statistics.mean(e2e_ms)

    res["parity"] = {"bytes": c_bytes, "sha256": hashlib.sha256(gpu_comp).hexdigest(), "equal": None,
                     "checked_against": "not run"}
    if with_oracle:
        from oracle import pyoracle as po

        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        comp = po.lzss_compress_async(data, WINDOW, literal=True, threads=cores)
        backb = po.lzss_decompress(comp)
        dt = time.perf_counter() - t0
        assert backb == data
        res["parity"] = {"bytes": len(comp), "sha256": hashlib.sha256(comp).hexdigest(), "equal": comp == gpu_comp,
                         "checked_against": "CPU oracle (literal mode) over the whole stream"}
        res["cpu"] = {"value": n / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                      "sample": f"the whole {n >> 20} MiB stream, compress+decompress, C port of the Go reference in "
                                f"literal mode, {cores} threads, {dt:.1f} s"}
    del d_in, flush
    torch.cuda.empty_cache()
    return res


def stream_summary(r, world=1):
    n = r["n"]
    out = {"config": stream_config(world), "value": world * n / (r["ms"] * 1e-3) / 1e9, "unit": "GB/s",
           "ms_per_step": r["ms"], "encode_GBps": world * n / (r["enc_ms"] * 1e-3) / 1e9,
           "decode_GBps": world * n / (r["dec_ms"] * 1e-3) / 1e9, "compressed_bytes": r["c"],
           "huffman_layer": {"input_bytes": r["c"], "output_bytes": r["hn"],
                             "encode_GBps": world * r["c"] / (r["henc_ms"] * 1e-3) / 1e9,
                             "decode_GBps": world * r["c"] / (r["hdec_ms"] * 1e-3) / 1e9},
           "gpu_launches_per_step": r["launches_per_step"], "parity": r["parity"]}
    if "e2e_ms" in r:
        out["e2e"] = {"value": world * n / (r["e2e_ms"] * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": r["h2d"],
                      "d2h_bytes_per_step": r["d2h"], "ms_per_step": r["e2e_ms"]}
    return out


def stream_roofline(r):
    """The dominant kernel of one compress + decompress of the stream."""
    peak, peak_src = peaks()
    name, cnt, ms = r["kernels"][0]
    n, c = r["n"], r["c"]
    stage_bytes = {"lzss_compress": n + c, "lzss_decompress": c + n}
    st = stage_of(name)
    algo = stage_bytes.get(st, n + c)
    per_launch_ms = ms / cnt
    achieved = algo / cnt / (per_launch_ms * 1e-3) / 1e9
    tot = sum(k[2] for k in r["kernels"])
    return {"bound": "hbm", "kernel": name, "stage": st, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": committed_traffic(name), "peak_source": peak_src,
            "launches": cnt, "kernel_ms_per_launch": per_launch_ms, "algorithmic_bytes_per_launch": algo / cnt,
            "share_of_kernel_time": ms / tot if tot else None,
            "kernels": [{"name": k[0], "launches": k[1], "ms": round(k[2], 4)} for k in r["kernels"][:10]],
            "note": "match search is integer/shared-memory bound, not HBM bound; see DESIGN.md"}


# ------------------------------------------------------------------------------------------ batch workload (c4)


def run_batch(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from raisin_b200 import parallel

    if args.workers <= 0 and world > 1:
        # the ranks of one box share its host cores: 8 workers per rank on a 32-core box with 8 ranks is
        # two threads per core before any helper thread
        args.workers = 8  # (8 ranks x 4 workers on 32 cores: 45-47 ms per step; 8 x 8: 36-43 ms)
        os.environ.setdefault("RSN_HOST_CORES", str(max(1, (os.cpu_count() or 8) // world)))
    mine = parallel.partition_files(args.files, world, rank)
    files = make_files(mine)  # before CUDA is initialised (the generator pool forks)

    import raisin_b200 as rsn

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = rsn._lib.lib()
    rsn._lib.check(lib.rsn_init(local_rank))
    n = len(files)
    total = sum(len(f) for f in files)
    total_all = args.files * FILE_BYTES

    # ---- inputs resident in HBM: one buffer, files at 256 KiB strides
    blob = b"".join(files)
    d_blob = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    offs = [0]
    for f in files:
        offs.append(offs[-1] + len(f))
    d_ins = (C.c_void_p * n)(*[d_blob.data_ptr() + o for o in offs[:-1]])
    ns = (C.c_size_t * n)(*[len(f) for f in files])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    del blob

    def dev_pass(keep=False, algos=ALGOS):
        outs, out_ns, rcs = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        rsn._lib.check(lib.rsn_batch_layers(algos, 1, n, d_ins, ns, outs, out_ns, rcs, args.workers, 1))
        b_outs, b_ns, rcs2 = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        rsn._lib.check(lib.rsn_batch_layers(algos, 0, n, outs, out_ns, b_outs, b_ns, rcs2, args.workers, 1))
        if keep:
            return outs, out_ns, b_outs, b_ns
        lib.rsn_dev_free_many(outs, n, None)
        lib.rsn_dev_free_many(b_outs, n, None)
        return sum(out_ns)

    def download(ptrs, sizes, which):
        res = []
        for i in which:
            h = (C.c_uint8 * max(1, sizes[i]))()
            rsn._lib.check(lib.rsn_dev_download(ptrs[i], sizes[i], h, None))
            res.append(bytes(h)[:sizes[i]])
        return res

    # ---- parity of what is being timed (untimed): every compressed file hashed, a sample compared
    # with the CPU oracle byte for byte, both directions
    outs, out_ns, b_outs, b_ns = dev_pass(keep=True)
    csum = sum(out_ns)
    comp_all = download(outs, out_ns, range(n))
    sha_all = hashlib.sha256(b"".join(comp_all)).hexdigest()
    sample = sorted(set(range(0, n, max(1, n // args.parity_files))))[:args.parity_files] if rank == 0 else []
    back_sample = download(b_outs, b_ns, sample)
    lossless = 0
    if rank == 0:
        back_all = download(b_outs, b_ns, range(n)) if n <= 1024 else None
        if back_all is not None:
            lossless = sum(back_all[i] == files[i] for i in range(n))
            del back_all
    lib.rsn_dev_free_many(outs, n, None)
    lib.rsn_dev_free_many(b_outs, n, None)
    parity = None
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        dt, want_c, want_b = oracle_batch([files[i] for i in sample], True, cores)
        eq_c = sum(comp_all[i] == w for i, w in zip(sample, want_c))
        eq_b = sum(b == w for b, w in zip(back_sample, want_b))
        sb = sum(len(files[i]) for i in sample)
        parity = {"files_checked": len(sample), "compressed_equal": eq_c, "decompressed_equal": eq_b,
                  "equal": eq_c == len(sample) and eq_b == len(sample),
                  "bytes": sum(len(w) for w in want_c),
                  "sha256": hashlib.sha256(b"".join(want_c)).hexdigest(),
                  "gpu_sha256_same_files": hashlib.sha256(b"".join(comp_all[i] for i in sample)).hexdigest(),
                  "checked_against": "CPU oracle (literal mode), lzss,huffman compress and decompress, files "
                                     f"{sample[0]}..{sample[-1]} step {sample[1] - sample[0] if len(sample) > 1 else 1} of rank 0's share",
                  "sha256_all_compressed_rank0": sha_all}
        if world == 1:
            cpu = {"value": sb / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                   "sample": f"{len(sample)} files of the batch ({sb >> 20} MiB), lzss,huffman compress+decompress, C "
                             f"port of the Go reference in literal mode, files side by side on {cores} threads, {dt:.1f} s"}
    del comp_all, back_sample

    # ---- timed: device resident
    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML initialisation, ~1 s) stalls CUDA
    # calls of other processes' threads now and then, and with the start inside the timed region one
    # run in three measured a step of 260-460 ms instead of 205-225.  It keeps sampling through the
    # timed region; only the samples taken there are reported.
    sampler = ClockSampler(local_rank, interval_ms=250)
    if not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        dev_pass()
    lib.rsn_reset_kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark()
    wall = []
    for a, b in ev:
        flush.fill_(1)  # L2 flush between timed iterations, outside the events
        torch.cuda.synchronize()
        a.record()
        t0 = time.perf_counter()
        if os.environ.get("RSN_ALLOC_TRACE"):
            print(f"[bench] timed step {len(wall)}", file=sys.stderr, flush=True)
        dev_pass()
        b.record()
        torch.cuda.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
    launches = int(lib.rsn_kernel_launches())
    ms = statistics.mean(a.elapsed_time(b) for a, b in ev)
    wall_ms = statistics.mean(wall)

    # ---- e2e: pinned host buffers in and out
    h_ptrs = []
    for f in files:
        p = lib.rsn_host_alloc(len(f))
        C.memmove(p, f, len(f))
        h_ptrs.append(p)
    h_ins = (C.c_void_p * n)(*h_ptrs)

    def host_pass():
        outs, out_ns, rcs = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        rsn._lib.check(lib.rsn_batch_layers(ALGOS, 1, n, h_ins, ns, outs, out_ns, rcs, args.workers, 0))
        b_outs, b_ns, rcs2 = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        rsn._lib.check(lib.rsn_batch_layers(ALGOS, 0, n, outs, out_ns, b_outs, b_ns, rcs2, args.workers, 0))
        c, d = sum(out_ns), sum(b_ns)
        lib.rsn_free_many(outs, n)
        lib.rsn_free_many(b_outs, n)
        return c, d

    e2e_steps = max(3, args.steps // 2)
    for _ in range(2):
        host_pass()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_list = []
    for _ in range(e2e_steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hc, hd = host_pass()
        e2e_list.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = statistics.mean(e2e_list)
    for p in h_ptrs:
        lib.rsn_host_free(p)

    # ---- per-kernel device times of one pass (one worker: kernels run alone on the GPU), and the
    # intermediate (LZSS) size the stage byte counts need
    outs, out_ns, rcs = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
    rsn._lib.check(lib.rsn_batch_layers(b"lzss", 1, n, d_ins, ns, outs, out_ns, rcs, args.workers, 1))
    c_lz = sum(out_ns)
    lib.rsn_dev_free_many(outs, n, None)
    saved_workers = args.workers
    args.workers = 1
    dev_pass()
    lib.rsn_kernel_timing(1)
    flush.fill_(1)
    torch.cuda.synchronize()
    dev_pass()
    kernels = kernel_report(lib)
    lib.rsn_kernel_timing(0)
    args.workers = saved_workers

    if world > 1:
        t = torch.tensor([ms, e2e_ms, wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, wall_ms = t.tolist()
        t = torch.tensor([float(launches), float(csum)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches, csum_all = int(t[0].item()), int(t[1].item())
    else:
        csum_all = csum

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        back_lz = total  # lz.Decompress output of this rank's files: as many bytes as went in (sizes, not content)
        stage_bytes = {"lzss_compress": total + c_lz, "lzss_decompress": c_lz + back_lz,
                       "huffman_compress": 2 * c_lz + csum, "huffman_decompress": csum + c_lz, "huffman_tree": csum}
        name, cnt, kms = kernels[0]
        st = stage_of(name)
        algo = stage_bytes.get(st, total)
        achieved = algo / (kms * 1e-3) / 1e9
        ktot = sum(k[2] for k in kernels)
        line = {
            "metric": METRIC, "value": total_all / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": dict(batch_config(world, args.files), workers=args.workers,
                                                               l2="flushed between timed iterations (256 MiB write)"),
            "timing": "CUDA events on the default stream around each step (the C-ABI calls return when the step's "
                      "work has finished), max over ranks; wall clock beside it",
            "wall_ms_per_step": wall_ms, "step_ms": [round(a.elapsed_time(b), 1) for a, b in ev],
            "compressed_bytes": csum_all, "lzss_stage_bytes_rank0": c_lz, "lossless_files_rank0": lossless,
            "files_rank0": n,
            "e2e": {"value": total_all / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": total + hc,
                    "d2h_bytes_per_step": hc + hd, "ms_per_step": e2e_ms,
                    "note": "rsn_batch_layers with pinned host buffers in and out; bytes are per rank"},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": name, "stage": st, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": committed_traffic(name), "peak_source": peak_src,
                         "launches": cnt, "kernel_ms_per_launch": kms / cnt,
                         "algorithmic_bytes_per_launch": algo / cnt,
                         "algorithmic_bytes": f"{st}: bytes in + bytes out of that stage over rank 0's {n} files (SURVEY 8(d)), "
                                              "divided over the kernel's launches",
                         "share_of_kernel_time": kms / ktot if ktot else None,
                         "kernels": [{"name": k[0], "launches": k[1], "ms": round(k[2], 4)} for k in kernels[:12]],
                         "measured": "cudaEvent pairs around every launch on its own stream (rsn_kernel_timing), one extra "
                                     "pass with one worker so kernels run alone"},
        }
        if parity:
            line["parity"] = parity
        if cpu:
            line["cpu_baseline"] = cpu
    return line, (torch, dist, rsn, lib)


# ------------------------------------------------------------------------------------------ main


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="batch", choices=["batch", "stream"],
                    help="batch: BASELINE configs[3] (default, the metric's own workload); stream: configs[1]")
    ap.add_argument("--files", type=int, default=N_FILES, help="batch workload: files in the whole batch (256 KiB each)")
    ap.add_argument("--bytes", type=int, default=N_STREAM, help="stream workload: bytes per stream")
    ap.add_argument("--workers", type=int, default=0, help="host threads/streams per rank for the batch call (0 = default)")
    ap.add_argument("--parity-files", type=int, default=48, help="files compared with the CPU oracle (rank 0)")
    ap.add_argument("--ref-files", type=int, default=24, help="--impl reference: files per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream", action="store_true", help="skip the configs[1] side measurement at N = 1")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if args.workload == "batch":
        line, (torch, dist, rsn, lib) = run_batch(args, rank, local_rank, world)
        if rank == 0 and world == 1 and not args.no_stream:
            r = measure_stream(args, torch, rsn, lib, 0, max(3, args.steps), 3, with_oracle=not args.no_cpu_baseline)
            line["stream_c2"] = stream_summary(r)
            line["stream_c2"]["roofline"] = stream_roofline(r)
            if "cpu" in r:
                line["stream_c2"]["cpu_baseline"] = r["cpu"]
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- stream workload as the headline (weak scaling: one independent 64 MiB stream per rank)
    import torch
    import torch.distributed as dist

    import raisin_b200 as rsn

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = rsn._lib.lib()
    rsn._lib.check(lib.rsn_init(local_rank))
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    r = measure_stream(args, torch, rsn, lib, rank, args.steps, max(3, args.warmup),
                       with_oracle=(rank == 0 and not args.no_cpu_baseline))
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([r["ms"], r["e2e_ms"], r["enc_ms"], r["dec_ms"], r["henc_ms"], r["hdec_ms"]], dtype=torch.float64,
                         device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r["ms"], r["e2e_ms"], r["enc_ms"], r["dec_ms"], r["henc_ms"], r["hdec_ms"] = t.tolist()
    if rank == 0:
        s = stream_summary(r, world)
        line = {"metric": METRIC_STREAM, "value": s["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": s["config"],
                "encode_GBps": s["encode_GBps"], "decode_GBps": s["decode_GBps"], "compressed_bytes": r["c"],
                "huffman_layer": s["huffman_layer"], "e2e": s["e2e"],
                "gpu_launches": r["launches_per_step"] * args.steps * world, "clocks": clocks,
                "roofline": stream_roofline(r), "parity": r["parity"]}
        if "cpu" in r:
            line["cpu_baseline"] = r["cpu"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
