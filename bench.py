#!/usr/bin/env python
"""bench.py — FLASHE hot path on N B200s: encode+encrypt (all clients) -> aggregate -> decrypt+decode.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] ...                # the reference's own CPU code (oracle/_ref)
    torchrun --nproc-per-node N bench.py --gpus N ...              # N > 1, one rank per GPU

Metric (BASELINE.json): client-elements/s = n_clients * L / time of one full round, whole job over all
N GPUs.  Workload = BASELINE config 5: L = 100M float32 elements, 64 clients, int_bits 32, double
masking, element-range sharded across the N GPUs (strong scaling: total work fixed; no data-path
collective — every element's mask depends only on (key, iter, client, index, L, n_jobs)).

One "step" = one round over synthetic gradients resident in HBM.  `value` is device-timed (CUDA
events on the launching stream, max over ranks); `e2e` is the same round through the public API with
HOST buffers: pinned float32 gradients copied host->device and the decoded float64 aggregate copied
device->host inside the timed region.  Inputs (25.6 GB at N=1) are far larger than L2 (126 MB), so no
L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "client_elements_per_sec_enc_agg_dec"
UNIT = "client-elements/s"
KEY = bytes(range(32))
ALPHA = 5.938345 * 0.1          # ACIQ(16 bit) * sigma, sigma = 0.1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elements", type=int, default=100_000_000)
    ap.add_argument("--clients", type=int, default=64)
    ap.add_argument("--int-bits", type=int, default=32)
    ap.add_argument("--n-jobs", type=int, default=0, help="reference N_JOBS baked into the ciphertext format (0 = cpu_count())")
    ap.add_argument("--share-streams", type=int, default=0, help="compute F(t,c+1) once for clients c and c+1")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the share-streams and precomputed-mask rounds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-elements", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-clients", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=2)
    return ap.parse_args()


def measured_traffic(n_client_elements):
    """DRAM bytes of the encode kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch), scaled to
    this launch's client-elements when the shapes differ.  None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))["k_stream_encode"]
        per = float(d["dram_bytes"]) / float(d["client_elements"])
        return per * n_client_elements, "%s (%.3f B per client-element, captured on %d client-elements)" % (
            d["source"], per, int(d["client_elements"]))
    except Exception:
        return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None
        self.first = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """Samples before this point (warm-up) are not part of the timed region."""
        self.first = len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        rows = self.rows[self.first:] or self.rows[-3:]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_round(elements, clients, int_bits, threads, keep=False):
    """One bounded round of the reference's CPU path on this host.  With oracle/_ref present (staged by
    oracle/make_ref.py from /root/reference) this is the reference's OWN code — jzf_flashe.FlasheCipher
    .encrypt / .decrypt with its multiprocessing.Pool(N_JOBS), jzf_quantize's encode / decode — driven like
    encrypt_test/final_big_table.ipynb:222-233, 408-411 (kind "reference"); otherwise the Python port
    oracle/flashe_port.py (kind "port").  Returns (seconds, phases, kind, outputs or None)."""
    import numpy as np
    xs = [(np.random.RandomState(1000 + c).standard_normal(elements) * 0.1).astype(np.float32) for c in range(clients)]
    seeds = [2000 + c for c in range(clients)]
    phases = {}
    from oracle import ref_driver as R
    if R.available():
        t0 = time.perf_counter()
        res = R.run_round(KEY, int_bits, 0, xs, float(ALPHA), 16, n_jobs=threads, seeds=seeds, timings=phases)
        dt = time.perf_counter() - t0
        return dt, phases, "reference", ((xs, seeds) + res if keep else None)
    from oracle import flashe_port as P
    np.random.seed(2000)
    t0 = time.perf_counter()
    P.run_round(KEY, int_bits, 0, xs, float(ALPHA), 16, n_jobs=threads, timings=phases)
    return time.perf_counter() - t0, phases, "port", None


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores
    (all of them: N_JOBS = cpu_count(), as jzf_flashe.py:7 does), each step a bounded sample of the
    workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    elements = min(args.cpu_sample_elements, args.elements)
    clients = min(args.cpu_sample_clients, args.clients)
    for _ in range(args.warmup):
        cpu_round(min(elements, 50_000), clients, args.int_bits, cores)
    times, phases, kind = [], {}, "port"
    for _ in range(args.steps):
        dt, phases, kind, _ = cpu_round(elements, clients, args.int_bits, cores)
        times.append(dt)
    total = sum(times)
    value = clients * elements * len(times) / total
    sample = "%d clients x %d elements per step (slice of the %d x %d workload), N_JOBS=%d, %s" % (
        clients, elements, args.clients, args.elements, cores,
        "the reference's own modules staged under oracle/_ref (FlasheCipher.encrypt/decrypt, multiprocessing.Pool per call)"
        if kind == "reference" else "Python port of the reference's loops (oracle/_ref absent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, cores),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "phases_s_last_step": phases},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, n_jobs):
    return {"workload": "BASELINE config 5: %dM-element float32 gradient, %d clients, encode+encrypt -> aggregate(element-wise) "
                        "-> decrypt+decode, element-range sharded over the GPUs" % (args.elements // 1_000_000, args.clients),
            "elements": args.elements, "clients": args.clients, "int_bits": args.int_bits, "element_bits": 16,
            "masking": "double", "n_jobs": n_jobs, "prf": "AES-256-ECB on the fly (no precomputed masks)",
            "share_streams": bool(args.share_streams), "noise": "device Philox4x32-10, res53",
            "l2": "inputs (%.1f GB per round) exceed L2; no flush needed" % (args.elements * args.clients * 4 / 1e9)}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import flashe_b200 as fb
    from flashe_b200.device import launch_count

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    all_cpus = os.sched_getaffinity(0)
    bound_cpus = bind_to_gpu_numa(local)          # pinned buffers and copy submissions on the GPU's own NUMA node
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L, n, bits = args.elements, args.clients, args.int_bits
    n_jobs = args.n_jobs or (os.cpu_count() or 1)
    # element-range shard of this rank, cut at multiples of 4 elements (16-byte aligned rows)
    per = (L // world) // 4 * 4
    begin = rank * per
    count = per if rank + 1 < world else L - begin
    span = fb.VectorSpan(L, n_jobs, begin, count)
    ctx = fb.DeviceContext(KEY, bits, dev)
    codec = fb.CodecSpec(alpha=float(ALPHA), element_bits=16, n_clients=n)
    noise = fb.NoiseSpec(seed=0x5EED, stream=0)
    scheme = fb.SCHEME_DOUBLE

    # synthetic gradients N(0, 0.1^2), generated on the device
    x = torch.empty((n, count), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    for c in range(n):
        g.manual_seed(1000 + c + 7919 * rank)
        x[c].normal_(0.0, 0.1, generator=g)
    cts = ctx.empty_words(count, rows=n)
    agg = ctx.empty_words(count)
    out = torch.empty(count, dtype=torch.float64, device=dev)

    def round_device(ev=None):
        if ev:
            ev[0].record()
        ctx.encode_encrypt_batch(0, 0, scheme, x, codec, noise, span, out=cts, share_streams=bool(args.share_streams))
        if ev:
            ev[1].record()
        ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
        if ev:
            ev[2].record()
        ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)
        if ev:
            ev[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        round_device()
    barrier()
    sampler.mark()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    l0 = launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for s in range(args.steps):
        round_device(evs[s])
    stop.record()
    barrier()
    launches = launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = start.elapsed_time(stop)
    ph = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / args.steps for i in range(3)]
    t = torch.tensor([ms_total] + ph, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ph = float(t[0]), [float(v) for v in t[1:]]
    ms_step = ms_total / args.steps
    value = n * L / (ms_step * 1e-3)

    # the reference's dense path sums the PACKED wire integers (carry leak, jzf_aggregator.py:406-419): same
    # bytes, timed beside the element-wise sum of the headline round
    pk = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ctx.aggregate(cts, fb.AGG_PACKED, out=agg)
    pk[0].record()
    for _ in range(3):
        ctx.aggregate(cts, fb.AGG_PACKED, out=agg)
    pk[1].record()
    torch.cuda.synchronize()
    packed_ms = pk[0].elapsed_time(pk[1]) / 3

    # the mask-only PRF kernel on this rank's shard: the measured ceiling of roofline_prf
    mask_only_elems = min(L, 50_000_000) // 4 * 4             # the same size on every N: a launch long enough to hide its ramp
    mspan = fb.VectorSpan(L, n_jobs, 0, mask_only_elems)
    mbuf = ctx.empty_words(mask_only_elems)
    for _ in range(2):
        ctx.masks(0, [0, 1], [1, -1], mspan, out=mbuf)
    pk[0].record()
    for _ in range(4):
        ctx.masks(0, [0, 1], [1, -1], mspan, out=mbuf)
    pk[1].record()
    torch.cuda.synchronize()
    del mbuf
    mask_only_blocks_per_s = 2 * (mask_only_elems / (128 // bits)) / (pk[0].elapsed_time(pk[1]) / 4 * 1e-3)

    # ---------------------------------------------------------------- same round, other legitimate schedules
    variants = None
    if not args.no_variants:
        variants = run_variants(args, ctx, fb, span, codec, noise, scheme, x, cts, agg, out, world, dev, barrier)

    # ---------------------------------------------------------------- end to end (host buffers)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, ctx, fb, span, codec, noise, scheme, x, cts, agg, out, world, dev, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = peaks()
    enc_bytes = n * count * 8                    # 4 B float in + 4 B ciphertext out per client-element
    enc_gbs = enc_bytes / (ph[0] * 1e-3) / 1e9
    m = 128 // bits
    blocks = (n + 1 if args.share_streams else 2 * n) * (count / m)
    blocks_per_s = blocks / (ph[0] * 1e-3)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    # PRF ceilings, measured: (1) the mask-only kernel (flashe_masks: the same AES-256 T-table PRF with nothing but a
    # 16-byte store per block), timed live on this GPU just above; (2) the shared-memory lookup rate of an SM from the
    # committed micro-benchmark (scripts/microbench.py, profiles/r2g_micro.jsonl: 31.87 of the nominal 32 conflict-free
    # 4-byte lookups per clock; a block costs 197 lookups with the hoisted round 1 and the counter-window factoring of
    # round 2, 224 for textbook T-table AES-256)
    lookups = 197.0
    lds_per_clk = 31.87
    lds_peak_blocks = 148 * sm_mhz * 1e6 * lds_per_clk / lookups
    agg_bytes = (n + 1) * count * 4
    dec_bytes = count * 12
    traffic, traffic_src = measured_traffic(n * count)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32" if bits <= 32 else ("u64" if bits <= 64 else "u128"), "data": "synthetic",
        "config": workload_config(args, n_jobs),
        "clocks": clocks,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_stream<WORDS=%d, MMAX=%d, M_ENCODE, SHARE=%d> (fused encode + AES-256 PRF masks + modular add)"
                               % (1 if bits <= 32 else (2 if bits <= 64 else 4), 4 if m <= 4 else (6 if m <= 6 else 16), int(bool(args.share_streams))),
                     "achieved": enc_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": enc_gbs / hbm_peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": enc_bytes, "ms_per_launch": ph[0],
                     "note": "this kernel is bound by the PRF (shared-memory table lookups), not HBM: see roofline_prf"},
        "roofline_prf": {"bound": "lds", "achieved": blocks_per_s / 1e9, "peak": mask_only_blocks_per_s / 1e9, "unit": "G AES-256 blocks/s",
                         "frac": blocks_per_s / mask_only_blocks_per_s,
                         "peak_source": "measured in this run: the mask-only kernel k_stream<M_MASKS> (flashe_masks, same PRF, no encode, "
                                        "16-byte store per 2 blocks) on %d M elements x 2 streams" % (mask_only_elems // 1_000_000),
                         "lookups_per_block": lookups,
                         "lds_lookup_ceiling": {"g_blocks_per_s": lds_peak_blocks / 1e9, "frac": blocks_per_s / lds_peak_blocks,
                                                "source": "148 SMs x sampled SM clock x 31.87 measured conflict-free LDS.32 per clock and SM "
                                                          "(k_lds_peak, profiles/r2g_micro.jsonl) / 197 lookups per block"},
                         "bitsliced_alternative_g_blocks_per_s": {"measured": 19.9, "scaled_to_113_gate_sbox": 24.6,
                                                                  "source": "k_aes_bitslice at 97.9 % ALU pipe, profiles/r2g_micro.jsonl"}},
        "phases": {"encode_encrypt_ms": ph[0], "aggregate_ms": ph[1], "decrypt_decode_ms": ph[2],
                   "aggregate_gbs": agg_bytes / (ph[1] * 1e-3) / 1e9, "aggregate_frac_of_hbm": agg_bytes / (ph[1] * 1e-3) / 1e9 / hbm_peak,
                   "decrypt_decode_gbs": dec_bytes / (ph[2] * 1e-3) / 1e9,
                   "aggregate_packed_carry_ms": packed_ms, "aggregate_packed_carry_gbs": agg_bytes / (packed_ms * 1e-3) / 1e9},
        "hbm_roofline_client_elements_per_s": hbm_peak * 1e9 / (12.0 + 16.0 / n) * world,
        "frac_of_hbm_roofline_end_to_end": value / (hbm_peak * 1e9 / (12.0 + 16.0 / n) * world),
    }
    if e2e:
        e2e["cpus_bound_to_gpu_numa_node"] = bound_cpus
        line["e2e"] = e2e
    if variants:
        for v in variants.values():
            if "ms_per_step" in v:
                v["value"] = n * L / (v["ms_per_step"] * 1e-3)
                v["frac_of_hbm_roofline_end_to_end"] = v["value"] / (hbm_peak * 1e9 / v["hbm_bytes_per_client_element"] * world)
        pm = variants.get("precomputed_masks")
        if pm and "online_encrypt_ms" in pm:
            pm["online_encrypt_gbs"] = n * count * 12 / (pm["online_encrypt_ms"] * 1e-3) / 1e9
            pm["online_encrypt_frac_of_hbm"] = pm["online_encrypt_gbs"] / hbm_peak
            pm["online_encrypt_noise32_frac_of_hbm"] = n * count * 12 / (pm["online_encrypt_noise32_ms"] * 1e-3) / 1e9 / hbm_peak
            # the same schedule with the 32-bit-resolution generator in the online step (aggregate and decrypt+decode unchanged)
            ms32 = pm["ms_per_step"] - pm["online_encrypt_ms"] + pm["online_encrypt_noise32_ms"]
            pm["noise32_ms_per_step"] = ms32
            pm["noise32_value"] = n * L / (ms32 * 1e-3)
            pm["noise32_frac_of_hbm_roofline_end_to_end"] = pm["noise32_value"] / (hbm_peak * 1e9 / pm["hbm_bytes_per_client_element"] * world)
        line["variants"] = variants
    if not args.no_cpu_baseline and world == 1:
        os.sched_setaffinity(0, all_cpus)         # the reference's Pool gets every host core again
        line["cpu_baseline"] = cpu_baseline(args, ctx, fb)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_variants(args, ctx, fb, span, codec, noise, scheme, x, cts, agg, out, world, dev, barrier):
    """The same round under the two other schedules the path offers; outputs are bit-identical to the
    headline round (checked here on the aggregate and on one client's ciphertext).
      share_streams      F(t,c+1) computed once for clients c and c+1 hosted on the same GPU.
      precomputed_masks  FLASHE's mask precomputation (jzf_flashe.py:596-666): the combined masks of the
                         round are generated ahead of time (fill, timed separately, off the critical
                         path in the reference's design) and the ONLINE round is encode + add of the
                         stored mask -> aggregate -> decrypt+decode: HBM-bound."""
    import torch
    import torch.distributed as dist
    n, count = x.shape
    res = {}

    def timed(fn, steps):
        fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # reference outputs of the headline (on-the-fly, unshared) round
    ctx.encode_encrypt_batch(0, 0, scheme, x, codec, noise, span, out=cts, share_streams=False)
    ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
    probe = min(5, n - 1)
    agg_ref, row_ref = agg.clone(), cts[probe].clone()

    def round_shared():
        ctx.encode_encrypt_batch(0, 0, scheme, x, codec, noise, span, out=cts, share_streams=True)
        ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
        ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)

    ms = timed(round_shared, args.steps)
    same = bool(torch.equal(agg.view(torch.int32), agg_ref.view(torch.int32)) and
                torch.equal(cts[probe].view(torch.int32), row_ref.view(torch.int32)))
    res["share_streams"] = {"ms_per_step": ms, "bit_exact_vs_headline_round": same, "hbm_bytes_per_client_element": 12.0 + 16.0 / n,
                            "aes_blocks_per_round": (n + 1) * count * world // (128 // args.int_bits)}

    # throughput-mode noise: one 32-bit Philox word per element instead of numpy's two (flashe_noise.resolution);
    # the reference's own noise is an unseeded MT19937 stream, so neither device stream is "the reference's" —
    # both are documented, both are reproduced bit for bit by the oracle when it is fed the same numbers
    noise32 = fb.NoiseSpec(seed=noise.seed, stream=noise.stream, resolution=32)

    def round_noise32():
        ctx.encode_encrypt_batch(0, 0, scheme, x, codec, noise32, span, out=cts, share_streams=False)
        ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
        ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)

    ms = timed(round_noise32, args.steps)
    res["noise_32bit_resolution"] = {"ms_per_step": ms, "hbm_bytes_per_client_element": 12.0 + 16.0 / n,
                                     "note": "same schedule as the headline round; stochastic rounding draws 32 random bits per element instead of 53"}

    try:
        masks = ctx.empty_words(count, rows=n)
    except RuntimeError as e:                                   # not enough HBM for the mask buffer
        res["precomputed_masks"] = {"skipped": str(e)[:120]}
        return res

    def fill():
        for c in range(n):
            ctx.masks(0, [c, c + 1], [1, -1], span, out=masks[c])

    fill_ms = timed(fill, 1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    enc_ms = []

    def round_online():
        ev[0].record()
        ctx.encode_add_premasked_batch(x, codec, noise, masks, span, out=cts)      # all clients: one launch
        ev[1].record()
        ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
        ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)

    ms = timed(round_online, args.steps)
    torch.cuda.synchronize()
    enc_ms = ev[0].elapsed_time(ev[1])
    same = bool(torch.equal(agg.view(torch.int32), agg_ref.view(torch.int32)) and
                torch.equal(cts[probe].view(torch.int32), row_ref.view(torch.int32)))
    e32 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for k in range(3):
        if k == 1:
            e32[0].record()
        ctx.encode_add_premasked_batch(x, codec, noise32, masks, span, out=cts)
    e32[1].record()
    torch.cuda.synchronize()
    enc32_ms = e32[0].elapsed_time(e32[1]) / 2
    t = torch.tensor([enc_ms, fill_ms, enc32_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["precomputed_masks"] = {"ms_per_step": ms, "online_encrypt_ms": float(t[0]), "fill_ms": float(t[1]),
                                "fill_g_aes_blocks_per_s": 2 * n * count * world / (128 // args.int_bits) / (float(t[1]) * 1e-3) / 1e9,
                                "mask_buffer_bytes_per_gpu": int(masks.numel() * masks.element_size()),
                                "online_encrypt_noise32_ms": float(t[2]),
                                "bit_exact_vs_headline_round": same, "hbm_bytes_per_client_element": 16.0 + 16.0 / n,
                                "gpu_launches_per_round": 3}
    del masks
    return res


def bind_to_gpu_numa(local):
    """Run this rank on the CPUs the driver reports as closest to its GPU, so that the pinned host buffers it
    allocates (first touch) and the copy submissions sit on the GPU's own NUMA node / PCIe root."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def h2d_probe(torch, dist, dev, world, barrier, nbytes=1 << 31, reps=3):
    """What the box's host -> device path gives with EVERY rank copying at once from pinned memory (the
    ceiling of any end-to-end number that has to bring the gradients in): per-rank GB/s (min over ranks) and
    the sum over ranks."""
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    devb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    devb.copy_(host, non_blocking=True)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        devb.copy_(host, non_blocking=True)
    b.record()
    barrier()
    gbs = nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9
    per_rank = [gbs]
    if world > 1:
        mine = torch.tensor([gbs], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [float(v[0]) for v in allr]
    del host, devb
    return {"per_rank_gbs": per_rank, "per_rank_min_gbs": min(per_rank), "per_rank_max_gbs": max(per_rank), "sum_gbs": sum(per_rank)}


def run_e2e(args, ctx, fb, span, codec, noise, scheme, x, cts, agg, out, world, dev, barrier):
    """The same round through the public API with HOST buffers: every step copies the float32 gradients from
    pinned host memory (groups of 4 clients, each group split over two copy streams, three staging buffers
    so the copy engines never wait for the kernels) and reads the decoded float64 aggregate back.  The timed
    region is bounded by the host -> device path, so the result is quoted against a concurrent H2D probe."""
    import torch
    import torch.distributed as dist
    n, count = x.shape
    probe = h2d_probe(torch, dist, dev, world, barrier)
    rank = dist.get_rank() if world > 1 else 0
    if os.environ.get("FLASHE_E2E_TEST_SKEW"):          # (exercise the uneven-range path on a box with even links)
        probe["per_rank_gbs"] = [g * (1.0 + 0.3 * i) for i, g in enumerate(probe["per_rank_gbs"])]
        probe.update(per_rank_min_gbs=min(probe["per_rank_gbs"]), per_rank_max_gbs=max(probe["per_rank_gbs"]),
                     sum_gbs=sum(probe["per_rank_gbs"]), skewed_for_test=True)
    if world > 1 and probe["per_rank_max_gbs"] > 1.1 * probe["per_rank_min_gbs"]:
        # The GPUs of the box do not see the same host bandwidth when all of them copy at once (PCIe switches shared by
        # GPU pairs): the end-to-end round is bound by those copies, so its element ranges are cut in proportion to what
        # each rank's link delivered in the probe (the device-timed round above keeps equal ranges).
        L, tot = span.total_len, probe["sum_gbs"]
        cuts, acc = [0], 0.0
        for g in probe["per_rank_gbs"][:-1]:
            acc += g
            cuts.append(int(L * acc / tot) // 4 * 4)
        cuts.append(L)
        begin, count = cuts[rank], cuts[rank + 1] - cuts[rank]
        span = fb.VectorSpan(L, span.n_jobs, begin, count)
        del x, cts, agg, out
        x = torch.empty((n, count), dtype=torch.float32, device=dev)
        g = torch.Generator(device=dev)
        for c in range(n):
            g.manual_seed(1000 + c + 7919 * rank)
            x[c].normal_(0.0, 0.1, generator=g)
        cts, agg = ctx.empty_words(count, rows=n), ctx.empty_words(count)
        out = torch.empty(count, dtype=torch.float64, device=dev)
        probe["shard_elements_per_rank"] = [cuts[i + 1] - cuts[i] for i in range(world)]
    group = 4 if n % 4 == 0 else 1
    host_x = torch.empty((n, count), dtype=torch.float32, pin_memory=True)
    host_x.copy_(x)                               # synthetic gradients, now "on the host"
    host_out = torch.empty(count, dtype=torch.float64, pin_memory=True)
    copy_streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    main = torch.cuda.current_stream(dev)
    depth = 3
    stage = [torch.empty((group, count), dtype=torch.float32, device=dev) for _ in range(depth)]
    staged = [[torch.cuda.Event() for _ in range(2)] for _ in range(depth)]
    consumed = [torch.cuda.Event() for _ in range(depth)]
    half = max(1, group // 2)

    def round_host():
        ngroups = n // group
        for gi in range(ngroups):
            b = gi % depth
            for k, cs in enumerate(copy_streams):
                lo, hi = (0, half) if k == 0 else (half, group)
                if lo >= hi:
                    continue
                with torch.cuda.stream(cs):
                    if gi >= depth:
                        cs.wait_event(consumed[b])
                    stage[b][lo:hi].copy_(host_x[gi * group + lo:gi * group + hi], non_blocking=True)
                    staged[b][k].record(cs)
                main.wait_event(staged[b][k])
            ns = fb.NoiseSpec(seed=noise.seed, stream=gi * group)
            ctx.encode_encrypt_batch(0, gi * group, scheme, stage[b], codec, ns, span,
                                     out=cts[gi * group:(gi + 1) * group], share_streams=bool(args.share_streams))
            consumed[b].record(main)
        ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
        ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)
        host_out.copy_(out, non_blocking=True)

    round_host()
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.e2e_steps):
        round_host()
    stop.record()
    barrier()
    t = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / args.e2e_steps
    h2d_total = int(n * span.total_len * 4)               # all ranks together, per step
    achieved = h2d_total / (ms * 1e-3) / 1e9
    return {"value": args.clients * args.elements / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "h2d_bytes_per_step": int(n * span.total_len * 4 // world), "d2h_bytes_per_step": int(span.total_len * 8 // world),
            "h2d_peak_gbs": probe["sum_gbs"], "h2d_probe": probe,
            "h2d_achieved_gbs_all_ranks": achieved, "frac_of_h2d_peak": achieved / probe["sum_gbs"],
            "bound": "host -> device copies (PCIe): %.1f GB per step over all ranks" % (h2d_total / 1e9),
            "api": "DeviceContext.encode_encrypt_batch/aggregate/decrypt_decode over the C ABI; pinned host float32 in (NUMA-local, "
                   "4-client groups over two copy streams, three staging buffers), float64 out",
            "steps": args.e2e_steps}


def cpu_baseline(args, ctx, fb):
    """BASELINE config 1 (1 M elements x 3 clients, int_bits 20, double masking) through the reference's CPU
    path on this box's host cores, timed — and the SAME run re-done on the device from the same inputs and
    the same rounding noise, every ciphertext, the aggregate, the decrypted integers and the decoded
    float64 compared with what the reference produced ("this config also is the CPU reference
    timing run")."""
    import numpy as np
    import torch
    cores = os.cpu_count() or 1
    elements, clients = min(args.cpu_sample_elements, args.elements), min(args.cpu_sample_clients, args.clients)
    bits = 20
    dt, phases, kind, kept = cpu_round(elements, clients, bits, cores, keep=True)
    out = {"value": clients * elements / dt, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": "BASELINE config 1: %d clients x %d elements, int_bits 20, one round (encode, encrypt, aggregate, decrypt, decode), "
                     "%s, multiprocessing.Pool(%d) per call as jzf_flashe.py does" % (
                         clients, elements, "the reference's own modules (oracle/_ref)" if kind == "reference" else "Python port", cores),
           "seconds": dt, "phases_s": phases}
    if kept is not None:
        xs, seeds, qs, cts, agg, dec, decoded = kept
        c20 = fb.DeviceContext(KEY, bits, ctx.device)
        span = fb.VectorSpan(elements, cores)
        codec = fb.CodecSpec(alpha=float(ALPHA), element_bits=16, n_clients=clients)
        us = []
        for sd in seeds:
            np.random.seed(sd)
            us.append(np.random.random(elements))
        x_d = torch.from_numpy(np.stack(xs)).to(ctx.device)
        u_d = torch.from_numpy(np.stack(us)).to(ctx.device)
        got_ct = c20.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x_d, codec, fb.NoiseSpec(u=u_d.reshape(-1)), span)
        got_agg = c20.aggregate(got_ct, fb.AGG_ELEMENTWISE)
        got_p = c20.empty_words(elements)
        got_out = c20.decrypt_decode(0, [clients], [0], got_agg, codec, span, p_out=got_p)
        torch.cuda.synchronize()
        want_ct = np.stack([np.asarray(ct, dtype=object).astype(np.uint32) for ct in cts])
        checks = {
            "ciphertexts": bool(np.array_equal(got_ct.cpu().numpy(), want_ct)),
            "aggregate": bool(np.array_equal(got_agg.cpu().numpy(), np.asarray(agg, dtype=object).astype(np.uint32))),
            "decrypted": bool(np.array_equal(got_p.cpu().numpy(), np.asarray(dec, dtype=object).astype(np.uint32))),
            "decoded_float64_bits": bool(np.array_equal(got_out.cpu().numpy().view(np.uint64),
                                                        np.asarray(decoded, dtype=np.float64).view(np.uint64))),
        }
        out["device_vs_reference_on_the_timed_run"] = "bit-exact" if all(checks.values()) else "MISMATCH"
        out["device_vs_reference_checks"] = checks
    return out


def main():
    args = parse_args()
    # stdout carries exactly one JSON line: anything a library prints there while we run (NCCL's version
    # banner, torchrun notices) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    real_stdout, sys.stdout = sys.stdout, os.fdopen(json_fd, "w")
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
        sys.stdout.flush()
    finally:
        sys.stdout = real_stdout


if __name__ == "__main__":
    main()
