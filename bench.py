#!/usr/bin/env python
"""bench.py — headline measurement of the polynomial-ring hot path on B200 (contract in the task statement, §④).

Default workload (BASELINE.json metric "FHEW/TFHE bootstraps/sec ...; NTT GB/s vs HBM roofline"):
  step  = one batch of FHEW NAND gate bootstraps (Fhew::op, scheme/fhew/src/fhew.rs:31-39) at the reference's own
          test parameter set FHEW-T (scheme/fhew/src/fhew/boolean.rs:225-239), `--batch` gates per GPU (default 16384),
          keys resident on the device, synthetic uniformly random key material / ciphertexts (timing does not depend
          on the values; parity is covered by tests/ and by the bit-exact sample check in the cpu_baseline leg).
  value = gates/s over all ranks, inputs resident in HBM (device pointers, fhe_fhew_bootstrap_batch).
  e2e   = the same through the host-slice C-ABI call fhe_fhew_bootstrap_batch_host (pinned host buffers, H2D + D2H
          inside the timed region).
  ntt   = BASELINE configs[1]: batched negacyclic NTT/iNTT sweep N = 2^10..2^16, 4096 polynomials, GB/s against the
          measured HBM copy peak (MEASURED_PEAKS.json) — extra keys on the same JSON line.
  --impl reference : the CPU restatement of the reference (oracle/liborc.so; the Rust reference cannot be built in
          this image) on all host cores, same metric/config, bounded sample per step.
Multi-GPU: one process per GPU (torchrun), batch sharded by ciphertext (weak scaling), no data-path collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fhew_nand_bootstraps_per_sec"
UNIT = "bootstraps/s"
FHEW_T_Q = 268409857  # first of two_adic_primes(28, 10) (boolean.rs:225-239)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []  # (host monotonic time, csv line)
        self.t0 = self.t1 = None

    def wait_ready(self, timeout=5.0):
        """Block until nvidia-smi has produced its first sample (it needs 0.1-0.5 s to start), so that the 20 ms sampling is
        already running when the timed region begins."""
        t_end = time.monotonic() + timeout
        while self.proc is not None and not self.lines and time.monotonic() < t_end:
            time.sleep(0.01)

    def begin(self):
        self.t0 = time.monotonic()

    def end(self):
        self.t1 = time.monotonic()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.monotonic(), ln.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        # the sampler runs from before the warm-up (nvidia-smi needs ~0.1 s to start); only samples taken between begin()
        # and end() - the timed region - are used, unless the region was too short to catch any
        inside = [ln for t, ln in self.lines if self.t0 is not None and self.t1 is not None and self.t0 <= t <= self.t1]
        note = None
        if not inside:
            inside = [ln for _, ln in self.lines]
            note = "timed region shorter than the sampling period: samples include the warm-up steps of the same workload"
        sm, mx, reasons, pw = [], [], set(), []
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "power_w_max": max(pw) if pw else None,
               "samples": len(sm), "reasons": sorted(reasons)}
        if note:
            out["note"] = note
        return out


def synth_fhew_key(param, seed):
    """Synthetic key material in the reference layout (uniform residues; bootstrapping.rs:92-99 shapes)."""
    n, q = param.n, param.big_q
    rng = np.random.default_rng(seed)
    ksk_a = rng.integers(0, param.q_ks, size=(n * param.ks_d, param.n_s), dtype=np.uint64)
    ksk_b = rng.integers(0, param.q_ks, size=(n * param.ks_d,), dtype=np.uint64)
    brk = rng.integers(0, q, size=(param.n_s, 2 * param.rgsw_d, 2, n), dtype=np.uint64)
    ak = rng.integers(0, q, size=(param.w + 1, param.rlwe_d, 2, n), dtype=np.uint64)
    g, m = 5, 2 * n
    ak_t = np.array([m - g] + [pow(g, v, m) for v in range(1, param.w + 1)], dtype=np.int64)
    return ksk_a, ksk_b, brk, ak, ak_t


def synth_cts(param, count, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, param.big_q, size=(count, param.n + 1), dtype=np.uint64)


def fhew_algorithmic_counts(param, steps_ext, steps_auto):
    """Closed-form integer work of one bootstrap in the fused dataflow: butterflies and MACs."""
    n, lg = param.n, param.log_n
    bf = (n // 2) * lg
    ext = steps_ext * ((2 * param.rgsw_d + 2) * bf)
    aut = steps_auto * ((param.rlwe_d + 2) * bf)
    mac = steps_ext * (2 * param.rgsw_d * 2 * n) + steps_auto * (param.rlwe_d * 2 * n)
    return ext + aut, mac


WORKLOAD = ("FHEW NAND gate bootstrap (Fhew::op), FHEW-T (boolean.rs:225-239): N=512 Q=268409857 RGSW/RLWE B=2^7 d=4, "
            "n_s=100 q_ks=2^16, w=10")


def fhew_config(batch, world, strong):
    """`config` of the JSON line, identical for the b200 arm and the reference arm of one invocation."""
    total = batch if strong else batch * world
    per = -(-total // world) if strong else batch
    ct_words = per * 513
    nset = max(2, int(np.ceil(160e6 / (2 * ct_words * 8))))
    return {"workload": WORKLOAD, "batch_per_gpu": per, "global_batch": total,
            "sharding": "by ciphertext, keys replicated (NCCL broadcast once)",
            "l2": "inputs rotate over %d distinct in/out sets (%.0f MB) > L2" % (nset, nset * 2 * ct_words * 8 / 1e6)}


def run_reference(args, rank, world):
    """CPU arm: the oracle (port of the reference's algorithm and dataflow: u128 % modmul, 3 transforms per product,
    coefficient-form keys) on all host cores.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import orc
    orc.build()
    orc.lib()
    cores = os.cpu_count() or 1
    P = orc.fhew_testing_param()
    K = orc.FhewKey(P, 0x5EED0001)
    sample = max(cores, 8 * cores if args.ref_sample is None else args.ref_sample)
    bits = np.random.default_rng(3).integers(0, 2, size=2 * sample).astype(np.int32)
    cts = K.encrypt(bits, 3)
    lin = (cts[:sample] + cts[sample:]) % np.uint64(P.big_q)
    for _ in range(args.warmup):  # W untimed warm-up steps, each one gate per host thread
        K.op([1, 1, 1, 0], lin[:cores], threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = K.op([1, 1, 1, 0], lin, threads=cores)
    dt = time.perf_counter() - t0
    assert (K.decrypt(out) == 1 - (bits[:sample] & bits[sample:])).all()
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.scaling == "strong" else "weak",
            "vs_baseline": None, "dtype": "u64 (u128 % modmul)", "data": "synthetic",
            "config": fhew_config(args.batch, world, args.scaling == "strong"),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "each step = %d gates of the workload (bounded sample) on %d threads, %d steps; CPU restatement of the "
                                       "reference (no Rust toolchain in the image)" % (sample, cores, args.steps)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def ncu_traffic(kernel, units):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py) when that capture processed the same number of units."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        rec = json.load(open(p)).get(kernel)
    except Exception:
        return None
    return rec["dram_bytes"] if rec and rec.get("units") == units else None


def ntt_sweep(pkg, ctx, torch, hbm_peak, reps, log_ns, batch):
    """BASELINE configs[1]: forward + inverse, in place, device resident; buffers rotate through a pool larger than L2."""
    from learn_fhe_b200 import util
    out = []
    pool_bytes = 4 << 30  # the largest case (N = 2^16, 4096 polys, u64) is 2 GiB; smaller cases rotate through the pool
    pool = torch.empty(pool_bytes // 8, dtype=torch.int64, device="cuda:%d" % ctx.device)
    stream = torch.cuda.current_stream(ctx.device)
    for log_n in log_ns:
        for bits, w in ((64, 8), (32, 4)):
            q = pkg.first_two_adic_prime(55 if bits == 64 else 28, log_n + 1)
            n = 1 << log_n
            words = batch * n
            view_words = words if bits == 64 else words // 2  # int64 words backing the u32 view
            nbuf = max(1, min(16, (pool_bytes // 8) // view_words))
            assert view_words <= pool_bytes // 8
            bufs = [pool[i * view_words:(i + 1) * view_words] for i in range(nbuf)]
            for b in bufs:  # valid residues: zero is fine for timing (data independent), but use a pattern
                b.random_(0, 1 << 27)
                if bits == 32:
                    b.bitwise_and_((((1 << 27) - 1) << 32) | ((1 << 27) - 1))
            name_f = "fhe_ntt_fwd_u64" if bits == 64 else "fhe_ntt_fwd_u32"
            name_i = "fhe_ntt_inv_u64" if bits == 64 else "fhe_ntt_inv_u32"
            res = {}
            for name, key in ((name_f, "fwd"), (name_i, "inv")):
                for i in range(3):
                    ctx.call(name, q, log_n, batch, pkg.dptr(bufs[i % nbuf]))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                stream.synchronize()
                e0.record(stream)
                for i in range(reps):
                    ctx.call(name, q, log_n, batch, pkg.dptr(bufs[i % nbuf]))
                e1.record(stream)
                e1.synchronize()
                ms = e0.elapsed_time(e1) / reps
                gbs = 2.0 * words * w / (ms * 1e-3) / 1e9
                res[key] = {"ms": ms, "gbs": gbs, "frac": gbs / hbm_peak}
            out.append({"log_n": log_n, "word_bits": bits, "batch": batch, "q": q, "buffers_rotated": nbuf,
                        "fwd_gbs": round(res["fwd"]["gbs"], 1), "inv_gbs": round(res["inv"]["gbs"], 1),
                        "fwd_frac_hbm": round(res["fwd"]["frac"], 4), "inv_frac_hbm": round(res["inv"]["frac"], 4),
                        "fwd_ms": round(res["fwd"]["ms"], 4), "inv_ms": round(res["inv"]["ms"], 4)})
    del pool
    return out


def device_ms(torch, stream, fn, warmup, steps):
    """Device time of `steps` calls of fn (CUDA events on the launching stream, after `warmup` untimed calls)."""
    for _ in range(warmup):
        fn()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1)


def wall_s(torch, dist, world, fn, warmup, steps):
    """Wall-clock seconds of `steps` calls of a host-slice (`_host`) entry point, barrier + synchronize on both sides."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    barrier()
    return time.perf_counter() - t0


def max_over_ranks(torch, dist, world, dev, x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def tfhe_param(pkg, synthetic_n1024):
    from learn_fhe_b200 import tfhe
    if synthetic_n1024:  # BASELINE configs[2] also names N=1024, for which the reference has no parameter set
        return pkg.TfheParam(log_p=2, padding=1, n=630, ks_log_b=2, ks_d=8, log_big_n=10, k=1, bs_log_b=7, bs_d=3)
    return tfhe.bootstrapping_testing_param()


def tfhe_leg(pkg, ctx, torch, dist, world, rank, local, batch, steps, synthetic_n1024=False, strong=False):
    """BASELINE configs[2]: TFHE programmable bootstrapping at the reference parameter set (tfhe/bootstrapping.rs:141-152:
    n=1024, N=2048, k=1, TGGSW B=2^23 d=1, key switch B=2^4 d=5), `batch` synthetic LWE ciphertexts per GPU (strong=True:
    `batch` in total, split contiguously over the ranks), keys uploaded on rank 0 and broadcast once.  Headline = mode 2, the
    fused bounded-error blind rotation (north_star: floating-point path within a stated bound, decryptions identical); the
    bit-identical reference dataflow (mode 0) is timed beside it.  Timing is value independent; parity of both modes is in
    tests/test_gpu_tfhe.py and re-checked on a sample in the cpu_baseline leg."""
    from learn_fhe_b200 import shard, tfhe
    dev = "cuda:%d" % local
    stream = torch.cuda.current_stream(local)
    P = tfhe_param(pkg, synthetic_n1024)
    n, N, k = P.n, P.big_n, P.k
    rng = np.random.default_rng(0x5EED0002)
    shapes = [(n, (k + 1) * P.bs_d, k + 1, N), (k * N * P.ks_d, n), (k * N * P.ks_d,)]
    if rank == 0:
        key_np = [rng.integers(0, 1 << 63, size=sh, dtype=np.uint64) for sh in shapes]
    else:
        key_np = [np.zeros(sh, dtype=np.uint64) for sh in shapes]
    bk = tfhe.BootstrappingKey(ctx, P, *key_np)
    del key_np
    if world > 1:
        bk.broadcast(dist, root=0)
    total = batch if strong else batch * world
    mine = (lambda lo_hi: lo_hi[1] - lo_hi[0])(shard.shard_range(total, rank, world)) if strong else batch
    lut_np = np.random.default_rng(5).integers(0, 1 << 63, size=N, dtype=np.uint64)
    lut = pkg.to_dev(lut_np, local)
    cts = torch.randint(-(1 << 62), 1 << 62, (mine, n + 1), dtype=torch.int64, device=dev)
    out = torch.empty_like(cts)
    step = lambda: tfhe.Bootstrapping.bootstrap_dev(bk, lut, cts, out)
    res = {"metric": "tfhe_pbs_per_sec", "unit": "PBS/s", "batch_per_gpu": mine, "global_batch": total, "steps": steps,
           "scaling": "strong" if strong else "weak",
           "config": ("synthetic, not from the reference: n=630 N=1024 k=1 B=2^7 d=3, ks B=2^2 d=8" if synthetic_n1024 else
                      "TFHE-T (tfhe/bootstrapping.rs:141-152): n=1024 N=2048 k=1 B=2^23 d=1, ks B=2^4 d=5"),
           "key_bytes": bk.nbytes}
    m, lg, nl = N // 2, (N // 2).bit_length() - 1, (k + 1) * P.bs_d
    pk = ctx.fp64_peak()
    for mode, name in ((3, "fused32"), (2, "fused"), (0, "bit_exact")):
        bk.set_mode(mode)
        step()  # warm-up (allocations, tables)
        ctx.prof_begin()
        ms = device_ms(torch, stream, step, 1, steps)
        prof = ctx.prof_end()
        ms = max_over_ranks(torch, dist, world, dev, ms)
        kname = "tfhe_blind_rotate_fast_kernel" if mode >= 2 else "tfhe_blind_rotate_kernel"
        br = prof.get(kname, {"ms": 0.0, "launches": 1})
        br_ms = br["ms"] / max(1, br["launches"])
        # algorithmic f64 work per PBS: n CMUX x [(k+1)d forward + I inverse] FFTs of N/2 points, 10 flops per radix-2 butterfly
        # (4 mul + 6 add) + pointwise products (6 flops) and sums (2); I = (k+1) when the products are summed in the Fourier
        # domain (mode 2), (k+1)^2 d in the reference dataflow (mode 0, every product rounded on its own)
        inv = (k + 1) if mode >= 2 else (k + 1) * nl
        flops = n * ((nl + inv) * (m // 2) * lg * 10 + (k + 1) * nl * m * 8)
        # peak: MEASURED on this device by fhe_diag_fp64_peak (csrc/diag.cu): DFMA counts two flops; the bit-identical mode may
        # not contract (the reference rounds every product and sum), so its binding rate is the DADD / DMUL issue rate
        peak = 2e12 * pk["dfma"] if mode >= 2 else 1e12 * min(pk["dadd"], pk["dmul"])
        ach = flops * mine / (br_ms * 1e-3) if br_ms else None
        r = {"value": total * steps / (ms * 1e-3), "unit": "PBS/s", "ms_per_step": ms / steps,
             "kernels": {kk: {"ms_per_launch": v["ms"] / v["launches"], "launches": v["launches"]} for kk, v in prof.items()},
             "roofline": {"bound": "fp64", "kernel": kname, "achieved": ach / 1e12 if ach else None, "peak": peak / 1e12,
                          "unit": "Tflop/s f64 (algorithmic; peak = measured %s rate of this device, csrc/diag.cu)" % ("DFMA x 2" if mode >= 2 else "DADD/DMUL"),
                          "frac": ach / peak if ach else None, "traffic": ncu_traffic(kname, mine) if mode != 2 else None, "flops_per_pbs": flops,
                          "fp64_peaks_tinstr": pk,
                          "ncu": ("profiles/r02_ncu_full_tfhe_fused_mode%d_b16384.csv: the busiest unit is the shared-memory / L1 data path "
                                  "(LSU wavefronts %s), FP64 pipe %s of issue slots" % ((3, "60 %", "46 %") if mode == 3 else (2, "74 %", "43 %"))) if mode >= 2 else "profiles/r01_ncu_full_tfhe_b16384.csv"}}
        if mode == 3:
            r["parity"] = ("same digits and exact sums as the reference; one rounding per output; |CMUX output - reference| < (k+1) d 2^(64+log_b+log_n-53) "
                           "+ 2^31 (accumulator words keep their top 32 bits: the f64 increments carry nothing below 2^35); decryptions identical")
            res.update(r)
            res["mode"] = "fused bounded-error blind rotation, 32-bit accumulator words (fhe_tfhe_key_set_mode 3)"
        elif mode == 2:
            r["parity"] = "as mode 3 with full 64-bit accumulator words (fhe_tfhe_key_set_mode 2)"
            res["fused_u64_accumulator"] = r
        else:
            r["parity"] = "raw torus words bit-identical to the reference dataflow"
            res["bit_exact_mode"] = r
    # e2e: the host-slice C-ABI call with pinned host buffers, H2D + D2H inside the timed region (the headline mode)
    bk.set_mode(3)
    h_in = torch.empty((mine, n + 1), dtype=torch.int64).pin_memory()
    h_out = torch.empty((mine, n + 1), dtype=torch.int64).pin_memory()
    h_in.copy_(cts)
    h_lut = torch.from_numpy(lut_np.view(np.int64)).pin_memory()
    host = lambda: ctx.call("fhe_tfhe_pbs_batch_host", bk.h, pkg.hptr(h_lut.numpy()), mine, pkg.hptr(h_in.numpy()), pkg.hptr(h_out.numpy()))
    sec = max_over_ranks(torch, dist, world, dev, wall_s(torch, dist, world, host, 1, steps))
    res["e2e"] = {"value": total * steps / sec, "unit": "PBS/s", "h2d_bytes_per_step": int(mine * (n + 1) * 8 + N * 8),
                  "d2h_bytes_per_step": int(mine * (n + 1) * 8), "api": "fhe_tfhe_pbs_batch_host"}
    step()
    torch.cuda.synchronize()
    assert np.array_equal(pkg.to_host(out), h_out.numpy().view(np.uint64)), "TFHE device and host paths disagree"
    bk.free()
    del cts, out
    torch.cuda.empty_cache()
    return res


def tfhe_cpu_leg(pkg, ctx, torch, local, cores, synthetic_n1024=False):
    """cpu_baseline of the TFHE leg: the oracle's PBS (reference dataflow, f64 FFT of c64.rs) on all host threads, on real
    encryptions under a real key; the same ciphertexts then go through the GPU: mode 0 must be bit-identical, mode 2 must
    decrypt identically."""
    from oracle import orc
    from learn_fhe_b200 import tfhe
    P = orc.tfhe_testing_param()
    pp = tfhe_param(pkg, synthetic_n1024)
    P.n, P.big_n, P.k, P.bs_log_b, P.bs_d, P.ks_log_b, P.ks_d, P.log_p, P.padding = (pp.n, pp.big_n, pp.k, pp.bs_log_b, pp.bs_d,
                                                                                   pp.ks_log_b, pp.ks_d, pp.log_p, pp.padding)
    K = orc.TfheKey(P, 0x5EED0003)
    ex = K.export()
    sample = 3 * cores
    msgs = (np.arange(sample, dtype=np.uint64) * np.uint64(7)) % np.uint64(1 << P.log_p)
    cts = K.encrypt(msgs, 11)
    table = ((3 * np.arange(1 << P.log_p) + 1) % (1 << P.log_p)).astype(np.uint64)
    v = K.lut_poly(table)
    t0 = time.perf_counter()
    ref = K.bootstrap(v, cts, threads=cores)
    dt = time.perf_counter() - t0
    bk = tfhe.BootstrappingKey(ctx, pp, ex["brk"], ex["ksk_a"], ex["ksk_b"])
    lut = tfhe.encode_lut(pp, v)
    exact = tfhe.Bootstrapping.bootstrap(bk, lut, cts)
    assert np.array_equal(exact, ref), "TFHE mode 0 differs from the oracle"
    want = table[msgs.astype(np.int64)]
    assert np.array_equal(K.decrypt(ref)[0], want)
    for mode in (2, 3):
        bk.set_mode(mode)
        fused = tfhe.Bootstrapping.bootstrap(bk, lut, cts)
        assert np.array_equal(K.decrypt(fused)[0], want), "TFHE mode %d decrypts differently" % mode
    bk.free()
    return {"value": sample / dt, "unit": "PBS/s", "cores": cores, "kind": "port",
            "sample": "%d PBS of the same parameter set on %d threads (%.1f s); GPU mode 0 bit-identical, modes 2 and 3 decryptions identical" % (sample, cores, dt)}


def ckks_leg(pkg, ctx, torch, dist, world, rank, local, batch, steps):
    """BASELINE configs[3]: Ckks::mul (tensor product + relinearise + rescale, ckks.rs:255-272) at N=2^16, log_qi=55, L=8,
    full level, `batch` ciphertext pairs per GPU.  Parity (bit-exact vs the oracle) is in tests/test_gpu_ckks.py."""
    from learn_fhe_b200 import ckks
    dev = "cuda:%d" % local
    stream = torch.cuda.current_stream(local)
    log_n, L = 16, 8
    P = ckks.CkksParam.new(ctx, log_n, 55, L)
    rng = np.random.default_rng(0x5EED0003)
    if rank == 0:  # the relinearisation key exists on rank 0 only and reaches the other ranks by one NCCL broadcast
        ksk = np.stack([np.stack([rng.integers(0, q, size=P.n, dtype=np.uint64) for q in P.qs + P.ps]) for _ in range(2)])
    else:
        ksk = np.zeros((2, 2 * L, P.n), dtype=np.uint64)
    rlk = ckks.CkksKeySwitchingKey(P, ksk)
    del ksk
    if world > 1:
        rlk.broadcast(dist, root=0)

    def rand_ct():
        t = torch.empty((batch, 2, L, P.n), dtype=torch.int64, device=dev)
        for i, q in enumerate(P.qs):
            t[:, :, i, :].random_(0, q)
        return t

    ct0, ct1 = rand_ct(), rand_ct()
    out = torch.empty((batch, 2, L - 1, P.n), dtype=torch.int64, device=dev)
    for _ in range(2):
        ckks.Ckks.mul_dev(P, rlk, L, ct0, ct1, out)  # warm-up (workspace allocation, tables)
    ctx.prof_begin()
    ms = device_ms(torch, stream, lambda: ckks.Ckks.mul_dev(P, rlk, L, ct0, ct1, out), 0, steps)
    prof = ctx.prof_end()
    ms = max_over_ranks(torch, dist, world, dev, ms)
    io_bytes = batch * (2 * 2 * L + 2 * (L - 1)) * P.n * 8
    res = {"metric": "ckks_mul_relin_rescale_per_sec", "value": batch * world * steps / (ms * 1e-3), "unit": "mult/s",
           "batch_per_gpu": batch, "steps": steps, "ms_per_step": ms / steps,
           "config": "CKKS-T: N=2^16, log_qi=55, L=8 (+8 special primes), level 8 -> 7, rlk resident",
           "compulsory_hbm_gbs": io_bytes * steps / (ms * 1e-3) / 1e9, "ntt_per_mult": 7 * L + 3 * L,
           "kernels": {kk: {"ms_per_step": v["ms"] / steps, "launches": v["launches"]} for kk, v in prof.items()}}
    # INT32-pipe roofline of the whole operation (same instruction-mix rule as the NTT lines: 9 32-bit products per 64-bit
    # Shoup butterfly or modular multiply).  Work per Ckks::mul (ckks.rs:255-293, rns.rs:99-158), l = L:
    #   (7l + 3L) NTTs of N/2 log N butterflies (4l forward of the inputs, l inverse + L forward around the base conversion of d2,
    #   2(l + L) inverse of the key products with P (d0, d1) folded in); tensor 4 l N; key products 2 (l + L) N + fold 2 l N; base conversion l -> L of d2:
    #   (l + l L) N; rescale_k by the L special primes of 2 polynomials: 2 (L + L l + l) N; final rescale: 2 (1 + 2 (l - 1)) N
    pk = ctx.int32_peak()
    if pk.get("imad"):
        l = L
        n_bf = (7 * l + 3 * L) * (P.n // 2) * log_n
        n_mm = (4 * l + 2 * (l + L) + 2 * l + (l + l * L) + 2 * (L + L * l + l) + 2 * (1 + 2 * (l - 1))) * P.n
        per = (4 / pk["imad"] + 2 / pk["imad_hi"] + 3 / pk["imad_wide"]) / 1e12
        t_min = (n_bf + n_mm) * per * batch
        res["roofline"] = {"bound": "int32", "kernel": "whole Ckks::mul (ntt_fast_* 60 %, rns_rescale 20 %)", "achieved": 9 * (n_bf + n_mm) * batch / (ms / steps * 1e-3) / 1e12,
                           "peak": 9 / per / 1e12, "unit": "Tmul/s (algorithmic 32-bit products; peak = measured INT32 multiply-pipe rate for 3 IMAD.WIDE + 2 IMAD.HI + 4 IMAD)",
                           "frac": t_min / (ms / steps * 1e-3), "traffic": ncu_traffic("ckks_mul_whole_op", batch),
                           "traffic_note": "sum of dram read + write bytes over every launch of one Ckks::mul batch (profiles/ncu_traffic.json)",
                           "algorithmic_bytes": io_bytes, "butterflies_per_mult": n_bf, "pointwise_modmuls_per_mult": n_mm}
    # e2e: host-slice C-ABI call, pinned host ciphertexts in and out (bounded to 64 pairs: 1 GiB in + 0.44 GiB out per step)
    eb = min(batch, 64)
    h0 = torch.empty((eb, 2, L, P.n), dtype=torch.int64).pin_memory()
    h1 = torch.empty((eb, 2, L, P.n), dtype=torch.int64).pin_memory()
    ho = torch.empty((eb, 2, L - 1, P.n), dtype=torch.int64).pin_memory()
    h0.copy_(ct0[:eb])
    h1.copy_(ct1[:eb])
    host = lambda: ctx.call("fhe_ckks_mul_relin_rescale_batch_host", P.h, rlk.h, L, eb, pkg.hptr(h0.numpy()), pkg.hptr(h1.numpy()), pkg.hptr(ho.numpy()))
    sec = max_over_ranks(torch, dist, world, dev, wall_s(torch, dist, world, host, 1, steps))
    res["e2e"] = {"value": eb * world * steps / sec, "unit": "mult/s", "batch_per_gpu": eb, "h2d_bytes_per_step": int(h0.numel() * 16),
                  "d2h_bytes_per_step": int(ho.numel() * 8), "api": "fhe_ckks_mul_relin_rescale_batch_host"}
    torch.cuda.synchronize()
    assert torch.equal(ho, out[:eb].cpu()), "CKKS device and host paths disagree"
    del h0, h1, ho
    rlk.free()
    P.free()
    del ct0, ct1, out
    torch.cuda.empty_cache()
    return res


def ckks_cpu_leg(pkg, ctx, cores):
    """cpu_baseline of the CKKS leg: the oracle's Ckks::mul (reference dataflow: coefficient-form products of 3 transforms
    each, u128 % arithmetic) at N=2^16, L=8 on all host threads, one pair per thread; the same pairs then go through the GPU
    path and must come back bit-identical."""
    from oracle import orc
    from learn_fhe_b200 import ckks
    log_n, L = 16, 8
    K = orc.CkksKey(log_n, 55, L, 0x5EED0004)
    n = 1 << log_n
    c0 = np.stack([K.encrypt((np.arange(n, dtype=np.int64) * (i + 3)) % 17 - 8, L, 70 + i) for i in range(cores)])
    c1 = np.stack([K.encrypt((np.arange(n, dtype=np.int64) * (i + 5)) % 5 - 2, L, 90 + i) for i in range(cores)])
    t0 = time.perf_counter()
    ref = K.mul(c0, c1, threads=cores)
    dt = time.perf_counter() - t0
    P = ckks.CkksParam(ctx, log_n, K.qs, K.ps)
    rlk = ckks.CkksKeySwitchingKey(P, K.ksk(-1))
    got = ckks.Ckks.mul(P, rlk, c0, c1)
    assert np.array_equal(got, ref), "CKKS GPU output differs from the oracle"
    rlk.free()
    P.free()
    return {"value": cores / dt, "unit": "mult/s", "cores": cores, "kind": "port",
            "sample": "%d ciphertext pairs at N=2^16, L=8 on %d threads (%.1f s); GPU outputs bit-identical" % (cores, cores, dt)}


def ntt_cpu_leg(pkg, ctx, cores, log_ns):
    """cpu_baseline of the NTT sweep: the oracle's radix-2 in-place transform (util/src/ring/fft.rs:40-54, u128 % butterflies)
    on all host threads, 16 polynomials per thread per size; the GPU transform of the same data must be bit-identical."""
    from oracle import orc
    from learn_fhe_b200 import util
    out = {}
    for log_n in log_ns:
        q = pkg.first_two_adic_prime(55, log_n + 1)
        n, polys = 1 << log_n, 16 * cores
        a = orc.residues(0x5EED0000 + log_n, polys * n, q).reshape(polys, n)
        t0 = time.perf_counter()
        ref = orc.ntt_fwd(q, a, threads=cores)
        dt = time.perf_counter() - t0
        x = a.copy()
        util.nega_cyclic_ntt_in_place(ctx, q, x)
        assert np.array_equal(x, ref), "GPU NTT differs from the oracle at log_n %d" % log_n
        out[log_n] = {"value": 2.0 * polys * n * 8 / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                      "sample": "%d polynomials on %d threads (%.3f s); GPU output bit-identical" % (polys, cores, dt)}
    return out


def ntt_e2e(pkg, ctx, torch, log_n, batch, reps):
    """fhe_ntt_fwd_host on pinned host polynomials (H2D + transform + D2H inside the timed region), u64 words."""
    q = pkg.first_two_adic_prime(55, log_n + 1)
    n = 1 << log_n
    h = torch.empty((batch, n), dtype=torch.int64).pin_memory()
    h.random_(0, 1 << 27)
    fn = lambda: ctx.call("fhe_ntt_fwd_host", q, pkg.hptr(h.numpy()), n, batch)
    sec = wall_s(torch, None, 1, fn, 1, reps)
    return {"value": 2.0 * batch * n * 8 * reps / sec / 1e9, "unit": "GB/s", "log_n": log_n, "batch": batch, "h2d_bytes_per_step": batch * n * 8,
            "d2h_bytes_per_step": batch * n * 8, "api": "fhe_ntt_fwd_host"}


def ckks_s2c_leg(pkg, ctx, torch, dist, world, rank, local, count, steps):
    """BASELINE configs[4], the part the reference implements (scheme/ckks/src/bootstrapping.rs:73-108; EvalMod / ModRaise do not
    exist there): Bootstrapping::slot_to_coeff at N = 2^16, L = 8, r = 3 - five grouped factor matrices, level 8 -> 3 - on `count`
    ciphertexts per GPU.  Plans and diagonals are computed once on the host (complex128 encodings: timing only; the 256-bit path
    and bit-exact parity are in tests/test_gpu_ckks_boot.py at log_n <= 9), the rotation keys exist on rank 0 and reach the other
    ranks by NCCL broadcast."""
    from learn_fhe_b200 import ckks, ckks_bootstrapping as cb
    dev = "cuda:%d" % local
    stream = torch.cuda.current_stream(local)
    log_n, L = 16, 8
    P = ckks.CkksParam.new(ctx, log_n, 55, L)
    t0 = time.perf_counter()
    bp = cb.BootstrappingParam(P, 3, "f64")
    rng = np.random.default_rng(0x5EED0006)
    mods = P.qs + P.ps

    def ksk_for(j):
        if rank != 0:
            return np.zeros((2, 2 * L, P.n), dtype=np.uint64)
        return np.stack([np.stack([rng.integers(0, m, size=P.n, dtype=np.uint64) for m in mods]) for _ in range(2)])

    bk = cb.BootstrappingKey(bp, ksk_for, chains=("sfft",))
    if world > 1:
        bk.broadcast_keys(dist, root=0)
    ct = torch.empty((count, 2, L, P.n), dtype=torch.int64, device=dev)
    for i, m in enumerate(P.qs):
        ct[:, :, i, :].random_(0, m)
    out = cb.Bootstrapping.chain_dev(bk, "sfft", ct)  # warm-up: uploads the encoded diagonals of the five plans
    setup_s = time.perf_counter() - t0
    ms = device_ms(torch, stream, lambda: cb.Bootstrapping.chain_dev(bk, "sfft", ct), 1, steps)
    ms = max_over_ranks(torch, dist, world, dev, ms)
    plans = [bk.plan("sfft", i, L - (len(bp.sfft_fmats) - 1 - i)) for i in range(len(bp.sfft_fmats))]
    res = {"metric": "ckks_slot_to_coeff_per_sec", "value": count * world * steps / (ms * 1e-3), "unit": "ciphertexts/s", "ciphertexts_per_gpu": count,
           "ms_per_step": ms / steps, "config": "N=2^16, L=8, r=3: %d grouped matrices, level %d -> %d; synthetic keys, complex128-encoded diagonals" % (
               len(bp.sfft_fmats), L, out.shape[2]),
           "rotation_keys": len(bk.rtk), "rotation_key_bytes": bk.nbytes, "keys": "rank 0, NCCL broadcast" if world > 1 else "single GPU",
           "rotations_per_ciphertext": sum(len(p["baby"]) - 1 + len(p["giant"]) - 1 for p in plans),
           "plain_mults_per_ciphertext": int(sum(p["present"].sum() for p in plans)), "host_setup_s": setup_s}
    bk.free()
    P.free()
    del ct, out
    torch.cuda.empty_cache()
    return res


def next_rows_leg(pkg, ctx, torch, bk, param, local):
    """Throughput of the SURVEY.md §8(f) rows built so far, one GPU (rank 0), small fixed sizes; parity of each is in tests/."""
    from learn_fhe_b200 import ckks, circuits, fhew
    dev = "cuda:%d" % local
    stream = torch.cuda.current_stream(local)
    res = {}
    # rank 1: FhewU8::wrapping_mul (uint8.rs:123-135) on a vector of 512 encrypted bytes, level-batched gate DAG
    B = 512

    def u8_mul():
        eng = circuits.GateEngine(bk)
        a = circuits.FhewU8.from_ciphertexts(eng, xa)
        b = circuits.FhewU8.from_ciphertexts(eng, xb)
        prod = a * b
        eng.evaluate([bit.node for bit in prod.bits])
        return eng

    xa = [pkg.to_dev(synth_cts(param, B, 900 + i), local) for i in range(8)]
    xb = [pkg.to_dev(synth_cts(param, B, 950 + i), local) for i in range(8)]
    eng = u8_mul()
    ms = device_ms(torch, stream, u8_mul, 0, 2) / 2
    res["fhew_u8_wrapping_mul"] = {"bytes_per_call": B, "gate_bootstraps_per_call": eng.gates, "bootstrap_batches_per_call": eng.launches,
                                   "ms_per_call": ms, "u8_mul_per_sec": B / (ms * 1e-3), "gate_bootstraps_per_sec": eng.gates / (ms * 1e-3)}
    del xa, xb, eng
    # rank 4: the 64-bit-modulus FHEW path at the multi-key parameter size (examples/multi_key_uint8.rs:15-29), synthetic key
    q = pkg.first_two_adic_prime(55, 12)
    wp = pkg.FhewParam(log_n=11, big_q=q, p=4, rlwe_log_b=11, rlwe_d=5, rgsw_log_b=11, rgsw_d=5, n_s=600, q_ks=1 << 20, ks_log_b=4, ks_d=5, w=10)
    wk = fhew.BootstrappingKey(ctx, wp, *synth_fhew_key(wp, 5))
    wf = pkg.to_dev(fhew.gate_poly(wp, [1, 1, 1, 0]), local)
    wb = 2 * ctx.sm_count
    win = pkg.to_dev(synth_cts(wp, wb, 6), local)
    wout = torch.empty_like(win)
    step = lambda: fhew.Bootstrapping.bootstrap_dev(wk, wf, win, wout, post_add=fhew.big_q_by_8(wp))
    step()
    ms = device_ms(torch, stream, step, 0, 1)
    res["fhew_64bit_modulus"] = {"config": "N=2048, 55-bit Q, decomposors (11,5), LWE n=600 q=2^20; synthetic key; generic kernels", "batch": wb,
                                 "ms_per_step": ms, "gates_per_sec": wb / (ms * 1e-3), "key_bytes": wk.nbytes}
    wk.free()
    del win, wout
    # rank 2: Bootstrapping::mul_mat (ckks/bootstrapping.rs:92-108) at N = 2^16, level 8: 8 baby x 4 giant steps, dense
    log_n, L, count, nb, ng = 16, 8, 16, 8, 4
    P = ckks.CkksParam.new(ctx, log_n, 55, L)
    rng = np.random.default_rng(0x5EED0005)
    mk = lambda: ckks.CkksKeySwitchingKey(P, np.stack([np.stack([rng.integers(0, m, size=P.n, dtype=np.uint64) for m in P.qs + P.ps]) for _ in range(2)]))
    keys = [mk() for _ in range(nb - 1 + ng - 1)]
    baby = [(0, None)] + [(pow(5, j, 2 * P.n), keys[j - 1]) for j in range(1, nb)]
    giant = [(0, None)] + [(pow(5, nb * i, 2 * P.n), keys[nb - 1 + i - 1]) for i in range(1, ng)]
    rot = lambda lst: (pkg.CkksRot * len(lst))(*[pkg.CkksRot(t, k.h if k is not None else None) for t, k in lst])
    b_arr, g_arr = rot(baby), rot(giant)
    present = np.ones((ng, nb), dtype=np.uint8)
    pts = torch.empty((ng * nb, L, P.n), dtype=torch.int64, device=dev)
    ct = torch.empty((count, 2, L, P.n), dtype=torch.int64, device=dev)
    for i, m in enumerate(P.qs):
        pts[:, i, :].random_(0, m)
        ct[:, :, i, :].random_(0, m)
    out = torch.empty((count, 2, L - 1, P.n), dtype=torch.int64, device=dev)
    import ctypes as C
    step = lambda: ctx.call("fhe_ckks_mul_mat", P.h, L, count, nb, C.cast(b_arr, C.c_void_p), ng, C.cast(g_arr, C.c_void_p), pkg.hptr(present),
                            pkg.dptr(pts), pkg.dptr(ct), pkg.dptr(out))
    step()
    ms = device_ms(torch, stream, step, 0, 2) / 2
    res["ckks_mul_mat"] = {"config": "N=2^16, level 8 -> 7, dense 32-diagonal BSGS (8 baby x 4 giant), 10 rotation keys resident, synthetic", "ciphertexts": count,
                           "ms_per_call": ms, "matrix_vector_products_per_sec": count / (ms * 1e-3),
                           "rotations_per_call": count * (nb - 1 + ng - 1), "plain_mults_per_call": count * nb * ng}
    for k in keys:
        k.free()
    P.free()
    del pts, ct, out
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="gate bootstraps per GPU per step")
    ap.add_argument("--ref-sample", type=int, default=None, help="gates per step of the reference arm (default 8 x cores)")
    ap.add_argument("--no-ntt", action="store_true", help="skip the NTT sweep leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ntt-reps", type=int, default=20)
    ap.add_argument("--no-tfhe", action="store_true", help="skip the TFHE PBS leg (BASELINE configs[2])")
    ap.add_argument("--no-ckks", action="store_true", help="skip the CKKS hom-mult leg (BASELINE configs[3])")
    ap.add_argument("--no-next", action="store_true", help="skip the SURVEY 8(f) legs (u8 circuits, 64-bit FHEW, CKKS mul_mat)")
    ap.add_argument("--no-strong", action="store_true", help="multi-GPU runs: skip the second (other) scaling curve")
    ap.add_argument("--tfhe-batch", type=int, default=16384, help="PBS per GPU per step")
    ap.add_argument("--ckks-batch", type=int, default=512, help="ciphertext pairs per GPU per step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch gates per GPU; strong: --batch gates in total, split contiguously over the ranks (configs[2]: 16384/R)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import _pkg
    pkg = _pkg.load_package()
    from learn_fhe_b200 import fhew
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG is left to the caller (default WARN); its log goes to stderr so that stdout stays the one JSON line
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    hbm_peak, peak_src, peak_json = peaks()
    ctx = pkg.Context(local)
    ctx.use_torch_stream()
    stream = torch.cuda.current_stream(local)

    param = fhew.single_key_testing_param(FHEW_T_Q)
    # keys: generated on rank 0, uploaded + transformed there, then broadcast once over NCCL
    key_np = synth_fhew_key(param, 0x5EED0000)
    if rank == 0:
        bk = fhew.BootstrappingKey(ctx, param, *key_np)
    else:  # placeholder of the right shape; contents arrive by broadcast
        bk = fhew.BootstrappingKey(ctx, param, *[np.zeros_like(x) if i < 4 else x for i, x in enumerate(key_np)])
    if world > 1:
        bk.broadcast(dist, root=0)
    table = [1, 1, 1, 0]
    f_np = fhew.gate_poly(param, table)
    post = fhew.big_q_by_8(param)
    f_dev = pkg.to_dev(f_np, local)
    from learn_fhe_b200 import shard
    strong = args.scaling == "strong"
    total = args.batch if strong else args.batch * world
    lo, hi = shard.shard_range(total, rank, world) if strong else (0, args.batch)
    B = hi - lo  # gates of this rank per step
    ct_words = B * (param.n + 1)
    # rotate over enough distinct input/output sets to exceed L2 (126 MB)
    nset = max(2, int(np.ceil(160e6 / (2 * ct_words * 8))))
    ins = [pkg.to_dev(synth_cts(param, B, 1000 * rank + i), local) for i in range(nset)]
    outs = [torch.empty_like(ins[0]) for _ in range(nset)]

    def step(i):
        fhew.Bootstrapping.bootstrap_dev(bk, f_dev, ins[i % nset], outs[i % nset], post_add=post)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler.begin()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ker_evs = []
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    sampler.end()
    launches = ctx.launches - l0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = total * args.steps / (ms_total * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # dominant kernel timed alone on its stream (blind rotation): live CUDA events around the kernel launch only
    kt = bk.time_kernels(f_dev, ins[0], outs[0], post, reps=max(2, args.steps))

    # e2e: host-slice C ABI with pinned host buffers
    h_in = torch.empty((B, param.n + 1), dtype=torch.int64).pin_memory()
    h_out = torch.empty((B, param.n + 1), dtype=torch.int64).pin_memory()
    h_in.numpy().view(np.uint64)[:] = synth_cts(param, B, 77 + rank)
    h_f = torch.from_numpy(f_np.view(np.int64)).pin_memory()

    def step_host():
        ctx.call("fhe_fhew_bootstrap_batch_host", bk.h, pkg.hptr(h_f.numpy()), post, B, pkg.hptr(h_in.numpy()), pkg.hptr(h_out.numpy()))

    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total * args.steps / float(t.item())
    # device-path outputs must equal host-path outputs on the same inputs (consistency, cheap)
    chk_in = pkg.to_dev(h_in.numpy().view(np.uint64)[:64].copy(), local)
    chk_out = torch.empty_like(chk_in)
    fhew.Bootstrapping.bootstrap_dev(bk, f_dev, chk_in, chk_out, post_add=post)
    torch.cuda.synchronize()
    assert np.array_equal(pkg.to_host(chk_out), h_out.numpy().view(np.uint64)[:64]), "device and host paths disagree"
    if world > 1:  # every rank must hold rank 0's key: same probe input -> same output everywhere
        probe = pkg.to_dev(synth_cts(param, 8, 4242), local)
        pout = torch.empty_like(probe)
        fhew.Bootstrapping.bootstrap_dev(bk, f_dev, probe, pout, post_add=post)
        torch.cuda.synchronize()
        gathered = [torch.empty_like(pout) for _ in range(world)]
        dist.all_gather(gathered, pout)
        assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree after key broadcast"

    ntt = None
    if rank == 0 and not args.no_ntt:
        ntt = ntt_sweep(pkg, ctx, torch, hbm_peak, args.ntt_reps, list(range(10, 17)), 4096)
        # the configured 4096 polynomials are 16-68 us launches for N <= 2^12 (one partly filled wave: launch, ramp and tail are a
        # third of the time); the same kernels on 65 536 polynomials show their steady-state rate
        ntt_large = ntt_sweep(pkg, ctx, torch, hbm_peak, max(3, args.ntt_reps // 4), [10, 11, 12], 65536)
        ntt_host = ntt_e2e(pkg, ctx, torch, 16, 1024, 3)

    # free the FHEW batch buffers before the wider legs
    del ins, outs
    torch.cuda.empty_cache()
    tsteps = max(1, min(args.steps, 2))
    tfhe_res = None if args.no_tfhe else tfhe_leg(pkg, ctx, torch, dist, world, rank, local, args.tfhe_batch, tsteps, False, strong)
    tfhe_res_1024 = None if args.no_tfhe else tfhe_leg(pkg, ctx, torch, dist, world, rank, local, args.tfhe_batch, tsteps, True, strong)
    ckks_res = None if args.no_ckks else ckks_leg(pkg, ctx, torch, dist, world, rank, local, args.ckks_batch, max(1, min(args.steps, 3)))
    s2c_res = None if (args.no_ckks or args.no_next) else ckks_s2c_leg(pkg, ctx, torch, dist, world, rank, local, 8, 2)

    # the other scaling curve (configs[2]: the 16 384 batch split 16 384 / R): measured in the same run when there is more than one
    # rank, so that one driver sweep over N yields both the weak and the strong line
    other = None
    if world > 1 and not args.no_strong:
        o_strong = not strong
        o_total = args.batch if o_strong else args.batch * world
        olo, ohi = shard.shard_range(o_total, rank, world) if o_strong else (0, args.batch)
        ob = ohi - olo
        o_in = pkg.to_dev(synth_cts(param, ob, 5000 + rank), local)
        o_out = torch.empty_like(o_in)
        ostep = lambda: fhew.Bootstrapping.bootstrap_dev(bk, f_dev, o_in, o_out, post_add=post)
        for _ in range(args.warmup):
            ostep()
        barrier()
        oms = max_over_ranks(torch, dist, world, dev, device_ms(torch, stream, ostep, 0, args.steps))
        other = {"scaling": "strong" if o_strong else "weak", "metric": METRIC, "value": o_total * args.steps / (oms * 1e-3), "unit": UNIT,
                 "global_batch": o_total, "batch_per_gpu": ob, "ms_per_step": oms / args.steps,
                 "resident_ctas_per_gpu": 7 * ctx.sm_count,
                 "note": "strong scaling is limited by the partial last wave: %d gates on %d resident CTAs per GPU = %.2f waves" % (ob, 7 * ctx.sm_count, ob / (7.0 * ctx.sm_count))}
        if not args.no_tfhe:
            t_o = tfhe_leg(pkg, ctx, torch, dist, world, rank, local, args.tfhe_batch, tsteps, False, o_strong)
            other["tfhe_pbs"] = {kk: t_o[kk] for kk in ("value", "unit", "global_batch", "batch_per_gpu", "ms_per_step", "scaling")}
        del o_in, o_out

    next_res = None
    if rank == 0 and not args.no_next:
        next_res = next_rows_leg(pkg, ctx, torch, bk, param, local)

    cpu = None
    if rank == 0 and not args.no_cpu:
        from oracle import orc  # cpu_baseline leg: the oracle as the timed CPU port + bit-exact checker of a GPU sample
        orc.build()
        cores = os.cpu_count() or 1
        P = orc.fhew_testing_param()
        K = orc.FhewKey.from_arrays(P, *key_np) if hasattr(orc.FhewKey, "from_arrays") else None
        sample = 40 * cores  # about 10 s of CPU work at ~4 gates/s/thread
        if K is not None:
            lin = h_in.numpy().view(np.uint64)[:sample].copy()
            t0 = time.perf_counter()
            ref = K.op(table, lin, threads=cores)
            dt = time.perf_counter() - t0
            assert np.array_equal(ref, h_out.numpy().view(np.uint64)[:sample]), "GPU output differs from the oracle"
            checked = True
        else:
            K = orc.FhewKey(P, 0x5EED0001)
            bits = np.random.default_rng(3).integers(0, 2, size=2 * sample).astype(np.int32)
            cts = K.encrypt(bits, 3)
            lin = (cts[:sample] + cts[sample:]) % np.uint64(P.big_q)
            t0 = time.perf_counter()
            K.op(table, lin, threads=cores)
            dt = time.perf_counter() - t0
            checked = False
        cpu = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d gates of the same workload on %d threads (%.1f s)%s" % (sample, cores, dt, ", GPU outputs bit-identical" if checked else "")}
        # the CPU path beside every other parameter set, same run, same box (north_star); each also re-checks GPU parity on its sample
        if tfhe_res is not None:
            tfhe_res["cpu_baseline"] = tfhe_cpu_leg(pkg, ctx, torch, local, cores, False)
            tfhe_res_1024["cpu_baseline"] = tfhe_cpu_leg(pkg, ctx, torch, local, cores, True)
        if ckks_res is not None:
            ckks_res["cpu_baseline"] = ckks_cpu_leg(pkg, ctx, cores)
        if ntt is not None:
            ntt_cpu = ntt_cpu_leg(pkg, ctx, cores, [r["log_n"] for r in ntt if r["word_bits"] == 64])
            for r in ntt:
                r["cpu_baseline"] = ntt_cpu[r["log_n"]]

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "u32 (Q < 2^30 residues, u64 MAC accumulators)", "data": "synthetic",
                "config": fhew_config(args.batch, world, strong),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ct_words * 8 + param.n * 8),
                        "d2h_bytes_per_step": int(ct_words * 8)},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": dict(kt["roofline"], traffic=ncu_traffic("fhew_blind_rotate_fast_kernel", B)), "kernels": kt["kernels"], "cpu_baseline": cpu, "peak_source": peak_src}
        if tfhe_res is not None:
            line["tfhe_pbs"] = tfhe_res
            line["tfhe_pbs_n1024_synthetic"] = tfhe_res_1024
        if ckks_res is not None:
            line["ckks_mul"] = ckks_res
        if s2c_res is not None:
            line["ckks_slot_to_coeff"] = s2c_res
        if other is not None:
            line["other_scaling"] = other
        if next_res is not None:
            line["next_rows"] = next_res
        if ntt is not None:
            line["ntt"] = ntt
            best = max(ntt, key=lambda r: (r["log_n"], r["word_bits"]))
            line["roofline_ntt"] = {"bound": "hbm", "achieved": best["fwd_gbs"], "peak": hbm_peak, "unit": "GB/s",
                                    "frac": best["fwd_frac_hbm"], "traffic": ncu_traffic("ntt_fwd_u64_2^16", 4096),
                                    "algorithmic_bytes": 2 * 4096 * (1 << best["log_n"]) * 8,
                                    "kernel": "ntt fwd u64 N=2^%d batch 4096" % best["log_n"], "e2e": ntt_host}
            # the transforms are bound by the INT32 multiply pipe, not by HBM: a 64-bit Shoup butterfly is 9 32-bit products
            # (3 IMAD.WIDE + 2 IMAD.HI + 4 IMAD), a 32-bit one 3 (2 IMAD + 1 IMAD.HI); same measured rates as `roofline`
            pk = kt["roofline"]["int32_peaks_tops"]
            if pk.get("imad"):
                for r in ntt:
                    bf = r["batch"] * (1 << r["log_n"]) // 2 * r["log_n"]
                    per = (4 / pk["imad"] + 2 / pk["imad_hi"] + 3 / pk["imad_wide"]) if r["word_bits"] == 64 else (2 / pk["imad"] + 1 / pk["imad_hi"])
                    r["fwd_frac_int32"] = round(bf * per / 1e12 / (r["fwd_ms"] * 1e-3), 4)
                    r["inv_frac_int32"] = round(bf * per / 1e12 / (r["inv_ms"] * 1e-3), 4)
                line["roofline_ntt"]["frac_int32_pipe"] = best["fwd_frac_int32"]
            line["ntt_steady_state_65536_polys"] = ntt_large
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
