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
    for _ in range(args.warmup_ref):
        K.op([1, 1, 1, 0], lin[:cores], threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = K.op([1, 1, 1, 0], lin, threads=cores)
    dt = time.perf_counter() - t0
    assert (K.decrypt(out) == 1 - (bits[:sample] & bits[sample:])).all()
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup_ref, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 (u128 % modmul)", "data": "synthetic",
            "config": {"workload": "FHEW NAND gate bootstrap, FHEW-T (boolean.rs:225-239): N=512 Q=268409857 d=4 n_s=100 w=10",
                       "batch_per_step": sample, "note": "CPU restatement of the reference (Rust toolchain absent); bounded sample"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d gates/step x %d steps" % (sample, args.steps)},
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


def max_over_ranks(torch, dist, world, dev, x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def tfhe_leg(pkg, ctx, torch, dist, world, rank, local, batch, steps, synthetic_n1024=False):
    """BASELINE configs[2]: TFHE programmable bootstrapping at the reference parameter set (tfhe/bootstrapping.rs:141-152:
    n=1024, N=2048, k=1, TGGSW B=2^23 d=1, key switch B=2^4 d=5), `batch` synthetic LWE ciphertexts per GPU, keys uploaded on
    rank 0 and broadcast once.  Timing is value independent; parity (bit-exact vs the oracle) is in tests/test_gpu_tfhe.py."""
    from learn_fhe_b200 import tfhe
    dev = "cuda:%d" % local
    stream = torch.cuda.current_stream(local)
    P = tfhe.bootstrapping_testing_param()
    if synthetic_n1024:  # BASELINE configs[2] also names N=1024, for which the reference has no parameter set
        P = pkg.TfheParam(log_p=2, padding=1, n=630, ks_log_b=2, ks_d=8, log_big_n=10, k=1, bs_log_b=7, bs_d=3)
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
    lut = pkg.to_dev(np.random.default_rng(5).integers(0, 1 << 63, size=N, dtype=np.uint64), local)
    cts = torch.randint(-(1 << 62), 1 << 62, (batch, n + 1), dtype=torch.int64, device=dev)
    out = torch.empty_like(cts)
    tfhe.Bootstrapping.bootstrap_dev(bk, lut, cts, out)  # warm-up (allocations, tables)
    ctx.prof_begin()
    ms = device_ms(torch, stream, lambda: tfhe.Bootstrapping.bootstrap_dev(bk, lut, cts, out), 0, steps)
    prof = ctx.prof_end()
    ms = max_over_ranks(torch, dist, world, dev, ms)
    br = prof.get("tfhe_blind_rotate_kernel", {"ms": 0.0, "launches": 1})
    br_ms = br["ms"] / max(1, br["launches"])
    # f64 operations of the reference dataflow per PBS: n CMUX x [(k+1)d forward + (k+1)^2 d inverse] FFTs of N/2 points,
    # 10 flops per radix-2 butterfly (4 mul + 6 add, never fused) + twist / pointwise / untwist
    m, lg, nl = N // 2, (N // 2).bit_length() - 1, (k + 1) * P.bs_d
    ffts = n * (nl + (k + 1) * nl)
    flops = ffts * (m // 2) * lg * 10 + n * (nl * m * 6 + (k + 1) * nl * m * (6 + 8))
    # The reference rounds every product and every sum separately (c64.rs / fft.rs use plain f64 * and +), so a bit-identical
    # kernel cannot contract them into FMAs: the binding rate is one f64 operation per lane per clock.  Nominal B200 figure
    # (64 FP64 lanes per SM per clock at the maximum SM clock; MEASURED_PEAKS.json has no f64 entry).
    fp64_ops_peak = 148 * 64 * 1.965e9
    res = {"metric": "tfhe_pbs_per_sec", "value": batch * world * steps / (ms * 1e-3), "unit": "PBS/s", "batch_per_gpu": batch,
           "steps": steps, "ms_per_step": ms / steps,
           "config": ("synthetic, not from the reference: n=630 N=1024 k=1 B=2^7 d=3, ks B=2^2 d=8; bit-exact f64 FFT dataflow" if synthetic_n1024 else
                      "TFHE-T (tfhe/bootstrapping.rs:141-152): n=1024 N=2048 k=1 B=2^23 d=1, ks B=2^4 d=5; bit-exact f64 FFT dataflow"),
           "key_bytes": bk.nbytes,
           "kernels": {kk: {"ms_per_launch": v["ms"] / v["launches"], "launches": v["launches"]} for kk, v in prof.items()},
           "roofline": {"bound": "fp64", "kernel": "tfhe_blind_rotate_kernel", "achieved": flops * batch / (br_ms * 1e-3) / 1e12 if br_ms else None,
                        "peak": fp64_ops_peak / 1e12,
                        "unit": "Tflop/s f64 (algorithmic unfused multiplies and adds of the reference dataflow; peak = nominal "
                                "non-FMA issue rate, 64 lanes/SM/clk x 148 SMs x 1.965 GHz)",
                        "frac": flops * batch / (br_ms * 1e-3) / fp64_ops_peak if br_ms else None,
                        "frac_vs_fma_peak": flops * batch / (br_ms * 1e-3) / (2 * fp64_ops_peak) if br_ms else None,
                        "traffic": ncu_traffic("tfhe_blind_rotate_kernel", batch), "flops_per_pbs": flops}}
    # optional evaluation mode: products summed in the Fourier domain (within the reference's error bound, decryptions
    # identical; NOT bit-identical, so it is reported beside the headline, never as it)
    bk.set_mode(True)
    tfhe.Bootstrapping.bootstrap_dev(bk, lut, cts, out)
    ms2 = device_ms(torch, stream, lambda: tfhe.Bootstrapping.bootstrap_dev(bk, lut, cts, out), 0, steps)
    ms2 = max_over_ranks(torch, dist, world, dev, ms2)
    res["fourier_acc_mode"] = {"value": batch * world * steps / (ms2 * 1e-3), "unit": "PBS/s", "ms_per_step": ms2 / steps,
                               "parity": "decryptions identical; torus words within 2^52 of the reference dataflow (tests/test_gpu_tfhe.py)"}
    bk.set_mode(False)
    bk.free()
    del cts, out
    torch.cuda.empty_cache()
    return res


def ckks_leg(pkg, ctx, torch, dist, world, rank, local, batch, steps):
    """BASELINE configs[3]: Ckks::mul (tensor product + relinearise + rescale, ckks.rs:255-272) at N=2^16, log_qi=55, L=8,
    full level, `batch` ciphertext pairs per GPU.  Parity (bit-exact vs the oracle) is in tests/test_gpu_ckks.py."""
    from learn_fhe_b200 import ckks
    dev = "cuda:%d" % local
    stream = torch.cuda.current_stream(local)
    log_n, L = 16, 8
    P = ckks.CkksParam.new(ctx, log_n, 55, L)
    rng = np.random.default_rng(0x5EED0003)
    ksk = np.stack([np.stack([rng.integers(0, q, size=P.n, dtype=np.uint64) for q in P.qs + P.ps]) for _ in range(2)])
    rlk = ckks.CkksKeySwitchingKey(P, ksk)

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
                           "frac": t_min / (ms / steps * 1e-3), "traffic": None,
                           "butterflies_per_mult": n_bf, "pointwise_modmuls_per_mult": n_mm}
    rlk.free()
    P.free()
    del ct0, ct1, out
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
    ap.add_argument("--tfhe-batch", type=int, default=16384, help="PBS per GPU per step")
    ap.add_argument("--ckks-batch", type=int, default=512, help="ciphertext pairs per GPU per step")
    args = ap.parse_args()
    args.warmup_ref = max(1, min(args.warmup, 1))
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
        # keep stdout to the one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION) goes to stdout
        os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "WARN")
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
    B = args.batch
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
    value = B * world * args.steps / (ms_total * 1e-3)
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
    e2e_value = B * world * args.steps / float(t.item())
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

    # free the FHEW batch buffers before the wider legs
    del ins, outs
    torch.cuda.empty_cache()
    tfhe_res = None if args.no_tfhe else tfhe_leg(pkg, ctx, torch, dist, world, rank, local, args.tfhe_batch, max(1, min(args.steps, 2)))
    tfhe_res_1024 = None if args.no_tfhe else tfhe_leg(pkg, ctx, torch, dist, world, rank, local, args.tfhe_batch, max(1, min(args.steps, 2)), True)
    ckks_res = None if args.no_ckks else ckks_leg(pkg, ctx, torch, dist, world, rank, local, args.ckks_batch, max(1, min(args.steps, 3)))

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

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32 (Q < 2^30 residues, u64 MAC accumulators)", "data": "synthetic",
                "config": {"workload": "FHEW NAND gate bootstrap (Fhew::op), FHEW-T (boolean.rs:225-239): N=512 Q=268409857 "
                                       "RGSW/RLWE B=2^7 d=4, n_s=100 q_ks=2^16, w=10",
                           "batch_per_gpu": B, "global_batch": B * world, "sharding": "by ciphertext, keys replicated (NCCL broadcast once)",
                           "l2": "inputs rotate over %d distinct in/out sets (%.0f MB) > L2" % (nset, nset * 2 * ct_words * 8 / 1e6)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ct_words * 8 + param.n * 8),
                        "d2h_bytes_per_step": int(ct_words * 8)},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": dict(kt["roofline"], traffic=ncu_traffic("fhew_blind_rotate_fast_kernel", B)), "kernels": kt["kernels"], "cpu_baseline": cpu, "peak_source": peak_src}
        if tfhe_res is not None:
            line["tfhe_pbs"] = tfhe_res
            line["tfhe_pbs_n1024_synthetic"] = tfhe_res_1024
        if ckks_res is not None:
            line["ckks_mul"] = ckks_res
        if next_res is not None:
            line["next_rows"] = next_res
        if ntt is not None:
            line["ntt"] = ntt
            best = max(ntt, key=lambda r: (r["log_n"], r["word_bits"]))
            line["roofline_ntt"] = {"bound": "hbm", "achieved": best["fwd_gbs"], "peak": hbm_peak, "unit": "GB/s",
                                    "frac": best["fwd_frac_hbm"], "traffic": None,
                                    "kernel": "ntt fwd u64 N=2^%d batch 4096" % best["log_n"]}
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
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
