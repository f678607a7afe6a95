#!/usr/bin/env python3
"""bench.py -- ChaCha20 stwo proofs/sec on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA backend (through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's own CPU prover (oracle/_ref = its shipped WASM
                                                           build compiled natively; single-threaded like the reference)

A step = one ChaCha20 stream proof (prove only) of 2^log_n_rows 64-byte blocks of synthetic data, PcsConfig::default().
`value`  : whole-job proofs/s with plaintext/ciphertext already resident in HBM (device-input entry point).
`e2e`    : same through the host-buffer C-ABI call (pinned host inputs -> H2D inside, proof bytes D2H inside).
Independent proofs shard across ranks with no data-path collective ("weak" scaling: one proof per GPU per step).
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_COLS = 33280
N_CONSTRAINTS = 54784


def synth_inputs(log_n, rank):
    """Deterministic synthetic workload: key 00..1f, nonce per rank, counter 1, plaintext = seeded random bytes,
    ciphertext = ChaCha20 keystream xor plaintext (computed with numpy; test data only)."""
    import numpy as np
    n = 1 << log_n
    key = bytes(range(32))
    nonce = bytes([0, 0, 0, rank, 0, 0, 0, 0x4A, 0, 0, 0, 0])
    counter = 1
    rng = np.random.default_rng(1000 + rank)
    pt = rng.integers(0, 2 ** 32, size=(n, 16), dtype=np.uint64).astype(np.uint32)
    kw = np.frombuffer(key, dtype="<u4").astype(np.uint32)
    nw = np.frombuffer(nonce, dtype="<u4").astype(np.uint32)
    init = np.empty((16, n), dtype=np.uint32)
    init[0], init[1], init[2], init[3] = 0x61707865, 0x3320646E, 0x79622D32, 0x6B206574
    for i in range(8):
        init[4 + i] = kw[i]
    init[12] = (counter + np.arange(n, dtype=np.uint64)).astype(np.uint32)
    for i in range(3):
        init[13 + i] = nw[i]
    v = init.copy()

    def rotl(x, r):
        return (x << np.uint32(r)) | (x >> np.uint32(32 - r))

    def qr(a, b, c, d):
        v[a] += v[b]; v[d] = rotl(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = rotl(v[b] ^ v[c], 12)
        v[a] += v[b]; v[d] = rotl(v[d] ^ v[a], 8); v[c] += v[d]; v[b] = rotl(v[b] ^ v[c], 7)

    with np.errstate(over="ignore"):
        for _ in range(10):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        ks = (v + init).T
    ct = pt ^ ks
    return key, nonce, counter, np.ascontiguousarray(pt), np.ascontiguousarray(ct)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.lines = []
        self.proc = None
        self.gpu = gpu

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY, "--format=csv,noheader",
                                          "-lms", "250"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1].split()[0])); mx = max(mx, float(f[2].split()[0]))
            except Exception:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def _ref_worker(args):
    """One reference proof in this process (the reference binary is single-threaded and keeps global state: one process each)."""
    log_sample, rank = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_wasm
    key, nonce, counter, pt, ct = synth_inputs(log_sample, rank)
    ptb, ctb = pt.tobytes(), ct.tobytes()
    t0 = time.perf_counter()
    res = ref_wasm.generate_chacha20_proof(key, nonce, counter, ptb, ctb)
    dt = time.perf_counter() - t0
    assert res.get("success") is True, res
    return dt


def cpu_reference_run(log_sample, steps, warmup, procs=1):
    """Times the reference's own prover (oracle/_ref) on this box's host cores: 2^log_sample blocks per proof.
    procs > 1: that many independent reference proofs at once, one process per host core (the reference has no threads of its
    own -- no rayon in its Cargo.lock -- so independent proofs are the only way it uses more than one core).
    Returns (seconds per step, proofs per step)."""
    if procs <= 1:
        times = []
        for i in range(warmup + steps):
            dt = _ref_worker((log_sample, 0))
            if i >= warmup:
                times.append(dt)
        return sum(times) / len(times), 1
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    times = []
    with ctx.Pool(procs) as pool:
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_ref_worker, [(log_sample, r) for r in range(procs)])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), procs


def host_parallelism(log_sample):
    """Processes the reference arm can run at once: one per host core, bounded by memory (a 2^log_sample-block reference proof
    needs ~0.4 MB per block of its wasm32 address space)."""
    cores = os.cpu_count() or 1
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 8 << 30
    per_proc = max(int(0.45e6 * (1 << log_sample)), 256 << 20)
    return max(1, min(cores, int(avail * 0.7) // per_proc))


def synth_aes_inputs(key_len, log_n, rank):
    """Deterministic AES-CTR workload (numpy AES, test data only): key/nonce from a seeded rng, counter 1."""
    import numpy as np
    sbox = np.zeros(256, dtype=np.uint8)   # FIPS-197 S-box from the field inverse + affine map
    p = q = 1
    while True:
        p = (p ^ ((p << 1) & 0xFF) ^ (0x1B if p & 0x80 else 0)) & 0xFF
        q ^= q << 1; q ^= q << 2; q ^= q << 4; q &= 0xFF
        if q & 0x80:
            q ^= 0x09
        x = q ^ ((q << 1 | q >> 7) & 0xFF) ^ ((q << 2 | q >> 6) & 0xFF) ^ ((q << 3 | q >> 5) & 0xFF) ^ ((q << 4 | q >> 4) & 0xFF)
        sbox[p] = (x ^ 0x63) & 0xFF
        if p == 1:
            break
    sbox[0] = 0x63

    def expand_key(key):
        nk = len(key) // 4; nrr = nk + 6
        w = [list(key[4 * i:4 * i + 4]) for i in range(nk)]
        rc = 1
        for i in range(nk, 4 * (nrr + 1)):
            t = list(w[i - 1])
            if i % nk == 0:
                t = [int(sbox[b]) for b in t[1:] + t[:1]]
                t[0] ^= rc
                rc = ((rc << 1) & 0xFF) ^ (0x1B if rc & 0x80 else 0)
            elif nk > 6 and i % nk == 4:
                t = [int(sbox[b]) for b in t]
            w.append([a ^ b for a, b in zip(w[i - nk], t)])
        return [sum(w[4 * r:4 * r + 4], []) for r in range(nrr + 1)]
    SR = [0, 5, 10, 15, 4, 9, 14, 3, 8, 13, 2, 7, 12, 1, 6, 11]
    nb = 1 << log_n
    rng = np.random.default_rng(2000 + rank)
    key = rng.bytes(key_len); nonce = rng.bytes(12); counter = 1
    pt = np.frombuffer(rng.bytes(16 * nb), dtype=np.uint8).reshape(nb, 16)
    rk = np.array(expand_key(key), dtype=np.uint8)
    nr = rk.shape[0] - 1
    blk = np.zeros((nb, 16), dtype=np.uint8)
    blk[:, :12] = np.frombuffer(nonce, dtype=np.uint8)
    ctrs = (counter + np.arange(nb, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)
    for i in range(4):
        blk[:, 12 + i] = (ctrs >> np.uint64(8 * (3 - i))) & np.uint64(0xFF)

    def xt(a):
        return ((a << 1) ^ ((a >> 7) * 0x1B)).astype(np.uint8)
    s = blk ^ rk[0]
    for r in range(1, nr + 1):
        s = sbox[s][:, SR]
        if r < nr:
            o = np.empty_like(s)
            for c in range(4):
                a0, a1, a2, a3 = (s[:, 4 * c + j] for j in range(4))
                o[:, 4 * c] = xt(a0) ^ xt(a1) ^ a1 ^ a2 ^ a3
                o[:, 4 * c + 1] = a0 ^ xt(a1) ^ xt(a2) ^ a2 ^ a3
                o[:, 4 * c + 2] = a0 ^ a1 ^ xt(a2) ^ xt(a3) ^ a3
                o[:, 4 * c + 3] = xt(a0) ^ a0 ^ a1 ^ a2 ^ xt(a3)
            s = o
        s = s ^ rk[r]
    ct = s ^ pt
    return key, nonce, counter, pt.tobytes(), ct.tobytes()


def aes_main(args, rank, local_rank, world):
    """BASELINE configs[2]: AES-128/256-CTR AIR proofs.  Inputs are 16 bytes per row, so the only meaningful figure is the
    end-to-end one through the host-buffer C ABI (reported as both value and e2e)."""
    key_len = 16 if args.workload == "aes128" else 32
    L = args.log_size if "--log-size" in " ".join(sys.argv) else 16
    cols, cons = (24480, 34464) if key_len == 16 else (34784, 49024)
    config = {"workload": "%s_ctr log_n_rows=%d blowup=2 (one proof of %d 16-byte blocks per GPU per step)" % (args.workload, L, 1 << L),
              "log_n_rows": L, "columns": cols, "constraints": cons, "pcs": "pow_bits=10,n_queries=3,log_blowup=1,last_layer=0",
              "l2": "inputs larger than L2 (LDE %.1f GB)" % (cols * (2 << L) * 4 / 1e9),
              "sharding": "independent proofs, one per rank, no data-path collective"}
    if args.impl == "reference":
        if rank != 0:
            return
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_wasm
        S = min(L, 9)
        key, nonce, counter, pt, ct = synth_aes_inputs(key_len, S, 0)
        fn = ref_wasm.generate_aes128_ctr_proof if key_len == 16 else ref_wasm.generate_aes256_ctr_proof
        t0 = time.perf_counter()
        res = fn(key, nonce, counter, pt, ct)
        sec = time.perf_counter() - t0
        assert res.get("success") is True, res
        scaled = 1.0 / (sec * (1 << (L - S)))
        print(json.dumps({"impl": "reference", "metric": "%s_ctr_proofs_per_sec" % args.workload, "value": scaled, "unit": "proofs/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / scaled,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32(M31)", "data": "synthetic",
                          "config": config,
                          "cpu_baseline": {"value": scaled, "unit": "proofs/s", "cores": 1, "kind": "reference",
                                           "sample": "reference prover on 2^%d blocks: %.3f s/proof, linearly scaled to 2^%d" % (S, sec, L)},
                          "e2e": {"value": scaled, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    os.environ["NCCL_DEBUG"] = os.environ.get("S2C_NCCL_DEBUG", "WARN")
    import torch
    import torch.distributed as dist
    import zk_symmetric_crypto_b200 as z
    from zk_symmetric_crypto_b200 import sharding
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    be = z.Backend(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    be.set_stream(stream.cuda_stream)
    key, nonce, counter, pt, ct = synth_aes_inputs(key_len, L, rank)

    def step():
        return be.prove_aes_ctr_raw(key, nonce, counter, pt, ct)
    sampler = ClockSampler(local_rank)   # started before the warm-up: see the ChaCha path
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        proof = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = be.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        proof = step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = sharding.max_over_ranks([e0.elapsed_time(e1)], device="cuda")[0]
    launches = be.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    be.set_profile(True)
    step()
    stages = be.stage_times()
    if rank == 0:
        value = world * args.steps / (ms / 1000.0)
        N = 1 << L
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        lde_ms = stages.get("trace_lde", 0.0)
        alg = (cols + 1) * N * 20
        print(json.dumps({"metric": "%s_ctr_proofs_per_sec" % args.workload, "value": value, "unit": "proofs/s", "n_gpus": world,
                          "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "u32(M31)", "data": "synthetic", "config": config,
                          "clocks": clocks, "gpu_launches": launches,
                          "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 2 * len(pt), "d2h_bytes_per_step": len(proof)},
                          "stage_ms": stages,
                          "roofline": {"kernel": "trace_lde", "bound": "hbm", "achieved": alg / (lde_ms / 1000.0) / 1e9 if lde_ms else None,
                                       "peak": hbm_peak, "unit": "GB/s", "frac": alg / (lde_ms / 1000.0) / 1e9 / hbm_peak if lde_ms else None,
                                       "traffic": None, "algorithmic_bytes_per_proof": alg}}))
    be.close()
    if world > 1:
        dist.destroy_process_group()


def load_json(*parts):
    try:
        return json.load(open(os.path.join(ROOT, *parts)))
    except Exception:
        return None


def int_peaks():
    """Measured integer roofs of this GPU family (profiles/microbench/int_peak.cu run on a B200, result committed as
    profiles/int_peak_r02.json): register-resident M31 butterflies/s and Blake2s half-G/s with every SM busy, plus the raw
    ALU-pipe / FMA-pipe instruction rates they follow from."""
    doc = load_json("profiles", "int_peak_r02.json")
    if not doc:
        return None
    r = {x["op"]: x for x in doc["results"]}
    return {"butterflies_per_s": r["BFLY_CUR"]["items_per_s"], "blake2s_half_g_per_s": r["BLAKE_G"]["items_per_s"],
            "alu_pipe_thread_instr_per_s": r["LOP3"]["thread_instr_per_s"], "fma_pipe_imad_per_s": r["IMAD"]["thread_instr_per_s"],
            "imad_wide_per_s": r["IMADWIDE"]["thread_instr_per_s"], "source": "profiles/int_peak_r02.json (profiles/microbench/int_peak.cu)"}


def verify_proof(proof, nonce, counter, ptb, ctb):
    """Checks the timed proof outside the timed region: the reference's own verifier when oracle/_ref travelled to this box
    (checker only), else this repo's host verifier."""
    import base64
    import zk_symmetric_crypto_b200 as z
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_wasm
        if ref_wasm.available():
            t0 = time.perf_counter()
            v = ref_wasm.verify_chacha20_proof(base64.b64encode(proof).decode(), nonce, counter, ptb, ctb)
            return {"by": "reference", "ok": v == {"algorithm": "chacha20", "valid": True}, "seconds": round(time.perf_counter() - t0, 2),
                    "verifier": "oracle/_ref verify_chacha20_proof (wasm_api.rs:609)"}
    except Exception as ex:
        note = "reference verifier unavailable: %s" % ex
    else:
        note = "oracle/_ref not on this box"
    ok, err = z.verify_chacha20_raw(proof, nonce, counter, ptb, ctb)
    return {"by": "host", "ok": bool(ok), "error": err, "note": note}


def aes_measure(z, torch, local_rank, key_len, L, steps, warmup, rank=0):
    """One AES-CTR workload on its own context: device-event time of `steps` proofs through the host-buffer C ABI (inputs are 16
    bytes per row, so end to end is the only meaningful figure) and the per-stage times of one profiled proof."""
    be = z.Backend(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    be.set_stream(stream.cuda_stream)
    key, nonce, counter, pt, ct = synth_aes_inputs(key_len, L, rank)
    try:
        for _ in range(warmup):
            proof = be.prove_aes_ctr_raw(key, nonce, counter, pt, ct)
        torch.cuda.synchronize()
        l0 = be.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            proof = be.prove_aes_ctr_raw(key, nonce, counter, pt, ct)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = (be.launch_count() - l0) // steps
        be.set_profile(True)
        be.prove_aes_ctr_raw(key, nonce, counter, pt, ct)
        stages = be.stage_times()
        ok, err = z.verify_aes_ctr_raw(proof, nonce, counter, pt, ct)
    finally:
        be.close()
    return {"ms_per_proof": ms, "launches_per_proof": launches, "stage_ms": stages, "proof_bytes": len(proof),
            "h2d_bytes": 2 * len(pt), "verified": {"by": "host", "ok": bool(ok), "error": err}}


def aes_roofline(cols, L, stages, hbm_peak):
    lde_ms = stages.get("trace_lde", 0.0)
    alg = (cols + 1) * (1 << L) * 20   # 8CN (interpolate) + 12CN (evaluate on the blown-up domain) per column
    gbs = alg / (lde_ms / 1000.0) / 1e9 if lde_ms else None
    nk = (load_json("profiles", "ncu_kernels_r02.json") or {}).get("aes_trace_lde")
    return {"kernel": "trace_lde", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak if gbs else None,
            "traffic": nk["dram_bytes"] if nk else None, "algorithmic_bytes_per_proof": alg}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--log-size", type=int, default=int(os.environ.get("S2C_BENCH_LOG", "20")))
    ap.add_argument("--cpu-log-size", type=int, default=None, help="size of the bounded CPU-reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the AES-CTR / proof-batch / sharded sub-records")
    ap.add_argument("--workload", default="chacha20", choices=["chacha20", "chacha20_sharded", "aes128", "aes256"],
                    help="chacha20 (BASELINE configs[1], the headline) or an AES-CTR AIR (configs[2])")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload in ("aes128", "aes256"):
        return aes_main(args, rank, local_rank, world)
    sharded = args.workload == "chacha20_sharded"   # cfg-5 mechanism: all ranks prove ONE trace together ("strong" scaling)
    L = args.log_size
    workload = "chacha20_stream log_n_rows=%d blowup=2 (one proof of %d blocks per GPU per step)" % (L, 1 << L)
    config = {"workload": workload, "log_n_rows": L, "columns": N_COLS, "constraints": N_CONSTRAINTS,
              "pcs": "pow_bits=10,n_queries=3,log_blowup=1,last_layer=0", "l2": "inputs larger than L2 (LDE %.1f GB)" %
              (N_COLS * (2 << L) * 4 / 1e9),
              "sharding": ("one trace over all ranks: column-sharded transforms, NCCL all-to-all of LDE row shards, row-sharded "
                           "leaf hashing / constraints" if sharded else "independent proofs, one per rank, no data-path collective")}

    if args.impl == "reference":
        # The reference's own prover on this box's host cores.  It cannot hold the workload (2^20 blocks need ~560 GB; its
        # wasm32 build stops at 2^13), so each step is a bounded sample: the LARGEST trace it proves within the time budget
        # (2^12 blocks, ~40 s), on every host core at once (independent proofs, one process per core: it has no threads of its
        # own), and the line says that the proofs/s figure is an extrapolation (linear in rows, optimistic for the CPU).
        if rank != 0:
            return
        S = args.cpu_log_size if args.cpu_log_size is not None else min(L, 12)
        procs = host_parallelism(S)
        sec, per_step = cpu_reference_run(S, max(1, min(args.steps, 2)), 0, procs)
        sec1, _ = cpu_reference_run(min(S, 10), 1, 0, 1)
        scale = (1 << (L - S)) if L >= S else 1
        scaled = per_step / (sec * scale)
        line = {"impl": "reference", "metric": "chacha20_proofs_per_sec", "value": scaled, "unit": "proofs/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / scaled, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32(M31)", "data": "synthetic", "config": config,
                "extrapolated": True, "from_log": S, "extrapolation": "measured at 2^%d blocks per proof, divided by %d (linear in rows)" % (S, scale),
                "cpu_baseline": {"value": scaled, "unit": "proofs/s", "cores": procs, "host_cores": os.cpu_count(), "kind": "reference",
                                 "extrapolated": True, "from_log": S,
                                 "single_process_2^%d_blocks_s" % min(S, 10): sec1,
                                 "sample": "reference prover (shipped WASM build compiled natively; no SIMD, no threads of its own) on "
                                           "2^%d blocks per proof, %d independent proofs at once on %d of %d host cores: %.2f s per batch; "
                                           "scaled linearly to 2^%d blocks" % (S, per_step, procs, os.cpu_count() or 0, sec, L)},
                "e2e": {"value": scaled, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    os.environ["NCCL_DEBUG"] = os.environ.get("S2C_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
    import numpy as np
    import torch
    import torch.distributed as dist
    import zk_symmetric_crypto_b200 as z
    from zk_symmetric_crypto_b200 import sharding
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    be = z.Backend(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    be.set_stream(stream.cuda_stream)
    key, nonce, counter, pt, ct = synth_inputs(L, 0 if sharded else rank)

    def join_comm(backend_obj):
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(z.backend.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        backend_obj.comm_init(rank, world, bytes(uid.cpu().tolist()))

    if sharded and world > 1:
        join_comm(be)
    nbytes = pt.nbytes
    # pinned host copies (e2e path) and device-resident copies (value path)
    pt_pin = torch.from_numpy(pt.view(np.int32)).pin_memory()
    ct_pin = torch.from_numpy(ct.view(np.int32)).pin_memory()
    pt_dev = pt_pin.to("cuda:%d" % local_rank)
    ct_dev = ct_pin.to("cuda:%d" % local_rank)

    def step_dev():
        # inputs resident in HBM; the two public-input Blake2s hashes (host work in the reference too) are computed inside the
        # call from a read-back of the buffers, so `value` skips nothing that `e2e` does except the H2D copies
        return be.prove_chacha20_ptr(key, nonce, counter, pt_dev.data_ptr(), ct_dev.data_ptr(), nbytes, on_device=True)

    def step_e2e():
        return be.prove_chacha20_ptr(key, nonce, counter, pt_pin.data_ptr(), ct_pin.data_ptr(), nbytes)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            proof = fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        ms = dev_ms  # device clock on the launching stream (the proof call is host-synchronous, so this is the step time)
        ms, wall_ms = sharding.max_over_ranks([ms, wall * 1000.0], device="cuda")  # multi-GPU numbers: max over ranks
        return ms, wall_ms / 1000.0, proof

    # the clock sampler (nvidia-smi polling) is started before the warm-up so that its start-up (NVML initialisation takes
    # driver locks for a few hundred ms) does not land inside the timed region; it keeps sampling through the timed steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        proof = step_dev()
    l0 = be.launch_count()
    ms, wall, proof = timed(step_dev, args.steps)
    launches = be.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e, wall_e2e, proof_e2e = timed(step_e2e, args.steps)
    if rank == 0 or not sharded:
        assert proof_e2e == proof, "host-input and device-input paths must give the same proof"
    # the proof that was timed, checked outside the timed region
    verified = verify_proof(proof, nonce, counter, pt.tobytes(), ct.tobytes()) if rank == 0 else None
    # one profiled step for the per-kernel breakdown (CUDA events on the launching stream around each kernel)
    be.set_profile(True)
    step_dev()
    stages = be.stage_times()
    cnt = be.counters()
    be.set_profile(False)

    # ---- cfg-5 on the driver's hardware: under torchrun the same ranks additionally prove ONE trace together (rank 0's inputs)
    #      and rank 0 checks the bytes against its own single-GPU proof of those inputs
    sharded_rec = None
    if world > 1 and not sharded and not args.no_extras and L >= 16:
        try:
            key0, nonce0, counter0, pt0, ct0 = synth_inputs(L, 0) if rank else (key, nonce, counter, pt, ct)
            p0_pin = torch.from_numpy(pt0.view(np.int32)).pin_memory()   # rank 0's inputs, pinned on every rank
            c0_pin = torch.from_numpy(ct0.view(np.int32)).pin_memory()
            join_comm(be)
            times = []
            for it in range(3):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                sp = be.prove_chacha20_ptr(key0, nonce0, counter0, p0_pin.data_ptr(), c0_pin.data_ptr(), nbytes)
                e1.record(stream)
                barrier()
                if it:
                    times.append(sharding.max_over_ranks([e0.elapsed_time(e1)], device="cuda")[0])
            be.set_profile(True)
            be.prove_chacha20_ptr(key0, nonce0, counter0, p0_pin.data_ptr(), c0_pin.data_ptr(), nbytes)
            sst = be.stage_times()
            be.set_profile(False)
            a2a = sharding.max_over_ranks([sst.get("all_to_all", 0.0) + sst.get("group_barrier", 0.0)], device="cuda")[0]
            peer_windows = bool(be.counters().get("peer_windows", 0))
            be.comm_destroy()
            if rank == 0:
                sharded_rec = {"workload": "ONE chacha20 trace log_n_rows=%d over %d GPUs (column-sharded transforms, row shards exchanged, "
                                           "row-sharded hashing / constraints)" % (L, world),
                               "exchange": ("peer-window stores over NVLink fused into the last transform pass (CUDA IPC), one 4-byte "
                                            "all-reduce per plan group as the barrier" if peer_windows else
                                            "NCCL grouped send/recv all-to-all of staged tiles"),
                               "ms": min(times), "ms_all": times, "parity": sp == proof, "single_gpu_ms": ms / args.steps,
                               "speedup_vs_single_gpu": (ms / args.steps) / min(times), "all_to_all_ms": a2a,
                               "stage_ms_rank0_profiled": sst, "h2d_bytes": 2 * nbytes,
                               "timing": "CUDA events on the launching stream around the host-buffer C-ABI call, barrier on both sides, max over ranks"}
        except Exception as ex:  # never lose the headline line to the extra record
            if rank == 0:
                sharded_rec = {"error": str(ex)[:300], "parity": False}

    line = None
    if rank == 0:
        peaks = load_json("MEASURED_PEAKS.json") or {}
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        N = 1 << L
        cols_t = 32 * cnt.get("fft_words", 0)           # columns transformed from the packed witness (both passes)
        # ALGORITHMIC bytes (SURVEY 8d / DESIGN.md section 5).  Transforms: per column interpolate reads N + writes N words and
        # evaluate reads N + writes 2N = 20 N bytes (the three-pass pipeline's own scratch traffic is NOT algorithmic: it is
        # reported as `pipeline_bytes`); columns recomputed on half of the domain count 14 N.  Leaves and constraints read every
        # LDE value they use once.
        half_words = cnt.get("fft_words_half", 0)
        full_words = cnt.get("fft_words", 0) - half_words
        fft_alg = 32 * N * (20 * full_words + 14 * half_words)
        fft_pipeline = cols_t * N * ((1 / 8 + 4) + (4 + 8) + (8 + 8))
        # butterflies per column: interpolate L layers of N/2, evaluate L layers of N/2 on each half of the 2N-point domain
        butterflies = 32 * N * (full_words * 1.5 * L + half_words * 1.0 * L)
        alg = {
            "fft": fft_alg,
            "trace_merkle_leaves": N_COLS * 2 * N * 4 + 2 * N * 32,
            "constraints": 22528 * N * 4 + 4 * 2 * N * 4,   # single-GPU mode evaluates storage rows [0, N)
        }
        fft_ms = sum(stages.get(k, 0.0) for k in ("ifft_low", "fft_mid", "fft_low", "fft_small"))
        kern_ms = {"fft": fft_ms, "trace_merkle_leaves": stages.get("trace_merkle_leaves", 0.0), "constraints": stages.get("constraints", 0.0)}
        top = max(kern_ms, key=lambda k: kern_ms[k])
        ncu = load_json("profiles", "ncu_kernels_r02.json") or {}
        ip = int_peaks()
        compressions = 2 * N * (N_COLS * 4 // 64)
        int_roof = {}
        if ip:
            # integer roofs, measured (profiles/int_peak_r02.json): a Blake2s compression is 160 half-G; a butterfly is the 7-instruction
            # M31 butterfly of kernels_fft2.cu timed register-resident with every SM busy
            if kern_ms["trace_merkle_leaves"]:
                a = compressions / (kern_ms["trace_merkle_leaves"] / 1000.0)
                pk = ip["blake2s_half_g_per_s"] / 160.0
                int_roof["trace_merkle_leaves"] = {"bound": "int", "achieved": a / 1e9, "peak": pk / 1e9, "unit": "Gcompress/s", "frac": a / pk}
            if fft_ms:
                a = butterflies / (fft_ms / 1000.0)
                pk = ip["butterflies_per_s"]
                int_roof["fft"] = {"bound": "int", "achieved": a / 1e9, "peak": pk / 1e9, "unit": "Gbutterfly/s", "frac": a / pk}
        achieved = alg[top] / (kern_ms[top] / 1000.0) / 1e9 if kern_ms[top] else None
        nk = ncu.get({"fft": "fft_passes", "trace_merkle_leaves": "leaves_kernel", "constraints": "constraints_tiles_kernel"}[top])
        roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak if achieved else None,
                    "traffic": nk["traffic_over_algorithmic"] * alg[top] if nk else None,
                    "traffic_source": ("DRAM bytes / algorithmic bytes of one captured launch (%s; %s: %.3f) x the algorithmic bytes of "
                                       "this run" % (nk.get("launch"), nk.get("source"), nk["traffic_over_algorithmic"])) if nk else None,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_proof": alg[top], "kernel_ms_per_proof": kern_ms[top],
                    "int": int_roof.get(top),
                    "note": "the dominant kernels are integer-pipe bound (Blake2s: ALU pipe; M31 butterflies: ALU + FMA pipes), so the "
                            "integer fraction is the one that says how close the kernel is to its roof; see DESIGN.md section 5",
                    "all_kernels": {k: {"ms": kern_ms[k], "algorithmic_bytes": alg[k],
                                        "hbm_frac": alg[k] / (kern_ms[k] / 1000.0) / 1e9 / hbm_peak if kern_ms[k] else None,
                                        "int": int_roof.get(k)} for k in kern_ms},
                    "fft_all_passes": {"ms": fft_ms, "algorithmic_GB/s": fft_alg / (fft_ms / 1000.0) / 1e9 if fft_ms else None,
                                       "frac": fft_alg / (fft_ms / 1000.0) / 1e9 / hbm_peak if fft_ms else None,
                                       "pipeline_bytes": fft_pipeline, "columns_transformed": cols_t, "butterflies": butterflies},
                    "int_peaks": ip, "counters": cnt}
        per_step = 1 if sharded else world
        value = per_step * args.steps / (ms / 1000.0)
        e2e_value = per_step * args.steps / (ms_e2e / 1000.0)
        line = {"metric": "chacha20_proofs_per_sec", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if sharded else "weak",
                "vs_baseline": None, "dtype": "u32(M31)", "data": "synthetic", "config": config,
                "blocks_per_sec": value * N, "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": "proofs/s", "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": len(proof_e2e),
                        "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": wall_e2e * 1000.0 / args.steps},
                "wall_ms_per_step": wall * 1000.0 / args.steps, "verified": verified, "stage_ms": stages, "roofline": roofline}
        if sharded_rec:
            line["sharded"] = sharded_rec
    be.close()

    # ---- sub-records the driver would otherwise never see (each on fresh contexts, after the headline's arena is released)
    if not args.no_extras and not sharded:
        extras = {}
        try:
            extras = extra_records(z, torch, dist, sharding, local_rank, rank, world, line)
        except Exception as ex:
            extras = {"error": str(ex)[:300]}
        if rank == 0:
            line.update(extras)

    if rank == 0:
        if not args.no_cpu_baseline:
            S = args.cpu_log_size if args.cpu_log_size is not None else min(L, 10)
            try:
                sec, _ = cpu_reference_run(S, 1, 0)
                scale = (1 << (L - S)) if L >= S else 1
                scaled = 1.0 / (sec * scale)
                line["cpu_baseline"] = {"value": scaled, "unit": "proofs/s", "cores": 1, "host_cores": os.cpu_count(), "kind": "reference",
                                        "extrapolated": True, "from_log": S,
                                        "sample": "reference prover (oracle/_ref: shipped WASM build compiled natively, 1 thread as "
                                                  "shipped) on 2^%d blocks: %.3f s/proof, linearly scaled to 2^%d blocks (the reference "
                                                  "cannot hold more than 2^13)" % (S, sec, L)}
            except Exception as ex:  # oracle/_ref absent on this box
                line["cpu_baseline"] = {"value": None, "unit": "proofs/s", "cores": 1, "kind": "reference", "sample": "unavailable: %s" % ex}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extra_records(z, torch, dist, sharding, local_rank, rank, world, line):
    """BASELINE configs[2] (AES-128/256-CTR, log 16) and SURVEY 8(d) cfg-4 (batches of independent product-size proofs sharded
    round-robin over the ranks) as sub-records of the default run."""
    import hashlib
    out = {}
    peaks = load_json("MEASURED_PEAKS.json") or {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    if rank == 0:
        aes = {}
        for name, key_len, cols in (("aes128", 16, 24480), ("aes256", 32, 34784)):
            r = aes_measure(z, torch, local_rank, key_len, 16, 3, 2)
            r["workload"] = "%s_ctr log_n_rows=16 blowup=2, one proof of 65,536 16-byte blocks through the host-buffer C ABI" % name
            r["proofs_per_sec"] = 1000.0 / r["ms_per_proof"]
            r["roofline"] = aes_roofline(cols, 16, r["stage_ms"], hbm_peak)
            aes[name + "_log16"] = r
        out["aes_ctr"] = aes
    # cfg-4: n=4 x 4096 and n=12 x 64 independent ChaCha20 proofs, seeds 0..B-1, round-robin over the ranks, several contexts per GPU
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_golden import case_inputs
    from zk_symmetric_crypto_b200.pool import ProverPool
    batches = {}
    for label, nb, count, nctx in (("log4_x4096", 2, 4096, 16), ("log12_x64", 4096, 64, 4)):
        inputs = [case_inputs(nb, seed % 8) for seed in range(min(count, 8))]   # 8 distinct inputs cycled (host-side generation is slow)
        mine = sharding.shard_indices(count, world, rank)
        pool = ProverPool(local_rank, nctx)
        try:
            jobs = [("chacha20_raw",) + tuple(inputs[i % len(inputs)]) for i in mine]
            pool.prove_many(jobs[:nctx])    # warm-up: twiddles, arenas
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            proofs = pool.prove_many(jobs)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            dt = time.perf_counter() - t0
        finally:
            pool.close()
        dt = sharding.max_over_ranks([dt])[0] if world == 1 else sharding.max_over_ranks([dt], device="cuda")[0]
        first = {}   # determinism check: equal inputs give equal proofs, whichever context proved them
        same = all(first.setdefault(i % len(inputs), hashlib.sha256(p).digest()) == hashlib.sha256(p).digest() for i, p in zip(mine, proofs))
        batches[label] = {"proofs": count, "blocks_per_proof": nb, "contexts_per_gpu": nctx, "seconds": dt, "proofs_per_sec": count / dt,
                          "identical_inputs_identical_proofs": same,
                          "timing": "host wall clock around the whole batch, barrier + synchronize on both sides, max over ranks"}
    if rank == 0:
        out["proof_batches_cfg4"] = batches
    return out


if __name__ == "__main__":
    main()
