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


def cpu_reference_run(log_sample, steps, warmup):
    """Times the reference's own prover (oracle/_ref) on this box's host cores: 2^log_sample blocks per proof."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_wasm
    import numpy as np
    key, nonce, counter, pt, ct = synth_inputs(log_sample, 0)
    ptb, ctb = pt.tobytes(), ct.tobytes()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        res = ref_wasm.generate_chacha20_proof(key, nonce, counter, ptb, ctb)
        dt = time.perf_counter() - t0
        assert res.get("success") is True, res
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times)


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--log-size", type=int, default=int(os.environ.get("S2C_BENCH_LOG", "20")))
    ap.add_argument("--cpu-log-size", type=int, default=10, help="size of the bounded CPU-reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        if rank != 0:
            return
        S = args.cpu_log_size
        sec = cpu_reference_run(S, max(1, min(args.steps, 3)), 1 if args.warmup else 0)
        # scale to the workload's unit: prover cost is ~linear in rows at fixed column count (optimistic for the CPU)
        scaled = 1.0 / (sec * (1 << (L - S))) if L >= S else 1.0 / sec
        line = {"impl": "reference", "metric": "chacha20_proofs_per_sec", "value": scaled, "unit": "proofs/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / scaled, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32(M31)", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": scaled, "unit": "proofs/s", "cores": 1, "kind": "reference",
                                 "sample": "reference prover (shipped WASM build compiled natively, single-threaded as shipped) on "
                                           "2^%d blocks: %.3f s/proof, linearly scaled to 2^%d blocks" % (S, sec, L)},
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
    if sharded and world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(z.backend.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        be.comm_init(rank, world, bytes(uid.cpu().tolist()))
    nbytes = pt.nbytes
    # pinned host copies (e2e path) and device-resident copies (value path)
    pt_pin = torch.from_numpy(pt.view(np.int32)).pin_memory()
    ct_pin = torch.from_numpy(ct.view(np.int32)).pin_memory()
    pt_dev = pt_pin.to("cuda:%d" % local_rank)
    ct_dev = ct_pin.to("cuda:%d" % local_rank)
    pt_hash = hashlib.blake2s(pt.tobytes()).digest()
    ct_hash = hashlib.blake2s(ct.tobytes()).digest()

    def step_dev():
        return be.prove_chacha20_ptr(key, nonce, counter, pt_dev.data_ptr(), ct_dev.data_ptr(), nbytes, on_device=True,
                                     pt_hash=pt_hash, ct_hash=ct_hash)

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
    # one profiled step for the per-kernel breakdown (CUDA events on the launching stream around each kernel)
    be.set_profile(True)
    step_dev()
    stages = be.stage_times()
    be.set_profile(False)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        N = 1 << L
        cnt = be.counters()
        cols_t = 32 * cnt.get("fft_words", 0)           # columns transformed from the packed witness (both passes)
        n_dep_tiles = 2 * 336                            # adder-sum tiles combined (both passes)
        tile_b = 32 * 2 * N * 4
        # ALGORITHMIC bytes per proof of each kernel family: per column, iFFT reads the packed bits
        # and writes N words, the strided pass reads N and writes 2N, the last pass reads and writes 2N (SURVEY 8d:
        # 8CN + 12CN); leaves and constraints read every LDE value once.
        alg = {
            "ifft_low": cols_t * N * (1 / 8 + 4),
            "fft_mid": cols_t * N * (4 + 8),
            "fft_low": cols_t * N * (8 + 8),
            "fft_small": cols_t * N * (1 / 8 + 8),
            "combine": n_dep_tiles * 4 * tile_b,
            "trace_merkle_leaves": N_COLS * 2 * N * 4 + 2 * (2 * N * 32) * 85,
            "constraints": N_COLS * 2 * N * 4 + 2 * (4 * 2 * N * 4) * 85,
        }
        fft_ms = sum(stages.get(k, 0.0) for k in ("ifft_low", "fft_mid", "fft_low", "fft_small"))
        fft_bytes = sum(alg[k] for k in ("ifft_low", "fft_mid", "fft_low", "fft_small") if stages.get(k))
        top = max((k for k in stages if k in alg), key=lambda k: stages[k], default=None)
        roofline = None
        if top:
            achieved = alg[top] / (stages[top] / 1000.0) / 1e9
            roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_proof": alg[top], "kernel_ms_per_proof": stages[top],
                        "note": "integer-issue bound, not HBM bound: see DESIGN.md section 5 for instruction counts",
                        "fft_all_passes": {"ms": fft_ms, "GB/s": fft_bytes / (fft_ms / 1000.0) / 1e9 if fft_ms else None,
                                           "frac": fft_bytes / (fft_ms / 1000.0) / 1e9 / hbm_peak if fft_ms else None,
                                           "columns_transformed": cols_t},
                        "all_kernels_gbs": {k: alg[k] / (stages[k] / 1000.0) / 1e9 for k in stages if k in alg and stages[k] > 0},
                        "counters": cnt}
        if roofline:
            # measured DRAM traffic / pipe utilisation of the same kernel from the committed ncu capture (one launch)
            try:
                nk = json.load(open(os.path.join(ROOT, "profiles", "ncu_kernels_r01.json"))).get(top)
            except Exception:
                nk = None
            if nk and L == 20:
                roofline["traffic"] = nk["dram_bytes"]
                roofline["traffic_unit"] = "DRAM bytes read+written by ONE launch (ncu --set full); that launch's algorithmic bytes: %s" % (
                    nk.get("launch_algorithmic_bytes"))
                roofline["int_pipe"] = {k: nk[k] for k in ("kernel", "launch_ms", "issue_active_pct", "alu_pipe_pct", "fma_pipe_pct",
                                                           "dram_pct", "registers") if k in nk}
                roofline["int_pipe"]["source"] = "profiles/ncu_kernels_r01.json"
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
                "wall_ms_per_step": wall * 1000.0 / args.steps, "stage_ms": stages, "roofline": roofline}
        if not args.no_cpu_baseline:
            S = args.cpu_log_size
            try:
                sec = cpu_reference_run(S, 1, 0)
                scaled = 1.0 / (sec * (1 << (L - S))) if L >= S else 1.0 / sec
                line["cpu_baseline"] = {"value": scaled, "unit": "proofs/s", "cores": 1, "kind": "reference",
                                        "sample": "reference prover (oracle/_ref: shipped WASM build compiled natively, 1 thread as "
                                                  "shipped) on 2^%d blocks: %.3f s/proof, linearly scaled to 2^%d blocks" % (S, sec, L)}
            except Exception as ex:  # oracle/_ref absent on this box
                line["cpu_baseline"] = {"value": None, "unit": "proofs/s", "cores": 1, "kind": "reference", "sample": "unavailable: %s" % ex}
        print(json.dumps(line))
    be.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
