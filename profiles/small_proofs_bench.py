"""Throughput of product-size proofs (ChaCha20: 2 blocks -> log 4, AES-CTR: 5 blocks -> log 8; js/src/config.ts chunk sizes) with
several backend contexts driven by host threads on ONE GPU (each context owns its streams; ctypes releases the GIL)."""
import sys, os, time, threading
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import zk_symmetric_crypto_b200 as z
from make_golden import case_inputs
from make_golden_aes import aes_case_inputs

algo = sys.argv[1] if len(sys.argv) > 1 else "chacha20"
n_proofs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
if algo == "chacha20":
    inp = case_inputs(2, 0)
    call = lambda be: be.prove_chacha20_raw(*inp)
else:
    inp = aes_case_inputs(16 if algo == "aes128" else 32, 5, 1)
    call = lambda be: be.prove_aes_ctr_raw(*inp)
for workers in (1, 4, 8, 16, 32):
    bes = [z.Backend(0) for _ in range(workers)]
    local = threading.local()
    idx = iter(range(workers)); lock = threading.Lock()
    def work(i):
        if not hasattr(local, "be"):
            with lock:
                local.be = bes[next(idx)]
        return len(call(local.be))
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(work, range(workers * 2)))          # warm-up
        t0 = time.perf_counter()
        out = list(ex.map(work, range(n_proofs)))
        dt = time.perf_counter() - t0
    print("%s: %2d contexts: %6.1f proofs/s (%.2f ms/proof amortised), proof %d bytes" % (algo, workers, n_proofs / dt, dt / n_proofs * 1e3, out[0]))
    for b in bes: b.close()
