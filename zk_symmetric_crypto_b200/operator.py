"""Host-side mirror of the reference's `ZKOperator` for the stwo engine
(/root/reference/js/src/stwo/operator.ts:87-191, interface /root/reference/js/src/types.ts:220-240):
generate_witness -> JSON bytes, groth16_prove -> {"proof": <bincode bytes>} on the CUDA backend, groth16_verify -> bool on the
host verifier."""
import base64
import json

from . import backend


def make_stwo_zk_operator(algorithm="chacha20", device=0):
    if algorithm not in ("chacha20", "aes-128-ctr", "aes-256-ctr"):
        raise backend.BackendError("algorithm %r not available in this build" % algorithm)
    be = None

    def _be():
        nonlocal be
        if be is None:  # verification alone never opens a CUDA context
            be = backend.Backend(device)
        return be
    prove_name = {"chacha20": "generate_chacha20_proof", "aes-128-ctr": "generate_aes128_ctr_proof",
                  "aes-256-ctr": "generate_aes256_ctr_proof"}[algorithm]

    class _Op:
        def generate_witness(self, inp):
            """operator.ts:91 -- witness = JSON{algorithm,key,nonce,counter,plaintext(out),ciphertext(in)} as bytes."""
            w = {"algorithm": algorithm, "key": base64.b64encode(bytes(inp["key"])).decode(),
                 "nonce": base64.b64encode(bytes(inp["nonce"])).decode(), "counter": int(inp["counter"]),
                 "plaintext": base64.b64encode(bytes(inp["out"])).decode(),
                 "ciphertext": base64.b64encode(bytes(inp["in"])).decode()}
            return json.dumps(w).encode()

        def groth16_prove(self, witness):
            """operator.ts:97-133 (name kept from the ZKOperator interface)."""
            w = json.loads(bytes(witness).decode())
            if w.get("algorithm") != algorithm:
                raise backend.BackendError("Unsupported algorithm: %s" % w.get("algorithm"))
            res = getattr(_be(), prove_name)(base64.b64decode(w["key"]), base64.b64decode(w["nonce"]), w["counter"],
                                             base64.b64decode(w["plaintext"]), base64.b64decode(w["ciphertext"]))
            if "error" in res:
                raise backend.BackendError("Stwo proof generation failed: %s" % res["error"])
            if not res.get("proof"):
                raise backend.BackendError("Stwo proof generation failed: no proof returned")
            return {"proof": base64.b64decode(res["proof"])}  # operator.ts:131: binary, like the gnark operator

        def groth16_verify(self, public_signals, proof, logger=None):
            """operator.ts:135-180: public_signals = {"noncesAndCounters": [{"nonce", "counter"}], "in": ciphertext,
            "out": plaintext}; proof = bytes or base64 str.  False on any error (logged through logger.warn if given)."""
            nc = (public_signals.get("noncesAndCounters") or [{}])[0]
            nonce, counter = nc.get("nonce"), nc.get("counter")
            if not nonce or counter is None:
                if logger:
                    logger.warn("Invalid publicSignals: missing nonce or counter")
                return False
            if not isinstance(counter, int) or counter < 0 or counter > 0xFFFFFFFF:
                raise backend.BackendError("counter must be a u32")  # operator.ts:150 assertU32Counter
            b64 = proof if isinstance(proof, str) else base64.b64encode(bytes(proof)).decode()
            fn = backend.verify_chacha20_proof if algorithm == "chacha20" else backend.verify_aes_ctr_proof
            res = fn(b64, nonce, counter, public_signals["out"], public_signals["in"])
            if res.get("error"):
                if logger:
                    logger.warn("Stwo STARK verification failed: %s" % res["error"])
                return False
            return res.get("valid") is True

        def release(self):
            if be is not None:
                be.close()

    return _Op()
