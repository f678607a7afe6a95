"""Host-side mirror of the reference's `ZKOperator` for the stwo engine
(/root/reference/js/src/stwo/operator.ts:87-191, interface /root/reference/js/src/types.ts:220-240):
generate_witness -> JSON bytes, groth16_prove -> {"proof": <base64 str>}, over the CUDA backend."""
import base64
import json

from . import backend


def make_stwo_zk_operator(algorithm="chacha20", device=0):
    if algorithm not in ("chacha20", "aes-128-ctr", "aes-256-ctr"):
        raise backend.BackendError("algorithm %r not available in this build" % algorithm)
    be = backend.Backend(device)
    prove = {"chacha20": be.generate_chacha20_proof, "aes-128-ctr": be.generate_aes128_ctr_proof,
             "aes-256-ctr": be.generate_aes256_ctr_proof}[algorithm]

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
            res = prove(base64.b64decode(w["key"]), base64.b64decode(w["nonce"]), w["counter"],
                        base64.b64decode(w["plaintext"]), base64.b64decode(w["ciphertext"]))
            if "error" in res:
                raise backend.BackendError(res["error"])
            return {"proof": res["proof"]}

        def release(self):
            be.close()

    return _Op()
