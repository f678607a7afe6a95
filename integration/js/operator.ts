/**
 * makeStwoCudaZkOperator -- the reference's stwo ZKOperator (js/src/stwo/operator.ts:87-191) backed by the B200 prover
 * library libs2c_b200.so (C ABI: include/s2c_b200.h) instead of the WASM module.  Drop this file into
 * js/src/stwo-cuda/operator.ts of reclaimprotocol/zk-symmetric-crypto; it binds the shared library with koffi exactly like the
 * gnark operator binds libprove/libverify (js/src/gnark/utils.ts:31-80).
 *
 * NOT BUILT OR RUN IN THIS REPOSITORY: the build image has no node/npm.  The Python mirror of the same logic
 * (zk_symmetric_crypto_b200/operator.py) is what the tests exercise; argument order, JSON results and error strings are the
 * reference's (stwo/src/wasm_api.rs:467-648, 652-946).
 */
import { Base64 } from 'js-base64'
import koffi from 'koffi'
import type { EncryptionAlgorithm, MakeZKOperatorOpts, ZKOperator, ZKProofInput } from '../types.ts'

type StwoWitnessData = {
	algorithm: EncryptionAlgorithm
	key: string // base64
	nonce: string // base64
	counter: number
	plaintext: string // base64
	ciphertext: string // base64
}

type ProveResult = { success?: boolean, error?: string, proof?: string, blocks?: number, algorithm?: string, proof_size_bytes?: number }
type VerifyResult = { valid?: boolean, error?: string, algorithm?: string }

const PROVE_SIG = '(void* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter, '
	+ 'const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, _Out_ void** json, _Out_ size_t* json_len)'
const VERIFY_SIG = '(const char* proof_b64, size_t proof_b64_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter, '
	+ 'const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, _Out_ void** json, _Out_ size_t* json_len)'

type Lib = {
	prove: Record<EncryptionAlgorithm, (...args: unknown[]) => number>
	verifyChaCha: (...args: unknown[]) => number
	verifyAes: (...args: unknown[]) => number
	free: (p: unknown) => void
}

let lib: Lib | undefined

/** Loads libs2c_b200.so once per process (S2C_B200_LIB overrides the path, as in the Python binding). */
function ensureLibLoaded(): Lib {
	if(lib) {
		return lib
	}

	const so = koffi.load(process.env.S2C_B200_LIB ?? 'libs2c_b200.so')
	lib = {
		prove: {
			'chacha20': so.func('int s2c_generate_chacha20_proof' + PROVE_SIG),
			'aes-128-ctr': so.func('int s2c_generate_aes128_ctr_proof' + PROVE_SIG),
			'aes-256-ctr': so.func('int s2c_generate_aes256_ctr_proof' + PROVE_SIG),
		},
		verifyChaCha: so.func('int s2c_verify_chacha20_proof' + VERIFY_SIG),
		verifyAes: so.func('int s2c_verify_aes_ctr_proof' + VERIFY_SIG),
		free: so.func('void s2c_free(void* p)'),
	}
	return lib
}

/** Reads the malloc'd (ptr, len) UTF-8 result and releases it -- the counterpart of __wbindgen_free in s2circuits.cjs:100-119. */
function takeJson<T>(l: Lib, out: [unknown], len: [number | bigint]): T {
	if(!out[0]) {
		throw new Error('libs2c_b200: no result buffer returned')
	}

	try {
		const bytes = koffi.decode(out[0], koffi.array('uint8_t', Number(len[0])))
		return JSON.parse(Buffer.from(bytes as Uint8Array).toString('utf8'))
	} finally {
		l.free(out[0])
	}
}

function assertU32Counter(counter: number): void {
	if(!Number.isInteger(counter) || counter < 0 || counter > 0xFFFFFFFF) {
		throw new RangeError('counter must be a uint32 integer (0 to 4294967295)')
	}
}

function serializeWitness(algorithm: EncryptionAlgorithm, input: ZKProofInput): Uint8Array {
	if(!input.noncesAndCounters?.length) {
		throw new Error('noncesAndCounters must be a non-empty array')
	}

	const { noncesAndCounters: [{ nonce, counter }] } = input
	assertU32Counter(counter)
	// 'in' is ciphertext and 'out' is plaintext in the JS library; stwo expects (key, nonce, counter, plaintext, ciphertext)
	const data: StwoWitnessData = {
		algorithm,
		key: Base64.fromUint8Array(input.key),
		nonce: Base64.fromUint8Array(nonce),
		counter,
		plaintext: Base64.fromUint8Array(input.out),
		ciphertext: Base64.fromUint8Array(input.in),
	}
	return new TextEncoder().encode(JSON.stringify(data))
}

export function makeStwoCudaZkOperator({ algorithm }: MakeZKOperatorOpts<{}>): ZKOperator {
	return {
		generateWitness(input) {
			return serializeWitness(algorithm, input)
		},

		async groth16Prove(witness) {
			const l = ensureLibLoaded()
			const data: StwoWitnessData = JSON.parse(new TextDecoder().decode(witness))
			const prove = l.prove[data.algorithm]
			if(!prove) {
				throw new Error(`Unsupported algorithm: ${data.algorithm}`)
			}

			const key = Base64.toUint8Array(data.key)
			const nonce = Base64.toUint8Array(data.nonce)
			const plaintext = Base64.toUint8Array(data.plaintext)
			const ciphertext = Base64.toUint8Array(data.ciphertext)
			const out: [unknown] = [null]
			const len: [number] = [0]
			// ctx = null: the library's process-wide context on device 0 (calls are serialised inside the library); a server
			// that owns several GPUs creates one context per (GPU, stream) with cb_init and passes it here
			prove(null, key, key.length, nonce, nonce.length, data.counter, plaintext, plaintext.length,
				ciphertext, ciphertext.length, out, len)
			const result = takeJson<ProveResult>(l, out, len)
			if(result.error) {
				throw new Error(`Stwo proof generation failed: ${result.error}`)
			}

			if(!result.proof) {
				throw new Error('Stwo proof generation failed: no proof returned')
			}

			return { proof: Base64.toUint8Array(result.proof) }
		},

		async groth16Verify(publicSignals, proof, logger) {
			const l = ensureLibLoaded()
			const expectedNonce = publicSignals.noncesAndCounters[0]?.nonce
			const expectedCounter = publicSignals.noncesAndCounters[0]?.counter
			const expectedCiphertext = publicSignals.in
			const expectedPlaintext = publicSignals.out
			if(!expectedNonce || expectedCounter === undefined) {
				logger?.warn('Invalid publicSignals: missing nonce or counter')
				return false
			}

			assertU32Counter(expectedCounter)
			const proofStr = typeof proof === 'string' ? proof : Base64.fromUint8Array(proof)
			const out: [unknown] = [null]
			const len: [number] = [0]
			const verify = algorithm === 'chacha20' ? l.verifyChaCha : l.verifyAes
			// verification is host code inside the library (no GPU needed); verdicts equal the WASM verifier's
			verify(proofStr, Buffer.byteLength(proofStr), expectedNonce, expectedNonce.length, expectedCounter,
				expectedPlaintext, expectedPlaintext.length, expectedCiphertext, expectedCiphertext.length, out, len)
			const result = takeJson<VerifyResult>(l, out, len)
			if(result.error) {
				logger?.warn({ error: result.error }, 'Stwo STARK verification failed')
				return false
			}

			return result.valid === true
		},

		release() {
			// the shared library stays mapped (like the WASM module); contexts created with cb_init are the caller's to destroy
		}
	}
}
