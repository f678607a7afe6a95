//! Out-of-container confirmation harness: runs the REFERENCE's own test-data generator
//! `prove_stream::<Blake2sMerkleChannel>(n, PcsConfig::default())` (stwo/src/chacha/bitwise/air_stream.rs:237-289) and the
//! B200 library's reproduction of it, and compares the bincode bytes and the three commitment roots.
//!     cargo run --release --features reference --example prove_stream_dump -- 6
use s2circuits::chacha::bitwise::air_stream::prove_stream;
use stwo::core::pcs::PcsConfig;
use stwo::core::vcs_lifted::blake2_merkle::Blake2sMerkleChannel;

fn main() {
    let n: u32 = std::env::args().nth(1).and_then(|s| s.parse().ok()).unwrap_or(4);
    let reference = bincode::serialize(&prove_stream::<Blake2sMerkleChannel>(n, PcsConfig::default())).expect("serialize");
    let ctx = s2c_b200::Ctx::new(0).expect("CUDA device");
    let ours = ctx.prove_stream_testdata(n).expect("prove");
    // StreamStatement = u32 log_size + 12 nonce + u32 counter + 2 x 32 hash bytes = 84 bytes; PcsConfig = 25 bytes; then the
    // commitments vector (u64 length + 32-byte roots)
    let roots = |p: &[u8]| (0..3).map(|i| hex(&p[117 + 32 * i..149 + 32 * i])).collect::<Vec<_>>();
    println!("log_size {n}: reference {} bytes, b200 {} bytes", reference.len(), ours.len());
    println!("reference roots {:?}\nb200 roots      {:?}", roots(&reference), roots(&ours));
    assert_eq!(reference, ours, "proof bytes differ");
    println!("byte-identical");
}

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{x:02x}")).collect()
}
