//! Exports the R1CS instance of the reference's verify circuit (the `ConstraintSystem<Fq>` built exactly as in
//! src/constraints.rs:335-370) in the on-disk format `blsgpu_r1cs_load_file` / `blsgpu_r1cs_check_file` read -- the way
//! arkworks-produced matrices and assignments reach the GPU kernel (SURVEY 8(c) "parity unpinned (2)").
//!
//!   cargo run --release --example export_r1cs -- out.r1cs
//!
//! File layout (all little-endian; `tests/test_r1cs_file.py` writes the same bytes from the in-repo builder):
//!   magic "BLSR1CS1" | u32 version = 1 | u32 field_bytes = 48 | u64 nrows | u64 ncols | u64 ninstance | u64 nnz[3] | u64 nwit
//!   then for each of A, B, C:  rowptr u64[nrows + 1] | col u32[nnz] (zero-padded to a multiple of 8 bytes) | coeff [nnz][48] canonical
//!   then nwit assignments z = [instance (starting with 1) .., witness ..], each ncols x 48 bytes canonical
//! Not compiled in the authoring image (no Rust toolchain there).
use ark_bls12_381::{Config as BlsConfig, Fq};
use ark_crypto_primitives::signature::SigVerifyGadget;
use ark_ff::{BigInteger, PrimeField};
use ark_r1cs_std::prelude::*;
use ark_relations::r1cs::{ConstraintSystem, ConstraintSystemRef, OptimizationGoal, SynthesisMode};
use bls_verify_gadget::bls::{Parameters, PublicKey, Signature, BLS};
use bls_verify_gadget::constraints::{BlsSignatureVerifyGadget, ParametersVar, PublicKeyVar, SignatureVar};
use std::io::Write;

fn fq_le48(x: &Fq) -> [u8; 48] { let mut o = [0u8; 48]; o.copy_from_slice(&x.into_bigint().to_bytes_le()[..48]); o }

fn synthesize(pk_hex: &str, msg: &[u8], sig_hex: &str) -> ConstraintSystemRef<Fq> {
    let cs = ConstraintSystem::<Fq>::new_ref();
    cs.set_optimization_goal(OptimizationGoal::Constraints);
    cs.set_mode(SynthesisMode::Prove { construct_matrices: true });
    let pp = Parameters::<BlsConfig>::default();
    let pk = PublicKey::<BlsConfig>::try_from(pk_hex).unwrap();
    let sig = Signature::<BlsConfig>::try_from(sig_hex).unwrap();
    // the allocation order of src/constraints.rs:346-367: parameters constant, public key / message / signature as witnesses
    let pp_var = ParametersVar::new_constant(cs.clone(), pp).unwrap();
    let pk_var = PublicKeyVar::new_witness(cs.clone(), || Ok(pk)).unwrap();
    let msg_var: Vec<UInt8<Fq>> = msg.iter().map(|b| UInt8::new_witness(cs.clone(), || Ok(*b)).unwrap()).collect();
    let sig_var = SignatureVar::new_witness(cs.clone(), || Ok(sig)).unwrap();
    let _ok = <BlsSignatureVerifyGadget<BlsConfig> as SigVerifyGadget<BLS<BlsConfig>, Fq>>::verify(&pp_var, &pk_var, &msg_var, &sig_var).unwrap();
    cs.finalize();
    cs
}

fn main() {
    let out = std::env::args().nth(1).expect("output path");
    // the cases of src/constraints.rs:326-332 (true, false, false): three assignments of the same matrices
    let pk = "a491d1b0ecd9bb917989f0e74f0dea0422eac4a873e5e2644f368dffb9a6e20fd6e10c1b77654d067c0618f6e5a7f79a";
    let sig = "882730e5d03f6b42c3abc26d3372625034e1d871b65a8a6b900a56dae22da98abbe1b68f85e49fe7652a55ec3d0591c20767677e33e5cbb1207315c41a9ac03be39c2e7668edc043d6cb1d9fd93033caa8a1c5b0e84bedaeb6c64972503a43eb";
    let mut m1 = [0x56u8; 32]; let m0 = m1; m1[31] = 0x57; let m2 = [0x78u8; 32];
    let systems: Vec<ConstraintSystemRef<Fq>> = [&m0[..], &m1[..], &m2[..]].iter().map(|m| synthesize(pk, m, sig)).collect();
    let mats = systems[0].to_matrices().expect("matrices");
    let (nrows, ninst, ncols) = (mats.num_constraints, mats.num_instance_variables, mats.num_instance_variables + mats.num_witness_variables);
    let mut f = std::io::BufWriter::new(std::fs::File::create(out).unwrap());
    f.write_all(b"BLSR1CS1").unwrap();
    f.write_all(&1u32.to_le_bytes()).unwrap(); f.write_all(&48u32.to_le_bytes()).unwrap();
    for v in [nrows as u64, ncols as u64, ninst as u64, mats.a_num_non_zero as u64, mats.b_num_non_zero as u64, mats.c_num_non_zero as u64, systems.len() as u64] { f.write_all(&v.to_le_bytes()).unwrap(); }
    for m in [&mats.a, &mats.b, &mats.c] {
        let mut acc = 0u64; f.write_all(&acc.to_le_bytes()).unwrap();
        for row in m.iter() { acc += row.len() as u64; f.write_all(&acc.to_le_bytes()).unwrap(); }
        for row in m.iter() { for (_, col) in row { f.write_all(&(*col as u32).to_le_bytes()).unwrap(); } }
        if acc % 2 == 1 { f.write_all(&0u32.to_le_bytes()).unwrap(); }
        for row in m.iter() { for (coeff, _) in row { f.write_all(&fq_le48(coeff)).unwrap(); } }
    }
    for cs in &systems {
        let b = cs.borrow().unwrap();                                   // z = [instance_assignment (1 first), witness_assignment]: arkworks' column numbering
        for x in b.instance_assignment.iter().chain(b.witness_assignment.iter()) { f.write_all(&fq_le48(x)).unwrap(); }
        println!("is_satisfied = {:?}", cs.is_satisfied());             // all_sat of blsgpu_r1cs_check_file must agree
    }
}
