//! Drop-in bodies for `groth16::{setup, prove, verify}` on one B200 (libzkb200.so), for the concrete BN254
//! instantiation `T = FrLocal, U = G1Local, V = G2Local` (src/groth16/fr.rs:9-16).  Lives at `src/groth16/fr/b200.rs`
//! as a CHILD module of `groth16::fr` (`mod b200;` in fr.rs): `SigmaG1`, `SigmaG2` and `Proof` have private fields and
//! no accessors (src/groth16/mod.rs:105-128) and `FrLocal(Fr)`, `G1Local(G1)`, `G2Local(G2)` have private tuple fields
//! (fr.rs:8-14); a descendant module of both sees them, so nothing in the reference has to become `pub`.
//!
//! Contract kept from the reference: synchronous calls, borrowed inputs, owned outputs, and PANICS instead of
//! `Result`s (fr.rs:54, field/mod.rs:440) -- every non-zero status of the C ABI becomes a panic with the library's text.
//!
//! NOT compiled in the zkb200 build image (no cargo / rustc there); see integration/rust/README.md.
extern crate bn;

use self::bn::{AffineG1, AffineG2, Fq, Fq2, Group, G1, G2};
use super::super::circuit::RootRepresentation;
use super::super::zkb200_sys::*;
use super::super::{Proof, Random, SigmaG1, SigmaG2};
use super::{FrLocal, G1Local, G2Local};
use std::collections::HashMap;

// ---- element conversion: canonical residues, 4 x u64 little-endian limbs -----------------------------------------
fn be32_to_limbs(be: &[u8; 32]) -> [u64; 4] {
    let mut l = [0u64; 4];
    for i in 0..4 {
        for j in 0..8 { l[i] |= (be[31 - (8 * i + j)] as u64) << (8 * j); }
    }
    l
}
fn limbs_to_be32(l: &[u64]) -> [u8; 32] {
    let mut be = [0u8; 32];
    for i in 0..4 {
        for j in 0..8 { be[31 - (8 * i + j)] = (l[i] >> (8 * j)) as u8; }
    }
    be
}
fn fr_limbs(a: &FrLocal) -> [u64; 4] {
    let mut be = [0u8; 32];
    (a.0).to_big_endian(&mut be).expect("Fr::to_big_endian");
    be32_to_limbs(&be)
}
fn fq_limbs(a: Fq) -> [u64; 4] {
    let mut be = [0u8; 32];
    a.to_big_endian(&mut be).expect("Fq::to_big_endian");
    be32_to_limbs(&be)
}
fn fq_from(l: &[u64]) -> Fq { Fq::from_slice(&limbs_to_be32(l)).expect("Fq::from_slice: residue out of range") }

/// G1Local -> 8 limbs (x, y); the identity (z = 0) -> all zero.
fn g1_limbs(p: &G1Local) -> [u64; 8] {
    let mut out = [0u64; 8];
    if let Some(a) = AffineG1::from_jacobian(p.0) {
        out[..4].copy_from_slice(&fq_limbs(a.x()));
        out[4..].copy_from_slice(&fq_limbs(a.y()));
    }
    out
}
fn g1_from_limbs(l: &[u64]) -> G1Local {
    if l.iter().all(|&w| w == 0) { return G1Local(G1::zero()); }
    G1Local(G1::from(AffineG1::new(fq_from(&l[0..4]), fq_from(&l[4..8])).expect("device returned a point off G1")))
}
/// G2Local -> 16 limbs (x.c0, x.c1, y.c0, y.c1) with Fq2 = c0 + c1 * u.
fn g2_limbs(p: &G2Local) -> [u64; 16] {
    let mut out = [0u64; 16];
    if let Some(a) = AffineG2::from_jacobian(p.0) {
        out[0..4].copy_from_slice(&fq_limbs(a.x().real()));
        out[4..8].copy_from_slice(&fq_limbs(a.x().imaginary()));
        out[8..12].copy_from_slice(&fq_limbs(a.y().real()));
        out[12..16].copy_from_slice(&fq_limbs(a.y().imaginary()));
    }
    out
}
fn g2_from_limbs(l: &[u64]) -> G2Local {
    if l.iter().all(|&w| w == 0) { return G2Local(G2::zero()); }
    let x = Fq2::new(fq_from(&l[0..4]), fq_from(&l[4..8]));
    let y = Fq2::new(fq_from(&l[8..12]), fq_from(&l[12..16]));
    G2Local(G2::from(AffineG2::new(x, y).expect("device returned a point off the twist")))
}
fn g1_vec(ps: &[G1Local]) -> Vec<u64> { ps.iter().flat_map(|p| g1_limbs(p).to_vec()).collect() }
fn g2_vec(ps: &[G2Local]) -> Vec<u64> { ps.iter().flat_map(|p| g2_limbs(p).to_vec()).collect() }

// ---- handles -----------------------------------------------------------------------------------------------------
pub struct B200 { ctx: *mut zkb_ctx }
pub struct DeviceQap<'a> { gpu: &'a B200, h: *mut zkb_qap, m: usize }
pub struct DeviceCrs<'a> { gpu: &'a B200, h: *mut zkb_crs }

fn check(ctx: *const zkb_ctx, rc: i32) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(zkb_last_error(ctx)) };
        panic!("zkb200 error {}: {}", rc, msg.to_string_lossy());
    }
}

impl B200 {
    /// One context per GPU; not thread-safe, distinct contexts are.  Panics without an sm_100 device: no CPU fallback.
    pub fn new(device: i32) -> Self {
        let mut ctx = std::ptr::null_mut();
        check(std::ptr::null(), unsafe { zkb_ctx_create(&mut ctx, device) });
        B200 { ctx }
    }

    /// `QAP::from(root_rep)` (fr.rs:140-173) without densifying: the sparse (root, value) rows go to the device as
    /// CSR by wire.  `omega_domain`: the caller asserts that `roots` are w^0 .. w^(n-1) for the n-th root of unity
    /// w = 5^((r-1)/n) (n a power of two): NTT path, any n up to 2^27.  Otherwise the roots are passed explicitly
    /// (the ASTParser's 1..=n, circuit/mod.rs:517) and the dense O(n^2) path serves n <= 32768.
    pub fn upload_qap<R: RootRepresentation<FrLocal>>(&self, rep: &R, omega_domain: bool) -> DeviceQap {
        let roots: Vec<FrLocal> = rep.roots().collect();
        let index: HashMap<[u64; 4], u32> = roots.iter().enumerate().map(|(k, r)| (fr_limbs(r), k as u32)).collect();
        let mut row_ptr: [Vec<u64>; 3] = [vec![0], vec![0], vec![0]];
        let mut gate: [Vec<u32>; 3] = [Vec::new(), Vec::new(), Vec::new()];
        let mut coeff: [Vec<u64>; 3] = [Vec::new(), Vec::new(), Vec::new()];
        let mats = vec![rep.u(), rep.v(), rep.w()];  // (a Vec: by-value iteration on the 2015 edition)
        for (k, rows) in mats.into_iter().enumerate() {
            for row in rows {
                for (root, value) in row {
                    let g = *index.get(&fr_limbs(&root)).expect("row references a value that is not a root");
                    gate[k].push(g);
                    coeff[k].extend_from_slice(&fr_limbs(&value));
                }
                row_ptr[k].push(gate[k].len() as u64);
            }
        }
        let m = row_ptr[0].len() - 1;
        assert_eq!(m, row_ptr[1].len() - 1);   // mod.rs:88-89
        assert_eq!(m, row_ptr[2].len() - 1);
        let roots_limbs: Vec<u64> = roots.iter().flat_map(|r| fr_limbs(r).to_vec()).collect();
        let host = zkb_qap_host {
            n: roots.len() as u64, m: m as u64, n_input: rep.input() as u64,
            row_ptr: [row_ptr[0].as_ptr(), row_ptr[1].as_ptr(), row_ptr[2].as_ptr()],
            gate: [gate[0].as_ptr(), gate[1].as_ptr(), gate[2].as_ptr()],
            coeff: [coeff[0].as_ptr(), coeff[1].as_ptr(), coeff[2].as_ptr()],
            roots: if omega_domain { std::ptr::null() } else { roots_limbs.as_ptr() },
        };
        let mut h = std::ptr::null_mut();
        check(self.ctx, unsafe { zkb_qap_upload(self.ctx, &host, &mut h) });
        DeviceQap { gpu: self, h, m }
    }

    /// A CRS produced by the reference's own `setup` (mod.rs:134-197): uploaded once, expanded into the fixed-base
    /// window tables, never touched on the host again.
    pub fn upload_crs(&self, s1: &SigmaG1<G1Local>, s2: &SigmaG2<G2Local>) -> DeviceCrs {
        let (a1, b1, d1) = (g1_limbs(&s1.alpha), g1_limbs(&s1.beta), g1_limbs(&s1.delta));
        let (xi1, xit, sg, sd) = (g1_vec(&s1.xi), g1_vec(&s1.xi_t), g1_vec(&s1.sum_gamma), g1_vec(&s1.sum_delta));
        let (b2, g2, d2, xi2) = (g2_limbs(&s2.beta), g2_limbs(&s2.gamma), g2_limbs(&s2.delta), g2_vec(&s2.xi));
        let host = zkb_crs_host {
            n: s1.xi.len() as u64, n_sum_gamma: s1.sum_gamma.len() as u64, n_sum_delta: s1.sum_delta.len() as u64,
            alpha1: a1.as_ptr(), beta1: b1.as_ptr(), delta1: d1.as_ptr(),
            xi1: xi1.as_ptr(), xi_t: xit.as_ptr(), sum_gamma: sg.as_ptr(), sum_delta: sd.as_ptr(),
            beta2: b2.as_ptr(), gamma2: g2.as_ptr(), delta2: d2.as_ptr(), xi2: xi2.as_ptr(),
        };
        let mut h = std::ptr::null_mut();
        check(self.ctx, unsafe { zkb_crs_upload(self.ctx, &host, 0, 1, &mut h) });
        DeviceCrs { gpu: self, h }
    }

    /// groth16::setup (mod.rs:134-197) on the device: fresh non-zero alpha, beta, gamma, delta, x (fr.rs:90-99).
    pub fn setup(&self, qap: &DeviceQap) -> DeviceCrs {
        let toxic: Vec<u64> = (0..5).flat_map(|_| fr_limbs(&FrLocal::random_elem()).to_vec()).collect();
        let mut h = std::ptr::null_mut();
        check(self.ctx, unsafe { zkb_setup(self.ctx, qap.h, toxic.as_ptr(), 0, 1, &mut h) });
        DeviceCrs { gpu: self, h }
    }

    /// The device-resident CRS back as the reference's types (e.g. to hand it to the reference's `verify`).
    pub fn download_crs(&self, crs: &DeviceCrs) -> (SigmaG1<G1Local>, SigmaG2<G2Local>) {
        let (mut n, mut nsg, mut nsd) = (0u64, 0u64, 0u64);
        check(self.ctx, unsafe { zkb_crs_dims(crs.h, &mut n, &mut nsg, &mut nsd) });
        let (n, nsg, nsd) = (n as usize, nsg as usize, nsd as usize);
        let (mut a1, mut b1, mut d1) = ([0u64; 8], [0u64; 8], [0u64; 8]);
        let (mut b2, mut g2, mut d2) = ([0u64; 16], [0u64; 16], [0u64; 16]);
        let mut xi1 = vec![0u64; 8 * n];
        let mut xit = vec![0u64; 8 * n.saturating_sub(1)];
        let mut sg = vec![0u64; 8 * nsg];
        let mut sd = vec![0u64; 8 * nsd];
        let mut xi2 = vec![0u64; 16 * n];
        let mut host = zkb_crs_host {
            n: n as u64, n_sum_gamma: nsg as u64, n_sum_delta: nsd as u64,
            alpha1: a1.as_mut_ptr(), beta1: b1.as_mut_ptr(), delta1: d1.as_mut_ptr(),
            xi1: xi1.as_mut_ptr(), xi_t: xit.as_mut_ptr(), sum_gamma: sg.as_mut_ptr(), sum_delta: sd.as_mut_ptr(),
            beta2: b2.as_mut_ptr(), gamma2: g2.as_mut_ptr(), delta2: d2.as_mut_ptr(), xi2: xi2.as_mut_ptr(),
        };
        check(self.ctx, unsafe { zkb_crs_download(self.ctx, crs.h, &mut host) });
        let g1s = |v: &Vec<u64>| v.chunks(8).map(g1_from_limbs).collect::<Vec<_>>();
        (SigmaG1 { alpha: g1_from_limbs(&a1), beta: g1_from_limbs(&b1), delta: g1_from_limbs(&d1), xi: g1s(&xi1),
                   sum_gamma: g1s(&sg), sum_delta: g1s(&sd), xi_t: g1s(&xit) },
         SigmaG2 { beta: g2_from_limbs(&b2), gamma: g2_from_limbs(&g2), delta: g2_from_limbs(&d2),
                   xi: xi2.chunks(16).map(g2_from_limbs).collect() })
    }

    /// groth16::prove (mod.rs:213-296): fresh r, s (mod.rs:231).
    pub fn prove(&self, qap: &DeviceQap, crs: &DeviceCrs, weights: &[FrLocal]) -> Proof<G1Local, G2Local> {
        self.prove_with_rs(qap, crs, weights, FrLocal::random_elem(), FrLocal::random_elem())
    }

    /// The same with the two blinding scalars injected: the seam the parity tests use.
    pub fn prove_with_rs(&self, qap: &DeviceQap, crs: &DeviceCrs, weights: &[FrLocal], r: FrLocal, s: FrLocal)
        -> Proof<G1Local, G2Local>
    {
        let w = Self::weight_limbs(qap, weights);
        let mut out = zkb_proof::default();
        check(self.ctx, unsafe {
            zkb_prove(self.ctx, qap.h, crs.h, w.as_ptr(), fr_limbs(&r).as_ptr(), fr_limbs(&s).as_ptr(), &mut out)
        });
        Proof { a: g1_from_limbs(&out.a), b: g2_from_limbs(&out.b), c: g1_from_limbs(&out.c) }
    }

    /// Throughput mode: many witnesses against one QAP / CRS, several proofs in flight on the device.
    /// Same results as `prove_with_rs` in a loop.
    pub fn prove_many(&self, qap: &DeviceQap, crs: &DeviceCrs, weights: &[&[FrLocal]], rs: &[(FrLocal, FrLocal)])
        -> Vec<Proof<G1Local, G2Local>>
    {
        assert_eq!(weights.len(), rs.len());
        let bufs: Vec<Vec<u64>> = weights.iter().map(|w| Self::weight_limbs(qap, w)).collect();
        let ptrs: Vec<*const u64> = bufs.iter().map(|b| b.as_ptr()).collect();
        let r: Vec<u64> = rs.iter().flat_map(|p| fr_limbs(&p.0).to_vec()).collect();
        let s: Vec<u64> = rs.iter().flat_map(|p| fr_limbs(&p.1).to_vec()).collect();
        let mut out = vec![zkb_proof::default(); weights.len()];
        check(self.ctx, unsafe {
            zkb_prove_batch(self.ctx, qap.h, crs.h, ptrs.as_ptr(), 0, r.as_ptr(), s.as_ptr(), weights.len(), out.as_mut_ptr())
        });
        out.iter().map(|p| Proof { a: g1_from_limbs(&p.a), b: g2_from_limbs(&p.b), c: g1_from_limbs(&p.c) }).collect()
    }

    /// groth16::verify (mod.rs:299-320) with the CRS resident on the device.  `inputs`: the verifier-visible weights
    /// without the leading 1 (mod.rs:313 prepends it; so does the library).
    pub fn verify(&self, crs: &DeviceCrs, inputs: &[FrLocal], proof: &Proof<G1Local, G2Local>) -> bool {
        let inp: Vec<u64> = inputs.iter().flat_map(|a| fr_limbs(a).to_vec()).collect();
        let pc = zkb_proof { a: g1_limbs(&proof.a), b: g2_limbs(&proof.b), c: g1_limbs(&proof.c) };
        let mut ok = 0;
        check(self.ctx, unsafe { zkb_verify(self.ctx, crs.h, inp.as_ptr(), inputs.len(), &pc, &mut ok) });
        ok == 1
    }

    /// Many (inputs, proof) pairs against one CRS, every pair with the same number of inputs: one verdict each.
    pub fn verify_many(&self, crs: &DeviceCrs, inputs: &[&[FrLocal]], proofs: &[Proof<G1Local, G2Local>]) -> Vec<bool> {
        assert_eq!(inputs.len(), proofs.len());
        let k = inputs.first().map_or(0, |i| i.len());
        assert!(inputs.iter().all(|i| i.len() == k));
        let inp: Vec<u64> = inputs.iter().flat_map(|i| i.iter().flat_map(|a| fr_limbs(a).to_vec())).collect();
        let pcs: Vec<zkb_proof> = proofs.iter()
            .map(|p| zkb_proof { a: g1_limbs(&p.a), b: g2_limbs(&p.b), c: g1_limbs(&p.c) }).collect();
        let mut ok = vec![0i32; proofs.len()];
        check(self.ctx, unsafe { zkb_verify_batch(self.ctx, crs.h, inp.as_ptr(), k, pcs.as_ptr(), proofs.len(), ok.as_mut_ptr()) });
        ok.into_iter().map(|v| v == 1).collect()
    }

    /// `.zip()` truncation of mod.rs:237-288 expressed as zero padding: rows beyond the weights contribute nothing.
    fn weight_limbs(qap: &DeviceQap, weights: &[FrLocal]) -> Vec<u64> {
        let mut w = vec![0u64; 4 * qap.m];
        for (dst, a) in w.chunks_mut(4).zip(weights.iter()) { dst.copy_from_slice(&fr_limbs(a)); }
        w
    }
}

impl Drop for B200 { fn drop(&mut self) { unsafe { zkb_ctx_destroy(self.ctx) } } }
impl<'a> Drop for DeviceQap<'a> { fn drop(&mut self) { unsafe { zkb_qap_free(self.gpu.ctx, self.h) } } }
impl<'a> Drop for DeviceCrs<'a> { fn drop(&mut self) { unsafe { zkb_crs_free(self.gpu.ctx, self.h) } } }

#[cfg(test)]
mod tests {
    //! The reference's own end-to-end check (fr.rs:361-416, `verify == true`) through the device.
    use super::super::super::circuit::dummy_rep::DummyRep;
    use super::super::super::{setup, verify, CoefficientPoly, QAP};
    use super::*;

    #[test]
    fn b200_prove_verifies_with_the_reference_verifier() {
        // single multiplication gate a * b = c on the root {1} (fr.rs:248-271)
        let one = FrLocal::from(1usize);
        let rep = || DummyRep {
            u: vec![vec![], vec![(one, one)], vec![], vec![]],
            v: vec![vec![], vec![], vec![(one, one)], vec![]],
            w: vec![vec![], vec![], vec![], vec![(one, one)]],
            roots: vec![one],
            input: 2,
        };
        let qap: QAP<CoefficientPoly<FrLocal>> = rep().into();
        let (s1, s2) = setup(&qap);
        let weights = vec![one, FrLocal::from(3usize), FrLocal::from(4usize), FrLocal::from(12usize)];
        let gpu = B200::new(0);
        let dq = gpu.upload_qap(&rep(), false);
        let dc = gpu.upload_crs(&s1, &s2);
        let proof = gpu.prove(&dq, &dc, &weights);
        assert!(gpu.verify(&dc, &weights[1..3], &proof));
        assert!(verify::<CoefficientPoly<FrLocal>, _, _, _, _>((s1, s2), &weights[1..3], proof));
    }
}
