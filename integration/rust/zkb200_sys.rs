//! `extern "C"` view of include/zkb200.h (libzkb200.so).  Field elements: 4 x u64 little-endian limbs of the canonical
//! residue; G1 points: 8 limbs (x, y); G2 points: 16 limbs (x.c0, x.c1, y.c0, y.c1); identity = all zero.
//! Every function returns 0 on success, a negative code otherwise (`zkb_last_error` has the text).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct zkb_ctx { _p: [u8; 0] }
#[repr(C)] pub struct zkb_qap { _p: [u8; 0] }
#[repr(C)] pub struct zkb_crs { _p: [u8; 0] }
#[repr(C)] pub struct zkb_bases { _p: [u8; 0] }

/// QAP by non-zero evaluations on the root domain (CSR by wire) -- the `DummyRep` data model.
#[repr(C)]
pub struct zkb_qap_host {
    pub n: u64,                      // gates = qap.degree
    pub m: u64,                      // rows incl. the unity wire = qap.u.len()
    pub n_input: u64,                // qap.input
    pub row_ptr: [*const u64; 3],    // u, v, w: m + 1 offsets
    pub gate: [*const u32; 3],       // nnz gate indices (0-based position in `roots`)
    pub coeff: [*const u64; 3],      // nnz x 4 limbs
    pub roots: *const u64,           // null: the n-th roots of unity; else n x 4 limbs (n <= 4096)
}

#[repr(C)]
pub struct zkb_crs_host {
    pub n: u64, pub n_sum_gamma: u64, pub n_sum_delta: u64,
    pub alpha1: *const u64, pub beta1: *const u64, pub delta1: *const u64,
    pub xi1: *const u64, pub xi_t: *const u64, pub sum_gamma: *const u64, pub sum_delta: *const u64,
    pub beta2: *const u64, pub gamma2: *const u64, pub delta2: *const u64, pub xi2: *const u64,
}

#[repr(C)] #[derive(Clone, Copy)]
pub struct zkb_proof { pub a: [u64; 8], pub b: [u64; 16], pub c: [u64; 8] }
impl Default for zkb_proof { fn default() -> Self { zkb_proof { a: [0; 8], b: [0; 16], c: [0; 8] } } }

extern "C" {
    pub fn zkb_ctx_create(out: *mut *mut zkb_ctx, device_id: c_int) -> c_int;
    pub fn zkb_ctx_destroy(ctx: *mut zkb_ctx);
    pub fn zkb_last_error(ctx: *const zkb_ctx) -> *const c_char;
    pub fn zkb_host_alloc(out: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn zkb_host_free(p: *mut c_void);

    pub fn zkb_qap_upload(ctx: *mut zkb_ctx, qap: *const zkb_qap_host, out: *mut *mut zkb_qap) -> c_int;
    pub fn zkb_qap_free(ctx: *mut zkb_ctx, qap: *mut zkb_qap);

    pub fn zkb_crs_upload(ctx: *mut zkb_ctx, crs: *const zkb_crs_host, rank: c_int, world: c_int, out: *mut *mut zkb_crs) -> c_int;
    pub fn zkb_setup(ctx: *mut zkb_ctx, qap: *const zkb_qap, toxic: *const u64, rank: c_int, world: c_int, out: *mut *mut zkb_crs) -> c_int;
    pub fn zkb_crs_dims(crs: *const zkb_crs, n: *mut u64, n_sum_gamma: *mut u64, n_sum_delta: *mut u64) -> c_int;
    pub fn zkb_crs_download(ctx: *mut zkb_ctx, crs: *const zkb_crs, dst: *mut zkb_crs_host) -> c_int;
    pub fn zkb_crs_free(ctx: *mut zkb_ctx, crs: *mut zkb_crs);

    pub fn zkb_prove(ctx: *mut zkb_ctx, qap: *const zkb_qap, crs: *const zkb_crs, weights: *const u64,
                     r: *const u64, s: *const u64, out: *mut zkb_proof) -> c_int;
    pub fn zkb_prove_batch(ctx: *mut zkb_ctx, qap: *const zkb_qap, crs: *const zkb_crs, weights: *const *const u64,
                           weights_on_device: c_int, r: *const u64, s: *const u64, count: usize, out: *mut zkb_proof) -> c_int;
    pub fn zkb_prove_partial(ctx: *mut zkb_ctx, qap: *const zkb_qap, crs: *const zkb_crs, weights: *const u64,
                             weights_on_device: c_int, r: *const u64, s: *const u64, out_partial: *mut u64) -> c_int;
    pub fn zkb_prove_combine(ctx: *mut zkb_ctx, partials: *const u64, world: c_int, out: *mut zkb_proof) -> c_int;

    pub fn zkb_verify(ctx: *mut zkb_ctx, crs: *const zkb_crs, inputs: *const u64, n_inputs: usize,
                      proof: *const zkb_proof, ok: *mut c_int) -> c_int;
    pub fn zkb_verify_batch(ctx: *mut zkb_ctx, crs: *const zkb_crs, inputs: *const u64, n_inputs: usize,
                            proofs: *const zkb_proof, count: usize, ok: *mut c_int) -> c_int;
    pub fn zkb_pairing(ctx: *mut zkb_ctx, g1s: *const u64, g2s: *const u64, n: usize, gt: *mut u64) -> c_int;
}
