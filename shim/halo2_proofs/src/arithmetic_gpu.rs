//! Bodies of `arithmetic::best_fft` / `best_multiexp` for bn256 (halo2_proofs 1.1.0 src/arithmetic.rs).  These are the entry points
//! code OUTSIDE create_proof reaches (keygen's commitments, `ParamsKZG::commit*`, snark-verifier's accumulation); the proving session
//! never goes through them (its data stays in HBM).
use crate::gpu::{check, gpu};
use crate::zkb200_sys::*;
use halo2curves::bn256::{Fr, G1Affine, G1};

pub fn best_fft_fr(a: &mut [Fr], omega: Fr, log_n: u32) {
    assert_eq!(a.len(), 1 << log_n);
    let g = gpu();
    check(unsafe { zkb_ntt_fr_host(g.ctx, a.as_mut_ptr() as *mut u64, log_n, &omega as *const Fr as *const u64, std::ptr::null(), 0) })
        .expect("zkb_ntt_fr_host");
}

pub fn best_multiexp_g1(coeffs: &[Fr], bases: &[G1Affine]) -> G1 {
    assert_eq!(coeffs.len(), bases.len());
    let g = gpu();
    let (mut aff, mut jac) = ([0u64; 8], [0u64; 12]);
    check(unsafe {
        zkb_msm_g1_host(g.ctx, coeffs.as_ptr() as *const u64, bases.as_ptr() as *const u64, coeffs.len() as u64, aff.as_mut_ptr(), jac.as_mut_ptr(),
                        std::ptr::null_mut())
    })
    .expect("zkb_msm_g1_host");
    unsafe { std::mem::transmute_copy(&jac) } // G1 { x, y, z }: z = 1, or 0 for the identity
}
