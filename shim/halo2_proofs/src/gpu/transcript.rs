//! `T: TranscriptWrite<G1Affine, Challenge255<G1Affine>>` as a zkb_transcript_vtable: create_proof is generic over the transcript
//! (Blake2bWrite in the benches, PoseidonTranscript in snark-verifier-sdk, EvmTranscript for the final layer), a generic type
//! cannot cross a C ABI, but its four operations can.  The session calls them in exactly the order halo2's prover does, so the
//! bytes end up in the caller's writer and no replay is needed.
use crate::transcript::{EncodedChallenge, TranscriptWrite};
use crate::zkb200_sys::zkb_transcript_vtable;
use halo2curves::bn256::{Fr, G1Affine};
use std::os::raw::c_void;

pub struct Bridge<'a, E: EncodedChallenge<G1Affine>, T: TranscriptWrite<G1Affine, E>> {
    pub transcript: &'a mut T,
    pub io_error: Option<std::io::Error>,
    _e: std::marker::PhantomData<E>,
}

impl<'a, E: EncodedChallenge<G1Affine>, T: TranscriptWrite<G1Affine, E>> Bridge<'a, E, T> {
    pub fn new(transcript: &'a mut T) -> Self {
        Bridge { transcript, io_error: None, _e: std::marker::PhantomData }
    }
    pub fn vtable(&mut self) -> zkb_transcript_vtable {
        zkb_transcript_vtable {
            user: self as *mut _ as *mut c_void,
            common_scalar: common_scalar::<E, T>,
            write_scalar: write_scalar::<E, T>,
            write_point: write_point::<E, T>,
            squeeze_challenge: squeeze::<E, T>,
        }
    }
    fn note(&mut self, r: std::io::Result<()>) -> i32 {
        match r {
            Ok(()) => 0,
            Err(e) => {
                self.io_error.get_or_insert(e);
                1
            }
        }
    }
}

// Fr / G1Affine are plain limb arrays in halo2curves (Montgomery form): reading them through the pointer is a copy, not a conversion
unsafe fn fr(p: *const u64) -> Fr { std::ptr::read(p as *const Fr) }
unsafe fn pt(p: *const u64) -> G1Affine { std::ptr::read(p as *const G1Affine) }

unsafe extern "C" fn common_scalar<E: EncodedChallenge<G1Affine>, T: TranscriptWrite<G1Affine, E>>(u: *mut c_void, s: *const u64) -> i32 {
    let b = &mut *(u as *mut Bridge<E, T>);
    let r = b.transcript.common_scalar(fr(s));
    b.note(r)
}
unsafe extern "C" fn write_scalar<E: EncodedChallenge<G1Affine>, T: TranscriptWrite<G1Affine, E>>(u: *mut c_void, s: *const u64) -> i32 {
    let b = &mut *(u as *mut Bridge<E, T>);
    let r = b.transcript.write_scalar(fr(s));
    b.note(r)
}
unsafe extern "C" fn write_point<E: EncodedChallenge<G1Affine>, T: TranscriptWrite<G1Affine, E>>(u: *mut c_void, p: *const u64) -> i32 {
    let b = &mut *(u as *mut Bridge<E, T>);
    let r = b.transcript.write_point(pt(p));
    b.note(r)
}
unsafe extern "C" fn squeeze<E: EncodedChallenge<G1Affine>, T: TranscriptWrite<G1Affine, E>>(u: *mut c_void, out: *mut u64) -> i32 {
    let b = &mut *(u as *mut Bridge<E, T>);
    let c: Fr = b.transcript.squeeze_challenge_scalar::<()>().into();   // ChallengeScalar<G1Affine, ()> -> Fr
    std::ptr::write(out as *mut Fr, c);
    0
}
