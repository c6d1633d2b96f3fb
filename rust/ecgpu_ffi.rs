//! ecgpu_ffi.rs — reference-side binding of libecgpu.so (SOURCE ONLY: this repository's environment has no Rust
//! toolchain, so this file is not compiled here; the compiled equivalents are elliptic-curves_b200/host/ecgpu.hpp
//! and the ctypes binding elliptic-curves_b200/ecgpu/__init__.py).  It would live in the reference as
//! `k256/src/arithmetic/gpu.rs` behind a `gpu` feature.  See INTEGRATION.md for where each wrapper plugs into the
//! existing trait surface.
#![allow(non_camel_case_types, dead_code)]

use core::ffi::c_char;

#[repr(C)]
pub struct ecg_ctx {
    _private: [u8; 0],
}

pub const ECG_OK: i32 = 0;
pub const ECG_EINVAL: i32 = 1;
pub const ECG_ESCALAR_RANGE: i32 = 2;
pub const ECG_ENOT_ON_CURVE: i32 = 3;
pub const ECG_ECUDA: i32 = 4;
pub const ECG_ENCCL: i32 = 5;
pub const ECG_ENOMEM: i32 = 6;
pub const ECG_SECP256K1: i32 = 0;
pub const ECG_NISTP256: i32 = 1;
pub const ECG_FLAG_DEVICE_PTRS: u32 = 1;

#[link(name = "ecgpu")]
unsafe extern "C" {
    pub fn ecg_ctx_create(device_ids: *const i32, n_devices: i32, flags: u32, out: *mut *mut ecg_ctx) -> i32;
    pub fn ecg_ctx_destroy(ctx: *mut ecg_ctx);
    pub fn ecg_last_error(ctx: *const ecg_ctx) -> *const c_char;
    pub fn ecg_last_error_index(ctx: *const ecg_ctx) -> usize;
    /// `ProjectivePoint * Scalar` over a batch (k256/src/arithmetic/mul.rs:236-295)
    pub fn ecg_mul_batch(ctx: *mut ecg_ctx, curve: i32, n: usize, k: *const u8, p_xy: *const u8, p_inf: *const u8,
                         out_xy: *mut u8, out_inf: *mut u8) -> i32;
    /// `ProjectivePoint::mul_by_generator` (mul.rs:180-232)
    pub fn ecg_mul_gen_batch(ctx: *mut ecg_ctx, curve: i32, n: usize, k: *const u8, out_xy: *mut u8, out_inf: *mut u8) -> i32;
    /// `LinearCombination::lincomb` (mul.rs:66-175)
    pub fn ecg_lincomb(ctx: *mut ecg_ctx, curve: i32, n: usize, k: *const u8, p_xy: *const u8, p_inf: *const u8,
                       out_xy: *mut u8, out_inf: *mut u8) -> i32;
    pub fn ecg_lincomb_partial(ctx: *mut ecg_ctx, curve: i32, n: usize, k: *const u8, p_xy: *const u8, p_inf: *const u8,
                               out_xyz: *mut u8) -> i32;
    pub fn ecg_point_sum(ctx: *mut ecg_ctx, curve: i32, m: usize, xyz: *const u8, out_xy: *mut u8, out_inf: *mut u8) -> i32;
    /// `MulByGeneratorVartime::mul_by_generator_and_mul_add_vartime` (mul.rs:303-310)
    pub fn ecg_mul_gen_add_batch(ctx: *mut ecg_ctx, curve: i32, n: usize, a: *const u8, b: *const u8, p_xy: *const u8,
                                 p_inf: *const u8, out_xy: *mut u8, out_inf: *mut u8) -> i32;
    /// `BatchNormalize::batch_normalize` (projective.rs:345-391)
    pub fn ecg_batch_normalize(ctx: *mut ecg_ctx, curve: i32, n: usize, xyz: *const u8, out_xy: *mut u8, out_inf: *mut u8) -> i32;
    /// `FieldElement::{add, sub, neg, mul, square, invert}` (field.rs:116-196)
    pub fn ecg_field_op_batch(ctx: *mut ecg_ctx, curve: i32, op: i32, n: usize, a: *const u8, b: *const u8, out: *mut u8) -> i32;
    /// `schnorr::VerifyingKey::verify_raw` over a batch (schnorr/verifying.rs:76-99)
    pub fn ecg_schnorr_verify_batch(ctx: *mut ecg_ctx, n: usize, pk_x: *const u8, msg32: *const u8, sig64: *const u8,
                                    valid: *mut u8) -> i32;
    /// `ecdsa::VerifyingKey::verify_prehash` over a batch (ecdsa.rs:93-121)
    pub fn ecg_ecdsa_verify_batch(ctx: *mut ecg_ctx, curve: i32, n: usize, z32: *const u8, sig64: *const u8, q_xy: *const u8,
                                  low_s_only: i32, valid: *mut u8) -> i32;
    /// `AffinePoint::decompress` over a batch (primeorder/src/affine.rs:179-198)
    pub fn ecg_decompress_batch(ctx: *mut ecg_ctx, curve: i32, n: usize, sec1_33: *const u8, out_xy: *mut u8,
                                out_inf: *mut u8, valid: *mut u8) -> i32;
}

#[derive(Debug)]
pub enum GpuError {
    Cuda(String),
    /// `Scalar::from_repr` would have returned `None` for this element
    ScalarRange(usize),
    /// `AffinePoint::from_coordinates` would have returned `None` for this element
    NotOnCurve(usize),
    Invalid,
}

/// One `ecg_ctx`; `!Sync` by construction (raw pointer): one engine per calling thread.
pub struct GpuEngine {
    ctx: *mut ecg_ctx,
    curve: i32,
}

impl GpuEngine {
    pub fn new(curve: i32, devices: &[i32]) -> Result<Self, GpuError> {
        let mut ctx = core::ptr::null_mut();
        // SAFETY: out-pointer is valid; the library copies `devices` before returning.
        match unsafe { ecg_ctx_create(devices.as_ptr(), devices.len() as i32, 0, &mut ctx) } {
            ECG_OK => Ok(Self { ctx, curve }),
            _ => Err(GpuError::Cuda("ecg_ctx_create failed: no CUDA device (there is no CPU fallback)".into())),
        }
    }

    /// `out[i] = k[i] * P[i]` on canonical byte encodings (`Scalar::to_bytes`, `AffineCoordinates::{x, y}`).
    pub fn mul_batch_bytes(&mut self, k: &[[u8; 32]], p_xy: &[[u8; 64]], p_inf: &[u8]) -> Result<(Vec<[u8; 64]>, Vec<u8>), GpuError> {
        let n = k.len();
        assert!(p_xy.len() == n && p_inf.len() == n);
        let mut out_xy = vec![[0u8; 64]; n];
        let mut out_inf = vec![0u8; n];
        // SAFETY: all slices hold exactly n elements of the layout include/ecgpu.h specifies.
        let rc = unsafe {
            ecg_mul_batch(self.ctx, self.curve, n, k.as_ptr().cast(), p_xy.as_ptr().cast(), p_inf.as_ptr(),
                          out_xy.as_mut_ptr().cast(), out_inf.as_mut_ptr())
        };
        self.check(rc).map(|_| (out_xy, out_inf))
    }

    /// BIP340 batch verification: one bool per (key, message, signature).
    pub fn schnorr_verify_batch(&mut self, pk: &[[u8; 32]], msg: &[[u8; 32]], sig: &[[u8; 64]]) -> Result<Vec<bool>, GpuError> {
        let n = pk.len();
        assert!(msg.len() == n && sig.len() == n);
        let mut valid = vec![0u8; n];
        // SAFETY: as above.
        let rc = unsafe { ecg_schnorr_verify_batch(self.ctx, n, pk.as_ptr().cast(), msg.as_ptr().cast(), sig.as_ptr().cast(), valid.as_mut_ptr()) };
        self.check(rc).map(|_| valid.into_iter().map(|v| v != 0).collect())
    }

    fn check(&self, rc: i32) -> Result<(), GpuError> {
        match rc {
            ECG_OK => Ok(()),
            // SAFETY: ctx is live for the lifetime of self.
            ECG_ESCALAR_RANGE => Err(GpuError::ScalarRange(unsafe { ecg_last_error_index(self.ctx) })),
            ECG_ENOT_ON_CURVE => Err(GpuError::NotOnCurve(unsafe { ecg_last_error_index(self.ctx) })),
            ECG_EINVAL => Err(GpuError::Invalid),
            _ => Err(GpuError::Cuda(unsafe { core::ffi::CStr::from_ptr(ecg_last_error(self.ctx)) }.to_string_lossy().into_owned())),
        }
    }
}

impl Drop for GpuEngine {
    fn drop(&mut self) {
        // SAFETY: created by ecg_ctx_create, destroyed exactly once.
        unsafe { ecg_ctx_destroy(self.ctx) }
    }
}
