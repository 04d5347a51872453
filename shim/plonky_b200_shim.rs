//! plonky -> libplonky_b200.so shim.  SOURCE ONLY: not compiled in this repository (no Rust toolchain in
//! the image).  Drop in as `src/gpu.rs`, add `mod gpu;` to `src/lib.rs`, and replace the bodies named at
//! the bottom.  Everything above L1 (`plonk.rs`, `halo.rs`, `verifier.rs`, `poly_commit.rs`, ...) is untouched.
#![allow(non_camel_case_types, dead_code)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_uint};

use once_cell::sync::OnceCell;

use crate::{AffinePoint, Curve, Field, ProjectivePoint};
use crate::{Bls12377, Bls12377Base, Bls12377Scalar, Tweedledee, TweedledeeBase, Tweedledum, TweedledumBase};

#[repr(C)] pub struct plk_msm_table { _p: [u8; 0] }
#[repr(C)] pub struct plk_fft_plan { _p: [u8; 0] }

#[link(name = "plonky_b200")]
extern "C" {
    fn plk_last_error_message() -> *const c_char;
    fn plk_msm_precompute(curve: c_int, xyz: *const u64, zero: *const u8, n: usize, w: c_uint, out: *mut *mut plk_msm_table) -> c_int;
    fn plk_msm_execute(t: *const plk_msm_table, scalars: *const u64, n: usize, out_xyz: *mut u64, out_zero: *mut u8) -> c_int;
    fn plk_msm_execute_batch(t: *const plk_msm_table, scalars: *const u64, n: usize, k: usize, out_xyz: *mut u64, out_zero: *mut u8) -> c_int;
    fn plk_msm_parallel(curve: c_int, scalars: *const u64, xyz: *const u64, zero: *const u8, n: usize, w: c_uint, out_xyz: *mut u64, out_zero: *mut u8) -> c_int;
    fn plk_msm_free(t: *mut plk_msm_table);
    fn plk_fft_precompute(field: c_int, degree: usize, out: *mut *mut plk_fft_plan) -> c_int;
    fn plk_fft_size(p: *const plk_fft_plan) -> usize;
    fn plk_fft_pow2(p: *const plk_fft_plan, input: *const u64, out: *mut u64, n: usize) -> c_int;
    fn plk_ifft_pow2(p: *const plk_fft_plan, input: *const u64, out: *mut u64, n: usize) -> c_int;
    fn plk_fft(p: *const plk_fft_plan, input: *const u64, n_in: usize, out: *mut u64) -> c_int;
    fn plk_fft_batch(p: *const plk_fft_plan, input: *const u64, n_in: usize, k: usize, inverse: c_int, out: *mut u64) -> c_int;
    fn plk_divide_by_z_h(p: *const plk_fft_plan, coeffs: *const u64, n_in: usize, n_gates: usize, out: *mut u64) -> c_int;
    fn plk_fft_free(p: *mut plk_fft_plan);
    fn plk_batch_inverse(field: c_int, input: *const u64, out: *mut u64, n: usize) -> c_int;
    // Halo IPA rounds kept on the device (src/halo.rs:63-124)
    fn plk_ipa_new(curve: c_int, a: *const u64, b: *const u64, g_xy: *const u64, g_zero: *const u8, n: usize, out: *mut *mut plk_ipa_state) -> c_int;
    fn plk_ipa_new_with_table(t: *const plk_msm_table, a: *const u64, b: *const u64, n: usize, out: *mut *mut plk_ipa_state) -> c_int;
    fn plk_ipa_round_lr(s: *mut plk_ipa_state, l_xyz: *mut u64, l_zero: *mut u8, r_xyz: *mut u64, r_zero: *mut u8, ip_l: *mut u64, ip_r: *mut u64) -> c_int;
    fn plk_ipa_fold(s: *mut plk_ipa_state, u: *const u64, u_inv: *const u64) -> c_int;
    fn plk_ipa_read(s: *const plk_ipa_state, a: *mut u64, b: *mut u64, g_xy: *mut u64, g_zero: *mut u8) -> c_int;
    fn plk_ipa_free(s: *mut plk_ipa_state);
    // generator derivation (src/hash_to_curve.rs:53-76) and the point wire format (src/serialization.rs:32-72)
    fn plk_blake_hash_usize_to_curve(curve: c_int, seed_start: u64, n: usize, points_xy: *mut u64) -> c_int;
    fn plk_points_compress(curve: c_int, points_xy: *const u64, zero: *const u8, n: usize, out: *mut u8) -> c_int;
    fn plk_points_decompress(curve: c_int, input: *const u8, n: usize, out_xy: *mut u64, out_zero: *mut u8, out_status: *mut u8) -> c_int;
    // round 2: one-call blinded commitments (src/poly_commit.rs:32-66), vanishing_poly (src/plonk.rs:375-456), field bytes, subgroup
    fn plk_commit_batch(t: *const plk_msm_table, scalars: *const u64, n: usize, k: usize, blinding: *const u64, h_xy: *const u64, h_zero: u8,
                        out_xy: *mut u64, out_zero: *mut u8) -> c_int;
    fn plk_vanishing_poly(plan_8n: *const plk_fft_plan, degree: usize, wires_8n: *const u64, constants_8n: *const u64, sigma_8n: *const u64,
                          plonk_z_coeffs: *const u64, k_is: *const u64, alpha: *const u64, beta: *const u64, gamma: *const u64,
                          inner_zeta: *const u64, inner_a: *const u64, out_coeffs_8n: *mut u64) -> c_int;
    fn plk_field_to_bytes(field: c_int, input: *const u64, n: usize, out: *mut u8) -> c_int;
    fn plk_field_from_bytes(field: c_int, input: *const u8, n: usize, out: *mut u64) -> c_int;
    fn plk_fft_subgroup(p: *const plk_fft_plan, out: *mut u64) -> c_int;
}

/// The reference panics on this path (assert_eq! curve_msm.rs:67,106; log2_strict util.rs:16-19;
/// expect("No inverse") field.rs:267); the C ABI returns a status -- turn it back into a panic.
fn check(status: c_int) {
    if status != 0 {
        let msg = unsafe { CStr::from_ptr(plk_last_error_message()) }.to_string_lossy().into_owned();
        panic!("{}", msg);
    }
}

pub trait GpuField: Field {
    const FIELD_ID: c_int;
    const LIMBS: usize;
    fn limbs(&self) -> &[u64];
    fn from_limbs(l: &[u64]) -> Self;
}
macro_rules! gpu_field {
    ($t:ty, $id:expr, $l:expr) => {
        impl GpuField for $t {
            const FIELD_ID: c_int = $id;
            const LIMBS: usize = $l;
            fn limbs(&self) -> &[u64] { &self.limbs }
            fn from_limbs(l: &[u64]) -> Self { let mut limbs = [0u64; $l]; limbs.copy_from_slice(l); Self { limbs } }
        }
    };
}
gpu_field!(TweedledeeBase, 0, 4);
gpu_field!(TweedledumBase, 1, 4);
gpu_field!(Bls12377Scalar, 2, 4);
gpu_field!(Bls12377Base, 3, 6);

pub trait GpuCurve: Curve where Self::BaseField: GpuField, Self::ScalarField: GpuField { const CURVE_ID: c_int; }
impl GpuCurve for Tweedledee { const CURVE_ID: c_int = 0; }
impl GpuCurve for Tweedledum { const CURVE_ID: c_int = 1; }
impl GpuCurve for Bls12377 { const CURVE_ID: c_int = 2; }

fn pack_fields<F: GpuField>(xs: &[F]) -> Vec<u64> {
    let mut v = Vec::with_capacity(xs.len() * F::LIMBS);
    for x in xs { v.extend_from_slice(x.limbs()); }
    v
}
fn unpack_fields<F: GpuField>(v: &[u64]) -> Vec<F> { v.chunks(F::LIMBS).map(F::from_limbs).collect() }

/// AffinePoint / ProjectivePoint are default-repr structs (curve.rs:73-78,175-181): pack, never transmute.
fn pack_proj<C: GpuCurve>(ps: &[ProjectivePoint<C>]) -> (Vec<u64>, Vec<u8>) where C::BaseField: GpuField, C::ScalarField: GpuField {
    let mut xyz = Vec::with_capacity(ps.len() * 3 * C::BaseField::LIMBS);
    let mut zero = Vec::with_capacity(ps.len());
    for p in ps {
        xyz.extend_from_slice(p.x.limbs());
        xyz.extend_from_slice(p.y.limbs());
        xyz.extend_from_slice(p.z.limbs());
        zero.push(p.zero as u8);
    }
    (xyz, zero)
}
fn unpack_proj<C: GpuCurve>(xyz: &[u64], zero: u8) -> ProjectivePoint<C> where C::BaseField: GpuField, C::ScalarField: GpuField {
    if zero != 0 { return ProjectivePoint::ZERO; }
    let l = C::BaseField::LIMBS;
    ProjectivePoint::nonzero(C::BaseField::from_limbs(&xyz[..l]), C::BaseField::from_limbs(&xyz[l..2 * l]), C::BaseField::from_limbs(&xyz[2 * l..3 * l]))
}

pub struct GpuTable(*mut plk_msm_table);
unsafe impl Send for GpuTable {}
unsafe impl Sync for GpuTable {}
impl Drop for GpuTable { fn drop(&mut self) { unsafe { plk_msm_free(self.0) } } }

pub struct GpuPlan(*mut plk_fft_plan);
unsafe impl Send for GpuPlan {}
unsafe impl Sync for GpuPlan {}
impl Drop for GpuPlan { fn drop(&mut self) { unsafe { plk_fft_free(self.0) } } }

// ---- bodies for src/curve/curve_msm.rs -----------------------------------------------------------
/// MsmPrecomputation keeps (generators, w) for Clone / Serialize / PartialEq (curve_msm.rs:16) and a lazily
/// built device table.
pub fn gpu_table<C: GpuCurve>(generators: &[ProjectivePoint<C>], w: usize, cell: &OnceCell<GpuTable>) -> *const plk_msm_table
where C::BaseField: GpuField, C::ScalarField: GpuField {
    cell.get_or_init(|| {
        let (xyz, zero) = pack_proj(generators);
        let mut t = std::ptr::null_mut();
        check(unsafe { plk_msm_precompute(C::CURVE_ID, xyz.as_ptr(), zero.as_ptr(), generators.len(), w as c_uint, &mut t) });
        GpuTable(t)
    }).0
}
/// body of msm_execute / msm_execute_parallel (curve_msm.rs:63, :102) and pedersen_hash (plonk_util.rs:193)
pub fn gpu_msm_execute<C: GpuCurve>(table: *const plk_msm_table, scalars: &[C::ScalarField]) -> ProjectivePoint<C>
where C::BaseField: GpuField, C::ScalarField: GpuField {
    let s = pack_fields(scalars);
    let mut out = vec![0u64; 3 * C::BaseField::LIMBS];
    let mut zero = 0u8;
    check(unsafe { plk_msm_execute(table, s.as_ptr(), scalars.len(), out.as_mut_ptr(), &mut zero) });
    unpack_proj::<C>(&out, zero)
}
/// body of commit_polynomials' MSM part (plonk_util.rs:215-231): k coefficient vectors, one table
pub fn gpu_msm_execute_batch<C: GpuCurve>(table: *const plk_msm_table, rows: &[&[C::ScalarField]]) -> Vec<ProjectivePoint<C>>
where C::BaseField: GpuField, C::ScalarField: GpuField {
    let (k, n, l) = (rows.len(), rows[0].len(), C::BaseField::LIMBS);
    let mut s = Vec::with_capacity(k * n * 4);
    for r in rows { assert_eq!(r.len(), n); s.extend(pack_fields(r)); }
    let mut out = vec![0u64; k * 3 * l];
    let mut zero = vec![0u8; k];
    check(unsafe { plk_msm_execute_batch(table, s.as_ptr(), n, k, out.as_mut_ptr(), zero.as_mut_ptr()) });
    (0..k).map(|i| unpack_proj::<C>(&out[i * 3 * l..(i + 1) * 3 * l], zero[i])).collect()
}
/// body of msm_parallel (curve_msm.rs:54-61)
pub fn gpu_msm_parallel<C: GpuCurve>(scalars: &[C::ScalarField], generators: &[ProjectivePoint<C>], w: usize) -> ProjectivePoint<C>
where C::BaseField: GpuField, C::ScalarField: GpuField {
    assert_eq!(scalars.len(), generators.len());
    let (xyz, zero) = pack_proj(generators);
    let s = pack_fields(scalars);
    let mut out = vec![0u64; 3 * C::BaseField::LIMBS];
    let mut oz = 0u8;
    check(unsafe { plk_msm_parallel(C::CURVE_ID, s.as_ptr(), xyz.as_ptr(), zero.as_ptr(), scalars.len(), w as c_uint, out.as_mut_ptr(), &mut oz) });
    unpack_proj::<C>(&out, oz)
}

// ---- bodies for src/fft.rs -----------------------------------------------------------------------
pub fn gpu_plan<F: GpuField>(degree: usize, cell: &OnceCell<GpuPlan>) -> *const plk_fft_plan {
    cell.get_or_init(|| {
        let mut p = std::ptr::null_mut();
        check(unsafe { plk_fft_precompute(F::FIELD_ID, degree, &mut p) });
        GpuPlan(p)
    }).0
}
/// body of fft_with_precomputation_power_of_2 (fft.rs:103) / ifft_with_precomputation_power_of_2 (fft.rs:82)
pub fn gpu_fft_pow2<F: GpuField>(plan: *const plk_fft_plan, values: &[F], inverse: bool) -> Vec<F> {
    let input = pack_fields(values);
    let mut out = vec![0u64; input.len()];
    let rc = unsafe {
        if inverse { plk_ifft_pow2(plan, input.as_ptr(), out.as_mut_ptr(), values.len()) }
        else { plk_fft_pow2(plan, input.as_ptr(), out.as_mut_ptr(), values.len()) }
    };
    check(rc);
    unpack_fields(&out)
}
/// body of fft_with_precomputation (fft.rs:61): zero-pads to the plan size on the device
pub fn gpu_fft_padded<F: GpuField>(plan: *const plk_fft_plan, coefficients: &[F]) -> Vec<F> {
    let input = pack_fields(coefficients);
    let mut out = vec![0u64; unsafe { plk_fft_size(plan) } * F::LIMBS];
    check(unsafe { plk_fft(plan, input.as_ptr(), coefficients.len(), out.as_mut_ptr()) });
    unpack_fields(&out)
}
/// body of Polynomial::divide_by_z_h (polynomial.rs:330-380); also removes the per-call fft_precompute(8n) at :345
pub fn gpu_divide_by_z_h<F: GpuField>(plan: *const plk_fft_plan, coeffs: &[F], n_gates: usize) -> Vec<F> {
    let input = pack_fields(coeffs);
    let mut out = vec![0u64; unsafe { plk_fft_size(plan) } * F::LIMBS];
    check(unsafe { plk_divide_by_z_h(plan, input.as_ptr(), coeffs.len(), n_gates, out.as_mut_ptr()) });
    unpack_fields(&out)
}
/// body of Field::batch_multiplicative_inverse (field.rs:251-278)
pub fn gpu_batch_inverse<F: GpuField>(x: &[F]) -> Vec<F> {
    let input = pack_fields(x);
    let mut out = vec![0u64; input.len()];
    check(unsafe { plk_batch_inverse(F::FIELD_ID, input.as_ptr(), out.as_mut_ptr(), x.len()) });
    unpack_fields(&out)
}

/// halo_a / halo_b / halo_g of batch_opening_proof (halo.rs:53-57) on the device.  `table` = the circuit's
/// pedersen_g_msm_precomputation: with it G is never folded (plk_ipa_new_with_table).
#[repr(C)] pub struct plk_ipa_state { _private: [u8; 0] }
pub struct GpuIpa<C: GpuCurve>(*mut plk_ipa_state, std::marker::PhantomData<C>);
impl<C: GpuCurve> Drop for GpuIpa<C> { fn drop(&mut self) { unsafe { plk_ipa_free(self.0) } } }
impl<C: GpuCurve> GpuIpa<C> where C::BaseField: GpuField, C::ScalarField: GpuField {
    pub fn new(table: *const plk_msm_table, a: &[C::ScalarField], b: &[C::ScalarField]) -> Self {
        let (pa, pb) = (pack_fields(a), pack_fields(b));
        let mut s = std::ptr::null_mut();
        check(unsafe { plk_ipa_new_with_table(table, pa.as_ptr(), pb.as_ptr(), a.len(), &mut s) });
        GpuIpa(s, std::marker::PhantomData)
    }
    /// (msm_parallel(a_lo, g_hi, 8), msm_parallel(a_hi, g_lo, 8), <a_lo, b_hi>, <a_hi, b_lo>)  -- halo.rs:87-93
    pub fn round_lr(&mut self) -> (ProjectivePoint<C>, ProjectivePoint<C>, C::ScalarField, C::ScalarField) {
        let l = <C::BaseField as GpuField>::LIMBS;
        let (mut lx, mut rx, mut lz, mut rz) = (vec![0u64; 3 * l], vec![0u64; 3 * l], 0u8, 0u8);
        let (mut il, mut ir) = (vec![0u64; 4], vec![0u64; 4]);
        check(unsafe { plk_ipa_round_lr(self.0, lx.as_mut_ptr(), &mut lz, rx.as_mut_ptr(), &mut rz, il.as_mut_ptr(), ir.as_mut_ptr()) });
        (unpack_proj::<C>(&lx, lz), unpack_proj::<C>(&rx, rz), GpuField::from_limbs(&il), GpuField::from_limbs(&ir))
    }
    /// halo.rs:117-123
    pub fn fold(&mut self, u: C::ScalarField, u_inv: C::ScalarField) {
        check(unsafe { plk_ipa_fold(self.0, u.limbs().as_ptr(), u_inv.limbs().as_ptr()) });
    }
}
/// body of `(0..degree).map(blake_hash_usize_to_curve::<C>)` (circuit_builder.rs:1127, verifier.rs:174)
pub fn gpu_pedersen_generators<C: GpuCurve>(start: usize, n: usize) -> Vec<u64> where C::BaseField: GpuField {
    let mut xy = vec![0u64; n * 2 * <C::BaseField as GpuField>::LIMBS];
    check(unsafe { plk_blake_hash_usize_to_curve(C::CURVE_ID, start as u64, n, xy.as_mut_ptr()) });
    xy          // n x (x, y) Montgomery limbs; AffinePoint::nonzero(x, y) each
}

/// body of PolynomialCommitment::coeffs_vec_to_commitments (poly_commit.rs:52-66): k MSMs + [blinding_i] H + batch_to_affine, one call
pub fn gpu_commit_batch<C: GpuCurve>(table: *const plk_msm_table, rows: &[&[C::ScalarField]], blinding: Option<&[C::ScalarField]>,
                                     h: &AffinePoint<C>) -> Vec<AffinePoint<C>>
where C::BaseField: GpuField, C::ScalarField: GpuField {
    let (k, n, l) = (rows.len(), rows[0].len(), C::BaseField::LIMBS);
    let scalars: Vec<u64> = rows.iter().flat_map(|r| r.iter().flat_map(|x| x.limbs().iter().copied())).collect();
    let blind: Option<Vec<u64>> = blinding.map(|b| b.iter().flat_map(|x| x.limbs().iter().copied()).collect());
    let mut hxy = Vec::with_capacity(2 * l);
    hxy.extend_from_slice(h.x.limbs());
    hxy.extend_from_slice(h.y.limbs());
    let (mut out, mut zero) = (vec![0u64; k * 2 * l], vec![0u8; k]);
    check(unsafe { plk_commit_batch(table, scalars.as_ptr(), n, k, blind.as_ref().map_or(std::ptr::null(), |b| b.as_ptr()), hxy.as_ptr(),
                                    h.zero as u8, out.as_mut_ptr(), zero.as_mut_ptr()) });
    (0..k).map(|i| if zero[i] != 0 { AffinePoint::ZERO } else {
        AffinePoint::nonzero(C::BaseField::from_limbs(&out[2 * l * i..2 * l * i + l]), C::BaseField::from_limbs(&out[2 * l * i + l..2 * l * (i + 1)])) }).collect()
}

/// body of Circuit::vanishing_poly (plonk.rs:375-456): rows are the 8n-point evaluations the circuit already holds
pub fn gpu_vanishing_poly<F: GpuField>(plan_8n: *const plk_fft_plan, degree: usize, wires_8n: &[Vec<F>], constants_8n: &[Vec<F>],
                                       sigma_8n: &[Vec<F>], z_coeffs: &[F], k_is: &[F], alpha: F, beta: F, gamma: F, zeta: F, a: F) -> Vec<F> {
    let rows = |m: &[Vec<F>]| -> Vec<u64> { m.iter().flat_map(|r| r.iter().flat_map(|x| x.limbs().iter().copied())).collect() };
    let flat = |v: &[F]| -> Vec<u64> { v.iter().flat_map(|x| x.limbs().iter().copied()).collect() };
    let (w, c, s, z, k) = (rows(wires_8n), rows(constants_8n), rows(sigma_8n), flat(z_coeffs), flat(k_is));
    let mut out = vec![0u64; 8 * degree * F::LIMBS];
    check(unsafe { plk_vanishing_poly(plan_8n, degree, w.as_ptr(), c.as_ptr(), s.as_ptr(), z.as_ptr(), k.as_ptr(), alpha.limbs().as_ptr(),
                                      beta.limbs().as_ptr(), gamma.limbs().as_ptr(), zeta.limbs().as_ptr(), a.limbs().as_ptr(), out.as_mut_ptr()) });
    out.chunks(F::LIMBS).map(F::from_limbs).collect()
}

// ---- edits in the reference -----------------------------------------------------------------------
// src/curve/curve_msm.rs
//   pub struct MsmPrecomputation<C> { generators: Vec<ProjectivePoint<C>>, w: usize, #[serde(skip)] gpu: OnceCell<GpuTable> }
//   pub fn msm_precompute(generators, w)        -> MsmPrecomputation { generators: generators.to_vec(), w, gpu: OnceCell::new() }
//   pub fn msm_execute[_parallel](pre, scalars) -> gpu::gpu_msm_execute(gpu::gpu_table(&pre.generators, pre.w, &pre.gpu), scalars)
//   pub fn msm_parallel(scalars, generators, w) -> gpu::gpu_msm_parallel(scalars, generators, w)
// src/fft.rs
//   pub struct FftPrecomputation<F> { degree: usize, #[serde(skip)] gpu: OnceCell<GpuPlan>, .. }
//   pub fn fft_with_precomputation_power_of_2(c, pre)  -> gpu::gpu_fft_pow2(gpu::gpu_plan::<F>(pre.degree, &pre.gpu), c, false)
//   pub fn ifft_with_precomputation_power_of_2(p, pre) -> gpu::gpu_fft_pow2(.., p, true)
//   pub fn fft_with_precomputation(c, pre)             -> gpu::gpu_fft_padded(.., c)
// src/polynomial.rs:330  divide_by_z_h -> gpu::gpu_divide_by_z_h;   src/field/field.rs:251 -> gpu::gpu_batch_inverse
// src/plonk_util.rs:215  commit_polynomials -> gpu::gpu_msm_execute_batch (one call for all polynomials)
// src/halo.rs:63-124      let mut ipa = gpu::GpuIpa::<C>::new(table, &halo_a, &halo_b); per round: let (l, r, ip_l, ip_r) = ipa.round_lr();
//                         halo_l_j = l + C::convert(l_j_blinding_factor) * pedersen_h + C::convert(ip_l) * u_prime  (unchanged host code);
//                         after the challenge: ipa.fold(u_j, u_j_inv); at the end plk_ipa_read gives halo_a[0], halo_b[0], halo_g
// src/circuit_builder.rs:1127, src/verifier.rs:174   pedersen_g -> gpu::gpu_pedersen_generators::<C>(0, degree)
// src/serialization.rs:32-72   Vec<AffinePoint<C>> (de)serialisation of proofs / VKs -> plk_points_compress / plk_points_decompress
// src/serialization.rs:17-30   Field ToBytes / FromBytes over slices -> plk_field_to_bytes / plk_field_from_bytes
// src/poly_commit.rs:52-66     coeffs_vec_to_commitments -> gpu::gpu_commit_batch (the caller draws the blinding factors, :39-43)
// src/plonk.rs:375-456         vanishing_poly -> gpu::gpu_vanishing_poly(self.fft_precomputation_8n plan, ..., InnerC::ZETA, InnerC::A)
// src/plonk.rs:47-51           subgroup_n / subgroup_8n -> plk_fft_subgroup
