/* plonky_b200 -- C ABI of the B200-native MSM / NTT prover core.
 *
 * This is the drop-in boundary for the L1 "bulk kernel" layer of 0xPolygonZero/plonky
 * (SURVEY.md section 8(b)).  The reference has no FFI today: the path sits behind plain generic
 * Rust functions re-exported from the crate root (src/lib.rs:21-23).  Each entry point below
 * names the reference function whose BODY it replaces (file:line relative to the reference root);
 * INTEGRATION.md shows the Rust shim a maintainer would add.
 *
 * Data layout (all entry points)
 *   field element : L little-endian u64 limbs in MONTGOMERY form, fully reduced, exactly the
 *                   `limbs` array of the reference structs (src/field/tweedledee_base.rs:14-18):
 *                   L = 4 for TweedledeeBase / TweedledumBase / Bls12377Scalar, L = 6 for Bls12377Base.
 *   affine point  : x then y (2*L limbs)  + one u8 zero flag  (AffinePoint, src/curve/curve.rs:73-78)
 *   projective    : x, y, z (3*L limbs)   + one u8 zero flag  (ProjectivePoint, src/curve/curve.rs:175-181;
 *                   homogeneous: the affine point is (x/z, y/z)).
 *   The Rust structs are not repr(C): the shim packs them into these SoA buffers, never transmutes.
 *   `zero` arrays may be NULL on input (= no identity points).
 *   MSM results are returned NORMALISED: z = ONE (R mod p) and (x, y) the unique affine
 *   coordinates, or x = y = z = 0 with *zero = 1.  ProjectivePoint equality in the reference is
 *   projective equivalence (src/curve/curve.rs:280-302), so this is a valid return value for
 *   every caller, and it makes results bit-comparable.
 *
 * Errors: the reference panics on this path (assert_eq! src/curve/curve_msm.rs:67,106; log2_strict
 *   src/util.rs:16-19; "No inverse" src/field/field.rs:267).  The C ABI never aborts: it returns a
 *   plk_status and the shim turns non-zero into panic!.
 *
 * Threading: every call is re-entrant.  Host-buffer calls run on a per-calling-thread CUDA stream;
 *   *_dev calls run on the stream the caller passes (cudaStream_t as void*).  Handles are immutable
 *   after creation except for a mutex-guarded pool of scratch buffers, one set per executing stream,
 *   so executes issued on different streams against the same table may overlap.
 *
 * Host-pointer entry points include the host<->device copies; *_dev entry points take device
 *   pointers (cudaMalloc'ed, 16-byte aligned) and neither copy nor synchronise.
 */
#ifndef PLONKY_B200_H
#define PLONKY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLK_ABI_VERSION 1

typedef enum plk_status {
  PLK_OK = 0,
  PLK_EINVAL = 1,    /* unknown field/curve id, NULL pointer, window out of range            */
  PLK_ELENGTH = 2,   /* scalars.len() != precomputation.len()  (curve_msm.rs:67,106)         */
  PLK_ENOTPOW2 = 3,  /* log2_strict on a non power of two     (util.rs:16-19)                */
  PLK_ESIZE = 4,     /* transform length != precomputation size (fft.rs:107-111)             */
  PLK_EZERO = 5,     /* inverse of zero requested              (field.rs:267 "No inverse")   */
  PLK_ECUDA = 6,     /* CUDA runtime error (plk_last_error_message has the text)             */
  PLK_ENOMEM = 7,
  PLK_ETOOBIG = 8    /* log2(n) exceeds the field's TWO_ADICITY (field.rs:430)               */
} plk_status;

/* Field::* implementors of the reference */
enum {
  PLK_FIELD_TWEEDLEDEE_BASE = 0, /* src/field/tweedledee_base.rs  (scalar field of Tweedledum) */
  PLK_FIELD_TWEEDLEDUM_BASE = 1, /* src/field/tweedledum_base.rs  (scalar field of Tweedledee) */
  PLK_FIELD_BLS12_377_SCALAR = 2, /* src/field/bls12_377_scalar.rs */
  PLK_FIELD_BLS12_377_BASE = 3   /* src/field/bls12_377_base.rs   (6 limbs)                    */
};
/* Curve implementors of the reference */
enum {
  PLK_CURVE_TWEEDLEDEE = 0, /* src/curve/tweedledee_curve.rs: base 0, scalars 1 */
  PLK_CURVE_TWEEDLEDUM = 1, /* src/curve/tweedledum_curve.rs: base 1, scalars 0 */
  PLK_CURVE_BLS12_377 = 2   /* src/curve/bls12_377_curve.rs:  base 3, scalars 2 */
};

const char* plk_status_string(int status);
const char* plk_last_error_message(void);       /* thread-local */
int plk_abi_version(void);
int plk_device_count(int* count);
int plk_set_device(int device);                 /* device used by the calling thread */
int plk_field_limbs(int field);                 /* u64 limbs per element, 0 if unknown */
int plk_curve_base_field(int curve);
int plk_curve_scalar_field(int curve);

/* ------------------------------------------------------------------------------------------
 * MSM  (src/curve/curve_msm.rs)
 * ---------------------------------------------------------------------------------------- */
typedef struct plk_msm_table plk_msm_table;     /* = MsmPrecomputation<C> (curve_msm.rs:16-25), device resident */

/* msm_precompute(generators: &[ProjectivePoint<C>], w) (curve_msm.rs:27-38).
 * `w` is accepted for interface fidelity (1 <= w <= 32); it is a tuning knob of the reference's
 * table layout, not semantics -- the device picks its own window for the same group element. */
int plk_msm_precompute(int curve, const uint64_t* points_xyz, const uint8_t* zero, size_t n,
                       unsigned w, plk_msm_table** out);
/* same from affine generators (the form `pedersen_g` is produced in, circuit_builder.rs:1127-1133) */
int plk_msm_precompute_affine(int curve, const uint64_t* points_xy, const uint8_t* zero, size_t n,
                              unsigned w, plk_msm_table** out);
size_t plk_msm_table_len(const plk_msm_table* t);
unsigned plk_msm_table_window(const plk_msm_table* t);   /* the w the caller passed */
void plk_msm_free(plk_msm_table* t);
/* Geometry the device chose for this table (measurement / bench.py): out[0] = window bits c, out[1] = number of
 * windows, out[2] = bucket accumulation mode (0: XYZZ mixed additions, 1: batched-affine tree), out[3] = Montgomery
 * products per mixed XYZZ addition (10; a batched-affine addition is 6), out[4] = batched-affine rounds in front of the
 * XYZZ stage.  Returns the number of values written (<= cap). */
int plk_msm_table_info(const plk_msm_table* t, unsigned* out, int cap);

/* msm_execute / msm_execute_parallel(&precomputation, scalars) (curve_msm.rs:63, :102) and
 * pedersen_hash (src/plonk_util.rs:193-198).  n must equal the table length (PLK_ELENGTH). */
int plk_msm_execute(const plk_msm_table* t, const uint64_t* scalars, size_t n,
                    uint64_t* out_xyz, uint8_t* out_zero);
/* k scalar vectors against one table: the body of commit_polynomials -> coeffs_vec_to_commitments
 * without blinding (src/plonk_util.rs:215-231, src/poly_commit.rs:52-66).  scalars is k*n*4 limbs,
 * out_xyz k*3*L limbs, out_zero k bytes. */
int plk_msm_execute_batch(const plk_msm_table* t, const uint64_t* scalars, size_t n, size_t k,
                          uint64_t* out_xyz, uint8_t* out_zero);
/* commit_polynomials -> PolynomialCommitment::coeffs_vec_to_commitments (src/plonk_util.rs:215-231,
 * src/poly_commit.rs:32-66) in ONE call: for each of the k coefficient vectors (k*n*4 limbs) the commitment
 * pedersen_hash(coeffs_i) + [blinding_i] * H, then batch_to_affine over all k.  blinding: k*4 limbs (the factors the
 * caller drew, poly_commit.rs:39-43: the RNG stays with the caller) or NULL for blinding = false; h_xy / h_zero: the
 * blinding point (pedersen_h).  out_xy: k affine points (2*L limbs each), out_zero: k flags. */
int plk_commit_batch(const plk_msm_table* t, const uint64_t* scalars, size_t n, size_t k, const uint64_t* blinding,
                     const uint64_t* h_xy, uint8_t h_zero, uint64_t* out_xy, uint8_t* out_zero);
/* msm_parallel(scalars, generators, w) = precompute + execute (curve_msm.rs:54-61), used with
 * changing bases by the Halo IPA (src/halo.rs:87,91,122): no table is kept. */
int plk_msm_parallel(int curve, const uint64_t* scalars, const uint64_t* points_xyz,
                     const uint8_t* zero, size_t n, unsigned w, uint64_t* out_xyz, uint8_t* out_zero);

/* Device-resident variants.  d_scalars: n*4 u64 (Montgomery).  d_out_xyz: 3*L u64, d_out_zero: 1 byte
 * (padded to 8).  Asynchronous on `stream`. */
int plk_msm_execute_dev(const plk_msm_table* t, const void* d_scalars, size_t n, void* d_out_xyz,
                        void* d_out_zero, void* stream);
/* k scalar vectors (k*n*4 u64, row-major) against one table on device buffers: d_out_xyz k*3*L u64,
 * d_out_zero k bytes.  The k executes are forked onto internal streams (reduction tails overlap with the
 * next accumulation) and joined back on `stream`. */
int plk_msm_execute_batch_dev(const plk_msm_table* t, const void* d_scalars, size_t n, size_t k,
                              void* d_out_xyz, void* d_out_zero, void* stream);
/* Multi-GPU: un-normalised partial sum (4*L u64, XYZZ coordinates) of this rank's shard ...      */
int plk_msm_execute_partial_dev(const plk_msm_table* t, const void* d_scalars, size_t n,
                                void* d_partial, void* stream);
/* ... and the reduction of `count` gathered partials to one normalised point. */
int plk_msm_combine_partials_dev(int curve, const void* d_partials, size_t count, void* d_out_xyz,
                                 void* d_out_zero, void* stream);
size_t plk_msm_partial_limbs(int curve);        /* u64 limbs of one partial (4*L) */
/* msm_parallel (curve_msm.rs:54-61) on device buffers: d_points_xy n affine points (2*L u64 each, identity = (0, 0)),
 * d_scalars n*4 u64.  Table-free variable-base path (per-window buckets + Horner); asynchronous on `stream`.  Shares no
 * table, window choice or reduction kernel with plk_msm_execute_dev, which is why bench.py cross-checks one against the
 * other. */
int plk_msm_parallel_dev(int curve, const void* d_scalars, const void* d_points_xy, size_t n, void* d_out_xyz,
                         void* d_out_zero, void* stream);

/* ------------------------------------------------------------------------------------------
 * NTT  (src/fft.rs)
 * ---------------------------------------------------------------------------------------- */
typedef struct plk_fft_plan plk_fft_plan;       /* = FftPrecomputation<F> (fft.rs:28-40), device resident */

/* fft_precompute(degree) (fft.rs:47-59): the plan size is 2^log2_ceil(degree). */
int plk_fft_precompute(int field, size_t degree, plk_fft_plan** out);
size_t plk_fft_size(const plk_fft_plan* p);     /* FftPrecomputation::size (fft.rs:36-40) */
/* Memory / speed knob.  The passes after the first multiply every element by an inter-pass twiddle w_N^e.  By default a
 * pass whose twiddle table has at most 2^24 entries reads it from a full table (built lazily, per direction: up to
 * N * 32 bytes for the last pass of a size-N transform); with log2_entries = 0 every twiddle is formed on the fly from two
 * tables of sqrt(N) entries and one extra product (no N-entry table ever exists, ~10 % slower at 2^24).  Applies to
 * transforms issued after the call; tables already built are kept. */
int plk_fft_set_direct_log(plk_fft_plan* p, int log2_entries);
void plk_fft_free(plk_fft_plan* p);

/* fft_with_precomputation_power_of_2(coefficients, &pre) (fft.rs:103-156): natural order in,
 * natural order out, out[k] = sum_j c_j w^(jk), w = primitive_root_of_unity(log2 n).
 * n must be a power of two (PLK_ENOTPOW2) equal to the plan size (PLK_ESIZE). */
int plk_fft_pow2(const plk_fft_plan* p, const uint64_t* in, uint64_t* out, size_t n);
/* ifft_with_precomputation_power_of_2 (fft.rs:82-101) */
int plk_ifft_pow2(const plk_fft_plan* p, const uint64_t* in, uint64_t* out, size_t n);
/* fft_with_precomputation (fft.rs:61-80): zero-pads n_in <= size coefficients; out has `size` elements */
int plk_fft(const plk_fft_plan* p, const uint64_t* in, size_t n_in, uint64_t* out);
/* k transforms of the same size (values_to_polynomials / polynomials_to_values_padded,
 * src/plonk_util.rs:169-190): in = k rows of n_in elements, out = k rows of `size` elements;
 * inverse != 0 selects the IFFT (then n_in must equal size). */
int plk_fft_batch(const plk_fft_plan* p, const uint64_t* in, size_t n_in, size_t k, int inverse,
                  uint64_t* out);
/* Coset low-degree extension, fused: out[k] = sum_j (c_j g^j) w^(jk) over `size` points with the
 * n_in coefficients zero-padded -- Polynomial::divide_by_z_h's first half (src/polynomial.rs:336-347)
 * when shift == NULL (g = MULTIPLICATIVE_SUBGROUP_GENERATOR); `shift` (L limbs, Montgomery) overrides g. */
int plk_coset_lde(const plk_fft_plan* p, const uint64_t* coeffs, size_t n_in, const uint64_t* shift,
                  uint64_t* out);
/* Inverse of the above: IFFT then scale coefficient i by g^-i (src/polynomial.rs:368-378). */
int plk_coset_ifft(const plk_fft_plan* p, const uint64_t* evals, const uint64_t* shift, uint64_t* out);
/* Polynomial::divide_by_z_h (src/polynomial.rs:330-380) for a coefficient vector of `size` elements and
 * the vanishing polynomial X^n_gates - 1: coset LDE, pointwise * 1/(g^n w^(n i) - 1), coset IFFT,
 * all on device. */
int plk_divide_by_z_h(const plk_fft_plan* p, const uint64_t* coeffs, size_t n_in, size_t n_gates,
                      uint64_t* out);

/* Polynomial::mul (src/polynomial.rs:209-227): out = a * b as 2^log2_ceil(deg a + deg b + 1) coefficients
 * (*out_len; the reference returns the un-trimmed IFFT output), or the single coefficient 0 when either operand is
 * zero.  Trailing zero coefficients of a and b are ignored like Polynomial::degree does.  out_cap elements must be
 * available (PLK_ESIZE otherwise).  The per-call fft_precompute of the reference (:217) becomes a cached plan. */
int plk_poly_mul(int field, const uint64_t* a, size_t na, const uint64_t* b, size_t nb, uint64_t* out,
                 size_t out_cap, size_t* out_len);

/* permutation_polynomial (src/plonk_util.rs:233-262): the values of Plonk's Z on the subgroup.
 *   out[0] = 1,  out[i] = out[i-1] * prod_j (w(i-1, j) + beta k_j x_{i-1} + gamma) / prod_j (w(i-1, j) + beta s_j(i-1) + gamma)
 * for j < num_routed (NUM_ROUTED_WIRES = 6, src/plonk.rs:22).  subgroup: degree elements x_i; wires: Witness::wire_values
 * row-major, wire (i, j) at i * wire_stride + j (wire_stride = NUM_WIRES = 9); sigma: num_routed rows of sigma_row_len
 * elements, s_j(i) at sigma_stride * i (the reference reads the 8n-point LDE with stride 8, :252); k_is: the routed
 * shifts get_subgroup_shift(j) (they come from a seeded ChaCha RNG in the reference: caller-supplied); beta, gamma.
 * One thread per gate, Montgomery's trick for the n divisions, prefix-product scan.  PLK_EZERO if a denominator is 0. */
int plk_permutation_polynomial(int field, size_t degree, unsigned num_routed, const uint64_t* subgroup,
                               const uint64_t* wires, size_t wire_stride, const uint64_t* sigma,
                               size_t sigma_row_len, size_t sigma_stride, const uint64_t* k_is,
                               const uint64_t* beta, const uint64_t* gamma, uint64_t* out);

/* Circuit::vanishing_poly (src/plonk.rs:375-456) for a 4-limb scalar field: the pointwise part (plk_vanishing_points*)
 * and the whole function (plk_vanishing_poly).  m = 8 * degree points x_i = w_8n^i; row-major inputs:
 *   wires_8n      9 rows (NUM_WIRES) of m evaluations, constants_8n 6 rows (NUM_CONSTANTS), sigma_8n 6 rows (NUM_ROUTED_WIRES),
 *   z_8n          m evaluations of Plonk's Z, subgroup_8n the m points (plk_fft_subgroup of the 8n plan),
 *   k_is          the 6 routed shifts get_subgroup_shift(j) (seeded ChaCha RNG in the reference: caller-supplied),
 *   alpha, beta, gamma the challenges; inner_zeta, inner_a = InnerC::ZETA and InnerC::A of the recursion's inner curve
 *   (InnerC::BaseField = this field; curve_endo.rs:55, curve_dbl.rs:45).
 * out_8n[i] = reduce_with_powers([L_1(x)(Z(x) - 1), Z(x) f'(x) - g'(x) Z(g x)] ++ evaluate_all_constraints(...), alpha)
 * with evaluate_all_constraints = src/gates/mod.rs:46-125 over the ten gates of src/gates/.  "right" is the point 8 steps
 * on, "below" 8 * GRID_WIDTH (65) steps on (plonk.rs:407-409).  d_params (device variant): k_is[6], alpha, beta, gamma,
 * inner_zeta, inner_a as 11 consecutive elements. */
int plk_vanishing_points(int field, size_t degree, const uint64_t* wires_8n, const uint64_t* constants_8n,
                         const uint64_t* sigma_8n, const uint64_t* z_8n, const uint64_t* subgroup_8n,
                         const uint64_t* k_is, const uint64_t* alpha, const uint64_t* beta, const uint64_t* gamma,
                         const uint64_t* inner_zeta, const uint64_t* inner_a, uint64_t* out_8n);
int plk_vanishing_points_dev(int field, size_t degree, const void* d_wires_8n, const void* d_constants_8n,
                             const void* d_sigma_8n, const void* d_z_8n, const void* d_subgroup_8n,
                             const void* d_params, void* d_out_8n, void* stream);
/* The whole of vanishing_poly against fft_precomputation_8n: pad_to_8n + FFT of the n Z coefficients (plonk.rs:388-391),
 * the pointwise evaluation, Polynomial::from_evaluations (plonk.rs:455) -- 8n coefficients out, nothing leaves the device
 * in between.  PLK_ESIZE unless the plan has size 8 * degree. */
int plk_vanishing_poly(const plk_fft_plan* plan_8n, size_t degree, const uint64_t* wires_8n,
                       const uint64_t* constants_8n, const uint64_t* sigma_8n, const uint64_t* plonk_z_coeffs,
                       const uint64_t* k_is, const uint64_t* alpha, const uint64_t* beta, const uint64_t* gamma,
                       const uint64_t* inner_zeta, const uint64_t* inner_a, uint64_t* out_coeffs_8n);
/* w^k for k < size: the plan's evaluation domain in natural order (Field::cyclic_subgroup_known_order,
 * src/field/field.rs:292-300; subgroup_n / subgroup_8n of Circuit, src/plonk.rs:47-51). */
int plk_fft_subgroup(const plk_fft_plan* p, uint64_t* out);

/* Device-resident variants: d_in / d_out hold n_in resp. `size` elements per row, k rows.
 * flags: bit0 inverse, bit1 coset shift by the field generator on the coefficient side. */
#define PLK_FFT_INVERSE 1u
#define PLK_FFT_COSET 2u
int plk_fft_dev(const plk_fft_plan* p, const void* d_in, size_t n_in, size_t k, unsigned flags,
                void* d_out, void* stream);

/* Domain-split transform of size N = R1 * M over `world` GPUs (one process per GPU; the caller performs
 * the single all-to-all between the phases, e.g. torch.distributed.all_to_all_single over NCCL):
 *   layout in : rank r holds rows j_1 in [row_base, row_base + rows) of x[j_1 + R1 * j'], each of length M
 *   phase A   : size-M transforms (plan_m) * w_N^(j_1 k') -> d_send in the send layout
 *               [dest][row][k' mod (M / world)]; d_work is an M * rows element work buffer
 *   exchange  : all-to-all with equal splits
 *   phase B   : in place on the received buffer [j_1][M / world]: transform over j_1 (R1 = 2^log_r1 <= 256)
 *   layout out: rank s holds X[k' + M * k_1] at [k_1][kl], k' = s * M / world + kl
 * flags: PLK_FFT_INVERSE.  plan_n is the plan of size N (its twiddle tables are used), plan_m of size M. */
int plk_fft_dist_phase_a(const plk_fft_plan* plan_m, const plk_fft_plan* plan_n, const void* d_in,
                         size_t rows, size_t row_base, unsigned world, unsigned flags, void* d_work,
                         void* d_send, void* stream);
int plk_fft_dist_phase_b(const plk_fft_plan* plan_n, void* d_recv, unsigned log_r1, unsigned log_cols,
                         unsigned flags, void* stream);
/* Phase A with the exchange fused into the kernel (NVLink peer stores instead of a collective): the last pass writes
 * each value directly into the receive buffer of the GPU that owns its column block.  peer_recv: `world` device
 * pointers (host array), peer_recv[d] = rank d's receive buffer ([R1][M / world] elements) as mapped into THIS process
 * (plk_ipc_open; the rank's own buffer for d == rank).  world <= 8.  The caller must order "every rank finished phase A"
 * before phase B reads (e.g. one 1-element all-reduce on the stream) and use two receive buffers alternately so that a
 * fast rank's next transform cannot overwrite a buffer a slow rank is still reading. */
int plk_fft_dist_phase_a_p2p(const plk_fft_plan* plan_m, const plk_fft_plan* plan_n, const void* d_in,
                             size_t rows, size_t row_base, unsigned world, unsigned flags, void* d_work,
                             void* const* peer_recv, void* stream);
/* cudaMalloc'ed buffers that the other processes of the node can map over NVLink (cudaIpcGetMemHandle /
 * cudaIpcOpenMemHandle): plk_ipc_alloc returns the buffer and its 64-byte handle (exchange it with the peers by any
 * means, e.g. torch.distributed.all_gather_object), plk_ipc_open maps a peer's handle, plk_ipc_close unmaps it,
 * plk_ipc_free releases an owned buffer.  plk_copy_dev: asynchronous device-to-device copy on `stream`. */
int plk_ipc_alloc(size_t bytes, void** d_ptr, uint8_t handle[64]);
int plk_ipc_open(const uint8_t handle[64], void** d_ptr);
int plk_ipc_close(void* d_ptr);
int plk_ipc_free(void* d_ptr);
int plk_copy_dev(void* d_dst, const void* d_src, size_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Helpers on the same arithmetic (used by the parity tests and by callers that stay on device)
 * ---------------------------------------------------------------------------------------- */
/* elementwise field ops over n elements (src/field/monty.rs:38-177):
 * 0 add, 1 sub, 2 mul, 3 square, 4 neg, 5 inverse (Fermat ladder), 6 to_canonical, 7 from_canonical, 8 double,
 * 9 inverse by the binary extended GCD (the reference's algorithm, src/bigint/bigint_inverse.rs:6-55) */
int plk_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* Field ToBytes / FromBytes (src/serialization.rs:17-30): n elements <-> n * 8*L bytes, the canonical value
 * little-endian (to_canonical_u8_vec).  from_bytes returns PLK_EINVAL ("Out of range") for a value >= the modulus,
 * where the reference returns io::Error. */
int plk_field_to_bytes(int field, const uint64_t* in, size_t n, uint8_t* out);
int plk_field_from_bytes(int field, const uint8_t* in, size_t n, uint64_t* out);
/* Field::batch_multiplicative_inverse (src/field/field.rs:251-278); PLK_EZERO if any input is 0 */
int plk_batch_inverse(int field, const uint64_t* in, uint64_t* out, size_t n);
/* ProjectivePoint::batch_to_affine (src/curve/curve.rs:216-232) */
int plk_batch_to_affine(int curve, const uint64_t* points_xyz, const uint8_t* zero, size_t n,
                        uint64_t* out_xy, uint8_t* out_zero);
/* affine_summation_best / affine_multisummation_best (src/curve/curve_summations.rs:18-35): sums of affine
 * points; `lists` lists are the slices [offsets[i], offsets[i+1]) of points_xy.  Results are normalised
 * projective points (3*L limbs each) + zero flags.  An empty list sums to ZERO (curve_summations.rs:170). */
int plk_affine_summation(int curve, const uint64_t* points_xy, const uint8_t* zero, size_t n,
                         uint64_t* out_xyz, uint8_t* out_zero);
int plk_affine_multisummation(int curve, const uint64_t* points_xy, const uint8_t* zero,
                              const uint64_t* offsets, size_t lists, uint64_t* out_xyz, uint8_t* out_zero);
/* n independent CurveScalar * ProjectivePoint products (src/curve/curve_multiplication.rs:63-70), e.g. the
 * blinding terms of PolynomialCommitment::coeffs_to_commitment (src/poly_commit.rs:44). */
int plk_curve_mul(int curve, const uint64_t* points_xyz, const uint8_t* zero, const uint64_t* scalars,
                  size_t n, uint64_t* out_xyz, uint8_t* out_zero);
/* Synthetic generator set for benchmarks: P_i = [splitmix64(seed + i)] * G, affine, written to device
 * (d_points_xy: n*2*L u64).  Stands in for blake_hash_usize_to_curve (src/hash_to_curve.rs:53-76). */
int plk_points_generate_dev(int curve, uint64_t seed, size_t n, void* d_points_xy, void* stream);
int plk_points_generate(int curve, uint64_t seed, size_t n, uint64_t* points_xy);
int plk_msm_precompute_affine_dev(int curve, const void* d_points_xy, size_t n, unsigned w,
                                  plk_msm_table** out);

/* ------------------------------------------------------------------------------------------
 * Halo inner-product argument rounds  (src/halo.rs:63-124, the loop of batch_opening_proof)
 * ---------------------------------------------------------------------------------------- */
/* halo_a, halo_b (n scalar-field elements each) and halo_g (n points) kept in HBM across the log2(n) rounds
 * (halo.rs:53-57).  The challenger (Rescue sponge), the blinding terms [l_j] H / [r_j] H and the
 * [<a, b>] U' terms stay with the caller (plk_curve_mul covers the single scalar multiplications). */
typedef struct plk_ipa_state plk_ipa_state;
/* a, b: n*4 limbs (Montgomery); g_xy: n affine points (2*L limbs each) + optional zero flags.  n must be a power
 * of two (PLK_ENOTPOW2 = log2_strict(degree), halo.rs:62). */
int plk_ipa_new(int curve, const uint64_t* a, const uint64_t* b, const uint64_t* g_xy, const uint8_t* g_zero,
                size_t n, plk_ipa_state** out);
/* Same rounds against the fixed-base table of the n generators (the prover already holds
 * pedersen_g_msm_precomputation, src/plonk.rs:64-69): G is never folded.  The folded generators are kept as
 * coefficients c_i over the original ones, L_j / R_j become two fixed-base MSMs per round and halo_g one MSM of the
 * c_i at the end -- the same group elements as halo.rs:87-123 without 255 doublings per generator and round.
 * The table is borrowed and must outlive the state; n must equal its length (PLK_ELENGTH).  plk_ipa_read returns
 * g only once the length is 1. */
int plk_ipa_new_with_table(const plk_msm_table* table, const uint64_t* a, const uint64_t* b, size_t n,
                           plk_ipa_state** out);
size_t plk_ipa_len(const plk_ipa_state* s);     /* current vector length (halves with every fold) */
void plk_ipa_free(plk_ipa_state* s);
/* First half of a round (halo.rs:87-93): out_l = msm_parallel(a_lo, G_hi, 8), out_r = msm_parallel(a_hi, G_lo, 8)
 * as normalised projective points (3*L limbs + zero flag), out_ip_l = <a_lo, b_hi>, out_ip_r = <a_hi, b_lo>
 * (Field::inner_product, src/field/field.rs:214-221; 4 limbs each).  PLK_EINVAL once the length is 1. */
int plk_ipa_round_lr(plk_ipa_state* s, uint64_t* out_l_xyz, uint8_t* out_l_zero, uint64_t* out_r_xyz,
                     uint8_t* out_r_zero, uint64_t* out_ip_l, uint64_t* out_ip_r);
/* Second half (halo.rs:117-123) with the round challenge u = u_j and u_inv = u_j^-1 (4 limbs each):
 * a <- u_inv a_hi + u a_lo, b <- u_inv b_lo + u b_hi, G_i <- msm_parallel([u_inv, u], [G_lo_i, G_hi_i], 4). */
int plk_ipa_fold(plk_ipa_state* s, const uint64_t* u, const uint64_t* u_inv);
/* Copy the current vectors out (any pointer may be NULL): after the last fold these are halo_a[0], halo_b[0]
 * and halo_g[0].to_affine() (halo.rs:126-131).  g_xy: len*2*L limbs, g_zero: len bytes. */
int plk_ipa_read(const plk_ipa_state* s, uint64_t* a, uint64_t* b, uint64_t* g_xy, uint8_t* g_zero);

/* ------------------------------------------------------------------------------------------
 * Generator derivation and the point wire format  (src/hash_to_curve.rs, src/serialization.rs)
 * ---------------------------------------------------------------------------------------- */
/* blake_hash_usize_to_curve(seed) for seed = seed_start .. seed_start + n - 1 (src/hash_to_curve.rs:53-76): the
 * derivation of pedersen_g (seeds 0..degree), pedersen_h (degree) and U (degree + 1) in CircuitBuilder::build
 * (src/circuit_builder.rs:1127-1129) and verify_proof (src/verifier.rs:151-174).  BLAKE3 (the reference's blake3
 * crate) is implemented on device for the single-block inputs this needs.  Output: n affine points (2*L limbs). */
int plk_blake_hash_usize_to_curve(int curve, uint64_t seed_start, size_t n, uint64_t* points_xy);
int plk_blake_hash_usize_to_curve_dev(int curve, uint64_t seed_start, size_t n, void* d_points_xy, void* stream);
/* blake_hash_base_field_to_curve(seed) for n arbitrary base-field seeds (L limbs each, Montgomery) (:57-76) */
int plk_blake_hash_base_field_to_curve(int curve, const uint64_t* seeds, size_t n, uint64_t* points_xy);
/* AffinePoint ToBytes / FromBytes (src/serialization.rs:32-72): per point one mask byte (bit 0: zero, bit 1: y is
 * odd) followed by x as 8*L canonical little-endian bytes; plk_point_compressed_bytes = 1 + 8*L is the stride.
 * (The reference's `read` consumes only the mask byte of a zero point although `write` emits 1 + 8*L bytes; the
 * fixed stride here follows `write`.)  Decompression recomputes y with Field::square_root
 * (src/field/field.rs:440-473, the reference's Tonelli-Shanks loop, hence the same root).  out_status (optional,
 * n bytes): 0 ok, 1 "Out of range", 2 "Invalid x coordinate"; any non-zero status makes the call return PLK_EINVAL
 * (the reference returns io::Error). */
size_t plk_point_compressed_bytes(int curve);
int plk_points_compress(int curve, const uint64_t* points_xy, const uint8_t* zero, size_t n, uint8_t* out);
int plk_points_decompress(int curve, const uint8_t* in, size_t n, uint64_t* out_xy, uint8_t* out_zero,
                          uint8_t* out_status);

/* number of CUDA kernels this library has launched in the calling process (bench.py: gpu_launches) */
uint64_t plk_kernel_launch_count(void);

/* Measurement hooks (bench.py's live roofline numbers).  With profiling enabled every MSM execute /
 * transform records CUDA events between its kernels on the launching stream; the *_last_*_ms calls
 * synchronise on the last event and return the number of phases written (negative status on error).
 * MSM phases: count, scan, scatter, accumulate, bucket_sum, range, final.  NTT: one phase per pass. */
int plk_set_profiling(int enabled);
/* Montgomery products per second the device sustains in `field` (4 independent dependent chains per thread, one
 * resident wave): the arithmetic ceiling the MSM / NTT kernels are quoted against beside the HBM roofline. */
int plk_measure_mul_throughput(int field, double* products_per_s);
int plk_msm_last_phase_ms(const plk_msm_table* t, float* out_ms, int cap);
int plk_fft_last_pass_ms(const plk_fft_plan* p, float* out_ms, int cap);
int plk_fft_num_passes(const plk_fft_plan* p);

#ifdef __cplusplus
}
#endif
#endif /* PLONKY_B200_H */
