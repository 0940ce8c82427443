/* l2b.h -- C ABI of libl2b (l2hmc_b200/csrc), the B200 (sm_100a) kernels behind
 * the leapfrog-integrator hot path of saforem2/l2hmc-qcd.
 *
 * The reference has no FFI: its hot path is three Python classes built in
 * trainers/pytorch/trainer.py:490-506,540-562 (LatticeSU3/LatticeU1, Dynamics,
 * NetworkFactory).  This header is the thin native layer our mirrors of those
 * classes (l2hmc_b200/{group,lattice,dynamics}) call through ctypes; each entry
 * point names the reference code it replaces (paths relative to
 * /root/reference/src/l2hmc).
 *
 * Conventions
 *  - every function returns 0 on success or a negative L2B_ERR_* code; a
 *    human-readable message for the calling thread is in l2b_last_error();
 *  - all pointers are DEVICE pointers unless named host_*; the caller owns and
 *    allocates every buffer (no hidden allocation, no ownership transfer);
 *    scratch space is passed in (`ws`, `ws_bytes`) and sized by *_ws_bytes();
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it
 *    and re-entrant on distinct streams with distinct workspaces;
 *  - SU(3) fields use the reference layout [nb, 4, T, X, Y, Z, 3, 3] complex,
 *    interleaved (re, im)  (configs.py:501-507); U(1) fields are [nb, 2, T, X];
 *  - `dtype` is the REAL scalar type of the field: L2B_F64 (complex128 / float64)
 *    or L2B_F32.  SU(3) entry points currently implement L2B_F64 only (the only
 *    SU(3) precision the reference configures, conf/experiment/su3.yaml) and
 *    return L2B_ERR_UNSUPPORTED otherwise.
 *  - per-chain reductions are deterministic (fixed-order two-stage sums).
 *  - step sizes: every L2HMC update takes `double eps, const <real>* eps_dev`.  With eps_dev ==
 *    NULL the step size is `eps`; otherwise the kernel reads it from device memory as
 *    eps * (*eps_dev) (eps is then a host-side multiplier, +-1 in practice), so a trainable step
 *    size never has to visit the host and whole steps can be captured in CUDA graphs.
 */
#ifndef L2B_H
#define L2B_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L2B_OK 0
#define L2B_ERR_INVALID (-1)     /* bad argument (null pointer, non-positive size ...) */
#define L2B_ERR_UNSUPPORTED (-2) /* dtype / option not implemented                     */
#define L2B_ERR_WORKSPACE (-3)   /* ws_bytes smaller than *_ws_bytes()                  */
#define L2B_ERR_CUDA (-4)        /* a CUDA runtime call failed                          */

#define L2B_F32 0
#define L2B_F64 1
#define L2B_BF16 2 /* element type of net-side buffers only (vec8, s/t/q, GEMM operands) */

const char* l2b_last_error(void);
int l2b_version(void);
/* first 32 bits of the SHA-256 of the include/l2b.h this library was compiled against (set by the
 * build as -DL2B_ABI_HASH=...; 0 for a build that did not pass it).  The ctypes binding compares it
 * with the header next to it and refuses / rebuilds a stale binary: many entry points take
 * positional pointer arguments, a mismatch would be silent memory corruption */
uint32_t l2b_abi_hash(void);
/* same over the header and every file under csrc/: identifies the source tree the binary was built from */
uint32_t l2b_source_hash(void);
/* number of kernel launches issued by this library on behalf of the calling
 * process since load (bench.py reports it as gpu_launches) */
uint64_t l2b_launch_count(void);
/* tuning knobs (process-wide): "su3_force_variant" = launch geometry of the force
 * kernel, see kForceVariants in csrc/l2b_su3.cu; "su3_fuse_drift" = 0/1, run the drift as its
 * own kernel (6 transfers per step) or fused into the force kernel (4); "su3_fuse_conversions" = 0/1,
 * l2b_su3_hmc_trajectory converts all four fields with separate kernels, or reads / writes the momenta and
 * x_prop in the boundary layout directly from its first / last launches (same bits) */
int l2b_set_option(const char* key, int value);

/* ------------------------------------------------------------------------ */
/* SU(3)                                                                     */
/* ------------------------------------------------------------------------ */

/* (no reference counterpart: torch allocates its temporaries implicitly)
 * bytes of scratch needed by any SU(3) entry point for `nb` chains of a
 * T x X x Y x Z lattice (two planar field copies + reduction partials) */
size_t l2b_su3_ws_bytes(int nb, const int dims[4], int dtype);

/* reference layout [nb,4,T,X,Y,Z,3,3] (configs.py:501-507) <-> internal planar layout
 * U[b][mu][re/im of the 9 entries][site] (exposed for tests/benchmarks) */
int l2b_su3_aos_to_soa(const void* x_aos, void* x_soa, int nb, const int dims[4], int dtype, void* stream);
int l2b_su3_soa_to_aos(const void* x_soa, void* x_aos, int nb, const int dims[4], int dtype, void* stream);

/* LatticeSU3._wilson_loops (lattice/su3/pytorch/lattice.py:157-199, c1 == 0):
 * wloops[6, nb, T, X, Y, Z] complex, planes ordered (u=1,v=0),(2,0),(2,1),(3,0),(3,1),(3,2) */
int l2b_su3_wilson_loops(const void* x, void* wloops, int nb, const int dims[4], int dtype,
                         void* ws, size_t ws_bytes, void* stream);

/* LatticeSU3.action / _plaquettes / _int_charges / _sin_charges in one pass
 * (lattice.py:201-269): sums[nb, 2] = (sum Re tr P, sum Im tr P) per chain.
 *   action = -(beta/3) sums[:,0];  plaq = sums[:,0]/(18 V);
 *   intQ = sums[:,1]/(32 pi^2);    sinQ = sums[:,1]/(18 V)                    */
int l2b_su3_plaq_sums(const void* x, double* sums, int nb, const int dims[4], int dtype,
                      void* ws, size_t ws_bytes, void* stream);

/* LatticeSU3.grad_action (lattice.py:299-308): force[nb,4,T,X,Y,Z,3,3] =
 * (beta/3) TAH(U A) with analytic staples; optional act_sums[nb] receives
 * sum Re tr P (so action comes with the force at no cost; may be NULL).       */
int l2b_su3_force(const void* x, double beta, void* force, double* plaq_sum_or_null, int nb,
                  const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream);

/* LatticeSU3.action / grad_action of the improved action, c1 != 0 (lattice.py:96-112,180-196,252-269,
 * 299-308; the reference builds 12 rectangle traces per site from bmm + roll and gets the force from autograd):
 *   force_or_null = (beta/3) TAH(U [(1 - 8 c1) A + c1 R]),  R = the 18 rectangle staples of the link;
 *   sums_or_null[nb, 2] = (sum Re tr P, sum Re tr R), so  S = -(beta/3) ((1 - 8 c1) sums[:,0] + c1 sums[:,1]). */
int l2b_su3_force_c1(const void* x, double beta, double c1, void* force_or_null, double* sums_or_null, int nb,
                     const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream);

/* SU3.exp (group/su3/pytorch/group.py:88-90): out = matrix_exp(scale * p), n matrices */
int l2b_su3_exp(const void* p, double scale, void* out, size_t nmat, int dtype, void* stream);

/* SU3.update_gauge / Dynamics._update_x_{fwd,bwd} for SU(3)
 * (group.py:45-50, dynamics/pytorch/dynamics.py:1420-1425,1468-1474):
 *   mask == NULL :  x_out = exp(eps p) x
 *   mask != NULL :  x_out = m*x + exp(eps p) ((1-m)*x),  m = mask[4*V*9] float32,
 *                   element-wise and shared by all chains (dynamics.py:1101-1110);
 *                   if mask_complement != 0 the roles of m and 1-m are swapped.
 * x_out may alias x.                                                          */
int l2b_su3_update_gauge(const void* x, const void* p, double eps, const double* eps_dev, const float* mask,
                         int mask_complement, void* x_out, int nb, const int dims[4], int dtype,
                         void* stream);

/* SU3.projectSU / compat_proj (group/su3/pytorch/utils.py:341-346), and
 * group_to_vec = su3_to_vec(projectSU(x)) (group.py:138-147) fused:
 * either output may be NULL.  vec8[nmat, 8] real.                             */
int l2b_su3_project(const void* x, void* x_proj_or_null, void* vec8_or_null, size_t nmat, int dtype,
                    void* stream);
/* su3_to_vec (utils.py:394-420) / vec_to_su3 (utils.py:423-445) without projection */
int l2b_su3_to_vec(const void* x, void* vec8, size_t nmat, int dtype, void* stream);
int l2b_su3_from_vec(const void* vec8, void* x, size_t nmat, int dtype, void* stream);
/* SU3.projectTAH (group.py:92-103) */
int l2b_su3_tah(const void* x, void* out, size_t nmat, int dtype, void* stream);
/* SU3.kinetic_energy (group.py:125-126): ke[nb] = 0.5 sum_links(|P|_F^2 - 8) */
int l2b_su3_kinetic(const void* p, double* ke, int nb, const int dims[4], int dtype, void* ws,
                    size_t ws_bytes, void* stream);
/* checkSU (utils.py:376-391): avg[nb], max[nb] */
int l2b_su3_check(const void* x, double* avg, double* max, int nb, const int dims[4], int dtype,
                  void* ws, size_t ws_bytes, void* stream);
/* SU3.random_momentum / randTAH3 (utils.py:171-195): Gaussian traceless
 * anti-Hermitian momenta, <|P|_F^2> = 8, from a counter-based Philox4x32-10
 * stream keyed by (seed, offset, link index); ke_or_null[nb] as l2b_su3_kinetic.
 * offset_dev_or_null: device-resident call counter added to `offset` and incremented by
 * one after the draw, so that a CUDA-graph replay of this call draws fresh momenta. */
int l2b_su3_rand_momentum(uint64_t seed, uint64_t offset, uint64_t* offset_dev_or_null, void* p, double* ke_or_null,
                          int nb, const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream);

/* Dynamics._update_v_fwd / _update_v_bwd epilogue (dynamics.py:1266-1297) on
 * complex v, force with REAL s, t, q of shape [nb, 4*V*9] (network outputs):
 *   forward  (sign=+1): v' = exp(eps s/2) v - eps/2 (F exp(eps q) + t), logdet = +sum eps s/2
 *   backward (sign=-1): v' = exp(-eps s/2) (v + eps/2 (F exp(eps q) + t)), logdet = -sum eps s/2
 * s/t/q may be NULL (treated as 0, i.e. a plain HMC half kick). v_out may alias v. */
int l2b_su3_vupdate(const void* v, const void* force, const void* s, const void* t, const void* q,
                    double eps, const double* eps_dev, int sign, void* v_out, double* logdet, int nb, const int dims[4],
                    int dtype, void* ws, size_t ws_bytes, void* stream);

/* Dynamics.transition_kernel_hmc (dynamics.py:900-954) for SU(3): nlf leapfrog
 * steps of size eps from (x, v) at coupling beta; kicks between consecutive
 * drifts are merged (identical up to rounding to the reference's two half
 * kicks).  Outputs: x_prop, v_prop (reference layout), and
 * energies[nb, 4] = (KE0, S0, KE1, S1) so that H = KE + S as in
 * Dynamics.hamiltonian (dynamics.py:1479-1483).                               */
int l2b_su3_hmc_trajectory(const void* x, const void* v, double beta, double eps, int nlf,
                           void* x_prop, void* v_prop, double* energies, int nb,
                           const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream);

/* --- adjoints (L2HMC training; the reference relies on autograd for these).  Gradients
 * of complex fields use torch's convention G = dL/dRe + i dL/dIm for a real loss L. --- */
/* adjoint of LatticeSU3.action (lattice.py:252-269): gx = coef[b] * A^+ (A = staple sum); with
 * coef[b] = -(beta/3) * dL/dS[b] this is dL/dx through S = -(beta/3) sum Re tr P */
int l2b_su3_action_grad(const void* x, const double* coef, void* gx, int nb, const int dims[4], int dtype, void* ws,
                        size_t ws_bytes, void* stream);
/* adjoint of l2b_su3_force with the reference's graph semantics (lattice.py:299-308: dsdx =
 * autograd(S) is a constant, only the explicit `@ x.adjoint()` is attached):
 * gx = TAH(gforce)^+ dsdx,  dsdx = -(beta/3) A^+ */
int l2b_su3_force_bwd(const void* x, double beta, const void* gforce, void* gx, int nb, const int dims[4], int dtype,
                      void* ws, size_t ws_bytes, void* stream);
/* adjoints of the improved action / force (c1 != 0; what autograd does through lattice.py:96-112,252-269,299-308):
 *   gforce_or_null == NULL:  gx = coef[b] Aimp^+,  Aimp = (1 - 8 c1) A + c1 R; coef[b] = -(beta/3) dL/dS[b]
 *   gforce_or_null != NULL:  gx = TAH(gforce)^+ (scale Aimp^+), scale = -(beta/3)  (force adjoint at fixed dsdx) */
int l2b_su3_action_grad_c1(const void* x, const double* coef_or_null, double scale, double c1,
                           const void* gforce_or_null, void* gx, int nb, const int dims[4], int dtype, void* ws,
                           size_t ws_bytes, void* stream);
/* adjoint of l2b_su3_wilson_loops: gx from the cotangent gwloops[6, nb, T, X, Y, Z] complex of the
 * per-site loops (the reference back-propagates lattice.py:157-199 through 18 bmm + 12 roll) */
int l2b_su3_wilson_loops_bwd(const void* x, const void* gwloops, void* gx, int nb, const int dims[4], int dtype,
                             void* ws, size_t ws_bytes, void* stream);
/* adjoint of l2b_su3_vupdate (dynamics.py:1266-1297) w.r.t. v, force, s, t, q (gforce/gs/gt/gq may be NULL) and eps
 * (geps[nb], per chain) */
int l2b_su3_vupdate_bwd(const void* v, const void* force, const void* s, const void* t, const void* q, double eps, const double* eps_dev,
                        int sign, const void* gv_out, const double* glogdet, void* gv, void* gforce, void* gs, void* gt,
                        void* gq, double* geps, int nb, const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream);
/* adjoint of l2b_su3_update_gauge (group.py:45-50, dynamics.py:1420-1425,1468-1474) w.r.t. x, p and eps (matrix-exponential adjoint as a Taylor
 * series on Cayley-Hamilton coefficients; *bad_flag is set when ||eps p||_F > 3 somewhere) */
int l2b_su3_update_gauge_bwd(const void* x, const void* p, double eps, const double* eps_dev, const float* mask, int mask_complement,
                             const void* gx_out, void* gx, void* gp, double* geps, int* bad_flag, int nb,
                             const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream);
/* adjoint of l2b_su3_to_vec (utils.py:394-420): gx[nmat,3,3] from gvec8[nmat,8] */
int l2b_su3_to_vec_bwd(const void* gvec8, void* gx, size_t nmat, int dtype, void* stream);

/* group_to_vec = su3_to_vec(projectSU(x)) (dynamics.py:1154-1156, group.py:138-147) with the
 * 8 reals per link written directly in the vnet's element type `vec_dtype`
 * (L2B_F64 / L2B_F32 / L2B_BF16): no separate cast pass in front of the input GEMM */
int l2b_su3_project_vec(const void* x, void* vec8, int vec_dtype, size_t nmat, int dtype, void* stream);
/* adjoint of projectSU (cotangent gmat[nmat,3,3]) and/or of group_to_vec (cotangent
 * gvec8[nmat,8] of type vec_dtype); the reference differentiates utils.py:227-346 with autograd,
 * here the closed form: polar factor + 3x3 Sylvester solve (csrc/l2b_su3_math.cuh) */
int l2b_su3_project_bwd(const void* x, const void* gmat_or_null, const void* gvec8_or_null, int vec_dtype, void* gx,
                        size_t nmat, int dtype, void* stream);

/* The two kernels of one leapfrog step on fields ALREADY in the planar layout
 * (l2b_su3_aos_to_soa), for callers that keep the state planar between steps
 * and for per-kernel timing (bench.py); together they are Dynamics.leapfrog_hmc (dynamics.py:900-913):
 *   force_kick:  P <- P - eps_kick * (beta/3) TAH(U A); sums_or_null[nb, 2] =
 *                (sum Re tr P_plaq, sum_links(|P|_F^2 - 8)) after the kick
 *   drift:       U <- exp(eps P) U                                            */
int l2b_su3_force_kick_planar(const void* u_planar, void* p_planar, double beta, double eps_kick,
                              double* sums_or_null, int nb, const int dims[4], int dtype, void* ws,
                              size_t ws_bytes, void* stream);
int l2b_su3_drift_planar(void* u_planar, const void* p_planar, double eps, int nb, const int dims[4],
                         int dtype, void* stream);
/* one whole leapfrog step (dynamics.py:900-913, kicks merged) in ONE kernel (what l2b_su3_hmc_trajectory runs):
 *   P <- P - eps_kick (beta/3) TAH(U A);   u_out <- exp(eps_drift P) u_in
 * u_out must not alias u_in (other links still read their old neighbours):
 * 4 field transfers per step instead of 6.  sums_or_null as above.            */
int l2b_su3_force_kick_drift_planar(const void* u_in_planar, void* p_planar, void* u_out_planar, double beta,
                                    double eps_kick, double eps_drift, double* sums_or_null, int nb,
                                    const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream);

/* adjoint of l2b_su3_heads_vupdate (network.py:536-548 + dynamics.py:1266-1297), element-wise part (the GEMMs of
 * the Linear backward are l2b_gemm_bf16 calls): from (s, t, q) as dumped by the forward (stq f32 [3, nb, xdim]) and the
 * cotangents gv_out, glogdet it writes gv, gforce (may be NULL), geps[nb], the cotangents of the heads' PRE-activations
 * gpre [3, nb, xdim] in gpre_dtype (L2B_F32 / L2B_BF16: the dtype of the dz / dW GEMMs) and their sums over the
 * chains colsum f32 [5, xdim]: rows 0-2 = the three heads' bias gradients, rows 3, 4 = sum_b gs*s and sum_b gq*q,
 * the ScaledTanh.coeff gradients.  ws: nb * ceil(xdim/256) doubles. */
int l2b_su3_heads_vupdate_bwd(const void* v, const void* force, const float* stq, const float* scale_s,
                              const float* scale_q, float scale_t, double eps, const double* eps_dev, int sign,
                              const void* gv_out, const double* glogdet, void* gv, void* gforce_or_null, void* gpre,
                              int gpre_dtype, float* colsum, double* geps, int nb, int xdim, void* ws,
                              size_t ws_bytes, void* stream);

/* L2HMC sweep (Dynamics.transition_kernel_fb, dynamics.py:956-1029) with the state kept in the planar layout
 * (no conversion around the stencil kernels): lattice.py:299-308, group.py:138-147 and dynamics.py:1420-1425 as
 * force without kick, group_to_vec (vec8 in [b][mu][site][8] order, as the AoS version) and the masked
 * link update; the mask is the [xdim] element mask permuted to [4][9][V].  The heads kernel
 * l2b_su3_heads_vupdate is layout-agnostic: pack the head weights with their rows permuted the same way. */
int l2b_su3_force_planar(const void* u_planar, double beta, void* f_planar, int nb, const int dims[4], int dtype,
                         void* stream);
int l2b_su3_project_vec_planar(const void* x_planar, void* vec8, int vec_dtype, int nb, const int dims[4], int dtype,
                               void* stream);
int l2b_su3_update_gauge_planar(const void* x_planar, const void* p_planar, double eps, const double* eps_dev,
                                const float* mask_planar, int mask_complement, void* x_out_planar, int nb,
                                const int dims[4], int dtype, void* stream);

/* both masked link updates of one leapfrog layer (dynamics.py:1195-1198 forward: mask m then 1 - m; :1217-1220
 * backward: 1 - m then m -> first_complement = 1) in one pass: no momentum update separates them and they share the
 * step size, so exp(eps p) is formed once and x, p are read once.  Same arithmetic per update as
 * l2b_su3_update_gauge_planar applied twice. */
int l2b_su3_update_gauge_planar_pair(const void* x_planar, const void* p_planar, double eps, const double* eps_dev,
                                     const float* mask_planar, int first_complement, void* x_out_planar, int nb,
                                     const int dims[4], int dtype, void* stream);

/* ------------------------------------------------------------------------ */
/* vnet INPUT layer on the tensor cores (tcgen05)                             */
/* ------------------------------------------------------------------------ */
/* InputLayer of the SU(3) vnet (network/pytorch/network.py:349-451; its two Linears :415-422) with its inputs
 * as Dynamics._call_vnet builds them (dynamics.py:1142-1160):
 *   z[b, h] = act( W_x su3_to_vec(projectSU(x))[b] + b_x + W_v su3_to_vec(projectSU(F))[b] + b_v ),  h < hidden <= 256
 * as a split-K bf16 GEMM (K = 2 * 8 * nlinks) with fp32 accumulators in TMEM, one CTA per SM.  Both operands are
 * streamed by TMA bulk copies from K-major core-matrix images: the weights from l2b_su3_input_pack, the activations
 * from l2b_su3_project_vec_planar_lm, which writes vec8 "link-major" [link][chain < nb_pad][8] bf16 (one link = one
 * K core; the caller zero-fills the buffer once so that the pad rows nb..nb_pad-1 are zero).
 * activation: 0 identity, 1 tanh, 2 relu, 3 swish, 4 leaky_relu(0.01), 5 elu (network.py:40-46).
 * z_bf16: [nb, hidden].  nlinks % 8 == 0, nb_pad % 16 == 0, nb <= nb_pad <= 256, else L2B_ERR_UNSUPPORTED. */
size_t l2b_su3_input_packed_bytes(int nlinks, int hidden);
size_t l2b_su3_input_ws_bytes(int nb_pad, int hidden);
int l2b_su3_input_pack(const void* w_x, const void* w_v, int w_dtype, void* packed, int nlinks, int hidden,
                       void* stream);
int l2b_su3_project_vec_planar_lm(const void* x_planar, void* vec8_lm, int nb, int nb_pad, const int dims[4], int dtype,
                                  void* stream);
int l2b_su3_input_layer(const void* act_x, const void* act_f, const void* packed, const float* bias_x,
                        const float* bias_v, int activation, void* z_bf16, int nb, int nb_pad, int nlinks, int hidden,
                        void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------ */
/* dense layers: general bf16 GEMM on the tensor cores (tcgen05)              */
/* ------------------------------------------------------------------------ */
/* Every nn.Linear of the networks the fused kernels do not cover -- the hidden Linears (network/pytorch/
 * network.py:489-493, 538-541), the input Linears under autograd (:415-422) -- and the GEMMs of every Linear's
 * backward pass that the reference gets from ATen autograd (dX = dY W, dW = dY^T X; trainer.py:1326-1345 calls
 * loss.backward()):
 *     D[m][n] = act( sum_{seg < nseg} sum_{k < K} A_seg(m, k) B_seg(n, k) + bias[n] )  (+ D when accumulate)
 * bf16 operands, fp32 accumulation in TMEM.  Operands are plain row-major matrices; `*_kmajor` says which axis is
 * contracted: 1 = stored [MN][K] (leading dimension ld >= K), 0 = stored [K][MN] (ld >= MN) -- no transposed copy
 * is needed for any of the three GEMMs of a Linear.  a_ptrs / b_ptrs: HOST arrays of nseg (1..32) device pointers,
 * all segments share the shapes and leading dimensions.  out: [M][ldo] in out_dtype (L2B_BF16 or L2B_F32).
 * splits > 1 cuts the concatenated K axis over that many CTAs per tile (fp32 partials in ws, summed in a fixed
 * order); l2b_gemm_bf16_splits proposes a value that fills the GPU.  seg_inner: order of the (segment, K chunk) loop,
 * 0 = each segment front to back, 1 = all segments of a K chunk before the next chunk.  activation codes as l2b_su3_input_layer.
 * Stored row lengths, leading dimensions and N must be multiples of 8 (16-byte units), pointers 16-byte aligned,
 * else L2B_ERR_UNSUPPORTED / L2B_ERR_INVALID. */
int l2b_gemm_bf16_splits(int M, int N, int K, int nseg, int b_kmajor);
size_t l2b_gemm_bf16_ws_bytes(int M, int N, int splits);
int l2b_gemm_bf16(const void* const* a_ptrs, long long lda, int a_kmajor, const void* const* b_ptrs, long long ldb,
                  int b_kmajor, int nseg, int seg_inner, int M, int N, int K, void* out, int out_dtype, long long ldo,
                  int accumulate, const float* bias, int activation, int splits, void* ws, size_t ws_bytes,
                  void* stream);
/* fp32 nets without autocast (the reference's default precision, configs.py `precision: float32`): an fp32-accurate
 * Linear on the same kernel.  l2b_split_bf16x3 writes x = x1 + x2 + x3 as three bf16 matrices out[3][rows][out_ld]
 * (zero padded columns); the six products a_i b_j with i + j <= 4 are six segments of one l2b_gemm_bf16 launch
 * (seg_inner = 1: the segments of one K chunk run back to back, so the re-read operands hit L2).  Relative error
 * ~ 2^-22, that of an fp32 GEMM. */
int l2b_split_bf16x3(const float* x, long long rows, long long cols, long long ld, void* out, long long out_ld,
                     void* stream);

/* ------------------------------------------------------------------------ */
/* U(1) xnet convolution stack                                                */
/* ------------------------------------------------------------------------ */
/* ConvStack (network/pytorch/network.py:240-346): blocks of PeriodicPadding(n - 1) (:151-172: n - 1 wrapped rows /
 * columns on BOTH sides) + Conv2d(f, n), MaxPool2d after every second block, activation.  A block is
 *     out[(b, oh, ow)][co] = sum_k col[(b, oh, ow)][k] W[co][k] + bias[co],   OH = H + n - 1,  k = (ci, kh, kw),
 *     col[(b, oh, ow)][(ci, kh, kw)] = in[b, ci, (oh + kh - n + 1) mod H, (ow + kw - n + 1) mod W]
 * i.e. l2b_gemm_bf16 on the gathered matrix `col` and Conv2d's own weight viewed as [Cout, Cin n^2]; the output is
 * the next block's input in NHWC.  l2b_conv_im2col writes col[planes][nb OH OW][K8] bf16 (K8 = Cin n^2 rounded up
 * to a multiple of 8, zero padded; planes = 1: bf16 nets, 3: the bf16x3 split of fp32 nets) from an input with
 * element strides (batch, channel, row, column) -- NCHW for the network input, NHWC between blocks.
 * k_order = 0: columns in Conv2d's own order k = (ci, kh, kw); 1: tap-major k = (kh, kw, ci) (the caller passes the
 * weight as [Cout, n, n, Cin]) -- on NHWC activations with Cin % 8 == 0 the gather then moves 32-byte vectors.
 * l2b_conv_col2im is its adjoint (dcol [nb OH OW][ldc] f32 / bf16 -> din f32 with free output strides), a gather
 * with a fixed summation order.  l2b_pool_act = MaxPool2d(pool) (floor, first maximum wins as in ATen) followed by
 * the activation (codes as l2b_su3_input_layer) on NHWC; idx keeps the winning tap, pre the pooled pre-activation
 * (needed for swish only); l2b_pool_act_bwd scatters gy * act' back (gx is zero-filled first). */
int l2b_conv_im2col(const void* in, int in_dtype, int nb, int C, int H, int W, int n, const long long strides[4],
                    void* col, int planes, int k_order, void* stream);
int l2b_conv_col2im(const void* dcol, int dcol_dtype, long long ldc, int nb, int C, int H, int W, int n, float* din,
                    const long long out_strides[4], int k_order, void* stream);
int l2b_pool_act(const void* x, int dtype, int nb, int H, int W, int C, int pool, int activation, void* y,
                 unsigned char* idx, float* pre, void* stream);
int l2b_pool_act_bwd(const float* gy, const void* y, int dtype, const float* pre, const unsigned char* idx, int nb, int H,
                     int W, int C, int pool, int activation, float* gx, void* stream);

/* ------------------------------------------------------------------------ */
/* vnet output heads on the tensor cores (tcgen05), fused with the momentum update */
/* ------------------------------------------------------------------------ */
/* The three heads of the vnet LeapfrogLayer (network/pytorch/network.py:536-548:
 *   s = nw.s e^{c_s} tanh(W_s z + b_s), t = nw.t (W_t z + b_t), q = nw.q e^{c_q} tanh(W_q z + b_q),
 * each [nb, xdim]) and Dynamics._update_v_fwd/_bwd (dynamics.py:1266-1297) that consumes them,
 * as ONE kernel: bf16 tcgen05.mma with fp32 accumulators in TMEM, weights streamed by TMA bulk
 * copies from a pre-packed bf16 image, epilogue straight out of TMEM.  s, t, q never reach HBM. */
/* bytes of the packed image / of the logdet workspace */
size_t l2b_vnet_heads_packed_bytes(int xdim, int hidden);
size_t l2b_vnet_heads_ws_bytes(int nb, int xdim);
/* W_s, W_t, W_q: [xdim, hidden] row-major (nn.Linear.weight) of w_dtype (L2B_F32/F64/BF16) ->
 * bf16 UMMA tile image (also the fp32 -> bf16 cast autocast would do); redo when weights change */
int l2b_vnet_pack_heads(const void* w_s, const void* w_t, const void* w_q, int w_dtype, void* packed, int xdim,
                        int hidden, void* stream);
/* z: bf16 [nb, hidden] (output of the hidden stack); bias_*: f32 [xdim]; scale_s = nw.s e^{c_s},
 * scale_q = nw.q e^{c_q}: f32 [xdim]; scale_t = nw.t; v, force, v_out: complex128 [nb, xdim];
 * logdet[nb] (NULL: skip); stq_or_null: f32 [3, nb, xdim] dump of (s, t, q) for the backward pass
 * and for tests.  hidden % 8 == 0 and hidden <= 256, else L2B_ERR_UNSUPPORTED. */
int l2b_su3_heads_vupdate(const void* z, const void* packed, const float* bias_s, const float* bias_t,
                          const float* bias_q, const float* scale_s, const float* scale_q, float scale_t,
                          const void* v, const void* force, double eps, const double* eps_dev, int sign, void* v_out,
                          double* logdet, float* stq_or_null, int nb, int xdim, int hidden, void* ws, size_t ws_bytes, void* stream);
/* Two consecutive momentum updates between which the links do not move (dynamics.py:1187-1228: the second half of
 * leapfrog layer i and the first half of layer i+1; the two around the turn-around of the forward / backward sweep,
 * dynamics.py:1002, with v -> -v in between: negate_between) see the same (s, t, q) and the same force when the
 * layers share one vnet (use_separate_networks = false, conf/dynamics/su3.yaml).  One pass computes
 *   v'' = upd( [-] upd(v; eps1, sign1); eps2, sign2 ),   logdet = logdet_1 + logdet_2
 * -- v, F and the head weights are read once instead of twice. */
int l2b_su3_heads_vupdate_pair(const void* z, const void* packed, const float* bias_s, const float* bias_t,
                               const float* bias_q, const float* scale_s, const float* scale_q, float scale_t,
                               const void* v, const void* force, double eps1, const double* eps1_dev, int sign1,
                               double eps2, const double* eps2_dev, int sign2, int negate_between, void* v_out,
                               double* logdet, int nb, int xdim, int hidden, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------ */
/* U(1), x[nb, 2, T, X] real angles                                          */
/* ------------------------------------------------------------------------ */
/* scratch bytes for the U(1) entry points that take `ws` (no reference counterpart) */
size_t l2b_u1_ws_bytes(int nb, int T, int X, int dtype);

/* LatticeU1.wilson_loops (lattice/u1/pytorch/lattice.py:154-159): w[nb, T, X] */
int l2b_u1_wilson_loops(const void* x, void* w, int nb, int T, int X, int dtype, void* stream);
/* the reference's 4x4 loop angles (lattice/u1/pytorch/lattice.py:161-186, the sixteen rolled terms in its order;
 * `plaqs4x4` = mean cos of them, :205-219), w [nb, T, X] -- the caller applies upstream's trailing `.T` */
int l2b_u1_wilson_loops4x4(const void* x, void* w, int nb, int T, int X, int dtype, void* stream);
/* LatticeU1._action / plaqs / _sin_charges / _int_charges in one pass
 * (lattice.py:80-86,188-228): obs[nb, 4] = (action, plaq, sinQ, intQ), in `dtype` */
int l2b_u1_observables(const void* x, double beta, void* obs, int nb, int T, int X, int dtype,
                       void* stream);
/* LatticeU1.grad_action (lattice.py:102-117), analytic */
int l2b_u1_force(const void* x, double beta, void* force, int nb, int T, int X, int dtype,
                 void* stream);
/* Dynamics.transition_kernel_hmc + leapfrog_hmc for U(1) (dynamics.py:900-954): the whole trajectory of one chain
 * runs inside one thread block with x, v resident in shared memory.
 * energies[nb, 4] = (KE0, S0, KE1, S1) in `dtype`.                            */
int l2b_u1_hmc_trajectory(const void* x, const void* v, double beta, double eps, int nlf,
                          void* x_prop, void* v_prop, void* energies, int nb, int T, int X,
                          int dtype, void* stream);
/* Dynamics._update_v_{fwd,bwd} epilogue for real fields (dynamics.py:1266-1297) */
int l2b_u1_vupdate(const void* v, const void* force, const void* s, const void* t, const void* q,
                   double eps, const void* eps_dev, int sign, void* v_out, void* logdet, int nb, int xdim, int dtype,
                   void* stream);
/* Dynamics._update_x_{fwd,bwd} for U(1) (dynamics.py:1398-1419,1443-1467),
 * use_ncp selects the non-compact-projection update; m = mask[xdim] float32.
 * x_out is wrapped to [-pi, pi) like g.compat_proj.                           */
int l2b_u1_xupdate(const void* x, const void* v, const void* s, const void* t, const void* q,
                   const float* mask, double eps, const void* eps_dev, int sign, int use_ncp, void* x_out, void* logdet,
                   int nb, int xdim, int dtype, void* stream);

/* U1Phase.kinetic_energy (group/u1/pytorch/group.py:164-165): ke[nb] = 0.5 sum v^2 */
int l2b_u1_kinetic(const void* v, void* ke, int nb, int xdim, int dtype, void* stream);
/* U1Phase.compat_proj (group.py:130-131): ((x + pi) mod 2 pi) - pi, n elements */
int l2b_u1_compat_proj(const void* x, void* out, size_t n, int dtype, void* stream);

/* --- U(1) dense networks fused with the updates (inference; network/pytorch/network.py:349-551) --- */
/* Input layer of a U(1) LeapfrogLayer WITHOUT conv stack (network.py:349-451), xlayer and vlayer in one
 * pass over the two fields: pre[nb, units] = W_x . f(x) + b_x + W_v . v + b_v, with mode 1 (xnet)
 * f(x) = cat(cos(mask * x), sin(mask * x)) (dynamics.py:1169-1178; w_x is [units, 2 xdim]) and mode 0
 * (vnet) f(x) = x (w_x is [units, xdim]; pass the force as v).  The caller applies the activation.
 * units <= 16, else L2B_ERR_UNSUPPORTED. */
size_t l2b_u1_input_ws_bytes(int nb, int xdim);
int l2b_u1_input_layer(int mode, const void* x, const void* v, const float* mask, const void* w_x, const void* b_x,
                       const void* w_v, const void* b_v, int units, void* pre, int nb, int xdim, int dtype, void* ws,
                       size_t ws_bytes, void* stream);
/* The three output heads of a U(1) LeapfrogLayer (network/pytorch/network.py:536-548) fused with the
 * update that consumes them: mode 0 = Dynamics._update_v_fwd/_bwd (dynamics.py:1266-1297) on (a = v,
 * b = force), mode 1 = _update_x_fwd/_bwd (dynamics.py:1398-1467) on (a = x, b = v, mask).  z: [nb, hidden]
 * output of the hidden stack; w_*: [xdim, hidden] nn.Linear weights, b_*: [xdim], coeff_*: [xdim]
 * (ScaledTanh.coeff), nw_* the NetWeight factors; all of `dtype`.  hidden <= 32 (CUDA-core kernel: the
 * U(1) nets are 16 wide), else L2B_ERR_UNSUPPORTED.  s, t, q never reach HBM. */
size_t l2b_u1_heads_ws_bytes(int nb, int xdim);
int l2b_u1_heads_update(int mode, const void* z, int hidden, const void* w_s, const void* w_t, const void* w_q,
                        const void* b_s, const void* b_t, const void* b_q, const void* coeff_s, const void* coeff_q,
                        double nw_s, double nw_t, double nw_q, const void* a, const void* b, const float* mask,
                        double eps, const void* eps_dev, int sign, int use_ncp, void* out, void* logdet, int nb,
                        int xdim, int dtype, void* ws, size_t ws_bytes, void* stream);

/* --- adjoints (L2HMC training; the reference relies on autograd for these, trainers/pytorch/trainer.py:1284-1314) --- */
/* adjoint of l2b_u1_wilson_loops (lattice/u1/pytorch/lattice.py:154-159): gx[nb,2,T,X] from gw[nb,T,X] */
int l2b_u1_wilson_loops_bwd(const void* gw, void* gx, int nb, int T, int X, int dtype, void* stream);
/* adjoint of l2b_u1_force = Hessian-vector product of the action (the reference
 * differentiates through grad_action with create_graph=True, lattice.py:106,113-116) */
int l2b_u1_force_bwd(const void* x, double beta, const void* gforce, void* gx, int nb, int T, int X, int dtype,
                     void* stream);
/* adjoints of l2b_u1_vupdate / l2b_u1_xupdate (dynamics.py:1266-1297,1398-1467): gradients w.r.t. every input
 * (gs/gt/gq may be NULL) and geps[nb] = per-chain derivative w.r.t. eps */
int l2b_u1_vupdate_bwd(const void* v, const void* force, const void* s, const void* t, const void* q, double eps, const void* eps_dev,
                       int sign, const void* gv_out, const void* glogdet, void* gv, void* gforce, void* gs, void* gt,
                       void* gq, void* geps, int nb, int xdim, int dtype, void* stream);
int l2b_u1_xupdate_bwd(const void* x, const void* v, const void* s, const void* t, const void* q, const float* mask,
                       double eps, const void* eps_dev, int sign, int use_ncp, const void* gx_out, const void* glogdet, void* gx, void* gv,
                       void* gs, void* gt, void* gq, void* geps, int nb, int xdim, int dtype, void* stream);
/* out[b,:] = scale[b] * in[b,:]: adjoint of LatticeU1._action (lattice.py:80-86; in = force) and of
 * U1Phase.kinetic_energy (group.py:164-165; in = v) */
int l2b_rowscale(const void* in, const void* scale, void* out, int nb, int xdim, int dtype, void* stream);

/* ------------------------------------------------------------------------ */
/* Metropolis-Hastings accept / reject mix (dynamics.py:632-702,1065-1087)   */
/* ------------------------------------------------------------------------ */
/* out[b, :] = accept[b] ? prop[b, :] : init[b, :] for `nfields` field pairs of
 * `row_bytes[k]` bytes per chain; accept[nb] float32 0/1 = (acc > u).        */
int l2b_accept_mix(const void* const* host_init, const void* const* host_prop, void* const* host_out,
                   const size_t* host_row_bytes, int nfields, const float* accept, int nb,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* L2B_H */
