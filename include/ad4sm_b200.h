/* ad4sm_b200.h -- C ABI of libad4sm_b200.so: the B200-native `makeϕrKt` hot path of AD4SM.jl.
 *
 * The reference (AD4SM.jl v0.0.7, pure Julia) has no FFI of its own; the boundary it offers is
 * multiple dispatch on `Elements.makeϕrKt` / `Elements.makeϕrKt_d`.  Each entry point below names
 * the reference method(s) it replaces; `ad4sm.jl_b200/julia/AD4SMB200.jl` re-defines those methods
 * with `ccall` into this library.
 *
 * Conventions
 *   - plain C types only; all arrays are caller-owned host memory unless a name ends in `_dev`.
 *   - node ids and CSC indices are 1-based Int64, exactly as Julia holds them.
 *   - `u` is the reference's `u::Matrix{Float64}` (dofs_per_node x n_nodes, column-major);
 *     DOF id = (node-1)*dofs_per_node + comp   (elements.toolkit.jl:183,187).
 *   - every function returns 0 on success and a negative AD4SM_E* code otherwise;
 *     `ad4sm_last_error()` returns a thread-local message.  Nothing throws across the ABI.
 *   - one in-flight call per plan.  There is NO CPU fallback: without a CUDA device every
 *     compute entry point fails with AD4SM_ENODEV.
 */
#ifndef AD4SM_B200_H
#define AD4SM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AD4SM_ABI_VERSION 2

/* status codes */
#define AD4SM_OK 0
#define AD4SM_EINVAL (-1)  /* bad argument / unsupported element-material combination */
#define AD4SM_ENODEV (-2)  /* no CUDA device / CUDA runtime error */
#define AD4SM_ENOMEM (-3)
#define AD4SM_ESTATE (-4)  /* call order (e.g. fetch before eval) */
#define AD4SM_ENCCL (-5)

/* element kinds: CEElem{D,P,M,T,N} (elements.jl:71-77), CPElem (elements.jl:104-111) */
#define AD4SM_ELEM_CE 0
#define AD4SM_ELEM_CP 1
/* CAS: the axisymmetric continuum element north_star names.  The reference only ships its constructors (ASTria / ASQuad,
 * elasticelements.jl:235-284: N, Nx, Ny, the radius X0 and wgt = detJ*2π*X0 at every Gauss point); the struct `CAS` they
 * call and its getϕ do not exist, so the energy is this library's EXTENSION (parity unpinned): a 3x3
 * deformation gradient with the in-plane block I + ∇u (rows r, z) and the hoop stretch F33 = 1 + (N·u_r)/X0, fed to the
 * 3-D materials (Hooke, NeoHooke, MooneyRivlin, Ogden), Φ_e = Σ_gp wgt ϕ(F).  dim = 2, dofs (u_r, u_z) per node; gradN holds
 * (Nx, Ny) and `Nval` holds the HOOP ROW  N[gp][a] / X0[gp]. */
#define AD4SM_ELEM_CAS 2

/* material kinds (elasticmaterials.jl:25-120, phasefieldmaterials.jl:5-16) */
#define AD4SM_MAT_HOOKE 1        /* params: E, nu                flags: SMALL */
#define AD4SM_MAT_HOOKE2D 2      /* params: E, nu                flags: SMALL, PLANE_STRESS */
#define AD4SM_MAT_NEOHOOKE 3     /* params: C1, K   (K<0: incompressible / plane-stress form) */
#define AD4SM_MAT_MOONEYRIVLIN 4 /* params: C1, C2, K */
#define AD4SM_MAT_OGDEN 5        /* params: alpha, mu, K */
#define AD4SM_MAT_PHASEFIELD 6   /* params[0..3]: base material, params[4..6]: l0, Gc, n; mat_base = base kind */

/* material flags */
#define AD4SM_F_SMALL 1u              /* `small=true`: E = sym(F)-I instead of Green-Lagrange */
#define AD4SM_F_PLANE_STRESS 2u       /* Hooke2D{T,:plane_stress} */
#define AD4SM_F_DHN 4u                /* PhaseField{M,:DHn} (default :ATn) */
#define AD4SM_F_PF_GRAD_INTENDED 8u   /* d-phase: keep d/dd of l0^2|∇d|^2 (reference drops it, toolkit:159-166) */
#define AD4SM_F_NH2D_EXT 16u          /* PhaseField{NeoHooke} on 2x2 F: plane-strain embedding (extension) */

/* batch flags.  HALO: elements owned by another rank of a mesh-partitioned run.  They take part in the pattern and in
 * the assembly of this rank's rows, but are not evaluated here: their staged element blocks / residuals are written
 * into the plan's staging arrays by the exchange step (ad4sm_staging_view_get) between ad4sm_eval_elements_device and
 * ad4sm_eval_assemble_device.  gradN / wgt / Nval may be NULL for a HALO batch. */
#define AD4SM_BATCH_HALO 1

/* fields a plan can assemble for */
#define AD4SM_FIELD_U 0 /* displacement dofs: dofs_per_node per node */
#define AD4SM_FIELD_D 1 /* phase-field dofs: 1 per node (makeϕrKt_d) */

/* Element shapes of the batched constructors (ad4sm_construct, ad4sm_batch.mesh): the reference constructor each one
 * replaces for a whole group of elements at once. */
#define AD4SM_SHAPE_TRIA3 1 /* Tria, TriaP               elasticelements.jl:35-46,   phasefieldelements.jl:9-24    D=2 N=3 P=1 */
#define AD4SM_SHAPE_QUAD4 2 /* Quad, QuadR, QuadP[R]     elasticelements.jl:50-80,   phasefieldelements.jl:25-55   D=2 N=4 P=n^2 */
#define AD4SM_SHAPE_TET4 3  /* Tet04, Tet04P             elasticelements.jl:81-96,   phasefieldelements.jl:129-145 D=3 N=4 P=1 */
#define AD4SM_SHAPE_HEX8 4  /* Hex08, Hex08R, Hex08P[R]  elasticelements.jl:129-162, phasefieldelements.jl:56-94   D=3 N=8 P=n^3 */
#define AD4SM_SHAPE_WDG6 5  /* Wdg06, Wdg06P             elasticelements.jl:163-197, phasefieldelements.jl:95-128  D=3 N=6 P=6 */

/* Node coordinates + integration rule: what the reference constructors take as `p0` and `GP=`. */
typedef struct ad4sm_mesh {
  int32_t shape;        /* AD4SM_SHAPE_* */
  int32_t n_gp_1d;      /* QUAD4 / HEX8: points per direction of the tensor rule (1..3), the length of `GP=`; else ignored */
  double gp_xi[3];      /* the rule's abscissae and weights, GP = ((xi_1, w_1), ...) */
  double gp_w[3];
  int32_t coord_dim;    /* doubles per node in `coords` (>= D; extra components are ignored) */
  int32_t reserved;
  int64_t n_nodes;
  const double* coords; /* [n_nodes][coord_dim] host */
} ad4sm_mesh;

/* One homogeneous group of elements: same element type parameters (kind, D, P, N) and the same
 * material.  This is the SoA packing of a `Vector{CEElem}` / `Vector{CPElem}`
 * (elements.jl:71-77,104-111); array layouts follow the reference's nesting order. */
typedef struct ad4sm_batch {
  int32_t elem_kind;      /* AD4SM_ELEM_* */
  int32_t dim;            /* D: 2 or 3 */
  int32_t n_gauss;        /* P */
  int32_t n_elem_nodes;   /* N */
  int32_t mat_kind;       /* AD4SM_MAT_* */
  int32_t mat_base;       /* base material kind when mat_kind == PHASEFIELD, else 0 */
  uint32_t mat_flags;     /* AD4SM_F_* */
  int32_t batch_flags;    /* AD4SM_BATCH_* (0 for a plain single-GPU plan) */
  double mat_params[8];
  int64_t n_elems;
  const int64_t* nodes;      /* [n_elems][N]        elem.nodes, 1-based */
  const double* gradN;       /* [n_elems][D][P][N]  elem.∇N[d][gp][a] */
  const double* wgt;         /* [n_elems][P]        elem.wgt[gp] */
  const double* Nval;        /* [n_elems][P][N]     elem.N[gp][a]   (CPElem; NULL for CEElem; CAS: elem.N[gp][a] / elem.X0[gp]) */
  const int64_t* orig_index; /* [n_elems] 0-based position of each element in `elems` (fixes the
                                summation order); NULL = the element's slot, i.e. its position in the
                                concatenation of ALL batches in the order given (not "after the largest
                                key so far": give every batch an orig_index when any batch has one) */
  const ad4sm_mesh* mesh;    /* NULL, or: construct gradN / wgt / Nval on the device from the node coordinates (the three
                                pointers above are then ignored); dim, n_gauss, n_elem_nodes must match mesh->shape */
} ad4sm_batch;

typedef struct ad4sm_plan ad4sm_plan; /* opaque: device copies of the batches + CSC pattern + scatter maps */

/* Pointers into plan-owned DEVICE memory holding the result of the last evaluation of `field`
 * (valid until the next evaluation / plan destruction).  The pattern is the structural one;
 * `n_zero` tells whether `dropzeros` would remove anything. */
typedef struct ad4sm_device_view {
  int64_t n_dof, nnz;
  const double* phi_dev;    /* [1] */
  const double* r_dev;      /* [n_dof] */
  const double* nzval_dev;  /* [nnz]  structural CSC(=CSR, Kt is symmetric) values */
  const int64_t* colptr_dev; /* [n_dof+1] 1-based */
  const int64_t* rowval_dev; /* [nnz]     1-based */
  const int64_t* n_zero_dev; /* [1] number of stored exact zeros in nzval */
} ad4sm_device_view;

int ad4sm_abi_version(void);
const char* ad4sm_last_error(void);
int ad4sm_device_count(int* count);

/* Build a plan for `elems` (replaces the per-call boxing/allocation of _assemble_ϕrKt,
 * elasticelements.jl:208-221, and the triplet/sparse machinery of elements.toolkit.jl:169-203).
 * All batches must have the same N (the reference requires equal dofs/elem, elasticelements.jl:212-214).
 * dofs_per_node = size(u,1) (>= D; extra rows of u are ignored by getF, toolkit:13, and yield
 * structural zeros that dropzeros removes). `device` = CUDA ordinal, -1 = current. */
int ad4sm_plan_create(const ad4sm_batch* batches, int32_t n_batches, int64_t n_nodes, int32_t dofs_per_node,
                      int32_t device, ad4sm_plan** plan_out);
void ad4sm_plan_destroy(ad4sm_plan* plan);

/* The reference's element constructors for a whole group at once, on the device: writes what
 * `Hex08(nodes, p0; GP)` etc. store per element -- ∇N as [n_elems][D][P][N], wgt [n_elems][P], N values
 * [n_elems][P][N] (phase-field constructors; NULL to skip) and elem.V [n_elems] (NULL to skip) -- into host arrays.
 * `nodes` [n_elems][N] 1-based.  ad4sm_construct_device writes device arrays instead (same layouts) and returns after the
 * kernels have finished. */
int ad4sm_construct(const ad4sm_mesh* mesh, int64_t n_elems, const int64_t* nodes, int32_t device, double* gradN, double* wgt,
                    double* Nval, double* V);
int ad4sm_construct_device(const ad4sm_mesh* mesh, int64_t n_elems, const int64_t* nodes, int32_t device, double* gradN_dev,
                           double* wgt_dev, double* Nval_dev, double* V_dev);

/* Structural CSC pattern of Kt (before dropzeros) for a field: sizes, then the arrays. */
int ad4sm_plan_pattern_size(ad4sm_plan* plan, int32_t field, int64_t* n_dof, int64_t* nnz);
int ad4sm_plan_pattern(ad4sm_plan* plan, int32_t field, int64_t* colptr /*[n_dof+1]*/, int64_t* rowval /*[nnz]*/);

/* makeϕrKt(elems, u)  -- elasticelements.jl:356-358.  Writes ϕ and r (length(u)); leaves Kt on the
 * device and reports nnz AFTER dropzeros so the caller can size the arrays for ad4sm_fetch_csc. */
int ad4sm_eval(ad4sm_plan* plan, const double* u, double* phi, double* r, int64_t* nnz_out);

/* makeϕr: ϕ and r only -- elements.toolkit.jl:204-215 applied to the element loop of :216-218. */
int ad4sm_eval_phi_r(ad4sm_plan* plan, const double* u, double* phi, double* r);

/* makeϕrKt(elems::Vector{<:CPElem}, u, d) -- phasefieldelements.jl:228-238 (d indexed by node). */
int ad4sm_eval_pf_u(ad4sm_plan* plan, const double* u, const double* d, double* phi, double* r, int64_t* nnz_out);

/* makeϕrKt_d(elems, u, d) / makeϕrKt_d(elems, u, d, ϕmax) -- phasefieldelements.jl:239-262.
 * `hist`: NULL, or one pointer per batch to [n_elems][P][2] doubles (ϕmax[e][gp] tuples), updated
 * in place like the reference mutates ϕmax (:180).  r has n_nodes entries. */
int ad4sm_eval_pf_d(ad4sm_plan* plan, const double* u, const double* d, double* const* hist, double* phi, double* r,
                    int64_t* nnz_out);

/* Assembly only: makeϕrKt(Φ::Vector{<:adiff.D2}, elems, u) -- elements.toolkit.jl:169-203, the method
 * Solvers.solvestep! calls after building Φ itself (solvers.jl:213-216, 257-260).  `Phi`: ne records of `stride`
 * doubles, record e = element e of `elems` = (v | g[n] | h[n(n+1)/2]): exactly the memory of a Julia
 * Vector{D2{n,m,Float64}} (adiff.jl:29-42; stride = 1+n+m), n = N*size(u,1) local dofs ordered comp + size(u,1)*node
 * (toolkit:183,187), h = packed upper triangle, entry (i<=j), 1-based, at (j-1)j/2+i (adiff.jl:26,84).  The library
 * does what toolkit:182-202 does: ϕ = Σ v, r[idx] += g, Kt = dropzeros(sparse(...)) with duplicates summed in `elems`
 * order.  Needs size(u,1) == element dimension and no HALO batch.  Results as ad4sm_eval (+ ad4sm_fetch_csc). */
int ad4sm_assemble_duals(ad4sm_plan* plan, const void* Phi, int64_t stride, double* phi, double* r, int64_t* nnz_out);
/* makeϕr(Φ, elems, u) -- elements.toolkit.jl:204-215: ϕ and r only; records (v | g[n] | anything), i.e. a
 * Vector{D1{n,Float64}} (stride 1+n) or Vector{D2{n,m,Float64}} (stride 1+n+m). */
int ad4sm_assemble_duals_phi_r(ad4sm_plan* plan, const void* Phi, int64_t stride, double* phi, double* r);

/* SparseMatrixCSC arrays of the last evaluation, after dropzeros (toolkit:202).  Any pointer may be NULL. */
int ad4sm_fetch_csc(ad4sm_plan* plan, int64_t* colptr, int64_t* rowval, double* nzval);

/* ---- Newton-step glue on the device (Solvers.solvestep!, nEqs == 0 branch; solvers.jl:183-184, 211-224, 257-291) ----
 * The reference slices the assembled Kt on the host after every makeϕrKt: Kt[ifreeu,ifreeu] for lu(), Kt[ifreeu,icnstu]*
 * Δucnst for the predictor, fi[ifreeu] - fe[ifreeu] and two norms.  These entry points do that where Kt lives, so that
 * only res, the norms and the VALUES of Kt[ifree,ifree] cross PCIe (its pattern is built once per plan).
 *   setup   : `bfree` = bfreeu[:] (one byte per dof of the field, column-major order of u; non-zero = free).  Builds
 *             ifree = findall(bfree), icnst = findall(!bfree) and the CSC pattern of Kt[ifree,ifree] (free numbering,
 *             1-based, STRUCTURAL: exact zeros are kept -- nnz_out of the evaluation tells whether there are any).
 *   pattern : colptr_ff [nfree+1], rowval_ff [nnz_ff].
 *   reduce  : after an evaluation of `field`.  kind 0 (predictor, :220-221): res = fi[ifree] - fe[ifree] -
 *             Kt[ifree,icnst]*ducnst;  kind 1 (corrector, :262): res = fe[ifree] - fi[ifree].  fe NULL = zeros(n_dof)
 *             (the default of solvestep!); ducnst [ncnst] is read for kind 0 only (NULL = zeros).  Outputs (any may be
 *             NULL): res [nfree], nzval_ff [nnz_ff], norms[2] = { norm(res), norm(fi[icnst]) } (:263-264). */
int ad4sm_newton_setup(ad4sm_plan* plan, int32_t field, const uint8_t* bfree, int64_t* nfree, int64_t* ncnst, int64_t* nnz_ff);
int ad4sm_newton_pattern(ad4sm_plan* plan, int32_t field, int64_t* colptr_ff, int64_t* rowval_ff);
int ad4sm_newton_reduce(ad4sm_plan* plan, int32_t field, int32_t kind, const double* fe, const double* ducnst, double* res,
                        double* nzval_ff, double* norms);
/* the same on device-resident data (asynchronous on the plan's stream); results through ad4sm_newton_view_get */
typedef struct ad4sm_newton_view {
  int64_t nfree, ncnst, nnz_ff;
  const double* res_dev;         /* [nfree] */
  const double* nzval_ff_dev;    /* [nnz_ff] */
  const double* norms_dev;       /* [2] */
  const int64_t* colptr_ff_dev;  /* [nfree+1] 1-based */
  const int64_t* rowval_ff_dev;  /* [nnz_ff]  1-based */
} ad4sm_newton_view;
int ad4sm_newton_reduce_device(ad4sm_plan* plan, int32_t field, int32_t kind, const double* fe_dev, const double* ducnst_dev);
int ad4sm_newton_view_get(ad4sm_plan* plan, int32_t field, ad4sm_newton_view* view);

/* ---- constraint equations from a fixed catalogue (Solvers.makeϕrKt(eqns, u, λ), solvers.jl:17-22, 24-73) ----
 * The reference's ConstEq holds an arbitrary closure `func`, evaluated with D2 duals: veqs[ii] = val, reqs[iDoFs,ii] =
 * grad, Keqs[iDoFs,iDoFs] += λ[ii] hess.  A closure cannot cross a C ABI; the catalogue offered instead is every
 * constraint that is a polynomial of degree <= 2 in its dofs (periodic boundary conditions, ties, prescribed averages,
 * rigid links / distances):   ϕ(v) = 1/2 v'Qv + c'v - c0,  v = u[iDoFs].
 * Equation ii owns dofs[dof_ptr[ii] .. dof_ptr[ii+1]) (1-based linear indices into u[:], the reference's iDoFs), the
 * coefficients c at the same positions, and optionally a symmetric n x n matrix Q (column-major) at Q[q_ptr[ii] ..
 * q_ptr[ii+1]) -- an empty range, or q_ptr == NULL, makes the equation linear.  Outputs keep that layout:
 *   veqs [n_eqs];  grad aligned with dofs (= what is assigned to reqs[iDoFs, ii]);  hess aligned with Q (= λ[ii] hess(ϕ),
 *   the increment of Keqs[iDoFs, iDoFs]; NULL to skip).  All arrays host memory. */
typedef struct ad4sm_consteq {
  int64_t n_eqs;
  const int64_t* dof_ptr; /* [n_eqs+1] */
  const int64_t* dofs;
  const double* c;
  const double* c0;       /* [n_eqs] */
  const int64_t* q_ptr;   /* [n_eqs+1] or NULL */
  const double* Q;
} ad4sm_consteq;
int ad4sm_consteq_eval(const ad4sm_consteq* eqns, int64_t n_dof, const double* u, const double* lambda, int32_t device,
                       double* veqs, double* grad, double* hess);

/* ---- batched post-processing with first derivatives (elements.toolkit.jl:8-132), one value per element in `elems`
 * order; any output may be NULL.  D = element dimension, matrices column-major like Julia's SMatrix{D,D}:
 *   F     [ne][D*D]  getF(elem, u[:,elem.nodes]) = I + Σ_gp wgt ∇u_gp / V          (toolkit:8-27)
 *   detJ  [ne]       detJ(elem, u) = Σ_gp wgt det F_gp / V                           (:33-47)
 *   vol   [ne]       getV(elem, u) = Σ_gp wgt det F_gp; getV(elems,u) is their sum   (:52-59)
 *   sigma [ne][D*D]  getσ(elem, u): volume-averaged Cauchy stress Σ_gp wgt (1/J P Fᵀ)_gp / V   (:63-87)
 *   Pk    [ne][D*D]  getP(elem, u): volume average of P1K' (the transposed first Piola-Kirchhoff stress, :117-132)
 * Elastic materials only (the reference has no getϕ(F, ::PhaseField) without d).  `u` as for ad4sm_eval. */
int ad4sm_post(ad4sm_plan* plan, const double* u, double* F, double* detJ, double* vol, double* sigma, double* Pk);
/* the same with u and the outputs in device memory (asynchronous on the plan's stream) */
int ad4sm_post_device(ad4sm_plan* plan, const double* u_dev, double* F_dev, double* detJ_dev, double* vol_dev, double* sigma_dev,
                      double* Pk_dev);

/* ---- device-resident path (inputs/outputs stay in HBM; used by GPU-side callers and bench.py) ---- */
int ad4sm_set_stream(ad4sm_plan* plan, void* cuda_stream); /* cudaStream_t; NULL = plan's own stream */
/* mode: 0 = elastic (ad4sm_eval), 1 = PF u-phase, 2 = PF d-phase, 3 = elastic ϕ,r only.
 * u_dev [dpn*n_nodes], d_dev [n_nodes] or NULL, hist_dev: NULL or concatenated [sum n_elems*P*2].
 * Asynchronous on the plan's stream. */
int ad4sm_eval_device(ad4sm_plan* plan, int32_t mode, const double* u_dev, const double* d_dev, double* hist_dev);
int ad4sm_device_view_get(ad4sm_plan* plan, int32_t field, ad4sm_device_view* view);

/* ---- mesh-partitioned runs: the two halves of ad4sm_eval_device, and the staging they communicate through ----
 * The reference has no multi-device path; this is the seam where interface contributions cross ranks (DESIGN.md
 * "Multi-GPU").  elements: K1 over the non-HALO batches (element energy/gradient/Hessian -> staging).
 * assemble: K2/K3 over ALL batches' staging + the ϕ reduction (HALO slots contribute 0 to ϕ). */
int ad4sm_eval_elements_device(ad4sm_plan* plan, int32_t mode, const double* u_dev, const double* d_dev, double* hist_dev);
int ad4sm_eval_assemble_device(ad4sm_plan* plan, int32_t mode);
/* Optional first half of the assembly stage of a partitioned run: assembles the INTERIOR -- the node rows of r and the
 * node pairs of Kt that receive no contribution from a HALO slot (for contiguous partitions: everything but the
 * interface planes) -- and leaves the evaluation pending; it may therefore run while the interface rows are still in
 * flight.  The following ad4sm_eval_assemble_device then assembles only the remainder.  Results are bitwise those of the
 * one-call assembly.  (ϕ, which needs no exchange either, is reduced at the end of the element stage.)  On plans without
 * destination-sorted staging or without HALO batch it does nothing. */
int ad4sm_eval_assemble_split_device(ad4sm_plan* plan, int32_t mode);
/* Optional, after ad4sm_eval_assemble_split_device and once the interface rows have arrived: moves the HALO rows into the
 * sorted staging and assembles the interface rows of r on `cuda_stream` (cudaStream_t; the stream the exchange ran on, NULL =
 * the plan's copy stream), so that this work overlaps the interior assembly still running on the plan's stream.  The caller
 * orders `cuda_stream` behind the exchange and the plan's stream behind `cuda_stream` before ad4sm_eval_assemble_device,
 * which then only assembles the interface pairs of Kt.  A no-op when nothing was split off. */
int ad4sm_eval_halo_prepare_device(ad4sm_plan* plan, int32_t mode, void* cuda_stream);
/* The same two stages for callers whose u / r live in host memory (what makeϕrKt's arguments are, elasticelements.jl:
 * 356-358): elements_host uploads u (d, hist) and runs the element stage -- on large single-batch plans the upload goes
 * in node chunks and K1 starts on the element slice whose nodes have arrived; fetch_phi_r_async enqueues the copy of
 * this rank's ϕ and r to the host as soon as they are final, while the Kt assembly is still running.  Both are
 * asynchronous: the host buffers (pinned, to be truly asynchronous) must stay untouched until ad4sm_sync. */
int ad4sm_eval_elements_host(ad4sm_plan* plan, int32_t mode, const double* u, const double* d, double* const* hist);
int ad4sm_fetch_phi_r_async(ad4sm_plan* plan, int32_t field, double* phi, double* r);
/* Peer staging (fused compute + communication): instead of a send/recv after the element stage, K1's bulk-store
 * epilogue writes the staged blocks / residuals of interface elements ALSO into the owner rank's HALO slots, through
 * peer-mapped memory over NVLink (CUDA IPC handles), while it computes.  The ranks then only need a barrier between
 * the element and assembly stages.  Links are contiguous slot ranges (z-slabs and other contiguous partitions). */
#define AD4SM_PEER_HANDLE_BYTES 128 /* two cudaIpcMemHandle_t: Ke and re staging */
typedef struct ad4sm_peer_link {
  int64_t first_slot;         /* first local (own) staging slot of the range */
  int64_t count;              /* slots in the range */
  int64_t remote_first_slot;  /* where the range starts in the peer plan's staging (its HALO slots) */
  const void* peer_handles;   /* AD4SM_PEER_HANDLE_BYTES bytes exported by the peer's ad4sm_peer_export */
  int64_t remote_first_slot_b; /* AD4SM_OPT_PEER_PUSH: where the range starts in the peer's SECOND receive buffer, which
                                * lies behind its staging: peer n_slots + (remote_first_slot - peer's first HALO slot);
                                * ignored otherwise */
} ad4sm_peer_link;
#define AD4SM_MAX_PEER_LINKS 4
/* AD4SM_OPT_PEER_PUSH (set before ad4sm_peer_attach; sorted-staging plans only): K1 keeps the destination-sorted fast
 * path, evaluates the linked ranges FIRST (element-major copies in local memory) and the copy engine then pushes them
 * into the peer's HALO rows over NVLink while the rest of K1 runs -- no SM is spent on communication and no rank waits
 * for its neighbour's K1.  The HALO rows are double-buffered (a push for evaluation n+1 can never overwrite rows a slow
 * peer still reads for evaluation n), alternating with every ad4sm_eval_elements_* / ad4sm_eval_assemble_device pair,
 * so all ranks must issue those calls in lockstep, separated by a barrier between the two stages. */
int ad4sm_peer_export(ad4sm_plan* plan, void* handles_out /* AD4SM_PEER_HANDLE_BYTES */);
int ad4sm_peer_attach(ad4sm_plan* plan, int32_t n_links, const ad4sm_peer_link* links);
/* Without peer memory: the slot ranges whose element-major staging (Ke_dev / re_dev of ad4sm_staging_view_get) the
 * caller will read between the stages (the rows it sends).  Declaring them -- even zero of them -- lets the element
 * stage use the destination-sorted staging for everything else; undeclared, the whole staging is element-major. */
int ad4sm_export_ranges(ad4sm_plan* plan, int32_t n_ranges, const int64_t* first_slot, const int64_t* count);

/* ---- the partitioned evaluation as ONE call, NCCL inside the library (ad4sm_dist_init(nranks, rank, id)) ----
 * For hosts without torch.distributed (Julia with MPI.jl / Distributed, C): one process per GPU; rank 0 calls
 * ad4sm_dist_unique_id and ships the 128 bytes to the others by whatever means the host has; every rank calls
 * ad4sm_dist_init (ncclCommInitRank) and builds its plan as (own elements, HALO batch of the foreign elements touching its
 * owned nodes, sorted by (owner rank, position in `elems`)).  ad4sm_dist_attach names, per neighbour, the OWN slots whose
 * staged rows that neighbour needs (any subset in any order -- a general partition, not only slabs; non-contiguous lists
 * are packed by a gather kernel) and the range of this rank's HALO slots that neighbour fills.  ad4sm_dist_eval* then runs
 * element stage -> grouped ncclSend/ncclRecv of the staged rows on the plan's stream -> assembly -> ncclAllGather of the
 * per-rank ϕ, summed in rank order (run-to-run reproducible).  Owned rows of Kt and r are bitwise those of a single-GPU
 * evaluation of the whole mesh.  libnccl.so.2 is bound at run time (dlopen; AD4SM_NCCL_LIB overrides the name).
 * The exchange and the ϕ all-gather run on a side stream while the interior of the assembly (ad4sm_eval_assemble_split_device)
 * proceeds.  When every rank's send lists are contiguous slot ranges (z-slabs etc.) and CUDA IPC works between the ranks,
 * ad4sm_dist_attach -- a collective call -- sets up the copy-engine push transport of ad4sm_peer_attach by itself
 * (IPC handles and receive ranges all-gathered over NCCL) and ncclSend/ncclRecv is not used; AD4SM_DIST_PUSH=0 disables it. */
#define AD4SM_NCCL_ID_BYTES 128
typedef struct ad4sm_dist ad4sm_dist;
typedef struct ad4sm_dist_peer {
  int32_t rank;               /* the neighbour */
  int64_t n_send;             /* own elements that are HALO on that neighbour ... */
  const int64_t* send_slots;  /* ... their slots in this plan (0-based), in the neighbour's HALO order */
  int64_t recv_first_slot;    /* first staging slot (>= number of own elements) of the HALO elements it evaluates for us */
  int64_t n_recv;
} ad4sm_dist_peer;
int ad4sm_dist_unique_id(void* id_out /* AD4SM_NCCL_ID_BYTES */);
int ad4sm_dist_init(int32_t nranks, int32_t rank, const void* nccl_id, int32_t device, ad4sm_dist** comm_out);
void ad4sm_dist_destroy(ad4sm_dist* comm);
int ad4sm_dist_attach(ad4sm_plan* plan, ad4sm_dist* comm, int32_t n_peers, const ad4sm_dist_peer* peers);
/* device-resident (asynchronous on the plan's stream; the global ϕ is phi_dev of ad4sm_device_view_get, r and Kt the
 * plan's outputs, complete on the rows this rank owns) and host variants (u in; global ϕ and this rank's r out) */
int ad4sm_dist_eval_device(ad4sm_plan* plan, int32_t mode, const double* u_dev, const double* d_dev, double* hist_dev);
int ad4sm_dist_eval(ad4sm_plan* plan, int32_t mode, const double* u, const double* d, double* phi_global, double* r_local);

typedef struct ad4sm_staging_view {
  int64_t n_slots;       /* elements of all batches, in batch order (slot = position in that concatenation) */
  int64_t ke_width;      /* doubles per slot in Ke_dev: N(N+1)/2 node-pair blocks of DB*DB (DB = D for FIELD_U, 1 for FIELD_D) */
  int64_t re_width;      /* doubles per slot in re_dev: N*DB */
  double* Ke_dev;        /* [n_slots][ke_width] */
  double* re_dev;        /* [n_slots][re_width] */
  double* phie_dev;      /* [n_slots] */
} ad4sm_staging_view;
int ad4sm_staging_view_get(ad4sm_plan* plan, int32_t field, ad4sm_staging_view* view);
int ad4sm_sync(ad4sm_plan* plan);

/* Plan options / introspection.  (Options 1, 2 and infos 1, 2 belonged to the round-1 fused K1+assembly kernel, which
 * was measured slower than the separate kernels in every configuration and has been removed; the ids stay reserved.) */
#define AD4SM_OPT_SORTED 3        /* destination-sorted staging + streaming assembly (K2s); default on where available */
#define AD4SM_OPT_PEER_PUSH 4    /* 1: ad4sm_peer_attach sets up copy-engine pushes instead of K1 peer stores (see there) */
#define AD4SM_INFO_SORTED 3       /* 0 = not available for this plan, 1 = available but off, 2 = on */
#define AD4SM_INFO_SPLIT_PAIRS 4  /* node pairs ad4sm_eval_assemble_split_device assembles ahead of the exchange (0: none) */
#define AD4SM_INFO_DIST_TRANSPORT 5 /* after ad4sm_dist_attach: 1 = ncclSend/ncclRecv of the staged rows, 2 = copy-engine push (CUDA IPC); 0 = not attached */
int ad4sm_set_option(ad4sm_plan* plan, int32_t option, int64_t value);
int ad4sm_get_info(ad4sm_plan* plan, int32_t what, int64_t* value);

/* Per-kernel CUDA-event timing (bench.py roofline): enable, run evaluations, then read the
 * accumulated milliseconds and launch counts per kernel class. */
#define AD4SM_PROF_ELEMENT 0  /* K1 element energy/gradient/Hessian */
#define AD4SM_PROF_ASSEMBLE 1 /* K2 segmented assembly of Kt and r */
#define AD4SM_PROF_REDUCE 2   /* ϕ reduction */
#define AD4SM_PROF_GATHER 3   /* time between the element and assembly stages (interface exchange of a partitioned run) */
#define AD4SM_PROF_NCLASS 4
int ad4sm_profile_enable(ad4sm_plan* plan, int32_t on);
int ad4sm_profile_read(ad4sm_plan* plan, double ms[AD4SM_PROF_NCLASS], int64_t launches[AD4SM_PROF_NCLASS],
                       int32_t reset);

/* FP64 DFMA peak microbenchmark (roofline denominator; MEASURED_PEAKS.json has no FP64 entry). */
int ad4sm_measure_fp64_peak(int32_t device, double* tflops, double* sm_mhz_est);

#ifdef __cplusplus
}
#endif
#endif /* AD4SM_B200_H */
