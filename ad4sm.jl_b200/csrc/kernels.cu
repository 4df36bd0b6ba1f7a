// kernels.cu -- CUDA kernels of the makeϕrKt hot path (sm_100a).
//
//   K1  element energy / gradient / Hessian          replaces the @threads loop of _assemble_ϕrKt
//                                                     (/root/reference/src/elasticelements.jl:216-218)
//   K2  deterministic assembly of Kt (segmented left fold over a precomputed source map, no atomics)
//   K3  residual assembly                             replaces makeϕrKt(Φ, elems, u)
//   ϕ   fixed-tree reduction                          (/root/reference/src/elements.toolkit.jl:169-203)
//   pattern / dropzeros                               sparse(...) structure and dropzeros (toolkit:202)
//
// K1 thread mapping (Hex8: LP = 8 lanes per element, 16 elements per 128-thread CTA):
//   phase 1  lane q evaluates Gauss point q: gathers u, builds F, evaluates the material tangent in
//            registers and leaves (w·A[45], w·P[9], w·ψ) in a shared-memory record;
//   phase 2  lane q owns column node q: streams the 8 Gauss-point records (shared-memory broadcast) and
//            accumulates its 4-5 node blocks K(a,q) = Σ_gp ∇N_aᵀ (A ∇N_q) in registers -- FP64-FMA bound,
//            no cross-lane floating-point reduction.
// Elements with one Gauss point and N <= 4 (Tet04, Tria) run one thread per element, all in registers.
#include <cstdlib>
#include <cub/cub.cuh>

#include "element.cuh"
#include "internal.h"

namespace ad4sm {

constexpr int K1_THREADS = 128;

// Shared-memory layout of one element (doubles):  [P][rstride] GP records | [P][gstride] ∇N as [k][a] | ([P][nstride] N)
// Strides are even (16-byte aligned rows) with an odd number of 16-byte units, so the 8 lanes of an
// element, which write rows `stride` apart with STS.128 in phase 1, hit 8 different bank groups.
__host__ __device__ inline int odd_units_stride(int n) {
  int s = (n + 1) & ~1;
  if (((s >> 1) & 1) == 0) s += 2;
  return s;
}
struct SmemLayout {
  int rstride, gstride, nstride, goff, noff, estride;
};
__host__ __device__ inline SmemLayout k1_layout(int rec_size, int D, int N, int P, bool with_N) {
  SmemLayout L;
  L.rstride = odd_units_stride(rec_size);
  L.gstride = odd_units_stride(D * N);
  L.nstride = with_N ? odd_units_stride(N) : 0;
  L.goff = P * L.rstride;
  L.noff = L.goff + P * L.gstride;
  const int tot = L.noff + P * L.nstride;
  // phase 2 reads ∇N with LDS.64 (a half-warp spans two elements): skew successive elements by 8 doubles
  L.estride = ((tot + 15) & ~15) + 8;
  return L;
}

template <int D, int N> __device__ __forceinline__ void load_gp_inputs(const K1Args& a, long long e, int gp, double* G,
                                                                       double* ue, const int* nd) {
  const int P = a.P;
  const double* gN = a.gradN + e * (long long)(D * P * N);
#pragma unroll
  for (int k = 0; k < D; k++)
#pragma unroll
    for (int n = 0; n < N; n++) G[k * N + n] = __ldg(gN + (k * P + gp) * N + n);
#pragma unroll
  for (int n = 0; n < N; n++)
#pragma unroll
    for (int j = 0; j < D; j++) ue[n * D + j] = __ldg(a.u + (long long)nd[n] * a.dpn + j);
}
template <int CNT> __device__ __forceinline__ void store_row(double* dst, const double* src) {
#pragma unroll
  for (int i2 = 0; i2 < (CNT + 1) / 2; i2++) store2(dst, i2, src[2 * i2], (2 * i2 + 1 < CNT) ? src[2 * i2 + 1] : 0.0);
}

// ---- TMA / mbarrier helpers (sm_90+ PTX; SASS: UBLKCP, SYNCS) ----------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {  // SASS: LDGSTS
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// shared -> global bulk copy (TMA store, SASS: UBLKCP); completion tracked by the issuing thread's bulk groups
__device__ __forceinline__ void bulk_store(void* dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Per-warp shared-memory slice of the persistent K1 (byte offsets).  Each warp is autonomous: it owns
// 32/LP elements per iteration, a double-buffered input stage filled by TMA bulk copies (∇N, wgt, nodes, N),
// the GP records, and two mbarriers; no CTA-wide barrier is ever needed.
struct WarpLayout {
  int rstride;          // doubles between GP records
  int rec_el;           // bytes of records per element (P * rstride * 8)
  // two stages each; stage s lives at off + s * stg (plain arithmetic: indexing an array member of a kernel
  // parameter with a run-time index makes the compiler copy the struct to slow local memory)
  int g_off, g_stg, g_el;  // ∇N stage, element stride in bytes (raw [k][gp][a] + skew)
  int w_off, w_stg;        // wgt stage [EW][P] doubles
  int n_off, n_stg;        // node ids stage [EW][N] int32
  int v_off, v_stg;        // shape-function values stage [EW][P][N] doubles (phase field)
  int u_off, u_el;      // gathered nodal u (and d) of one element group: [EW][N*D (+N)] doubles, filled by cp.async
  int psi_off;          // [EW][P] w·ψ of the Gauss points
  int bar_off;          // 4 x 8-byte mbarriers: [0,1] ∇N/wgt/N stages, [2,3] node-id stages (loaded two groups ahead)
  int kout_el;          // epilogue staging stride per element (bytes): blocks, then r; 0 = store directly
  int re_off;           // byte offset of the element residual inside its epilogue staging
  int bytes;            // slice size, multiple of 128
};
static inline int rup(int x, int m) { return (x + m - 1) / m * m; }
static WarpLayout k1p_layout(int D, int N, int P, int LP, bool pf) {
  WarpLayout L;
  const int EW = 32 / LP, DD = D * D, NPB = n_pair_blocks(N);
  L.rstride = odd_units_stride(DD * (DD + 1) / 2 + DD);  // w·A and w·P; w·ψ lives in its own small array
  L.rec_el = P * L.rstride * 8;
  // phase 2 broadcasts a record chunk to the lanes of each element with LDS.128: the EW elements of a warp must sit
  // in different 16-byte bank groups, i.e. the element stride must be an odd number of 16-byte units
  if (((L.rec_el >> 4) & 1) == 0) L.rec_el += 16;
  int off = EW * L.rec_el;
  L.g_el = D * P * N * 8;
  if (LP == 8 && (L.g_el % 128) == 0) L.g_el += 64;  // half-warp = 2 elements read ∇N with LDS.64: skew by 64 B
  L.g_off = off;
  L.g_stg = EW * L.g_el;
  off += 2 * L.g_stg;
  L.w_off = off;
  L.w_stg = rup(EW * P * 8, 16);
  off += 2 * L.w_stg;
  L.n_off = off;
  L.n_stg = rup(EW * N * 4, 16);
  off += 2 * L.n_stg;
  L.v_off = off;
  L.v_stg = pf ? rup(EW * P * N * 8, 16) : 0;
  off += 2 * L.v_stg;
  L.u_el = rup((N * D + (pf ? N : 0)) * 8, 16);
  if (((L.u_el >> 4) & 1) == 0) L.u_el += 16;  // odd number of 16-byte units: conflict-free chunk reads
  L.u_off = off;  // one buffer: it is consumed into registers before the next group's gather is issued
  off += EW * L.u_el;
  L.psi_off = off;
  off += rup(EW * P * 8, 16);
  L.bar_off = off;
  off += 32;  // 4 mbarriers
  L.bytes = rup(off, 128);
  // epilogue: the element blocks and residual are staged in the (dead) record region when they fit, and leave
  // with one bulk copy each (16-byte granularity)
  L.kout_el = 0;
  L.re_off = NPB * DD * 8;
  if ((NPB * DD) % 2 == 0 && (N * D) % 2 == 0) {
    int k = L.re_off + N * D * 8;
    if (LP == 8)
      while (k % 128 != 64) k += 16;
    if (k <= L.rec_el) L.kout_el = k;
  }
  return L;
}

// ---- assembly of one node-pair block / one residual entry (K2 / K3) ----------
// Sums the DB x DB source blocks of node-pair block `blk` in `elems` order (left fold, like Julia's sparse(),
// toolkit:202) and writes DB rows of dpn consecutive values into CSC columns / CSR rows (i, 0..dpn-1).
// CG: read the staging with ld.global.cg (L2 only) -- required when it was written earlier in the same kernel.
template <int DB, bool CG> __device__ __forceinline__ int assemble_block(const AsmArgs& a, long long blk) {
  const int dpn = a.dpn;
  const unsigned s0 = __ldg(a.blk_ptr + blk), s1 = __ldg(a.blk_ptr + blk + 1);
  const unsigned i = __ldg(a.blk_row + blk);
  double acc[DB * DB];
#pragma unroll
  for (int t = 0; t < DB * DB; t++) acc[t] = 0.0;
  unsigned code = s0 < s1 ? __ldg(a.src + s0) : 0u;
  for (unsigned s = s0; s < s1; s++) {
    const unsigned cur = code;
    if (s + 1 < s1) code = __ldg(a.src + s + 1);  // prefetch the next source while this one is summed
    const double* B = a.Ke + (long long)(cur >> 1) * (DB * DB);
    double v[DB * DB];
#pragma unroll
    for (int t = 0; t < DB * DB; t++) v[t] = CG ? __ldcg(B + t) : B[t];
    if (cur & 1u) {
#pragma unroll
      for (int j = 0; j < DB; j++)
#pragma unroll
        for (int l = 0; l < DB; l++) acc[j * DB + l] += v[l * DB + j];
    } else {
#pragma unroll
      for (int t = 0; t < DB * DB; t++) acc[t] += v[t];
    }
  }
  const long long nb0 = __ldg(a.nbr_ptr + i);
  const int nn = (int)(__ldg(a.nbr_ptr + i + 1) - nb0);
  const int p = (int)(blk - nb0);
  double* out = a.nzval + nb0 * (dpn * dpn) + (long long)p * dpn;
  int nz = 0;
#pragma unroll
  for (int j = 0; j < DB; j++) {
    double* row = out + (long long)j * dpn * nn;
#pragma unroll
    for (int l = 0; l < DB; l++) {
      row[l] = acc[j * DB + l];
      nz += (acc[j * DB + l] == 0.0);
    }
    for (int l = DB; l < dpn; l++) {
      row[l] = 0.0;
      nz++;
    }
  }
  for (int j = DB; j < dpn; j++) {
    double* row = out + (long long)j * dpn * nn;
    for (int l = 0; l < dpn; l++) {
      row[l] = 0.0;
      nz++;
    }
  }
  return nz;
}
// r[(node, j)]: left fold over the node's elements in `elems` order (r[idxii] += grad, toolkit:190)
template <bool CG> __device__ __forceinline__ void assemble_r_entry(const AsmArgs& a, long long i, int j) {
  double acc = 0.0;
  if (j < a.DB) {
    const unsigned s0 = __ldg(a.adj_ptr + i), s1 = __ldg(a.adj_ptr + i + 1);
    for (unsigned s = s0; s < s1; s++) {
      unsigned long long v = __ldg(a.adj + s);
      if (v >= a.halo_adj0) v += a.halo_adj_shift;  // HALO rows of the receive buffer in use (peer push)
      const double* p = a.re + v * a.DB + j;
      acc += CG ? __ldcg(p) : *p;
    }
  }
  a.r[i * a.dpn + j] = acc;
}

// ---- phase 1 of the persistent K1: one Gauss point per lane ----------------------------------------------
// Register-allocation note (ptxas 12.9): both phases are called by ALL lanes of the warp, unconditionally, with the
// lane predicates passed in.  Wrapping them in `if (valid)` / `if (active)` made the allocator keep 12 phase-2
// accumulators in local memory inside the FP64-dense loop although ~80 registers were free; written this way the
// kernel needs 250 registers and no local memory at all.
// Operands are read in 16-byte chunks (one node pair of a [gp] row).  Lane q starts at chunk rho(q): the 8 lanes
// of a Hex8 element read rows 64 B apart, and the rotation spreads them over all eight 16-byte bank groups
// (conflict-free LDS.128).  Sums over nodes are permutation-invariant as long as G, u, d and N use the same node
// order, so no un-rotation is needed.
// Called by ALL lanes of the warp (it contains warp syncs).  After the nodal values are in registers it starts the
// asynchronous gather (cp.async) of the NEXT group's nodal u into the same buffer, so that it flies during the rest
// of the iteration.
//   ng_ids: this lane's element row of node ids of the next group (shared memory), nullptr = no next group
//   ng_dst: shared address of this lane's element row in the u buffer
template <int D, int N, int LP, int MF, bool PF>
__device__ __forceinline__ void k1_phase1(const MatDev* mat, int q, bool valid, int P, int rstride, const double* uS,
                                       const double* gS, const double* vS, const double* wS, double* recs, double* psiS,
                                       const int* ng_ids, unsigned ng_dst, const double* u, const double* dfield, int dpn) {
  constexpr bool VEC = (N % 2 == 0);
  constexpr int CH = N / 2;
  const int rho = (N == 8) ? ((q >> 1) & 3) : 0;
  double ue[N * D], de[N];
  if (valid) {
    if constexpr (VEC) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        int pc = c + rho;
        if (pc >= CH) pc -= CH;
        const double2* src = reinterpret_cast<const double2*>(uS + 2 * pc * D);
#pragma unroll
        for (int t = 0; t < D; t++) {
          const double2 v = src[t];
          ue[2 * c * D + 2 * t] = v.x;
          ue[2 * c * D + 2 * t + 1] = v.y;
        }
        if (PF) {
          const double2 v = *reinterpret_cast<const double2*>(uS + N * D + 2 * pc);
          de[2 * c] = v.x;
          de[2 * c + 1] = v.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < N * D; i++) ue[i] = uS[i];
      if (PF) {
#pragma unroll
        for (int n = 0; n < N; n++) de[n] = uS[N * D + n];
      }
    }
  }
  __syncwarp();  // the u buffer has been consumed
  if (ng_ids && q < N) {
    const long long node = ng_ids[q];
#pragma unroll
    for (int j = 0; j < D; j++) cp_async8(ng_dst + (q * D + j) * 8, u + node * dpn + j);
    if (PF) cp_async8(ng_dst + (N * D + q) * 8, dfield + node);
  }
  cp_async_commit();
  const MatDev m = *mat;
  for (int gp0 = 0; gp0 < P; gp0 += LP) {  // warp-uniform trip count
    const int gp = gp0 + q;
    const bool on = valid && gp < P;
    if (on) {
      double G[D * N], Nv[N];
      if constexpr (VEC) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
          int pc = c + rho;
          if (pc >= CH) pc -= CH;
#pragma unroll
          for (int k = 0; k < D; k++) {
            const double2 v = *reinterpret_cast<const double2*>(gS + (k * P + gp) * N + 2 * pc);
            G[k * N + 2 * c] = v.x;
            G[k * N + 2 * c + 1] = v.y;
          }
          if (PF) {
            const double2 v = *reinterpret_cast<const double2*>(vS + gp * N + 2 * pc);
            Nv[2 * c] = v.x;
            Nv[2 * c + 1] = v.y;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < D; k++)
#pragma unroll
          for (int n = 0; n < N; n++) G[k * N + n] = gS[(k * P + gp) * N + n];
        if (PF) {
#pragma unroll
          for (int n = 0; n < N; n++) Nv[n] = vS[gp * N + n];
        }
      }
      gauss_point_u<D, N, MF, PF>(m, G, ue, de, Nv, wS[gp], recs + gp * rstride, psiS + gp);
    }
  }
}

// ---- phase 2 of the persistent K1: one column node per lane ------------------------------------------------
// Called by ALL lanes of the warp (inactive lanes compute on valid shared memory and store nothing): see the note
// at k1_phase1, and the stores may alias the records other lanes are still reading: accumulate, __syncwarp, store.
template <int D, int N, int PT>
__device__ __forceinline__ void k1_phase2(int q, int P, int rstride, bool active, bool want_K, bool staged, const double* recs,
                                          const double* gS, double* kout_s, double* rout_s, double* kout_g, double* rout_g,
                                          const unsigned* cpos_e, double* C, const K1Args& a, long long slot,
                                          long long* t_acc = nullptr) {
  constexpr int NO = N / 2 + 1, DD = D * D;
  // destination-sorted staging: this lane's block positions, fetched now so that the loads fly during the contraction
  unsigned cp[NO];
  if (cpos_e && a.exp != 1) {
#pragma unroll
    for (int o = 0; o < NO; o++) cp[o] = (active && o < n_offsets(N, q)) ? __ldg(cpos_e + o * N + q) : 0xffffffffu;
  }
  double acc[NO][DD], rq[D];
  column_u_acc<D, N, PT, VecMem>(q, P, recs, rstride, gS, N, (PT ? PT : P) * N, want_K, acc, rq);
  if (t_acc) *t_acc = clock64();  // PROF only (nullptr is a compile-time constant otherwise)
  if (cpos_e) {  // (D == 3 only) every block goes to its own sorted slot
    if constexpr (D == 3) {
      __syncwarp();  // all lanes are done reading the records: their area becomes the staging
      if (a.exp == 1) {  // timing experiment: keep the accumulators alive, store (practically) nothing
        double t = rq[0] + rq[1] + rq[2];
#pragma unroll
        for (int o = 0; o < NO; o++)
#pragma unroll
          for (int i = 0; i < 9; i++) t += acc[o][i];
        if (t == 1.2345e300) C[0] = t;
        __syncwarp();
        return;
      }
      if (active) {
#pragma unroll
        for (int o = 0; o < NO; o++) {
          if (cp[o] != 0xffffffffu) {
            double v[10];
#pragma unroll
            for (int i = 0; i < 9; i++) v[i] = acc[o][i];
            v[9] = (cp[o] & 1u) ? 1.0 : 0.0;  // "stored transposed": the consumer (K2s) flips it
            if (kout_s) {  // via shared memory + one 80-byte bulk copy (TMA) per block: no store burst through the LSU
              double2* st = reinterpret_cast<double2*>(kout_s + (o * N + q) * CBLK3);
#pragma unroll
              for (int t = 0; t < 5; t++) st[t] = make_double2(v[2 * t], v[2 * t + 1]);
            } else {
              double2* dst = reinterpret_cast<double2*>(C + (long long)(cp[o] >> 1) * CBLK3);
#pragma unroll
              for (int t = 0; t < 5; t++) dst[t] = make_double2(v[2 * t], v[2 * t + 1]);
            }
          }
        }
        if (kout_s) {
          fence_proxy_async();  // this lane's own STS -> its own bulk copies: no cross-lane dependency
#pragma unroll
          for (int o = 0; o < NO; o++)
            if (cp[o] != 0xffffffffu)
              bulk_store(C + (long long)(cp[o] >> 1) * CBLK3, smem_u32(kout_s + (o * N + q) * CBLK3), (unsigned)(CBLK3 * 8));
        }
#pragma unroll
        for (int j = 0; j < D; j++) rout_g[q * D + j] = rq[j];
        // interface element of a partitioned run: element-major copy for the owner rank (peer memory, or the local Ke
        // that the send/recv exchange reads).  Looked up here, after the contraction, to keep phase 2 lean.
        for (int x = 0; x < a.n_xr; x++) {
          const long long k = slot - a.xr[x].first;
          if (k >= 0 && k < a.xr[x].count) {
            column_u_store<D, N>(q, acc, a.xr[x].ke + k * (long long)(n_pair_blocks(N) * DD));
#pragma unroll
            for (int j = 0; j < D; j++) (a.xr[x].re + k * (long long)(N * D))[q * D + j] = rq[j];
          }
        }
        if (a.xmap) {  // general partition: this element's row of the export buffers, if a peer owns some of its nodes
          const long long k = __ldg(a.xmap + slot);
          if (k >= 0) {
            column_u_store<D, N>(q, acc, a.xke + k * (long long)(n_pair_blocks(N) * DD));
#pragma unroll
            for (int j = 0; j < D; j++) (a.xre + k * (long long)(N * D))[q * D + j] = rq[j];
          }
        }
      }
      if (kout_s) bulk_commit();
    }
    __syncwarp();
    return;
  }
  __syncwarp();
  if (active) {
    if (staged) {  // shared-memory staging (STS), leaves with the bulk stores
      column_u_store<D, N>(q, acc, kout_s);
#pragma unroll
      for (int j = 0; j < D; j++) rout_s[q * D + j] = rq[j];
    } else {
      if (want_K) column_u_store<D, N>(q, acc, kout_g);
#pragma unroll
      for (int j = 0; j < D; j++) rout_g[q * D + j] = rq[j];
    }
  }
}


// ---- K1, u-dual, persistent, LP lanes per element, TMA-staged inputs ------------------------------
// PROF (developer tool, AD4SM_K1_PROF=1, Hex8 NeoHooke only): lane 0 of every warp accumulates the clock64() cycles of
// the four parts of its iteration (0 head: TMA issue + waits, 1 phase 1, 2 phase 2 up to the end of the contraction incl.
// the epilogue stores, 3 rest) in g_k1_prof[gwarp][4]; [4] = iterations.
__device__ long long g_k1_prof[148 * 8 * 16];
template <int D, int N, int LP, int MF, bool PF, bool PROF = false>
__global__ void __launch_bounds__(256, 1)
    k1_u_kernel(const __grid_constant__ K1Args a, const WarpLayout L, const int warps_per_cta) {
  using R = GpRec<D>;
  constexpr int EW = 32 / LP;
  constexpr int DD = D * D;
  constexpr int NPB = n_pair_blocks(N);
  extern __shared__ __align__(128) unsigned char smraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* const ws0 = smraw + (size_t)warp * L.bytes;
  const int el = lane / LP, q = lane % LP;
  const int P = a.P;
  const long long n_groups = (a.ne + EW - 1) / EW;
  const long long gstep = (long long)gridDim.x * warps_per_cta;
  const unsigned bar0 = smem_u32(ws0 + L.bar_off);
  if (lane == 0) {
    for (int i = 0; i < 4; i++) mbar_init(bar0 + 8 * i, 1);
    fence_mbar_init();
  }
  __syncwarp();
  unsigned char* const ws = ws0;

  // TMA bulk copies of one element group into stage s (∇N per element, wgt, N), all issued by lane 0 in straight-line
  // code.  (UBLKCP takes its operands from uniform registers: copies "issued by different lanes" are serialised by the
  // compiler into a loop over the lanes anyway, and cost 1 000+ cycles per iteration that way -- measured,
  // profiles/r02_k1_experiments.txt.)  No proxy fence: the stage buffers are only ever WRITTEN by the async proxy and
  // read by LDS; the reads of the previous use of the stage are ordered before this point by the __syncwarp that ends
  // every iteration (the write-after-read case needs no cross-proxy fence).
  auto issue = [&](long long grp, int s) {
    if (lane != 0) return;
    const long long e0 = grp * EW;
    const int cnt = (int)((a.ne - e0) < EW ? (a.ne - e0) : EW);
    const unsigned bar = bar0 + 8 * s;
    const unsigned gb = (unsigned)(D * P * N * 8);
    const unsigned wbytes = (unsigned)((cnt * P * 8 + 15) & ~15);
    const unsigned vbytes = PF ? (unsigned)((cnt * P * N * 8 + 15) & ~15) : 0u;
    mbar_expect_tx(bar, cnt * gb + wbytes + vbytes);
    const unsigned dst0 = smem_u32(ws + (L.g_off + (s) * L.g_stg));
    const double* src0 = a.gradN + e0 * (long long)(D * P * N);
#pragma unroll
    for (int i = 0; i < EW; i++)
      if (i < cnt) tma_load_1d(dst0 + i * L.g_el, src0 + (long long)i * (D * P * N), gb, bar);
    tma_load_1d(smem_u32(ws + (L.w_off + (s) * L.w_stg)), a.wgt + e0 * P, wbytes, bar);
    if (PF) tma_load_1d(smem_u32(ws + (L.v_off + (s) * L.v_stg)), a.Nval + e0 * (long long)(P * N), vbytes, bar);
  };
  // lane 0: node ids of group grp into id stage s.  They run one group further ahead than the rest, so that the
  // u gather of group i+1 can start at the beginning of iteration i.
  auto issue_ids = [&](long long grp, int s) {
    const long long e0 = grp * EW;
    const int cnt = (int)((a.ne - e0) < EW ? (a.ne - e0) : EW);
    const unsigned nbytes = (unsigned)((cnt * N * 4 + 15) & ~15);
    mbar_expect_tx(bar0 + 16 + 8 * s, nbytes);
    tma_load_1d(smem_u32(ws + (L.n_off + (s) * L.n_stg)), a.nodes + e0 * N, nbytes, bar0 + 16 + 8 * s);
  };

  const int gwarp = blockIdx.x * warps_per_cta + warp;  // global warp id = its first element group
  long long grp = gwarp;
  long long pt[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt0 = 0, ph0 = 0;
  auto htick = [&](int k) {  // PROF: sub-parts of the head (6 TMA issue, 7 wait ∇N, 8 wait ids, 9 wait u gather, 10 wait stores)
    if constexpr (PROF) {
      const long long t = clock64();
      pt[k] += t - ph0;
      ph0 = t;
    }
  };
  auto tick = [&](int k) {
    if constexpr (PROF) {
      const long long t = clock64();
      pt[k] += t - pt0;
      pt0 = t;
    }
  };
  if constexpr (PROF) pt0 = clock64();
  if (grp < n_groups) {
    issue(grp, 0);
    if (lane == 0) {
      issue_ids(grp, 0);
      if (grp + gstep < n_groups) issue_ids(grp + gstep, 1);
    }
    mbar_wait(bar0 + 16, 0u);
    const long long e = grp * EW + el;  // first group: gather its u synchronously with the pipeline start
    if (e < a.ne && q < N) {
      const long long node = (reinterpret_cast<const int*>(ws + (L.n_off + (0) * L.n_stg)) + el * N)[q];
      const unsigned dst = smem_u32(ws + L.u_off + el * L.u_el);
#pragma unroll
      for (int j = 0; j < D; j++) cp_async8(dst + (q * D + j) * 8, a.u + node * a.dpn + j);
      if (PF) cp_async8(dst + (N * D + q) * 8, a.dfield + node);
    }
    cp_async_commit();
  }
  for (int it = 0; grp < n_groups; grp += gstep, it++) {
    const int s = it & 1;
    const bool has_next = grp + gstep < n_groups;
    if constexpr (PROF) ph0 = clock64();
    if (has_next) issue(grp + gstep, s ^ 1);
    if (lane == 0 && grp + 2 * gstep < n_groups) issue_ids(grp + 2 * gstep, s);  // id stage s held this group's ids: gather done
    if constexpr (PROF) __syncwarp();
    htick(6);
    mbar_wait(bar0 + 8 * s, (unsigned)((it >> 1) & 1));                        // this group's ∇N, wgt, N
    htick(7);
    if (has_next) mbar_wait(bar0 + 16 + 8 * (s ^ 1), (unsigned)(((it + 1) >> 1) & 1));  // next group's node ids
    htick(8);
    cp_async_wait_all();                                                        // this group's u
    htick(9);
    if (a.cpos && a.sorted_tma) bulk_wait_read_all();  // sorted staging through per-lane bulk copies: their shared-memory source may be overwritten
    if (lane == 0) bulk_wait_read_all();  // the previous bulk stores have read the staging area (aliases the records)
    __syncwarp();
    htick(10);
    tick(0);

    const long long e = grp * EW + el;
    const bool valid = e < a.ne;
    const bool active = valid && q < N;
    const long long slot = a.slot0 + e;
    double* recs = reinterpret_cast<double*>(ws + (size_t)el * L.rec_el);
    const double* uS = reinterpret_cast<const double*>(ws + L.u_off + (size_t)el * L.u_el);
    double* psiS = reinterpret_cast<double*>(ws + L.psi_off) + el * P;
    const double* gS = reinterpret_cast<const double*>(ws + (L.g_off + (s) * L.g_stg) + (size_t)el * L.g_el);
    const double* wS = reinterpret_cast<const double*>(ws + (L.w_off + (s) * L.w_stg)) + el * P;
    const double* vS = reinterpret_cast<const double*>(ws + (L.v_off + (s) * L.v_stg)) + el * P * N;

    // phase 1: one Gauss point per lane; it also starts the next group's u gather
    const int* ng_ids =
        (has_next && (grp + gstep) * EW + el < a.ne) ? reinterpret_cast<const int*>(ws + (L.n_off + (s ^ 1) * L.n_stg)) + el * N : nullptr;
    k1_phase1<D, N, LP, MF, PF>(&a.mat, q, valid, P, L.rstride, uS, gS, vS, wS, recs, psiS, ng_ids,
                                smem_u32(ws + L.u_off + el * L.u_el), a.u, a.dfield, a.dpn);
    __syncwarp();
    tick(1);

    // phase 2: one column node per lane; the element blocks and residual go to the
    // staging area in the (then dead) record region, or straight to global memory when they do not fit
    const bool sorted = a.cpos != nullptr && a.want_K;
    const unsigned* cpos_e = sorted ? a.cpos + slot * (long long)NPB : nullptr;
    const bool staged = a.want_K && L.kout_el && !sorted;
    const bool sorted_tma = sorted && (NPB * CBLK3 * 8 <= L.rec_el) && a.sorted_tma;
    unsigned char* stg = ws + (size_t)el * (sorted ? L.rec_el : L.kout_el);
    constexpr int PTC = (N == 8) ? 8 : ((N == 4) ? 4 : 0);  // the usual Gauss count of the shape, compiled in
    long long t_acc = 0;
    if (PTC && P == PTC)
      k1_phase2<D, N, PTC>(q, P, L.rstride, active, a.want_K != 0, staged, recs, gS,
                           (sorted && !sorted_tma) ? nullptr : reinterpret_cast<double*>(stg),
                           reinterpret_cast<double*>(stg + L.re_off), a.Ke + slot * (long long)(NPB * DD),
                           a.re + slot * (long long)(N * D), cpos_e, a.C, a, slot, PROF ? &t_acc : nullptr);
    else
      k1_phase2<D, N, 0>(q, P, L.rstride, active, a.want_K != 0, staged, recs, gS,
                         (sorted && !sorted_tma) ? nullptr : reinterpret_cast<double*>(stg),
                         reinterpret_cast<double*>(stg + L.re_off), a.Ke + slot * (long long)(NPB * DD),
                         a.re + slot * (long long)(N * D), cpos_e, a.C, a, slot, PROF ? &t_acc : nullptr);
    if constexpr (PROF) {  // split phase 2 at the end of the contraction: [2] contraction, [5] epilogue stores
      if (t_acc) {
        pt[5] += clock64() - t_acc;
        pt[2] -= clock64() - t_acc;
      }
    }
    tick(2);
    if (valid && q == 0) {  // ϕ_e = Σ_gp w ψ, in Gauss-point order (elasticelements.jl:199-206)
      double ssum = 0.0;
      for (int gp = 0; gp < P; gp++) ssum += psiS[gp];
      a.phie[slot] = ssum;
    }
    if (staged) {  // one bulk copy (TMA store) per element and array
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        const long long e0 = grp * EW;
        const int cnt = (int)((a.ne - e0) < EW ? (a.ne - e0) : EW);
        for (int i = 0; i < cnt; i++) {
          const long long sl = a.slot0 + e0 + i;
          const unsigned src = smem_u32(ws + (size_t)i * L.kout_el);
          bulk_store(a.Ke + sl * (long long)(NPB * DD), src, (unsigned)(NPB * DD * 8));
          bulk_store(a.re + sl * (long long)(N * D), src + L.re_off, (unsigned)(N * D * 8));
          for (int x = 0; x < a.n_xr; x++) {  // interface element: the same bytes go to the owner rank over NVLink
            const long long k = sl - a.xr[x].first;
            if (k >= 0 && k < a.xr[x].count) {
              bulk_store(a.xr[x].ke + k * (long long)(NPB * DD), src, (unsigned)(NPB * DD * 8));
              bulk_store(a.xr[x].re + k * (long long)(N * D), src + L.re_off, (unsigned)(N * D * 8));
            }
          }
        }
        bulk_commit();
      }
    }
    __syncwarp();
    tick(3);
    pt[4]++;
  }
  if constexpr (PROF) {
    if (lane == 0)
      for (int k = 0; k < 12; k++) g_k1_prof[(size_t)gwarp * 16 + k] = pt[k];
  }
  if (a.cpos && a.sorted_tma) bulk_wait_read_all();
  if (lane == 0) {
    if (a.n_xr) {
      bulk_wait_all();  // peer stores must be complete, not only read, when the kernel ends
      __threadfence_system();
    } else {
      bulk_wait_read_all();  // shared memory must outlive the last bulk stores
    }
  }
}

// ---- K1, u-dual, one thread per element (P == 1) ------------------------------------------------
template <int D, int N, int MF, bool PF> __global__ void __launch_bounds__(K1_THREADS) k1_u_kernel_t1(const K1Args a) {
  using R = GpRec<D>;
  constexpr int DD = D * D;
  const long long e = (long long)blockIdx.x * K1_THREADS + threadIdx.x;
  if (e >= a.ne) return;
  int nd[N];
#pragma unroll
  for (int n = 0; n < N; n++) nd[n] = __ldg(a.nodes + e * N + n);
  double G[D * N], ue[N * D], de[N], Nv[N];
  load_gp_inputs<D, N>(a, e, 0, G, ue, nd);
  if (PF) {
#pragma unroll
    for (int n = 0; n < N; n++) {
      de[n] = __ldg(a.dfield + nd[n]);
      Nv[n] = __ldg(a.Nval + e * N + n);
    }
  }
  __align__(16) double rec[R::SIZE + 1];
  gauss_point_u<D, N, MF, PF>(a.mat, G, ue, de, Nv, __ldg(a.wgt + e), rec);
  const long long slot = a.slot0 + e;
  double* Ke = a.want_K ? a.Ke + slot * (long long)(n_pair_blocks(N) * DD) : nullptr;
  double* re = a.re + slot * (long long)(N * D);
#pragma unroll
  for (int q = 0; q < N; q++) column_u<D, N, 1, ScalarMem>(q, 1, rec, 0, G, 0, N, Ke, re);  // G is [k][a]
  a.phie[slot] = rec[R::OFF_PSI];
}


// ---- K1 for CAS batches (axisymmetric extension, element.cuh: cas_element): one thread per element, element-major
// staging (D = 2: the K2 gather assembles it like any 2-D element) ------------------------------------------------------
template <int N, int MF> __global__ void __launch_bounds__(64) k1_cas_kernel(const K1Args a) {
  const long long e = (long long)blockIdx.x * 64 + threadIdx.x;
  if (e >= a.ne) return;
  constexpr int ND = 2 * N;
  double ue[ND];
#pragma unroll
  for (int n = 0; n < N; n++) {
    const long long nd = __ldg(a.nodes + e * N + n);
#pragma unroll
    for (int j = 0; j < 2; j++) ue[n * 2 + j] = __ldg(a.u + nd * a.dpn + j);
  }
  double Kf[ND * ND], r[ND], phi;
  cas_element<N, MF>(a.mat, a.P, a.gradN + e * (long long)(2 * a.P * N), a.wgt + e * a.P, a.Nval + e * (long long)(a.P * N), ue,
                     a.want_K != 0, Kf, r, phi);
  const long long slot = a.slot0 + e;
  if (a.want_K) cas_store<N>(Kf, a.Ke + slot * (long long)(n_pair_blocks(N) * 4));
#pragma unroll
  for (int i = 0; i < ND; i++) a.re[slot * ND + i] = r[i];
  a.phie[slot] = phi;
}
template <int N> static cudaError_t launch_cas_n(const K1Args& a, cudaStream_t s) {
  const unsigned grid = (unsigned)((a.ne + 63) / 64);
  switch (a.mf) {
    case MF_HOOKE: k1_cas_kernel<N, MF_HOOKE><<<grid, 64, 0, s>>>(a); break;
    case MF_NEO: k1_cas_kernel<N, MF_NEO><<<grid, 64, 0, s>>>(a); break;
    case MF_MOONEY: k1_cas_kernel<N, MF_MOONEY><<<grid, 64, 0, s>>>(a); break;
    case MF_OGDEN: k1_cas_kernel<N, MF_OGDEN><<<grid, 64, 0, s>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
static cudaError_t launch_cas(const K1Args& a, cudaStream_t s) {
  if (a.D != 2 || !a.Nval || a.cpos) return cudaErrorInvalidValue;
  if (a.N == 3) return launch_cas_n<3>(a, s);
  if (a.N == 4) return launch_cas_n<4>(a, s);
  return cudaErrorInvalidValue;
}

// ---- K1, u-dual, one thread per element (P == 1), warp-cooperative output ------------------------------------------
// The thread-per-element kernel above leaves every store to the owning thread: 102 STG.64 per Tet4 element at a lane
// stride of 720 B, i.e. 32 cache lines per instruction -- the LSU's one-line-per-cycle front end, not DRAM, bounds it
// (C3: 1.40 ms for 2 M elements).  Here the threads of a warp put their element blocks / residuals into the warp's
// shared-memory slice (row stride 2 x odd doubles) and the warp copies them out together:
//   element-major staging: the 32 elements' rows are ONE contiguous range of Ke (and of re): 16-byte chunks, lane-contiguous;
//   destination-sorted staging (D = 3): every 80-byte block goes to its sorted slot (cpos), 5 consecutive lanes per block.
// Interface elements of a partitioned run are additionally copied element-major to their export rows (xr / xmap).
template <int D, int N> struct W1Layout {
  static constexpr int DD = D * D, NPB = N * (N + 1) / 2, KW = NPB * DD, RW = N * D;
  static constexpr int OS0 = KW + RW;
  static constexpr int OS = (OS0 % 2) ? OS0 + 1 : (((OS0 / 2) % 2) ? OS0 : OS0 + 2);  // even, OS/2 odd
  static constexpr int WARPS = 4;
  static constexpr size_t ROWS_BYTES = (size_t)WARPS * 32 * OS * sizeof(double);
  static constexpr size_t BYTES = ROWS_BYTES + (size_t)WARPS * 32 * NPB * sizeof(unsigned);  // + the warp's sorted positions
};
template <int D, int N, int MF, bool PF>
__global__ void __launch_bounds__(128) k1_u_kernel_w1(const __grid_constant__ K1Args a) {
  using R = GpRec<D>;
  using L = W1Layout<D, N>;
  constexpr int DD = L::DD, NPB = L::NPB, KW = L::KW, RW = L::RW, OS = L::OS;
  extern __shared__ __align__(16) double smw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long e0 = ((long long)blockIdx.x * L::WARPS + warp) * 32;
  if (e0 >= a.ne) return;  // whole warp; no CTA-wide barrier below
  const int cnt = (int)((a.ne - e0) < 32 ? (a.ne - e0) : 32);
  double* ws = smw + (size_t)warp * 32 * OS;
  double* row = ws + (size_t)lane * OS;
  const long long e = e0 + lane;
  const bool valid = lane < cnt;
  // destination-sorted staging: the positions of the warp's cnt * NPB blocks (contiguous in cpos) start travelling into
  // shared memory now (cp.async, no registers held), so that the store loop at the end does not wait on them
  if (a.cpos && a.want_K) {
    const unsigned* cp = a.cpos + (a.slot0 + e0) * (long long)NPB;
    const unsigned dst = smem_u32(reinterpret_cast<unsigned char*>(smw) + L::ROWS_BYTES) + (unsigned)(warp * 32 * NPB * 4);
#pragma unroll
    for (int k = 0; k < NPB; k++) {
      const int i = lane + 32 * k;
      if (i < cnt * NPB) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * i), "l"(cp + i) : "memory");
    }
    cp_async_commit();
  }
  if (valid) {
    int nd[N];
#pragma unroll
    for (int n = 0; n < N; n++) nd[n] = __ldg(a.nodes + e * N + n);
    double G[D * N], ue[N * D], de[N], Nv[N];
    load_gp_inputs<D, N>(a, e, 0, G, ue, nd);
    if (PF) {
#pragma unroll
      for (int n = 0; n < N; n++) {
        de[n] = __ldg(a.dfield + nd[n]);
        Nv[n] = __ldg(a.Nval + e * N + n);
      }
    }
    __align__(16) double rec[R::SIZE + 1];
    gauss_point_u<D, N, MF, PF>(a.mat, G, ue, de, Nv, __ldg(a.wgt + e), rec);
#pragma unroll
    for (int q = 0; q < N; q++) column_u<D, N, 1, ScalarMem>(q, 1, rec, 0, G, 0, N, a.want_K ? row : nullptr, row + KW);  // G is [k][a]
    a.phie[a.slot0 + e] = rec[R::OFF_PSI];
  }
  __syncwarp();
  const long long slot_w = a.slot0 + e0;
  // residual rows: contiguous [cnt][RW]
  {
    double* dst = a.re + slot_w * (long long)RW;
    if constexpr (RW % 2 == 0) {
      for (int c = lane; c < cnt * (RW / 2); c += 32) {
        const int el = c / (RW / 2), w = c - el * (RW / 2);
        reinterpret_cast<double2*>(dst)[c] = *reinterpret_cast<const double2*>(ws + (size_t)el * OS + KW + 2 * w);
      }
    } else {
      for (int c = lane; c < cnt * RW; c += 32) {
        const int el = c / RW, w = c - el * RW;
        dst[c] = ws[(size_t)el * OS + KW + w];
      }
    }
  }
  if (!a.want_K) return;
  if (a.cpos) {  // destination-sorted staging (D == 3): block b of element el -> C[cpos >> 1], 5 x 16 bytes
    if constexpr (D == 3) {
      cp_async_wait_all();
      __syncwarp();
      const unsigned* cps =
          reinterpret_cast<const unsigned*>(reinterpret_cast<const unsigned char*>(smw) + L::ROWS_BYTES) + (size_t)warp * 32 * NPB;
      for (int c = lane; c < cnt * NPB * 5; c += 32) {
        const int eb = c / 5, t = c - eb * 5;  // eb = el * NPB + b
        const int el = eb / NPB, b = eb - el * NPB;
        const unsigned code = cps[eb];
        const double* src = ws + (size_t)el * OS + b * 9 + 2 * t;
        const double v0 = src[0];
        const double v1 = t < 4 ? src[1] : ((code & 1u) ? 1.0 : 0.0);  // 10th double: "stored transposed" (K2s flips it)
        reinterpret_cast<double2*>(a.C + (long long)(code >> 1) * CBLK3)[t] = make_double2(v0, v1);
      }
    }
  } else {  // element-major staging: contiguous [cnt][KW]
    double* dst = a.Ke + slot_w * (long long)KW;
    if constexpr (KW % 2 == 0) {
      for (int c = lane; c < cnt * (KW / 2); c += 32) {
        const int el = c / (KW / 2), w = c - el * (KW / 2);
        reinterpret_cast<double2*>(dst)[c] = *reinterpret_cast<const double2*>(ws + (size_t)el * OS + 2 * w);
      }
    } else {
      for (int c = lane; c < cnt * KW; c += 32) {
        const int el = c / KW, w = c - el * KW;
        dst[c] = ws[(size_t)el * OS + w];
      }
    }
  }
  // interface elements of a partitioned run on the sorted path: element-major copies for the owner ranks
  if (a.cpos && (a.n_xr || a.xmap)) {
    for (int el = 0; el < cnt; el++) {
      const long long slot = slot_w + el;
      double *ke = nullptr, *re = nullptr;
      for (int x = 0; x < a.n_xr; x++) {
        const long long k = slot - a.xr[x].first;
        if (k >= 0 && k < a.xr[x].count) {
          ke = a.xr[x].ke + k * (long long)KW;
          re = a.xr[x].re + k * (long long)RW;
        }
      }
      if (!ke && a.xmap) {
        const long long k = __ldg(a.xmap + slot);
        if (k >= 0) {
          ke = a.xke + k * (long long)KW;
          re = a.xre + k * (long long)RW;
        }
      }
      if (ke) {
        for (int w = lane; w < KW; w += 32) ke[w] = ws[(size_t)el * OS + w];
        for (int w = lane; w < RW; w += 32) re[w] = ws[(size_t)el * OS + KW + w];
      }
    }
  }
}

// ---- K1, d-dual (makeϕrKt_d) -----------------------------------------------------------------------
template <int D, int N, int LP>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_d_kernel(const K1Args a, const SmemLayout L) {
  constexpr int EPC = K1_THREADS / LP;
  extern __shared__ __align__(16) double sm[];
  const int el = threadIdx.x / LP, q = threadIdx.x % LP;
  const long long e = (long long)blockIdx.x * EPC + el;
  const int P = a.P;
  double* recs = sm + el * L.estride;
  double* gS = recs + L.goff;
  double* nS = recs + L.noff;
  const bool valid = e < a.ne;
  if (valid) {
    int nd[N];
#pragma unroll
    for (int n = 0; n < N; n++) nd[n] = __ldg(a.nodes + e * N + n);
    for (int gp = q; gp < P; gp += LP) {
      double G[D * N], ue[N * D], de[N], Nv[N];
      load_gp_inputs<D, N>(a, e, gp, G, ue, nd);
#pragma unroll
      for (int n = 0; n < N; n++) {
        de[n] = __ldg(a.dfield + nd[n]);
        Nv[n] = __ldg(a.Nval + (e * P + gp) * N + n);
      }
      store_row<D * N>(gS + gp * L.gstride, G);
      store_row<N>(nS + gp * L.nstride, Nv);
      double* h = a.hist ? a.hist + (e * P + gp) * 2 : nullptr;
      gauss_point_d<D, N>(a.mat, G, Nv, ue, de, __ldg(a.wgt + e * P + gp), h, recs + gp * L.rstride);
    }
  }
  __syncthreads();
  if (valid && q < N) {
    const long long slot = a.slot0 + e;
    double* Ke = a.want_K ? a.Ke + slot * (long long)n_pair_blocks(N) : nullptr;
    column_d<D, N>(q, P, recs, L.rstride, gS, L.gstride, nS, L.nstride, Ke, a.re + slot * (long long)N);
    if (q == 0) {
      double s = 0.0;
      for (int gp = 0; gp < P; gp++) s += recs[gp * L.rstride + 2];
      a.phie[slot] = s;
    }
  }
}

template <int D, int N> __global__ void __launch_bounds__(K1_THREADS) k1_d_kernel_t1(const K1Args a) {
  const long long e = (long long)blockIdx.x * K1_THREADS + threadIdx.x;
  if (e >= a.ne) return;
  int nd[N];
#pragma unroll
  for (int n = 0; n < N; n++) nd[n] = __ldg(a.nodes + e * N + n);
  double G[D * N], ue[N * D], de[N], Nv[N];
  load_gp_inputs<D, N>(a, e, 0, G, ue, nd);
#pragma unroll
  for (int n = 0; n < N; n++) {
    de[n] = __ldg(a.dfield + nd[n]);
    Nv[n] = __ldg(a.Nval + e * N + n);
  }
  double rec[GpRecD<D>::SIZE];
  gauss_point_d<D, N>(a.mat, G, Nv, ue, de, __ldg(a.wgt + e), a.hist ? a.hist + e * 2 : nullptr, rec);
  const long long slot = a.slot0 + e;
  double* Ke = a.want_K ? a.Ke + slot * (long long)n_pair_blocks(N) : nullptr;
  double* re = a.re + slot * (long long)N;
#pragma unroll
  for (int q = 0; q < N; q++) column_d<D, N, 1>(q, 1, rec, 0, G, 0, Nv, 0, Ke, re);
  a.phie[slot] = rec[2];
}

// ---- K1 dispatch ---------------------------------------------------------------------------------
static bool lanes_for(int D, int N, int P, int& LP) {
  if ((D == 2 && N == 3) || (D == 3 && N == 4)) {
    // experiment knob: AD4SM_P1_PERSISTENT=1 runs one-Gauss-point simplices through the persistent 4-lanes-per-element
    // kernel (TMA-staged inputs, asynchronous u gather) instead of the thread-per-element one
    static const bool p1_persistent = getenv("AD4SM_P1_PERSISTENT") && atoi(getenv("AD4SM_P1_PERSISTENT")) != 0;
    LP = (P == 1 && !p1_persistent) ? 1 : 4;
    return true;
  }
  if (D == 2 && N == 4) {
    LP = 4;
    return true;
  }
  if (D == 3 && (N == 6 || N == 8)) {
    LP = 8;
    return true;
  }
  return false;
}

const char* k1_unsupported_reason(int D, int N, int P, int mf, int mode) {
  int LP;
  if (!lanes_for(D, N, P, LP)) return "unsupported element shape (D, N): supported are (2,3), (2,4), (3,4), (3,6), (3,8)";
  if (P < 1 || P > 27) return "number of Gauss points must be in 1..27";
  if (mf < 0 || mf > MF_OGDEN) return "unknown material family";
  (void)mode;
  return nullptr;
}

static int sm_count() {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  return n_sm;
}

bool k1_config(int D, int N, int P, bool pf, long long ne, K1Config* c) {
  int LP;
  if (!lanes_for(D, N, P, LP)) return false;
  c->persistent = LP > 1;
  c->EW = 32 / LP;
  c->warps = c->grid = c->staged = 0;
  if (LP == 1) return true;
  const WarpLayout L = k1p_layout(D, N, P, LP, pf);
  int warps = (227 * 1024) / L.bytes;
  if (warps < 1) return false;
  if (warps > 8) warps = 8;
  if (const char* e = getenv("AD4SM_K1_WARPS")) {  // experiment knob (profiles/): warps per CTA of the persistent K1
    const int w = atoi(e);
    if (w >= 1 && w < warps) warps = w;
  }
  // 2-D shapes need 140-190 registers and a few KB of shared memory per warp: two CTAs of 6 warps per SM (12 warps) instead
  // of one of 8 hide more of their latencies (the 3-D kernels fill the register file with 8 warps).  AD4SM_K1_CTAS overrides.
  int ctas = (D == 2) ? 2 : 1;
  if (const char* e = getenv("AD4SM_K1_CTAS")) ctas = atoi(e) >= 1 ? atoi(e) : ctas;
  if (ctas > 1 && warps > 6) warps = 6;
  const long long n_groups = (ne + c->EW - 1) / c->EW;
  if (n_groups < warps) warps = (int)(n_groups > 0 ? n_groups : 1);
  long long grid = (n_groups + warps - 1) / warps;
  if (grid > (long long)ctas * sm_count()) grid = (long long)ctas * sm_count();  // persistent: each warp strides over the element groups
  if (grid < 1) grid = 1;
  c->warps = warps;
  c->grid = (int)grid;
  c->staged = L.kout_el != 0;
  return true;
}

template <int D, int N, int LP, int MF, bool PF> static cudaError_t launch_u_t(const K1Args& a, cudaStream_t s) {
  if constexpr (LP == 1) {
    const char* t1e = getenv("AD4SM_K1_T1");  // 1: the round-1 kernel (per-thread stores; element-major staging only)
    const int t1 = t1e ? atoi(t1e) : 0;
    if (t1 && !a.cpos) {
      const unsigned grid = (unsigned)((a.ne + K1_THREADS - 1) / K1_THREADS);
      k1_u_kernel_t1<D, N, MF, PF><<<grid, K1_THREADS, 0, s>>>(a);
      return cudaGetLastError();
    }
    using WL = W1Layout<D, N>;
    auto kern = k1_u_kernel_w1<D, N, MF, PF>;
    if (WL::BYTES > 48 * 1024) {
      cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WL::BYTES);
      if (err != cudaSuccess) return err;
    }
    const long long per_cta = 32ll * WL::WARPS;
    kern<<<(unsigned)((a.ne + per_cta - 1) / per_cta), 32 * WL::WARPS, WL::BYTES, s>>>(a);
    return cudaGetLastError();
  } else {
    K1Config c;
    if (!k1_config(D, N, a.P, PF, a.ne, &c)) return cudaErrorInvalidValue;
    const WarpLayout L = k1p_layout(D, N, a.P, LP, PF);
    const int smem_max = 227 * 1024;
    const size_t smem = (size_t)c.warps * L.bytes;
    auto kern = k1_u_kernel<D, N, LP, MF, PF>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (err != cudaSuccess) return err;
    if constexpr (D == 3 && N == 8 && MF == MF_NEO && !PF) {
      static const bool prof = getenv("AD4SM_K1_PROF") != nullptr;
      if (prof) {  // developer tool: synchronous, prints the per-part cycle split of this launch to stderr
        auto pk = k1_u_kernel<D, N, LP, MF, PF, true>;
        if ((err = cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max)) != cudaSuccess) return err;
        pk<<<(unsigned)c.grid, c.warps * 32, smem, s>>>(a, L, c.warps);
        if ((err = cudaStreamSynchronize(s)) != cudaSuccess) return err;
        static long long h[148 * 8 * 16];
        cudaMemcpyFromSymbol(h, g_k1_prof, sizeof(h));
        double sum[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const int nw = c.grid * c.warps;
        for (int w = 0; w < nw && w < 148 * 8; w++)
          for (int k = 0; k < 12; k++) sum[k] += (double)h[w * 16 + k];
        fprintf(stderr, "K1PROF warps/cta %d iters/warp %.1f cycles/iter: head %.0f phase1 %.0f contraction %.0f epilogue %.0f rest %.0f\n",
                c.warps, sum[4] / nw, sum[0] / sum[4], sum[1] / sum[4], sum[2] / sum[4], sum[5] / sum[4], sum[3] / sum[4]);
        fprintf(stderr, "K1PROF head split: TMA issue %.0f wait gradN %.0f wait ids %.0f wait u gather %.0f wait stores %.0f\n",
                sum[6] / sum[4], sum[7] / sum[4], sum[8] / sum[4], sum[9] / sum[4], sum[10] / sum[4]);
        return cudaSuccess;
      }
    }
    kern<<<(unsigned)c.grid, c.warps * 32, smem, s>>>(a, L, c.warps);
    return cudaGetLastError();
  }
}

template <int D, int N, int LP> static cudaError_t launch_u_mf(const K1Args& a, bool pf, cudaStream_t s) {
  switch (a.mf) {
    case MF_HOOKE: return pf ? launch_u_t<D, N, LP, MF_HOOKE, true>(a, s) : launch_u_t<D, N, LP, MF_HOOKE, false>(a, s);
    case MF_NEO: return launch_u_t<D, N, LP, MF_NEO, false>(a, s);
    case MF_MOONEY: return launch_u_t<D, N, LP, MF_MOONEY, false>(a, s);
    case MF_PFNEO: return launch_u_t<D, N, LP, MF_PFNEO, true>(a, s);
    case MF_OGDEN: return launch_u_t<D, N, LP, MF_OGDEN, false>(a, s);
    default: return cudaErrorInvalidValue;
  }
}

template <int D, int N, int LP> static cudaError_t launch_d_t(const K1Args& a, cudaStream_t s) {
  if constexpr (LP == 1) {
    const unsigned grid = (unsigned)((a.ne + K1_THREADS - 1) / K1_THREADS);
    k1_d_kernel_t1<D, N><<<grid, K1_THREADS, 0, s>>>(a);
    return cudaGetLastError();
  } else {
    constexpr int EPC = K1_THREADS / LP;
    const SmemLayout L = k1_layout(GpRecD<D>::SIZE, D, N, a.P, true);
    const size_t smem = (size_t)EPC * L.estride * sizeof(double);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    auto kern = k1_d_kernel<D, N, LP>;
    if (smem > 48 * 1024) {
      cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return err;
    }
    const unsigned grid = (unsigned)((a.ne + EPC - 1) / EPC);
    kern<<<grid, K1_THREADS, smem, s>>>(a, L);
    return cudaGetLastError();
  }
}

// mode: MODE_ELASTIC / MODE_PF_U (u-dual) or MODE_PF_D (d-dual); a.want_K selects ϕ,r-only evaluation.
cudaError_t launch_k1(const K1Args& a, int mode, cudaStream_t s) {
  if (a.ne == 0) return cudaSuccess;
  int LP;
  if (!lanes_for(a.D, a.N, a.P, LP)) return cudaErrorInvalidValue;
  const bool is_pf = a.mat.kind == MAT_PHASEFIELD;
  if ((mode == MODE_ELASTIC) == is_pf) return cudaErrorInvalidValue;  // PhaseField needs d; elastic has none
  if (a.cas) return mode == MODE_ELASTIC ? launch_cas(a, s) : cudaErrorInvalidValue;
  const bool pf = (mode == MODE_PF_U);
#define SHAPE(DD, NN, LL)                 \
  if (a.D == DD && a.N == NN && LP == LL) \
    return mode == MODE_PF_D ? launch_d_t<DD, NN, LL>(a, s) : launch_u_mf<DD, NN, LL>(a, pf, s);
  SHAPE(2, 3, 1) SHAPE(2, 3, 4) SHAPE(2, 4, 4) SHAPE(3, 4, 1) SHAPE(3, 4, 4) SHAPE(3, 6, 8) SHAPE(3, 8, 8)
#undef SHAPE
  return cudaErrorInvalidValue;
}

// ---- K2: Kt values.  One thread per node-pair block (row node i, neighbour p); consecutive threads (p) write
// consecutive memory in each row.
template <int DB> __global__ void __launch_bounds__(256) k2_kernel(const AsmArgs a) {
  const long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= a.n_blocks) return;
  const int nz = assemble_block<DB, false>(a, blk);
  if (nz) atomicAdd(a.n_zero, (unsigned long long)nz);  // integer atomic: order-independent
}
// ---- K3: residual.  One thread per (node, component).
__global__ void __launch_bounds__(256) k3_kernel(const AsmArgs a) {
  const int dpn = a.dpn;
  const long long g = a.node0 * dpn + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.node1 * dpn) return;
  const long long i = g / dpn;
  assemble_r_entry<false>(a, i, (int)(g - i * dpn));
}

cudaError_t launch_k2(const AsmArgs& a, cudaStream_t s) {
  cudaError_t err = cudaMemsetAsync(a.n_zero, 0, sizeof(unsigned long long), s);
  if (err != cudaSuccess) return err;
  const long long nt = a.n_blocks;
  if (nt > 0) {
    const unsigned grid = (unsigned)((nt + 255) / 256);
    switch (a.DB) {
      case 1: k2_kernel<1><<<grid, 256, 0, s>>>(a); break;
      case 2: k2_kernel<2><<<grid, 256, 0, s>>>(a); break;
      case 3: k2_kernel<3><<<grid, 256, 0, s>>>(a); break;
      default: return cudaErrorInvalidValue;
    }
  }
  return cudaGetLastError();
}
// ---- K2s: Kt values from the destination-sorted staging.  One thread per unordered node pair {i <= p}: its
// contributions are contiguous in C and are summed in `elems` order like K2; the sum is written to block (i, p) and,
// transposed, to block (p, i) -- both triangles carry the same bits.
// What bounds it (measured, 1 M Hex8): not DRAM bytes but L1 wavefronts = instructions x distinct 128-byte lines.  With
// per-thread 16-byte loads every instruction touches 32 lines (reads alone 0.64 ms); so the 32 pairs of a warp fetch
// their one contiguous range of C (about 6.6 KB on a hex mesh) with a single TMA bulk copy into the warp's
// shared-memory slice -- no LSU work for the global reads, 0.50 ms = 6.5 TB/s read-only -- and each lane folds its
// blocks out of shared memory.  The 24-byte runs of the CSC output go out as one 16-byte and one 8-byte store
// (k2s_store3: 12 store instructions per pair instead of 18, -0.16 ms).  Left: the writes cost 0.22 ms (upper) and
// 0.30 ms (transposed copy, 24-byte pieces scattered over 13 columns) per GB where DRAM alone would take 0.16.
// Alternatives measured, all bitwise equal: streaming / write-through store hints (no change), a CTA-cooperative
// shared-memory stream with LDG (1.58 ms), five threads per pair (1.38 ms), smaller slices / more warps (slower).
#define AD4SM_K2S_MINB 4
constexpr int K2S_THREADS = 256;
constexpr int K2T_WARPS = 4, K2T_BLOCKS = 128;  // staged kernel: warps per CTA, blocks per warp slice (10 KB)
// acc += block, in canonical orientation.  The 10th double of a contribution block says whether the producer stored
// it transposed (K1 stores its blocks as computed and leaves the transposition to the memory-bound consumer).
__device__ __forceinline__ void k2s_add(double* acc, const double2* B) {
  const double2 v0 = B[0], v1 = B[1], v2 = B[2], v3 = B[3], v4 = B[4];
  const bool tr = v4.y != 0.0;
  // stored entries e = 3j + l: (v0.x v0.y v1.x | v1.y v2.x v2.y | v3.x v3.y v4.x)
  acc[0] += v0.x;
  acc[1] += tr ? v1.y : v0.y;
  acc[2] += tr ? v3.x : v1.x;
  acc[3] += tr ? v0.y : v1.y;
  acc[4] += v2.x;
  acc[5] += tr ? v3.y : v2.y;
  acc[6] += tr ? v1.x : v3.x;
  acc[7] += tr ? v2.y : v3.y;
  acc[8] += v4.x;
}
// three consecutive doubles at an 8-byte-aligned address as one 16-byte and one 8-byte store (lane-dependent split, no
// divergence)
__device__ __forceinline__ void k2s_store3(double* p, double v0, double v1, double v2) {
  const bool odd = (reinterpret_cast<unsigned long long>(p) & 8ull) != 0;
  *reinterpret_cast<double2*>(p + (odd ? 1 : 0)) = odd ? make_double2(v1, v2) : make_double2(v0, v1);
  p[odd ? 0 : 2] = odd ? v0 : v2;
}
// Block (i, p) and its transpose into the CSC values.  EXP is a TIMING EXPERIMENT switch (AD4SM_EXP_K2S; the results
// are wrong): 1 leaves the lower triangle unwritten, 2 writes nothing.
template <int EXP>
__device__ __forceinline__ void k2s_write(const AsmArgs& a, const double* acc, unsigned oup, unsigned olo, unsigned nn) {
  const long long su = 3ll * (nn & 0xffffu), sl = 3ll * (nn >> 16);  // column lengths of the two destinations
  if (EXP == 2) {  // keep the loads alive
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 9; j++) t += acc[j];
    if (t == 1.2345e300) a.nzval[oup] = t;
    return;
  }
  int nz = 0;
#pragma unroll
  for (int j = 0; j < 9; j++) nz += (acc[j] == 0.0);
#pragma unroll
  for (int j = 0; j < 3; j++) k2s_store3(a.nzval + oup + j * su, acc[j * 3], acc[j * 3 + 1], acc[j * 3 + 2]);
  if (olo != oup && EXP != 1) {
#pragma unroll
    for (int l = 0; l < 3; l++) k2s_store3(a.nzval + olo + l * sl, acc[l], acc[3 + l], acc[6 + l]);
    nz *= 2;
  }
  if (nz) atomicAdd(a.n_zero, (unsigned long long)nz);  // integer atomic: order-independent
}
// direct-load variant (AD4SM_K2S_DIRECT=1; also what ranges too large for a slice fall back to, per warp)
template <int EXP> __global__ void __launch_bounds__(K2S_THREADS, AD4SM_K2S_MINB) k2s_kernel3(const AsmArgs a) {
  const long long u = a.pair0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= a.pair1) return;
  const unsigned c0 = __ldg(a.cptr + u), c1 = __ldg(a.cptr + u + 1);
  const unsigned oup = __ldg(a.out_up + u), olo = __ldg(a.out_lo + u), nn = __ldg(a.nn_pack + u);
  double acc[9];
#pragma unroll
  for (int t = 0; t < 9; t++) acc[t] = 0.0;
  const double2* C2 = reinterpret_cast<const double2*>(a.C);
  unsigned c = c0;
  for (; c + 1 < c1; c += 2) {  // two blocks in flight; still folded in ascending (`elems`) order
    const double2* B = C2 + (long long)c * (CBLK3 / 2);
    double2 w[10];
#pragma unroll
    for (int t = 0; t < 10; t++) w[t] = B[t];
    k2s_add(acc, w);
    k2s_add(acc, w + 5);
  }
  if (c < c1) k2s_add(acc, C2 + (long long)c * (CBLK3 / 2));
  k2s_write<EXP>(a, acc, oup, olo, nn);
}
template <int EXP> __global__ void __launch_bounds__(K2T_WARPS * 32) k2s_tma_kernel3(const AsmArgs a) {
  __shared__ __align__(128) double2 buf[K2T_WARPS][K2T_BLOCKS * (CBLK3 / 2)];
  __shared__ __align__(8) unsigned long long bars[K2T_WARPS];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long u = a.pair0 + ((long long)blockIdx.x * K2T_WARPS + w) * 32 + lane;
  if (u - lane >= a.pair1) return;  // whole warp; no CTA-wide barrier below
  const bool valid = u < a.pair1;
  const unsigned c0 = __ldg(a.cptr + (valid ? u : a.pair1)), c1 = valid ? __ldg(a.cptr + u + 1) : c0;
  const unsigned bar = smem_u32(&bars[w]);
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  const long long um = valid ? u : 0;  // destinations: fetched while the copies fly
  const unsigned oup = __ldg(a.out_up + um), olo = __ldg(a.out_lo + um), nn = __ldg(a.nn_pack + um);
  double acc[9];
#pragma unroll
  for (int t = 0; t < 9; t++) acc[t] = 0.0;
  __syncwarp();
  // The warp's pairs own one contiguous range of C.  It is brought in by TMA bulk copies of at most K2T_BLOCKS blocks:
  // one copy on a hex mesh (~82 blocks per 32 pairs), several on meshes of higher valence (Kuhn tets: ~240), each
  // covering a run of consecutive lanes; a single pair with more contributions than a slice holds folds from global
  // memory directly.
  unsigned phase = 0;
  for (int s = 0; s < 32;) {
    const unsigned base = __shfl_sync(0xffffffffu, c0, s);
    const unsigned fits = __ballot_sync(0xffffffffu, lane >= s && (c1 - base) <= (unsigned)K2T_BLOCKS);
    const unsigned stop = ~(fits >> s);                      // first lane at or after s that does not fit
    const int cnt = stop ? __ffs((int)stop) - 1 : 32 - s;
    if (cnt == 0) {
      if (lane == s) {
        const double2* C2 = reinterpret_cast<const double2*>(a.C);
        for (unsigned c = c0; c < c1; c++) k2s_add(acc, C2 + (long long)c * (CBLK3 / 2));
      }
      s += 1;
      continue;
    }
    const unsigned nb = __shfl_sync(0xffffffffu, c1, s + cnt - 1) - base;
    if (nb > 0) {
      if (lane == 0) {
        fence_proxy_async();  // the previous chunk's reads of the slice precede this copy's writes
        mbar_expect_tx(bar, nb * (unsigned)(CBLK3 * sizeof(double)));
        tma_load_1d(smem_u32(buf[w]), a.C + (long long)base * CBLK3, nb * (unsigned)(CBLK3 * sizeof(double)), bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1u;
      if (lane >= s && lane < s + cnt) {
        const double2* B = buf[w] + (c0 - base) * (CBLK3 / 2);
        for (unsigned c = c0; c < c1; c++, B += CBLK3 / 2) k2s_add(acc, B);
      }
      __syncwarp();
    }
    s += cnt;
  }
  if (valid) k2s_write<EXP>(a, acc, oup, olo, nn);
}
cudaError_t launch_k2s(const AsmArgs& a, cudaStream_t s) {
  if (a.DB != 3 || a.dpn != 3 || a.n_pairs <= 0) return cudaErrorInvalidValue;
  if (!a.keep_zero_count) {
    cudaError_t err = cudaMemsetAsync(a.n_zero, 0, sizeof(unsigned long long), s);
    if (err != cudaSuccess) return err;
  }
  const long long np = a.pair1 - a.pair0;  // the launch's pair range (the whole pattern unless the assembly is split)
  if (np <= 0) return cudaSuccess;
  if (a.k2s_direct) {
    const unsigned grid = (unsigned)((np + K2S_THREADS - 1) / K2S_THREADS);
    switch (a.k2s_exp) {
      case 1: k2s_kernel3<1><<<grid, K2S_THREADS, 0, s>>>(a); break;
      case 2: k2s_kernel3<2><<<grid, K2S_THREADS, 0, s>>>(a); break;
      default: k2s_kernel3<0><<<grid, K2S_THREADS, 0, s>>>(a); break;
    }
  } else {
    const unsigned grid = (unsigned)((np + K2T_WARPS * 32 - 1) / (K2T_WARPS * 32));
    switch (a.k2s_exp) {
      case 1: k2s_tma_kernel3<1><<<grid, K2T_WARPS * 32, 0, s>>>(a); break;
      case 2: k2s_tma_kernel3<2><<<grid, K2T_WARPS * 32, 0, s>>>(a); break;
      default: k2s_tma_kernel3<0><<<grid, K2T_WARPS * 32, 0, s>>>(a); break;
    }
  }
  return cudaGetLastError();
}
// ---- partitioned runs with destination-sorted staging: the HALO elements' blocks arrive element-major (exchange /
// peer stores) and are moved to their sorted slots, transposed where the canonical orientation asks for it.
__global__ void __launch_bounds__(256) halo_scatter_kernel3(const double* Ke, const unsigned* cpos, double* C, long long slot0,
                                                            long long n_slots, int NPB) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_slots * NPB) return;
  const long long g = slot0 * NPB + t;
  const unsigned code = __ldg(cpos + g);
  const double* B = Ke + g * 9;
  double v[9];
#pragma unroll
  for (int i = 0; i < 9; i++) v[i] = B[i];
  double* dst = C + (long long)(code >> 1) * CBLK3;
#pragma unroll
  for (int i = 0; i < 9; i++) dst[i] = v[i];
  dst[9] = (code & 1u) ? 1.0 : 0.0;
}
cudaError_t launch_halo_scatter(const double* Ke, const unsigned* cpos, double* C, long long slot0, long long n_slots, int NPB,
                                cudaStream_t s) {
  const long long n = n_slots * NPB;
  if (n > 0) halo_scatter_kernel3<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(Ke, cpos, C, slot0, n_slots, NPB);
  return cudaGetLastError();
}
// ---- assembly-only entry (makeϕrKt(Φ, elems, u), elements.toolkit.jl:169-203): the caller's element duals -> staging.
// One thread per (slot, node-pair block): block (q, o) of the cyclic storage holds K((q+o)%N, j),(q, l); the Hessian is
// adiff's packed upper triangle, entry (i <= j), 0-based, at j(j+1)/2 + i (adiff.jl:26,84), local dof = comp + D*node.
__global__ void __launch_bounds__(256) duals_to_staging_kernel(const double* Phi, long long stride, const long long* index,
                                                               long long n_slots, int N, int D, double* Ke, double* re,
                                                               double* phie) {
  const int NPB = N * (N + 1) / 2, DD = D * D, n = N * D;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_slots * NPB) return;
  const long long slot = t / NPB;
  const int blk = (int)(t - slot * NPB), o = blk / N, q = blk - o * N;
  const double* rec = Phi + (index ? index[slot] : slot) * stride;
  const double* h = rec + 1 + n;
  const int a = (q + o) % N;
  if (Ke)  // nullptr: residual only (makeϕr, records may be D1: no h)
    for (int j = 0; j < D; j++)
      for (int l = 0; l < D; l++) {
        const int pr = a * D + j, pc = q * D + l, hi = pr > pc ? pr : pc, lo = pr > pc ? pc : pr;
        Ke[t * DD + j * D + l] = h[hi * (hi + 1) / 2 + lo];
      }
  if (blk < N)  // this thread also moves node blk's gradient entries, thread 0 the value
    for (int j = 0; j < D; j++) re[slot * n + blk * D + j] = rec[1 + blk * D + j];
  if (blk == 0) phie[slot] = rec[0];
}
cudaError_t launch_duals_to_staging(const double* Phi, long long stride, const long long* index, long long n_slots, int N, int D,
                                    double* Ke, double* re, double* phie, cudaStream_t s) {
  const long long n = n_slots * (N * (N + 1) / 2);
  if (n > 0) duals_to_staging_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(Phi, stride, index, n_slots, N, D, Ke, re, phie);
  return cudaGetLastError();
}
cudaError_t launch_k3(const AsmArgs& a, cudaStream_t s) {
  const long long nr = (a.node1 - a.node0) * a.dpn;
  if (nr > 0) k3_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}

// ---- ϕ = Σ_e ϕ_e: fixed two-level tree (chunk per block, then one block), run-to-run reproducible --------
constexpr int PHI_PARTIALS = 1024;
int phi_reduce_partials() { return PHI_PARTIALS; }

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  return sh[0];
}
__global__ void __launch_bounds__(256) phi_partial_kernel(const double* __restrict__ x, long long n, double* partials) {
  __shared__ double sh[256];
  const long long chunk = (n + PHI_PARTIALS - 1) / PHI_PARTIALS;
  const long long b0 = blockIdx.x * chunk;
  long long b1 = b0 + chunk;
  if (b1 > n) b1 = n;
  double v = 0.0;
  for (long long i = b0 + threadIdx.x; i < b1; i += 256) v += x[i];
  const double t = block_sum_256(v, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}
__global__ void __launch_bounds__(256) phi_final_kernel(const double* __restrict__ partials, double* phi) {
  __shared__ double sh[256];
  double v = 0.0;
  for (int i = threadIdx.x; i < PHI_PARTIALS; i += 256) v += partials[i];
  const double t = block_sum_256(v, sh);
  if (threadIdx.x == 0) *phi = t;
}
cudaError_t launch_phi_reduce(const double* phie, long long n, double* partials, double* phi, cudaStream_t s) {
  phi_partial_kernel<<<PHI_PARTIALS, 256, 0, s>>>(phie, n, partials);
  phi_final_kernel<<<1, 256, 0, s>>>(partials, phi);
  return cudaGetLastError();
}

// ---- structural CSC pattern (Int64, 1-based) of the assembled field ----------------------------------------
__global__ void __launch_bounds__(256) pattern_kernel(const AsmArgs a, long long* colptr, long long* rowval) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int dpn = a.dpn;
  if (g > a.n_nodes * dpn) return;
  if (g == a.n_nodes * dpn) {
    colptr[g] = a.n_blocks * dpn * dpn + 1;
    return;
  }
  const long long i = g / dpn;
  const int j = (int)(g - i * dpn);
  const long long nb0 = a.nbr_ptr[i];
  const int nn = (int)(a.nbr_ptr[i + 1] - nb0);
  const long long base = nb0 * (dpn * dpn) + (long long)j * dpn * nn;
  colptr[g] = base + 1;
  for (int p = 0; p < nn; p++) {
    const long long c = (long long)a.nbr[nb0 + p] * dpn;
    for (int l = 0; l < dpn; l++) rowval[base + p * dpn + l] = c + l + 1;
  }
}
cudaError_t launch_pattern(const AsmArgs& a, long long* colptr, long long* rowval, cudaStream_t s) {
  const long long n = a.n_nodes * a.dpn + 1;
  pattern_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, colptr, rowval);
  return cudaGetLastError();
}

// ---- dropzeros (toolkit:202): per-column non-zero counts -> exclusive scan -> compaction -----------------
__global__ void __launch_bounds__(256) dz_count_kernel(long long n_dof, const long long* colptr, const double* nzval,
                                                       long long* cnt) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_dof) return;
  long long k = 0;
  if (c < n_dof)
    for (long long s = colptr[c] - 1; s < colptr[c + 1] - 1; s++) k += (nzval[s] != 0.0);
  cnt[c] = k;
}
__global__ void __launch_bounds__(256) dz_fill_kernel(long long n_dof, const long long* colptr, const long long* rowval,
                                                      const double* nzval, long long* colptr_out, long long* rowval_out,
                                                      double* nzval_out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_dof) return;
  const long long o0 = colptr_out[c];  // exclusive scan of the counts (0-based)
  if (c < n_dof) {
    long long o = o0;
    for (long long s = colptr[c] - 1; s < colptr[c + 1] - 1; s++) {
      const double v = nzval[s];
      if (v != 0.0) {
        rowval_out[o] = rowval[s];
        nzval_out[o] = v;
        o++;
      }
    }
  }
}
__global__ void __launch_bounds__(256) dz_onebase_kernel(long long n, long long* colptr_out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) colptr_out[c] += 1;
}
size_t dropzeros_tmp_bytes(long long n_dof) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (long long*)nullptr, (long long*)nullptr, n_dof + 1);
  return bytes;
}
cudaError_t launch_dropzeros(long long n_dof, const long long* colptr, const long long* rowval, const double* nzval,
                             long long* colptr_out, long long* rowval_out, double* nzval_out, void* cub_tmp,
                             size_t cub_tmp_bytes, cudaStream_t s) {
  const unsigned grid = (unsigned)((n_dof + 1 + 255) / 256);
  // counts are staged in colptr_out's tail-compatible scratch: use rowval_out? no -- scan in place on colptr_out
  dz_count_kernel<<<grid, 256, 0, s>>>(n_dof, colptr, nzval, colptr_out);
  cudaError_t err = cub::DeviceScan::ExclusiveSum(cub_tmp, cub_tmp_bytes, colptr_out, colptr_out, n_dof + 1, s);
  if (err != cudaSuccess) return err;
  dz_fill_kernel<<<grid, 256, 0, s>>>(n_dof, colptr, rowval, nzval, colptr_out, rowval_out, nzval_out);
  dz_onebase_kernel<<<grid, 256, 0, s>>>(n_dof + 1, colptr_out);
  return cudaGetLastError();
}

// ---- FP64 FMA peak (roofline denominator for K1; MEASURED_PEAKS.json has no FP64 entry) --------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
cudaError_t launch_fp64_peak(double* tflops, double* mhz) {
  int dev = 0, sms = 0, khz = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  err = cudaMalloc(&out, sizeof(double) * blocks * threads);
  if (err != cudaSuccess) return err;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {  // first reps warm up clocks
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    err = cudaEventSynchronize(e1);
    if (err != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep >= 2 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (err != cudaSuccess) return err;
  *tflops = best;
  *mhz = khz / 1000.0;
  return cudaSuccess;
}

}  // namespace ad4sm
