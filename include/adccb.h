/* adccb.h - C ABI of the B200 (sm_100a) tensor backend that replaces adcc's libtensor
 * TensorImpl path for the ADC matvec + Davidson hot path.
 *
 * Conventions
 *   - every entry point returns 0 on success, non-zero on failure; the message of the last
 *     failure on the calling thread is returned by adccb_last_error() (the Python shim turns
 *     it into ValueError / RuntimeError as libadcc_src/pyiface/ExportAdcc.cc:63-71 does);
 *   - all pointers are DEVICE pointers to FP64 data unless a parameter is named host_*;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 *     asynchronous with respect to the host; nothing is owned or freed by the library;
 *   - tensors are dense, row-major; extents/strides are counted in elements (int64).
 *
 * Each function cites the reference interface it replaces (paths relative to the adcc
 * source tree, v0.18.0).
 */
#ifndef ADCCB_H
#define ADCCB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int adccb_version(void);
const char* adccb_last_error(void);
/* number of kernels launched by this library in the current process (bench: gpu_launches) */
long long adccb_launch_count(void);
/* SM count and compute capability of the current device */
int adccb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* K1 contraction core.  C[m,n] = alpha * sum_k A(m,k) * B(n,k) + beta * C[m,n],
 * A(m,k) = A[m*a_rs + k*a_cs], B(n,k) = B[n*b_rs + k*b_cs], C row-major with leading dim ldc.
 * K-contiguous, 16-byte aligned operands (a_cs == b_cs == 1, even row strides) take the TMA + DMMA
 * path; so does an N-major B (b_rs == 1, even b_cs: B stored as B[k][n], i.e. C = A . B without a
 * transposed copy - both ovvv couplings of adcc/adc_pp/matrix.py:270-276, 288-294 read ONE stored
 * <ja||bc> layout this way, the "a10 + a11 from one ovvv layout" of SURVEY 8(b)); everything else the
 * generic DMMA path.  `ws` is scratch for split-K partials (may be NULL: no split-K).
 * tile_hint: 0 = automatic, 112 / 128 = force the row-tile height.
 * Replaces libadcc::TensorImpl<N>::tensordot -> libtensor::expr::contract + eval
 * (libadcc_src/TensorImpl.cc:856-876, 901-1117; libadcc_src/Tensor.hh:189-196). */
int adccb_dgemm(void* stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                int64_t a_rs, int64_t a_cs, const double* B, int64_t b_rs, int64_t b_cs,
                double beta, double* C, int64_t ldc, double* ws, int64_t ws_bytes, int tile_hint);

/* Column-slab GEMM fused with the exchange that follows it in a sharded matvec:
 * C[m,n] = alpha * sum_k A[m*lda + k] * B[n*ldb + k] is written to the local C AND, tile by tile
 * from the same epilogue, to the same offsets of n_peers (<= 7) buffers that live in PEER GPUs'
 * memory (host array `host_peer_C` of device pointers obtained from adccb_peer_open).  Every rank
 * of a sharded sigma build calls this with its own column slab (C and the peer pointers already
 * offset to the slab's first column, common leading dimension ldc); after a stream-ordered
 * cross-rank synchronisation (any collective on the same stream) all ranks hold all slabs.  It
 * replaces "GEMM, then all-gather of the slabs" (SURVEY 8(e) fusion target); in the reference
 * the corresponding step is the shared-memory libtensor contraction writing one result tensor
 * (libadcc_src/TensorImpl.cc:901-1117).  A and B must be K-contiguous and 16-byte aligned. */
int adccb_dgemm_peers(void* stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                      int64_t lda, const double* B, int64_t ldb, double* C, int n_peers,
                      double* const* host_peer_C, int64_t ldc, double* ws, int64_t ws_bytes,
                      int tile_hint);
/* Peer-visible device memory for adccb_dgemm_peers (one process per GPU, CUDA IPC):
 * adccb_peer_alloc allocates `bytes` on the current device and writes the 64-byte IPC handle the
 * host side passes to the other ranks (any transport, e.g. torch.distributed all_gather_object);
 * adccb_peer_open maps a peer's buffer into this process and enables peer access to its device;
 * adccb_peer_close unmaps it; adccb_peer_free releases an own allocation.  These are the only
 * entry points that own memory, because IPC export needs a dedicated allocation. */
int adccb_peer_alloc(int64_t bytes, void** dev_ptr, void* host_handle64);
int adccb_peer_open(const void* host_handle64, void** dev_ptr);
int adccb_peer_close(void* dev_ptr);
int adccb_peer_free(void* dev_ptr);

/* out[o] = alpha * in[sum_d o_d * in_stride[d]] + gamma * add[o] + beta * out[o] for a
 * contiguous `out` of the given shape (rank <= 8); `add` may be NULL.
 * Covers transpose / copy / diagonal extraction / (anti)symmetrisation:
 * libadcc_src/TensorImpl.cc:681-717 (transpose), :667-678 (copy), :489-565 (diagonal),
 * :1147-1264 (symmetrise / antisymmetrise; libadcc_src/Tensor.hh:218-244). */
int adccb_strided_gather(void* stream, int ndim, const int64_t* shape, const int64_t* in_stride,
                         double alpha, const double* in, double gamma, const double* add,
                         double beta, double* out);

/* out = alpha * f(x, y) + beta * out over n elements; op: 0 x, 1 x*y, 2 x/y, 3 c (fill),
 * 4 x/(y - c) (Jacobi preconditioner, adcc/solver/preconditioner.py:65-82), 5 x + c.
 * Replaces scale/add/multiply/divide/set_mask: libadcc_src/TensorImpl.cc:568-664, :431-462. */
int adccb_elementwise(void* stream, int op, int64_t n, double alpha, const double* x,
                      const double* y, double c, double beta, double* out);

/* out[p*nb + q] = sa * a[p] + sb * b[q].  libadcc_src/TensorImpl.cc:741-789 (direct_sum). */
int adccb_direct_sum(void* stream, int64_t na, int64_t nb, double sa, const double* a, double sb,
                     const double* b, double* out);

/* S[i, col0 + c] += G[i, c] for i < nrows, c < ncols (row-major, leading dims lds / ldg): merges the
 * column slab a peer rank computed into the local sigma vector after the all-gather. */
int adccb_add_columns(void* stream, int64_t nrows, int64_t ncols, double* S, int64_t lds, int64_t col0,
                      const double* G, int64_t ldg);

/* out = beta * out + sum_k host_coeff[k] * host_ptrs[k][:] (one pass over all inputs).
 * libadcc_src/TensorImpl.cc:604-624 (add_linear_combination) reached from
 * libadcc_src/pyiface/export_Tensor.cc:260-277 (linear_combination_strict). */
int adccb_lincomb(void* stream, int nvec, const double* host_coeff, const double* const* host_ptrs,
                  int64_t n, double beta, double* out);

/* host_outs[j][:] = beta * host_outs[j][:] + sum_k host_coeff[j*nvec + k] * host_ptrs[k][:], j < nout:
 * nout linear combinations of the SAME nvec inputs in one pass over the inputs (each output summed
 * exactly as adccb_lincomb would).  The Davidson loop forms its n_block residuals, Ritz vectors and the
 * collapsed subspace this way (adcc/solver/davidson.py:139-140, 151, 168-170, where the reference
 * evaluates one libtensor linear combination per output). */
int adccb_lincomb_multi(void* stream, int nvec, int nout, const double* host_coeff,
                        const double* const* host_ptrs, int64_t n, double beta, double* const* host_outs);

/* out_dev[j] = beta * out_dev[j] + scale * <x, y_j>, j < nvec, in one pass over x and every
 * y_j with a fixed-order two-stage reduction.  ws: >= 32*4736*8 bytes of device scratch.
 * libadcc_src/TensorImpl.cc:1120-1144 (dot with a list), libadcc_src/Tensor.hh:198-206. */
int adccb_dot_many(void* stream, int nvec, const double* x, const double* const* host_ys, int64_t n,
                   double scale, double beta, double* out_dev, double* ws, int64_t ws_bytes);

/* out[p,q,r,s] = T[p',q',r',s'] * [spin(p)==spin(q)] * [spin(r)==spin(s)]: expands a spatial chemists'
 * (pq|rs) block of a restricted reference (extents n_spatial[4]) to spin orbitals, every output axis
 * being [alpha | beta] over the same spatial orbitals.  Device counterpart of the spin pattern
 * adcc/backends/EriBuilder.py:117-119 applies on the host during ReferenceState::eri import. */
int adccb_spin_expand4(void* stream, const int* n_spatial, const double* T, double* out);

/* Spin-block symmetry of restricted singlet (fac=+1) / triplet (fac=-1) vectors on dense
 * [alpha|beta]^rank tensors (rank 2 or 4): out = (T + fac * flip(T)) / 2, spin-changing blocks
 * zeroed.  Dense counterpart of the spin-block maps adcc/guess/guess_zero.py:121-161 that
 * libtensor keeps structurally (libadcc_src/TensorImpl/as_lt_symmetry.cc:100-172). */
int adccb_spin_project(void* stream, int ndim, const int* shape, const int* n_alpha, double fac,
                       const double* in, double* out);

/* In-place singlet purification of a doubles tensor T[i,j,a,b]:
 * T[aaaa] = T[abab] + T[abba], T[bbbb] = T[aaaa].
 * libadcc_src/amplitude_vector_enforce_spin_kind.cc:33-170. */
int adccb_spin_singlet(void* stream, const int* shape, const int* n_alpha, double* T);

/* Fused Davidson correction step on a dense rank-2 (ph) or rank-4 (pphh) block:
 *   out = Purify( Sym( SpinProject( x / (D - shift) ) ) )
 * D may be NULL (no preconditioning); spin_fac +1 singlet / -1 triplet / 0 none (adcc/guess/guess_zero.py:
 * 121-161 spin-block maps, as adccb_spin_project); sym_mode 0 none, 1 antisymmetrise(0,1), antisymmetrise(2,3),
 * symmetrise((0,1),(2,3)) (adcc/AdcMatrix.py:449-455), 2 antisymmetrise(2,3) (CVS, AdcMatrix.py:445-447);
 * purify != 0: singlet purification of the aaaa / bbbb blocks (libadcc_src/
 * amplitude_vector_enforce_spin_kind.cc:33-170).  The Python host divides first (adccb_elementwise op 4,
 * coalesced) and calls this with D == NULL: the gathered form of the division measured slower.  Replaces the chain
 * adcc/solver/preconditioner.py:65-82 -> adcc/solver/explicit_symmetrisation.py:49-98 the reference runs as
 * separate libtensor evaluations per new subspace vector (davidson.py:255-272). */
int adccb_precond_apply(void* stream, int ndim, const int* shape, const int* n_alpha, const double* x,
                        const double* D, double shift, int sym_mode, double spin_fac, int purify, double* out);

/* ---- antisymmetry-packed doubles store: U[IJ,AB], IJ = j(j-1)/2+i (i<j), AB = b(b-1)/2+a (a<b).
 * The doubles part of every ADC vector is antisymmetric in (i,j) and (a,b)
 * (adcc/guess/guess_zero.py:163-171); libtensor stores canonical blocks only, this store keeps
 * canonical elements only.  no / nv = number of occupied / virtual spin orbitals. */

/* out[i,j,a,b] (dense) from packed U, with signs and zero diagonals (Tensor::export_to,
 * libadcc_src/TensorImpl.cc:1369-1376). */
int adccb_packed_unpack(void* stream, int no, int nv, const double* U, double* out);
/* U[IJ,AB] = 1/4 (T[ijab] - T[jiab] - T[ijba] + T[jiba]) from dense T: import + the doubles
 * symmetrisation of adcc/AdcMatrix.py:449-455 in one pass. */
int adccb_packed_pack(void* stream, int no, int nv, const double* T, double* U);
/* GEMM-ready expansions of packed U: which = 0: X[i,a,k,c] = u[ikac]; 1: Ue[i,j,AB] = u[ij,AB];
 * 2: Uh[a,JK,b] = u[JK,ab] (no < 0: U holds a slab of -no JK rows). */
int adccb_packed_expand(void* stream, int which, int no, int nv, const double* U, double* out);
/* Slab forms of adccb_packed_expand for a matvec sharded over GPUs: which = 0: rows (il, a) of X for
 * the occupied slab i = first + il, il < count; which = 1: Ue[i, jl, AB] for the slab j = first + jl
 * of the second occupied index; which = 2: Uh[a, r, b] for `count` packed rows starting at U. */
int adccb_packed_expand_slab(void* stream, int which, int no, int nv, int first, int count,
                              const double* U, double* out);
/* Distribution of packed doubles vectors over `world` GPUs by virtual-pair (column) slabs: rank r owns
 * columns [host_cut[r], host_cut[r+1]) of the [nrows, npv] array and keeps them as a zero-padded
 * [nrows, wmax] slab.  dir 0: full -> [world, nrows, wmax] (the layout ncclReduceScatter sums the ranks'
 * partial sigma vectors in); dir 1: [world, nrows, wmax] (as all-gathered) -> full.  The reference
 * has no distributed store; this is the layout step of the sigma-vector all-gather / reduce-scatter of
 * SURVEY 8(e). */
int adccb_slab_layout(void* stream, int dir, int64_t nrows, int64_t npv, int world, const int64_t* host_cut,
                      int64_t wmax, double* full, double* slabs);
/* Stream-ordered strided 2-D copy between any two of {pinned host, device} buffers (pitches and width
 * in bytes): a rank's column slab of a vector that lives in shared pinned host memory <-> its device
 * slab (end-to-end form of the sharded matvec; Tensor::import_from / export_to,
 * libadcc_src/TensorImpl.cc:1369-1441, for one slab). */
int adccb_copy2d(void* stream, void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes,
                 int64_t width_bytes, int64_t height);
/* Spin-blocked operands for restricted references.  out[r, c] = in[rows[r]*ld + cols[c]] (gather) and
 * out[rows[r]*ld + cols[c]] += alpha * blk[r, c] (scatter-add); rows / cols: device int64 index lists,
 * NULL = identity.  <pq||rs> vanishes unless bra and ket carry the same spin projection, so the
 * pair-packed vvvv / oooo operands and the ovov operand are block diagonal over spin classes of their
 * composite indices; libtensor holds that structure as spin-block symmetry elements
 * (libadcc_src/TensorImpl/as_lt_symmetry.cc:100-172, adcc/guess/guess_zero.py:143-161), here the classes
 * are gathered into compact GEMM operands and the results scattered back. */
int adccb_block_gather(void* stream, int64_t n_rows, int64_t n_cols, const int64_t* rows, const int64_t* cols,
                       const double* in, int64_t ld, double* out);
int adccb_block_scatter_add(void* stream, int64_t n_rows, int64_t n_cols, const int64_t* rows, const int64_t* cols,
                            double alpha, const double* blk, double* out, int64_t ld);
/* same, assigning instead of accumulating: out[rows[r]*ld + cols[c]] = alpha * blk[r, c] */
int adccb_block_scatter(void* stream, int64_t n_rows, int64_t n_cols, const int64_t* rows, const int64_t* cols,
                        double alpha, const double* blk, double* out, int64_t ld);
/* out[al, b, c, d] = <a c||b d> (a = a_begin + al < a_begin + na) gathered from the pair-packed vvvv
 * block W[AB,CD]: the exchange-ordered operand slab of the o^2 v^4 builds of the ADC(3) singles block
 * (adcc/adc_pp/matrix.py:953 "acbd,cd->ab", :1029-1033 "icjd,acbd->iajb" / "idjc,acbd->iajb"), which on
 * the reference are libtensor contractions over the full vvvv tensor. */
int adccb_vvvv_exchange_slab(void* stream, int nv, int a_begin, int na, const double* W, double* out);
/* out[al, c, j, d] = <j c||a d> (a = a_begin + al) gathered from V2[c,(j,AD)] (the pair-packed ovvv
 * layout of adccb_synth_fill layout 4): operand slab of pi7 = t2[ijbd] <jc||ad>
 * (adcc/GroundState.py:239), which feeds adc3_pib (adcc/adc_pp/matrix.py:1087-1094). */
int adccb_ovvv_exchange_slab(void* stream, int no, int nv, int a_begin, int na, const double* V2, double* out);
/* Contraction of a stack of pair-packed ANTISYMMETRIC nv x nv matrices with vectors over ONE member of
 * the packed pair:  out[o, a] = beta*out[o, a] + alpha * sum_{r < n_red} sum_c A_{rho}[a,c] x_{rho}[c],
 * rho = o*row_stride_out + r*row_stride_red (in rows of nv(nv-1)/2 doubles of V), x_rho = X + (x_by_out ?
 * o : r)*ldx.  With V = V2[a',(j,BC)] = <ja'||bc> it evaluates  <ic||ab> r[cb]  (adcc/adc_pp/matrix.py:498),
 * <kb||cd> u[kd]  (:529) and  <ia||bc> p[ic]  (:951) without unpacking ovvv.  ws: scratch for the
 * per-chunk partial sums (n_chunks * n_out * nv doubles, n_chunks <= 4*SMs/n_out + 1). */
int adccb_pair_matvec(void* stream, int nv, int n_out, int n_red, int64_t row_stride_out, int64_t row_stride_red,
                      const double* V, const double* X, int64_t ldx, int x_by_out, double alpha, double beta,
                      double* out, int64_t ldo, double* ws, int64_t ws_bytes);
/* S[rows[r], AB] += alpha * (F_r[a,b] - F_r[b,a]), F_r[a,b] = P1[off1[r] + a*sa + b] - P2[off2[r] + a*sa + b]
 * (P2/off2 may be NULL; rows NULL = identity); off1/off2/rows: device int64 arrays of length nrows.  Folds the ovov-type
 * and ooov-type GEMM results into packed sigma (the antisymmetrise(2,3) of
 * adcc/adc_pp/matrix.py:241-243, 291-292). */
int adccb_packed_fold_virt(void* stream, int64_t nrows, int nv, int64_t sa, const double* P1,
                           const int64_t* off1, const double* P2, const int64_t* off2,
                           const int64_t* rows, double alpha, double* S);
/* S[pair(i,j), AB] += alpha * (C[i,j,AB] - C[j,i,AB])  (antisymmetrise(0,1), matrix.py:291).
 * C is [no, nj, npv] and holds the slab j = j_begin .. j_begin+nj-1 of the second occupied index
 * (nj = no, j_begin = 0: the full tensor); a slab contributes only the terms it owns, so that the
 * per-rank partial sigma vectors of a sharded matvec add up. */
int adccb_packed_fold_occ(void* stream, int no, int nv, int j_begin, int nj, const double* C,
                          double alpha, double* S);
/* D[IJ,AB] = 1/2 (dov[i,a]+dov[j,b]+dov[i,b]+dov[j,a]) + doo[i,j] + dvv[a,b] (doo, dvv may be
 * NULL): packed diagonal of the pphh-pphh block, adcc/adc_pp/matrix.py:109-124, 212-231. */
int adccb_packed_diagonal(void* stream, int no, int nv, const double* dov, const double* doo,
                          const double* dvv, double* D);
/* out[l, pair(p<q), r] = in[l, p, q, r]: keeps the canonical half of an antisymmetric index pair
 * while staging ERI blocks (ooov first pair, ovvv last pair) into their GEMM layout. */
int adccb_pair_select(void* stream, int64_t L, int n, int64_t R, const double* in, double* out);
/* Synthetic antisymmetrised ERIs of BASELINE config 5 generated in place from a counter-based
 * hash (no host data): layout 0 W[AB,CD] (vvvv, rows row_begin..), 1 O[IJ,KL] (oooo),
 * 2 Y[(j,b),(k,c)] = <kb||jc>, 3 V1[(j,AB),c] = <jc||ab>, 4 V2[a,(j,BC)] = <ja||bc>,
 * 5 G1[i,(JK,b)] = <jk||ib>, 6 G2[(IJ,a),k] = <ij||ka>, 7 dense block blk
 * (0 oooo, 1 ooov, 2 oovv, 3 ovov, 4 ovvv, 5 vvvv), 8 exchange-ordered vvvv slab [al,b,c,d] = <ac||bd>
 * and 9 exchange-ordered ovvv slab [al,c,j,d] = <jc||ad> with a = row_begin + al.  n = number of elements written.
 * Stands in for ReferenceState::eri -> import_eri_* (libadcc_src/ReferenceState.cc:478-509).
 * Slabs for sharding: row_begin = first AB row of layout 0; layouts 2, 3, 4 generate the occupied
 * slab j = j_begin .. j_begin+nj-1 (nj <= 0: all). */
int adccb_synth_fill(void* stream, int layout, int blk, int no, int nv, uint64_t seed, double scale,
                     int64_t n, int64_t row_begin, int j_begin, int nj, double* out);

#ifdef __cplusplus
}
#endif
#endif /* ADCCB_H */
