/* tbk.h — C ABI of libtbk_b200.so, the B200 (sm_100a) k-mesh engine behind the
 * PythTB API.
 *
 * The reference (PythTB 1.8.0, one pure-Python module) has no FFI; its boundary
 * is a set of Python methods.  Each entry point below replaces the arithmetic
 * of the reference function cited next to it (file:line relative to
 * /root/reference).  The Python host layer (pythtb_b200/) keeps the reference's
 * method signatures and calls these through ctypes; INTEGRATION.md shows the
 * stub a PythTB maintainer would add.
 *
 * Conventions
 *  - Plain C: pointers and sizes only.  "dev" pointers are CUDA device pointers
 *    owned by the caller (the Python layer gets them from torch tensors);
 *    "host" pointers are ordinary host memory.  Nothing is allocated inside a
 *    compute call except through the caller-provided workspace.
 *  - complex128 is two consecutive doubles (re, im) — numpy/LAPACK layout.
 *  - Every call takes a cudaStream_t as void* (NULL = legacy default stream)
 *    and is asynchronous with respect to the host.
 *  - Return value: 0 = success, < 0 = tbk_status error; the message is
 *    available from tbk_last_error() (thread-local).
 *  - Re-entrant: no global state besides the thread-local error string.
 */
#ifndef TBK_H_
#define TBK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  TBK_OK = 0,
  TBK_ERR_ARG = -1,       /* bad argument */
  TBK_ERR_CUDA = -2,      /* CUDA runtime error (message has the details) */
  TBK_ERR_WORKSPACE = -3, /* workspace too small (see tbk_*_workspace) */
  TBK_ERR_CONVERGE = -4,  /* an eigensolver did not converge */
  TBK_ERR_UNSUPPORTED = -5
} tbk_status;

#define TBK_MAX_DIM 4

/* Version of this ABI (major*100 + minor). */
int tbk_version(void);
/* Last error message of the calling thread ("" if none). */
const char* tbk_last_error(void);

/* ------------------------------------------------------------------------
 * Compiled model ("plan").  Replaces the Python objects read by
 * tb_model._gen_ham: _site_energies, _hoppings, _orb, _per (pythtb.py:140-180,
 * 475-478).  Layout documented in pythtb_b200/csrc/tbk_plan.cuh.  All arrays
 * are HOST pointers; tbk_model_create copies them to the current device.
 * ---------------------------------------------------------------------- */
typedef struct {
  int32_t dim_k;        /* number of periodic directions, 0..4 */
  int32_t nsta;         /* norb * nspin */
  int32_t nph;          /* entries of the lattice-vector phase table */
  int32_t nel;          /* lower-triangle matrix elements with at least one term */
  int32_t nterm;        /* scalar terms */
  int32_t convention;   /* 1 = PythTB (orbital positions in the phase), 2 = lattice vectors only */
  const double* ph_R;   /* [nph][dim_k] */
  const double* tau;    /* [nsta][dim_k] */
  const int32_t* el_ptr;  /* [nel+1] */
  const int32_t* el_row;  /* [nel] */
  const int32_t* el_col;  /* [nel] */
  const int32_t* t_ph;    /* [nterm] phase index | 1<<30 (conjugate), or -1 */
  const double* t_amp;    /* [nterm][2] */
  const int32_t* pm_ptr;  /* [nph+2] */
  const int32_t* pm_el;   /* [nterm] element index | 1<<30 */
  const double* pm_amp;   /* [nterm][2] */
} tbk_model_desc;

typedef struct tbk_model tbk_model;

int tbk_model_create(const tbk_model_desc* desc, tbk_model** out);
int tbk_model_destroy(tbk_model* model);

/* ------------------------------------------------------------------------
 * tb_model._gen_ham (pythtb.py:874-925), batched over k.
 *   k_dev  [nk][dim_k]  reduced coordinates          (ignored when dim_k == 0)
 *   ham_dev [nk][nsta][nsta] complex128, full Hermitian matrix, Convention
 *   chosen by the plan.
 * ---------------------------------------------------------------------- */
int tbk_gen_ham(const tbk_model* model, const double* k_dev, int64_t nk,
                double* ham_dev, void* stream);

/* ------------------------------------------------------------------------
 * tb_model._sol_ham + _nicefy_eig (pythtb.py:927-953, 3765-3775), batched.
 *   ham_dev [batch][n][n] complex128 row-major; only the lower triangle is
 *           read (numpy UPLO='L'); not modified.
 *   eval_dev [batch][n] ascending.
 *   evec_dev [batch][n(band)][n] rows = eigenvectors, or NULL.
 *   ws_dev / ws_bytes: workspace from tbk_eigh_workspace(n, batch, want_vec).
 * ---------------------------------------------------------------------- */
size_t tbk_eigh_workspace(int32_t n, int64_t batch, int32_t want_vec);
int tbk_eigh_batched(const double* ham_dev, int32_t n, int64_t batch, double* eval_dev,
                     double* evec_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * tb_model.solve_all (pythtb.py:955-1079): fused assembly + diagonalisation
 * for a list of k-points; H(k) never touches HBM for nsta below the
 * shared-memory limit.  Output in the reference layouts through strides
 * (in elements):  eval[b*ev_sb + k*ev_sk],  evec[b*vc_sb + k*vc_sk + orb]
 * so that solve_all's eval[band,k] / evec[band,k,orb(,spin)] are written
 * directly.  evec_dev may be NULL (eigenvalues only).
 * ---------------------------------------------------------------------- */
size_t tbk_solve_workspace(int32_t nsta, int64_t nk, int32_t want_vec);
int tbk_solve_k(const tbk_model* model, const double* k_dev, int64_t nk,
                double* eval_dev, int64_t ev_sb, int64_t ev_sk,
                double* evec_dev, int64_t vc_sb, int64_t vc_sk,
                void* ws_dev, size_t ws_bytes, void* stream);

/* tb_model.k_uniform_mesh (pythtb.py:1792-1861) on the device: k_dev[nk][nd], nk = prod(mesh), point
 * (i_0/mesh[0], i_1/mesh[1], ...) at the C-order index of (i_0, i_1, ...) — the same doubles as numpy's
 * arange(n)/float(n).  Lets solve_all run on a dense mesh (256^3 = 16.8 M points = 403 MB of k) without
 * building the list on the host and shipping it over PCIe. */
int tbk_kmesh_uniform(const int32_t* mesh_host, int32_t nd, double* k_dev, void* stream);

/* ------------------------------------------------------------------------
 * wf_array.solve_on_grid + impose_pbc (pythtb.py:2421-2532, 2674-2749).
 * k = start_k[d] + i_d/(mesh[d]-1) is generated on the device
 * (pythtb.py:2477).  The call fills rows [row0, row0+nrows) of mesh axis 0 of
 * a LOCAL wavefunction slab
 *     wfs_dev[local_row][i_1]..[i_{nd-1}][state][orb(,spin)]
 * whose axis-0 extent is nrows+1 (the extra row is the periodic image / halo
 * of the next shard) and whose other extents are mesh[d].  Periodic images
 * along axes d >= 1 are always written (x pbc_phase).  The closing row of
 * axis 0 depends on wrap0:
 *   1  single shard (row0 == 0, nrows == mesh[0]-1): the image of row 0 is written;
 *   0  left untouched — the caller fills it with the next shard's first row
 *      (tbk_halo_pack + an NCCL ring shift);
 *   2  solved in this launch as global row row0+nrows (the periodic image
 *      row 0 x pbc_phase when that index is mesh[0]-1): no communication, the
 *      values are bit-identical to the neighbour's first row.
 * ws_dev / ws_bytes: tbk_solve_workspace(nsta, npts, 1).
 *   pbc_phase_dev [nd][nsta] complex128 = exp(-2 pi i tau_j[per[d]])  (:2729)
 *   gaps_dev [nsta-1] minimal direct gaps over the solved rows (:2484,:2529);
 *            may be NULL.
 * ---------------------------------------------------------------------- */
int tbk_solve_grid(const tbk_model* model, const double* start_k_host, const int32_t* mesh_host,
                   int32_t nd, int32_t row0, int32_t nrows, int32_t wrap0,
                   double* wfs_dev, const double* pbc_phase_dev, double* gaps_dev,
                   void* ws_dev, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * Multi-GPU peer group (one process per GPU on one NVLink/NVSwitch box).
 * The reductions that end solve_on_grid (min direct gaps, pythtb.py:2529-2530)
 * and berry_flux (plane sum, pythtb.py:3148) are finished ACROSS ranks inside
 * the producing kernel: its last CTA stores the rank's partial result into a
 * mailbox in every peer's HBM (CUDA-IPC mapped, stores travel over NVLink),
 * waits for the peers' contributions and combines them in rank order — no
 * NCCL launch and no extra kernel (csrc/tbk_peer.cuh).
 *   tbk_peer_create   allocates this rank's mailbox, returns its 64-byte IPC handle;
 *   tbk_peer_connect  maps the peers' mailboxes (handles: [nranks][64] bytes, any
 *                     out-of-band all-gather — the Python layer uses torch.distributed);
 * All ranks must issue the same sequence of *_x calls with a non-NULL peer. */
typedef struct tbk_peer tbk_peer;
int tbk_peer_create(int32_t rank, int32_t nranks, tbk_peer** out, void* handle_out);
int tbk_peer_connect(tbk_peer* peer, const void* handles);
int tbk_peer_destroy(tbk_peer* peer);
/* Device-side barrier over the group on `stream` (a one-warp kernel): kernels enqueued after it start
 * only when every rank has reached the same point of its own stream.  Used by bench.py to start a
 * timed multi-GPU step together on all ranks (the analogue of torch.distributed.barrier(), but in
 * stream order and without a host round trip).  No-op for peer == NULL or a single rank. */
int tbk_peer_barrier(tbk_peer* peer, void* stream);
/* Deferred reduction.  tbk_peer_defer(peer, 1) makes the NEXT tbk_solve_grid_x keep this rank's minimal
 * gaps in its own memory (no exchange, no NVLink traffic) and return; gaps_dev is completed (minimum over
 * the ranks) by the next tbk_flux_plane_x issued with the same peer — the gaps travel in the same message as
 * the flux sums, so a solve + flux step is ONE exchange instead of two — or by tbk_peer_flush (a one-CTA
 * kernel), or implicitly before a later collective that cannot carry it.  gaps_dev must stay allocated
 * until then.  All ranks must make the same sequence of calls. */
int tbk_peer_defer(tbk_peer* peer, int32_t on);
int tbk_peer_flush(tbk_peer* peer, void* stream);

/* Multi-GPU: pack the first local row of a shard ([npoints][nsta_arr][n]
 * complex128) into a contiguous send buffer for the ring shift that closes the
 * neighbour's slab; phase_dev [n] (the pbc phase of pythtb.py:2729, applied by
 * rank 0 whose row becomes the last rank's periodic image) or NULL. */
int tbk_halo_pack(const double* row_dev, double* dst_dev, int64_t npoints, int32_t nsta_arr, int32_t n,
                  const double* phase_dev, void* stream);

/* wf_array.impose_pbc / impose_loop (pythtb.py:2674-2791) on a device array of
 * shape [outer][len][inner][nsta_arr][n]: slice len-1 = slice 0 (* phase[n]).
 * phase_dev may be NULL (impose_loop). */
int tbk_impose_boundary(double* wfs_dev, int64_t outer, int64_t len, int64_t inner,
                        int32_t nsta_arr, int32_t n, const double* phase_dev, void* stream);

/* ------------------------------------------------------------------------
 * Berry machinery on a wavefunction array viewed as
 *     wfs[slice][i0][i1][state][n]      (strides in complex elements)
 * ---------------------------------------------------------------------- */
typedef struct {
  const double* wfs_dev;  /* complex128 */
  int32_t n;              /* norb*nspin: contiguous length of one state */
  int32_t nsta_arr;       /* states stored per mesh point (stride n between states) */
  int32_t nocc;           /* number of selected states */
  const int32_t* occ_dev; /* [nocc] indices of the selected states */
  int64_t state_stride;   /* complex elements between consecutive states of one mesh point; 0 = n (the reference's
                           * [k..., state, orb] layout).  A STATE-MAJOR array [state][k...][orb] has
                           * state_stride = n * (number of mesh points) and point strides that are multiples of n:
                           * a kernel that needs one band of a two-band model then fetches half the bytes
                           * (64-byte DRAM lines would otherwise hold both bands of a k-point). */
} tbk_wf_view;

/* wf_array.berry_flux / _one_flux_plane (pythtb.py:3068-3205, 3840-3865).
 * Plaquette (i0,i1) of slice s uses the points (i0,i1),(i0+1,i1),(i0+1,i1+1),
 * (i0,i1+1) at offsets slice_off_dev[s] + i0*stride0 + i1*stride1.
 *   plaq_dev  [nslice][n0-1][n1-1] phases in [-pi,pi), or NULL
 *   total_dev [nslice] sum over the plane, or NULL
 * Each is -arg det(M1 M2 M3 M4) evaluated as the product of four link
 * determinants. */
size_t tbk_flux_workspace(int32_t nocc, int32_t n, int64_t nslice, int64_t n0, int64_t n1);
int tbk_flux_plane(const tbk_wf_view* view, const int64_t* slice_off_dev, int64_t nslice,
                   int64_t n0, int64_t stride0, int64_t n1, int64_t stride1,
                   double* plaq_dev, double* total_dev,
                   void* ws_dev, size_t ws_bytes, void* stream);

/* wf_array.berry_phase / _one_berry_loop (pythtb.py:2863-3066, 3798-3838) for
 * nstr strings of npts points: point t of string s at string_off_dev[s] +
 * t*stride.
 *   berry_evals == 0: out_dev[nstr]        = -arg det prod_t M_t      in [-pi,pi)
 *   berry_evals != 0: out_dev[nstr][nocc]  = sorted -arg eig(prod_t U_t), U_t the
 *                     unitary polar factor of M_t (numpy SVD U@Vh).
 * The 2 pi continuity post-processing (:3036-3065) stays on the host. */
size_t tbk_berry_workspace(int32_t nocc, int32_t n, int64_t nstr, int64_t npts, int32_t berry_evals);
int tbk_berry_strings(const tbk_wf_view* view, const int64_t* string_off_dev, int64_t nstr,
                      int64_t npts, int64_t stride, int32_t berry_evals, double* out_dev,
                      void* ws_dev, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * tb_model.position_matrix / position_hwf (pythtb.py:2034-2113, 2162-2279),
 * batched over k.
 *   evec_dev [batch][nocc][n] rows = states;  pos_dev [n] diagonal position
 *   operator (orbital coordinate, repeated over spin).
 *   xmat_dev [batch][nocc][nocc] = <m| r |n>.
 * tbk_position_hwf additionally diagonalises X:
 *   hwfc_dev [batch][nocc] ascending centres,
 *   hwf_dev  [batch][nocc][nocc] (orbital_basis == 0, rows = eigenvectors in the
 *            basis of the input states) or [batch][nocc][n] (orbital_basis != 0,
 *            rows = hwf @ evec, pythtb.py:2262-2274); may be NULL.
 * ---------------------------------------------------------------------- */
int tbk_position_matrix(const double* evec_dev, int64_t batch, int32_t nocc, int32_t n,
                        const double* pos_dev, double* xmat_dev, void* stream);
size_t tbk_position_hwf_workspace(int32_t nocc, int32_t n, int64_t batch);
int tbk_position_hwf(const double* evec_dev, int64_t batch, int32_t nocc, int32_t n,
                     const double* pos_dev, double* hwfc_dev, double* hwf_dev,
                     int32_t orbital_basis, void* ws_dev, size_t ws_bytes, void* stream);

/* tbk_solve_grid / tbk_flux_plane with the cross-rank reduction fused in: gaps_dev
 * (all ranks) receives the minimum over ranks, total_dev the sum over ranks (rank order).
 * Supported when the shard is solved by the register-resident mesh kernel (nsta <= 4) /
 * the row-marching flux kernel (nocc <= 2, n <= 4) and nslice <= 16; otherwise
 * TBK_ERR_UNSUPPORTED is returned before anything is launched and the caller reduces
 * with NCCL.  peer == NULL behaves exactly like the plain entry points. */
/* state_stride (tbk_solve_grid_x / tbk_solve_grid_prepare): 0 = the slab is stored [row][i_1]..[state][orb] as
 * described at tbk_solve_grid; > 0 = STATE-MAJOR slab [state][row][i_1]..[orb] with that many complex elements
 * between the states of a mesh point (= n * number of local mesh points for a compact array); see tbk_wf_view. */
int tbk_solve_grid_x(const tbk_model* model, const double* start_k_host, const int32_t* mesh_host,
                     int32_t nd, int32_t row0, int32_t nrows, int32_t wrap0,
                     double* wfs_dev, const double* pbc_phase_dev, double* gaps_dev,
                     void* ws_dev, size_t ws_bytes, int64_t state_stride, tbk_peer* peer, void* stream);
int tbk_flux_plane_x(const tbk_wf_view* view, const int64_t* slice_off_dev, int64_t nslice,
                     int64_t n0, int64_t stride0, int64_t n1, int64_t stride1,
                     double* plaq_dev, double* total_dev,
                     void* ws_dev, size_t ws_bytes, tbk_peer* peer, void* stream);

/* Wilson loops split across ranks (wf_array sharded along the string direction, berry_evals=True):
 * tbk_wilson_products returns, for every local string, the ORDERED product of the unitary (polar) link
 * matrices of its local links, prod_dev [nstr][nocc][nocc] (pythtb.py:3813-3826 restricted to the local
 * links; workspace as tbk_berry_workspace(..., berry_evals = 1)).  The host gathers the per-rank products
 * in rank order and tbk_wilson_phases multiplies mats_dev [nstr][nmat][nocc][nocc] (destroyed) along nmat
 * and returns the sorted eigenphases -angle(eigvals), out_dev [nstr][nocc] (pythtb.py:3834-3838). */
int tbk_wilson_products(const tbk_wf_view* view, const int64_t* string_off_dev, int64_t nstr, int64_t npts,
                        int64_t stride, double* prod_dev, void* ws_dev, size_t ws_bytes, void* stream);
size_t tbk_wilson_workspace(int32_t nocc, int64_t nstr, int64_t nmat);
int tbk_wilson_phases(double* mats_dev, int64_t nstr, int64_t nmat, int32_t nocc, double* out_dev,
                      void* ws_dev, size_t ws_bytes, void* stream);

/* The ordered product alone: prod_dev [nstr][nocc][nocc] = mats[s][0] mats[s][1] ... mats[s][nmat-1] (mats_dev
 * destroyed; workspace as tbk_wilson_workspace).  Used by the streamed 1-D Berry phase (config 4: a string of
 * 1e5 k-points is processed in chunks that never coexist in memory; every chunk contributes the ordered
 * product of its unitary link matrices, pythtb.py:3821-3826, and the chunk products are chained here). */
int tbk_wilson_chain(double* mats_dev, int64_t nstr, int64_t nmat, int32_t nocc, double* prod_dev,
                     void* ws_dev, size_t ws_bytes, void* stream);

/* Name of the kernel family the calling thread's last solve call dispatched to
 * (benchmark / profile bookkeeping). */
const char* tbk_last_kernel(void);
/* Number of kernels this library has launched in this process (all threads). */
int64_t tbk_launch_count(void);

/* Block the calling host thread until everything enqueued on `stream` has finished
 * (cudaStreamSynchronize); the host layer uses it after a call whose results are
 * written straight into pinned host memory. */
int tbk_stream_sync(void* stream);

/* Prepared calls.  tbk_solve_grid_prepare / tbk_flux_plane_prepare take exactly the arguments of
 * tbk_solve_grid_x / tbk_flux_plane_x (peer may be NULL) and keep them by value (start_k, mesh and the
 * view are copied; device buffers, the model and the peer group are referenced and must outlive the
 * handle).  tbk_prepared_run re-issues that call on `stream`; with sync != 0 it also waits for the
 * stream, so a repeated host-in / host-out step is one foreign call.  Nothing is launched by *_prepare. */
typedef struct tbk_prepared tbk_prepared;
int tbk_solve_grid_prepare(const tbk_model* model, const double* start_k, const int32_t* mesh, int32_t nd,
                           int32_t row0, int32_t nrows, int32_t wrap0, double* wfs_dev,
                           const double* pbc_phase_dev, double* gaps_dev, void* ws_dev, size_t ws_bytes,
                           int64_t state_stride, tbk_peer* peer, tbk_prepared** out);
int tbk_flux_plane_prepare(const tbk_wf_view* view, const int64_t* slice_off_dev, int64_t nslice, int64_t n0,
                           int64_t stride0, int64_t n1, int64_t stride1, double* plaq_dev, double* total_dev,
                           void* ws_dev, size_t ws_bytes, tbk_peer* peer, tbk_prepared** out);
int tbk_prepared_run(tbk_prepared* call, void* stream, int32_t sync);
int tbk_prepared_destroy(tbk_prepared* call);

/* Per-stage cycle counters of the blocked eigensolver, collected when the environment has TBK_PROF=1:
 * out8[0..3] = SM cycles spent in tridiagonalisation / bisection / inverse iteration / back-transformation
 * (summed over CTAs), out8[4] = matrices solved, out8[5] = matrices handed to the fallback solver,
 * out8[6] = cycles of the slowest single matrix.
 * Call after synchronising the stream.  Profiling aid, not part of the reference interface. */
int tbk_debug_profile(uint64_t* out8, int32_t reset);
/* Profiling aid (TBK_CTA_TRACE=1 in the environment before the first launch): per-CTA timeline of the
 * last mesh_small_kernel launch (entries [0, 4096)) and the last flux_rows_kernel launch (entries [4096, 8192)),
 * 4 words per CTA: SM id, begin and end (%globaltimer, ns), blockIdx.  out: [max_ctas][4], max_ctas <= 8192.
 * TBK_ERR_UNSUPPORTED when tracing is off. */
int tbk_debug_cta_trace(uint64_t* out, int64_t max_ctas, int32_t reset);

/* FP64 peak microkernels for the benchmark harness: one launch of a register-resident kernel that issues only
 * independent DFMA chains (kind 0, the FP64 FMA pipe) or independent mma.sync.m8n8k4.f64 (kind 1, the DMMA path
 * the overlap GEMMs use) on every SM.  tbk_bench_fp64_flops returns the flops of one such launch; the caller
 * times it with CUDA events: achieved flop/s = the FP64 roofline denominator MEASURED on this box in this run. */
int tbk_bench_fp64(int32_t kind, int32_t iters, double* sink_dev, void* stream);
double tbk_bench_fp64_flops(int32_t kind, int32_t iters);

/* L2 flush helper for benchmarks: overwrites buf_dev[bytes] (bytes > L2 size). */
int tbk_flush_l2(void* buf_dev, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TBK_H_ */
