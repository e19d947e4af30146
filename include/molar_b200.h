/* molar_b200.h — C ABI of libmolar_b200.so, the B200-native drop-in for MolAR's data-parallel
 * hot path (distance search, Kabsch fit / RMSD, centre-of-mass / gyration reductions).
 *
 * The boundary follows the only native-plugin precedent in the reference, molar_gromacs
 * (/root/reference/molar_gromacs/gromacs/wrapper.hpp:42-81): opaque handles, NULL / negative
 * return + a thread-local last-error string, exceptions never cross the ABI, count-then-fill
 * into caller-allocated arrays, plain pointers and sizes only.
 *
 * What each entry point replaces (paths relative to /root/reference/molar/src/):
 *   mb_set_frame          State.coords : Vec<Pos> + State.pbox            state.rs:21-28
 *   mb_set_masses         AtomStorage::masses()                           atom_storage.rs:272
 *   mb_search_single      distance_search_single[_pbc]                    distance_search.rs:892-954
 *   mb_search_double      distance_search_double[_pbc]                    distance_search.rs:659-754
 *   mb_search_within      distance_search_within[_pbc]                    distance_search.rs:519-598
 *   mb_search_double_vdw  distance_search_double_vdw[_pbc]                distance_search.rs:767-879
 *   mb_center_of_mass     Measure::center_of_mass                         measure.rs:60-75
 *   mb_gyration           Measure::gyration                               measure.rs:78-87,561-570
 *   mb_rmsd               rmsd / rmsd_mw                                  measure.rs:485-504,538-558
 *   mb_fit_transform      fit_transform / fit_transform_at_origin         measure.rs:507-535,613-643
 *   mb_apply_transform    Modify::apply_transform                         modify.rs:32-36
 *   mb_center_of_geometry Measure::center_of_geometry                     measure.rs:37-45
 *   mb_center_pbc         center_of_{mass,geometry}_pbc[_dims]            measure.rs:142-214
 *   mb_gyration_pbc       Measure::gyration_pbc                           measure.rs:216-226
 *   mb_inertia            Measure::inertia[_pbc] + do_inertia             measure.rs:88-98,228-238,573-610
 *   mb_principal_transform  Measure::principal_transform[_pbc]            measure.rs:100-108,240-252,645-649
 *   mb_batch_*            the per-frame loop AnalysisTask::run drives     analysis_task.rs:113-280
 *   mb_connectivity       SearchConnectivity::from_iter                   connectivity.rs:8-38
 *   mb_search_connectivity  the same, rows written by the search kernel    connectivity.rs:8-38, distance_search.rs:892-954
 *   mb_unwrap_connectivity  Modify::unwrap_connectivity[_dim]             modify.rs:64-131
 *   mb_batch_load_traj    DcdFileHandler::read_state / XtcFileHandler::read_state (+ molly's XTC codec)
 *                                                                         io/dcd_handler.rs:204-300,389-464; io/xtc_handler.rs:64-110
 *
 * Data conventions (identical to the Rust side, so a binding passes its buffers as they are):
 *   coordinates  Vec<Pos> = N x 3 f32, AoS, 12-byte stride                aliases.rs:23
 *   box          nalgebra Matrix3<f32> storage = COLUMN-major 9 floats; columns are the box
 *                vectors a,b,c (periodic_box.rs:9-13).  NULL = no box.
 *   selections   sorted global atom indices as usize = uint64_t (providers.rs:45-48);
 *                ids == NULL means the identity selection 0..n.  Every entry point checks that ids are
 *                STRICTLY INCREASING and < n_atoms (one O(n) host pass) and returns MB_ERR_ARG otherwise.
 *   pbc_dims     PbcDims bit mask, bit d = dimension d periodic (periodic_box.rs:70-128);
 *                0 selects the non-periodic variant of a search.
 *
 * Threading: a context is Send, not Sync (one context per calling thread, like TprHandle,
 * io/tpr_handler.rs:18).  There is no global mutable state except the thread-local error.
 * There is NO CPU fallback: without a CUDA device mb_open fails.
 */
#ifndef MOLAR_B200_H
#define MOLAR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct MbCtx MbCtx;

/* return codes; they map onto MeasureError (measure.rs:732-762) */
enum {
    MB_OK = 0,
    MB_ERR_ZERO_MASS = -1,  /* MeasureError::ZeroMass */
    MB_ERR_SIZES = -2,      /* MeasureError::Sizes */
    MB_ERR_SVD = -3,        /* MeasureError::Svd */
    MB_ERR_NO_PBC = -4,     /* MeasureError::Pbc(PeriodicBoxError::NoPbc) */
    MB_ERR_BOX = -5,        /* PeriodicBoxError::{ZeroLengthVector,InverseFailed} */
    MB_ERR_ARG = -6,        /* bad argument / no frame / index out of range */
    MB_ERR_CUDA = -7,       /* CUDA runtime error (message in mb_last_error) */
    MB_ERR_STATE = -8       /* call sequence error (e.g. fill before search) */
};

/* Message of the last failed call on this thread; valid until the next call on this thread. */
const char* mb_last_error(void);
/* ABI version of this header */
int mb_abi_version(void);

/* ---- context ---------------------------------------------------------------------------- */
MbCtx* mb_open(int device);            /* NULL on error */
void   mb_close(MbCtx* ctx);
/* cudaStream_t the context launches on (as void*), so a host can time it with its own events */
void*  mb_stream(MbCtx* ctx);
int    mb_synchronize(MbCtx* ctx);
/* Options (results never depend on them; unknown keys fail with MB_ERR_ARG):
 *   with_dist 0/1        pair searches also compute the distances (default 1)
 *   subdiv, subdiv_x/y/z, slice_x, atoms_per_cell   traversal grid of the pair kernel (0 = automatic)
 *   force_brute 0/1      general all-pairs kernel instead of the cell kernel
 *   exact_pbc 0/1        wrapped cell pairs always through the exact PeriodicBox expression (no shifted-image filter)
 *   two_set_cells_min    two-set searches use the cell kernel when n1 * n2 exceeds this
 *   batch_streams        streams the frames of mb_batch_search alternate over (0 = automatic)
 *   fused_fit            mb_batch_fit: 0 two kernels per frame group (default); 1, 2 TMA-staged single-pass kernels;
 *                        3 persistent kernel with an L2-served lagging second pass; 4 warp-specialised persistent
 *                        kernel (TMA rings, mbarrier hand-overs).  fit_lag / fit_teams / fit_group / fit_streams tune them
 *   profile 0/1          CUDA events around every pair-kernel launch (mb_stat) */
int    mb_set_option(MbCtx* ctx, const char* key, double value);

/* ---- frame ------------------------------------------------------------------------------- */
/* Copies the frame to the device (pinned staging + async H2D on the context stream). */
int mb_set_frame(MbCtx* ctx, const float* xyz, size_t n_atoms, const float* box9_colmajor);
/* Adopts a DEVICE pointer without copying (xyz_dev must stay valid until replaced). */
int mb_set_frame_device(MbCtx* ctx, const float* xyz_dev, size_t n_atoms, const float* box9_colmajor);
/* Copies the (possibly transformed) current frame back to the host. */
int mb_get_frame(MbCtx* ctx, float* xyz_out, size_t n_atoms);
/* Whole-system mass column. */
int mb_set_masses(MbCtx* ctx, const float* masses, size_t n_atoms);
/* Copies the mass column back (e.g. after mb_batch_synth_masses). */
int mb_get_masses(MbCtx* ctx, float* masses_out, size_t n_atoms);
/* A second coordinate set (e.g. the reference structure for rmsd / fit); same layout. */
int mb_set_frame2(MbCtx* ctx, const float* xyz, size_t n_atoms);

/* ---- distance search ----------------------------------------------------------------------
 * Each call runs the search on the device and returns the number of results (>= 0) or a
 * negative error code; results stay on the device until the next search on this context and
 * are fetched with mb_fill_*.  Results are the reference's result SET: duplicates the reference
 * would emit for degenerate grids are removed; `single` pairs are canonical (i < j); `double`
 * pairs are (i from set 1, j from set 2); `within` ids are sorted and unique. */
int64_t mb_search_single(MbCtx* ctx, float cutoff, const uint64_t* ids, size_t n, uint8_t pbc_dims);
/* set 2 positions are taken from frame2 when use_frame2 != 0, else from the current frame */
int64_t mb_search_double(MbCtx* ctx, float cutoff, const uint64_t* ids1, size_t n1,
                         const uint64_t* ids2, size_t n2, int use_frame2, uint8_t pbc_dims);
/* van der Waals contact search: pair (i,j) is reported when d <= vdw1[i] + vdw2[j] + EPSILON.  vdw1/vdw2
   hold one radius per SELECTED atom (host arrays of n1 / n2 floats); as in the reference the reported
   indices are LOCAL (positions within the two selections), fetched with mb_fill_pairs. */
int64_t mb_search_double_vdw(MbCtx* ctx, const uint64_t* ids1, size_t n1, const float* vdw1,
                             const uint64_t* ids2, size_t n2, const float* vdw2, int use_frame2, uint8_t pbc_dims);
/* lower3/upper3: grid bounds of the non-periodic variant, as the caller of
   distance_search_within passes them (selection/ast.rs:598-610); ignored when pbc_dims != 0;
   NULL => bounds of set 1 padded by cutoff + EPSILON (what the `within` AST node does). */
int64_t mb_search_within(MbCtx* ctx, float cutoff, const uint64_t* ids1, size_t n1,
                         const uint64_t* ids2, size_t n2, int use_frame2, uint8_t pbc_dims,
                         const float* lower3, const float* upper3);
/* Count only (no pair list is materialised): the contact count of config 5. */
int64_t mb_count_single(MbCtx* ctx, float cutoff, const uint64_t* ids, size_t n, uint8_t pbc_dims);
/* ij: 2*P entries (usize pairs) ; dist: P entries or NULL */
int mb_fill_pairs(MbCtx* ctx, uint64_t* ij, float* dist);
/* The same list in the device's own form, u32 x 2 per pair (8 B: half the PCIe bytes; BondStorage likewise caps
   systems at 2^32 atoms), single-set pairs canonical i < j.  ij32 should be page-locked (mb_host_alloc). */
int mb_fill_pairs_u32(MbCtx* ctx, uint32_t* ij32, float* dist);
int mb_fill_ids(MbCtx* ctx, uint64_t* ids);
/* Device-side view of the last pair list: packed (uint32, uint32), P entries.  For a single-set
   search the two ids of an entry are in no particular order (mb_fill_pairs and mb_pairs_checksum
   canonicalise to i < j); for a double search entry = (i from set 1, j from set 2). */
const void* mb_pairs_device(MbCtx* ctx, int64_t* n_pairs);
/* Order-independent checksum of the last pair list computed on the device:
   out[0] = sum over pairs of mix64(i<<32|j) (wrapping), out[1] = xor of the same. */
int mb_pairs_checksum(MbCtx* ctx, uint64_t out2[2]);
/* Reference grid dimensions used by the last search (Grid::dims). */
int mb_last_grid_dims(MbCtx* ctx, uint64_t dims3[3]);

/* ---- reductions / Kabsch (outputs f64 regardless of input width) ------------------------ */
int mb_center_of_mass(MbCtx* ctx, const uint64_t* ids, size_t n, double out3[3]);
int mb_gyration(MbCtx* ctx, const uint64_t* ids, size_t n, double* out);
/* sel1 on the current frame, sel2 on frame2 (use_frame2 != 0) or on the current frame */
int mb_rmsd(MbCtx* ctx, const uint64_t* ids1, size_t n1, const uint64_t* ids2, size_t n2,
            int use_frame2, int mass_weighted, double* out);
/* fits sel1 (current frame) ONTO sel2; R9 column-major, p' = R p + t */
int mb_fit_transform(MbCtx* ctx, const uint64_t* ids1, size_t n1, const uint64_t* ids2, size_t n2,
                     int use_frame2, int at_origin, double R9_colmajor[9], double t3[3]);
/* in place on the current frame (device copy; fetch with mb_get_frame) */
int mb_apply_transform(MbCtx* ctx, const uint64_t* ids, size_t n, const double R9_colmajor[9],
                       const double t3[3]);

/* ---- many small selections in one launch ---------------------------------------------------
 * The reference analyses per-residue / per-molecule quantities with a rayon loop over thousands of small selections,
 * each calling Measure::center_of_mass / center_of_geometry / gyration (selection.rs:318-322,
 * selection/par_split.rs:100-125).  Here all of them go to the device in one call: selection s is
 * ids[offsets[s] .. offsets[s+1]) (global atom ids, strictly increasing inside a selection; selections may overlap),
 * one warp (one CTA for large selections) per selection.  out: n_sel rows of
 *   MB_REDUCE_COM 3 doubles | MB_REDUCE_COG 3 | MB_REDUCE_GYRATION 1 | MB_REDUCE_COM_GYRATION 4 {com, rg}.
 * status_out (may be NULL): per selection 0 ok, 1 zero mass, 2 empty.  Returns MB_ERR_ZERO_MASS / MB_ERR_ARG if any
 * selection failed (its row is NaN; the other rows are valid). */
enum { MB_REDUCE_COM = 0, MB_REDUCE_COG = 1, MB_REDUCE_GYRATION = 2, MB_REDUCE_COM_GYRATION = 3 };
int mb_reduce_many(MbCtx* ctx, const uint64_t* ids, const uint64_t* offsets, size_t n_sel, int what, double* out,
                   int* status_out);

/* ---- periodic variants, inertia tensor, principal axes ----------------------------------- */
int mb_center_of_geometry(MbCtx* ctx, const uint64_t* ids, size_t n, double out3[3]);
/* center_of_mass_pbc[_dims] (mass_weighted != 0) / center_of_geometry_pbc[_dims]: every atom is replaced by
   its closest image next to the FIRST atom of the selection (pbc_dims: which dimensions wrap; 7 = the
   plain _pbc variants).  As in the reference, the first atom enters the mass-weighted numerator with
   weight one.  MB_ERR_NO_PBC without a box. */
int mb_center_pbc(MbCtx* ctx, const uint64_t* ids, size_t n, int mass_weighted, uint8_t pbc_dims, double out3[3]);
int mb_gyration_pbc(MbCtx* ctx, const uint64_t* ids, size_t n, double* out);
/* moments ascending; axes column-major, columns = principal axes (col2 = col0 x col1).  An axis is defined
   up to its sign: col0 and col1 are returned with their largest component positive.  pbc != 0: distances to
   the periodic centre of mass are minimum-image vectors (inertia_pbc). */
int mb_inertia(MbCtx* ctx, const uint64_t* ids, size_t n, int pbc, double moments3[3], double axes9_colmajor[9]);
/* p' = R p + t rotates the selection about its centre of mass onto its principal axes */
int mb_principal_transform(MbCtx* ctx, const uint64_t* ids, size_t n, int pbc, double R9_colmajor[9], double t3[3]);

/* ---- pair-list consumers that stay on the device --------------------------------------------
 * Adjacency (atom -> neighbours, both directions) of the LAST pair list of this context as CSR over the index
 * space [0, n_index): returns the number of entries (2 x pairs); row_ptr_out (n_index + 1 entries) may be NULL;
 * the neighbour lists follow with mb_fill_connectivity.  Order inside a row is unspecified (as in the reference). */
int64_t mb_connectivity(MbCtx* ctx, size_t n_index, uint64_t* row_ptr_out);
/* The same adjacency WITHOUT a pair list in between: SearchConnectivity::from_iter(distance_search_single[_pbc](cutoff,
 * sel)) (connectivity.rs:8-38 over distance_search.rs:892-954) as CSR over [0, n_atoms).  The search kernel walks the full
 * neighbour shell twice (count per atom, then every atom's row written at its own cursor), so no per-pair atomics and no
 * 8-byte pair ever reaches memory; small or degenerate grids fall back to pair list -> CSR.  Returns the number of
 * entries (2 x pairs); row_ptr_out (n_atoms + 1 entries) may be NULL; rows follow with mb_fill_connectivity.  The
 * context holds no pair list afterwards (mb_fill_pairs fails with MB_ERR_STATE). */
int64_t mb_search_connectivity(MbCtx* ctx, float cutoff, const uint64_t* ids, size_t n, uint8_t pbc_dims,
                               uint64_t* row_ptr_out);
int mb_fill_connectivity(MbCtx* ctx, uint64_t* cols_out);
/* Order-independent checksum of the adjacency on the device (parity checks at sizes whose lists should not travel):
   out4 = {sum, xor} of mix64((min << 32) | max) over the entries with row < column, then over those with row > column;
   for a symmetric list both halves equal mb_pairs_checksum of the pair list. */
int mb_connectivity_checksum(MbCtx* ctx, uint64_t out4[4]);
/* unwrap_connectivity_dim: contact graph of the selection at `cutoff` (periodic in all dims), every atom moved to
 * its closest image (image_dims) next to the atom it is reached from, walking from the lowest-index atom of each
 * component; coordinates change in place on the device (mb_get_frame).  roots_out[k] (may be NULL) = position within
 * the selection of the start atom of k's component.  Returns the number of components (start atoms).  Components and
 * start atoms are the reference's; positions agree with any traversal order up to f32 rounding. */
int64_t mb_unwrap_connectivity(MbCtx* ctx, float cutoff, const uint64_t* ids, size_t n, uint8_t image_dims,
                               int64_t* roots_out);

/* ---- batched, device-resident trajectory (what the benchmark drives) -------------------- */
/* Allocate n_frames x n_atoms x 3 f32 on the device and fill it with the synthetic generator
   of SURVEY.md §8(d) (same bits as the oracle's orc_synth_frame): frame f of this context is
   global frame first_frame + f. */
int mb_batch_synth(MbCtx* ctx, uint64_t seed, uint64_t first_frame, size_t n_frames, size_t n_atoms,
                   const float* box9_colmajor, int stray_permille);
/* Upload host frames instead (n_frames x n_atoms x 3). */
int mb_batch_upload(MbCtx* ctx, const float* xyz, size_t n_frames, size_t n_atoms,
                    const float* box9_colmajor);
int mb_batch_synth_masses(MbCtx* ctx, uint64_t seed, size_t n_atoms);
/* Trajectory ingest: decode frames [first_frame, first_frame + n_frames) of a DCD or XTC byte stream (the file
   contents, or any prefix holding whole frames) ON THE DEVICE into the resident batch.  The box of the first
   loaded frame becomes the batch box; per-frame boxes (9 floats, column-major, zeros = no box) and times are
   written to the optional host arrays.  DCD: both endiannesses, CHARMM extra/4D blocks, fixed atoms;
   coordinates are `x * 0.1` in f32 as the reference forms them.  XTC: magic 1995, precision as stored. */
enum { MB_TRAJ_DCD = 0, MB_TRAJ_XTC = 1 };
int mb_traj_probe(const void* bytes, size_t n_bytes, int format, size_t* n_frames, size_t* n_atoms);
int mb_batch_load_traj(MbCtx* ctx, const void* bytes, size_t n_bytes, int format, size_t first_frame, size_t n_frames,
                       float* boxes9_out, float* times_out);
/* Copy frames [f0, f1) of the resident batch to the host (n_atoms x 3 f32 each). */
int mb_batch_download(MbCtx* ctx, size_t f0, size_t f1, float* xyz_out);
/* Make batch frame f the current frame (no copy). */
int mb_batch_select(MbCtx* ctx, size_t frame);
/* Neighbour search on every frame [f0, f1): per-frame pair count and checksum written to
   host arrays (may be NULL); pair lists stay on the device and are overwritten frame by frame.
   mode 0 = enumerate pairs, 1 = count only. */
int mb_batch_search(MbCtx* ctx, float cutoff, uint8_t pbc_dims, size_t f0, size_t f1, int mode,
                    int64_t* counts, uint64_t* checksums2);
/* The same search over n_frames HOST frames ([n_frames][n_atoms][3] f32, ideally pinned): frames are uploaded on a
   copy stream in chunks while the previous chunk is searched — the role of the reference's IO thread + bounded
   channel (io.rs:209-233).  The pair list of the last frame stays on the device (mb_fill_pairs). */
int mb_stream_search(MbCtx* ctx, float cutoff, uint8_t pbc_dims, const float* frames, size_t n_frames, size_t n_atoms,
                     const float* box9_colmajor, int mode, int64_t* counts);
/* Same streaming for the other two per-frame loops: Kabsch fit of every host frame onto the first one (RMSD after
   the fit out; masses via mb_set_masses), and COM + gyration + contact count (n_frames x 5 doubles out). */
int mb_stream_fit(MbCtx* ctx, const float* frames, size_t n_frames, size_t n_atoms, double* rmsd_out);
int mb_stream_pipeline(MbCtx* ctx, float cutoff, uint8_t pbc_dims, const float* frames, size_t n_frames, size_t n_atoms,
                       const float* box9_colmajor, double* out);
/* Kabsch fit of every frame [f0,f1) onto frame `ref_frame`, superposition in place and
   unweighted RMSD after the fit (config 4).  rmsd_out: f1-f0 doubles (host, may be NULL). */
int mb_batch_fit(MbCtx* ctx, size_t ref_frame, size_t f0, size_t f1, int superpose, double* rmsd_out);
/* Per-frame COM + gyration + contact count (config 5): out is (f1-f0) x 5 doubles
   {com_x, com_y, com_z, rg, count}.  The same rows are left in device memory at
   mb_batch_scalars_device() for an NCCL gather by the host. */
int mb_batch_pipeline(MbCtx* ctx, float cutoff, uint8_t pbc_dims, size_t f0, size_t f1, double* out);
const void* mb_batch_scalars_device(MbCtx* ctx, size_t* n_rows, size_t* row_doubles);
/* Host only, no device needed: the traversal plan of the cell kernels for a periodic search of n atoms in this box.
   out_int: [0] cell kernel usable, [1..3] reference Grid::dims, [4..6] fine cells per reference cell, [7] x slices
   per home tile, [8..10] fine grid dims, [11] neighbour rows, [12] shifted-image filter enabled, [13] row capacity.
   rows4_out (4 x capacity signed chars): dy, dz, dxlo, dxhi per row, relative to the first cell of the home tile.
   out_band: [rc2_lo, rc2_hi] of the filter.  For tests of the planning logic. */
int mb_plan_describe(const float* box9_colmajor, float cutoff, uint8_t pbc_dims, size_t n, int full_shell,
                     int out_int[16], signed char* rows4_out, float out_band[2]);
/* Host only, no device needed: the periodic-box tables of the kernels — the reference's triclinic corrections
   (periodic_box.rs:25-66; ncorr vectors, 3 floats each) and the 13 (+v, -v) pair table (threshold per pair, bit
   1 << index-in-the-list per member, 0 when the reference's list does not hold it) the minimum-image code prunes them
   with.  For tests of the table logic. */
int mb_box_describe(const float* box9_colmajor, int* ncorr_out, float corr78_out[78], float pair_thr13_out[13],
                    uint32_t pair_bit26_out[26]);
/* ---- multi-GPU: frames shard across ranks, the per-frame scalars are gathered over NCCL ------------------------
 * The per-frame loop (analysis_task.rs:113-280) is embarrassingly parallel over frames: rank r (one GPU, one context)
 * processes its block of frames with the mb_batch_* / mb_stream_* calls and nothing is exchanged until the end of a
 * pass, when every rank contributes its rows of per-frame scalars to one NCCL all-gather (NVLink / NVSwitch).
 * libnccl.so.2 is opened with dlopen on first use (override the path with MOLAR_B200_NCCL).
 *   multi-process (one rank per GPU): rank 0 calls mb_comm_unique_id and hands the 128 bytes to the other ranks by
 *   any host channel (file, socket, MPI); every rank then calls mb_comm_init.
 *   single process: mb_comm_init_all over n contexts (one per device); collective calls must then come from one
 *   host thread per context, like any NCCL communicator created with ncclCommInitAll. */
#define MB_COMM_ID_BYTES 128
int mb_comm_unique_id(unsigned char* id128);
int mb_comm_init(MbCtx* ctx, int rank, int world, const unsigned char* id128);
int mb_comm_init_all(int n_ctx, MbCtx* const* ctxs);
void mb_comm_destroy(MbCtx* ctx);   /* also done by mb_close */
/* rank / world of the context's communicator (0 / 1 without one) and the NCCL version in use (0 = not loaded) */
int mb_comm_info(MbCtx* ctx, int* rank, int* world, int* nccl_version);
/* All-gather of per-frame scalar rows: every rank contributes n_rows x n_cols doubles; out_all (host) receives
   world x n_rows x n_cols doubles in rank order on EVERY rank.  rows == NULL contributes the rows the last
   mb_batch_pipeline / mb_batch_fit left on the device (mb_batch_scalars_device) without a host round trip.
   Without a communicator (single GPU) it degenerates to a copy. */
int mb_gather_scalars(MbCtx* ctx, const double* rows, size_t n_rows, size_t n_cols, double* out_all);
/* element-wise maximum over ranks (in place; how multi-GPU timings are reduced) and a barrier */
int mb_comm_max(MbCtx* ctx, double* inout, size_t n);
int mb_comm_barrier(MbCtx* ctx);
/* CUDA-event timing on the context stream for hosts without a CUDA binding of their own: slots 0..7 */
int mb_timer_record(MbCtx* ctx, int slot);
int mb_timer_elapsed_ms(MbCtx* ctx, int slot_begin, int slot_end, double* ms);
/* page-locked host memory (frames handed to mb_stream_* should live in it) */
void* mb_host_alloc(size_t bytes);
void mb_host_free(void* p);

/* number of kernels this context has launched since it was opened (for gpu_launches) */
uint64_t mb_launch_count(MbCtx* ctx);
/* Instrumentation.  With option "profile"=1 every pair-search kernel launch is bracketed by CUDA
   events on the context stream; "search_kernel_ms" / "search_kernel_launches" return the totals
   since the option was last set.  Other keys: "pair_capacity", "sm_count", "search_tests_per_frame" (distance tests
   per frame evaluated by the last count-only batch search: the numerator of its FP32-pipe roofline). */
int mb_get_stat(MbCtx* ctx, const char* key, double* out);

#ifdef __cplusplus
}
#endif
#endif /* MOLAR_B200_H */
