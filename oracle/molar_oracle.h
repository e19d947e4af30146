/* molar_oracle.h — C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement of MolAR's
 * (yesint/molar, /root/reference) algorithms for the hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it; the product (libmolar_b200.so) never links, loads or calls it.
 *
 * Parity status (see oracle/README.md):
 *   - search + PeriodicBox: PINNED by the reference's golden `within` vectors
 *     (molar/tests/generated_vmd_tests.in:27,35; generated_pteros_tests.in:21,27)
 *     and the PeriodicBox known-answer tests (molar/src/periodic_box.rs:456-620).
 *   - center_of_mass / gyration / rmsd / fit_transform / apply_transform:
 *     PARITY UNPINNED by the reference (its tests only print,
 *     molar/src/selection.rs:100-107,148-172); pinned here against an independent
 *     numpy/scipy f64 implementation and invariants instead.
 */
#ifndef MOLAR_ORACLE_H
#define MOLAR_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- PeriodicBox (molar/src/periodic_box.rs) ------------------------------ */
typedef struct OrcBox OrcBox;
/* m9: nalgebra Matrix3 storage order = column-major, columns are a,b,c.
   Returns NULL on ZeroLengthVector / InverseFailed (periodic_box.rs:156-176). */
OrcBox* orc_box_from_matrix(const float* m9_colmajor);
/* periodic_box.rs:188-235; NULL on error */
OrcBox* orc_box_from_vectors_angles(float a, float b, float c, float alpha, float beta, float gamma);
void orc_box_free(OrcBox*);
void orc_box_get(const OrcBox*, float* matrix9_colmajor, float* inv9_colmajor);
/* number of triclinic corrections, optionally copied out (3 floats each) */
int orc_box_corrections(const OrcBox*, float* out /*may be NULL*/);
void orc_box_lab_extents(const OrcBox*, float* out3);                              /* :369-375 */
void orc_box_shortest_vector(const OrcBox*, const float* v3, uint8_t dims, float* out3); /* :286-318 */
float orc_box_distance_squared(const OrcBox*, const float* p1, const float* p2, uint8_t dims); /* :379-381 */

/* ---- distance search (molar/src/distance_search.rs) ----------------------- */
/* Results are the reference's raw output: plan order, (i in cell1, j in cell2),
   duplicates included.  Canonicalisation (min,max / sort / unique) is the caller's job. */
typedef struct OrcResult OrcResult;
size_t orc_result_len(const OrcResult*);
/* ij: 2*len (may be NULL), d: len (may be NULL) */
void orc_result_fill(const OrcResult*, uint64_t* ij, float* d);
/* for `within` results: len ids */
void orc_result_fill_ids(const OrcResult*, uint64_t* ids);
void orc_result_grid_dims(const OrcResult*, uint64_t* dims3);
void orc_result_free(OrcResult*);

/* xyz: base pointer of the WHOLE frame (N x 3 f32 AoS, Vec<Pos>); ids: sorted global
   indices of the selection (NULL => identity 0..n).  nthreads<=1: serial plan loop;
   >1: thread pool over plan entries in chunks >=3 (mirrors rayon with_min_len(3),
   distance_search.rs:949-953). */
OrcResult* orc_search_single(float cutoff, const float* xyz, const uint64_t* ids, size_t n,
                             int nthreads);                                         /* :892-915 */
OrcResult* orc_search_single_pbc(float cutoff, const float* xyz, const uint64_t* ids, size_t n,
                                 const OrcBox* box, uint8_t pbc_dims, int nthreads); /* :928-954 */
OrcResult* orc_search_double(float cutoff, const float* xyz1, const uint64_t* ids1, size_t n1,
                             const float* xyz2, const uint64_t* ids2, size_t n2,
                             int nthreads);                                         /* :659-698 */
OrcResult* orc_search_double_pbc(float cutoff, const float* xyz1, const uint64_t* ids1, size_t n1,
                                 const float* xyz2, const uint64_t* ids2, size_t n2,
                                 const OrcBox* box, uint8_t pbc_dims, int nthreads); /* :713-754 */
OrcResult* orc_search_within(float cutoff, const float* xyz1, const uint64_t* ids1, size_t n1,
                             const float* xyz2, const uint64_t* ids2, size_t n2,
                             const float* lower3, const float* upper3, int nthreads); /* :519-558 */
OrcResult* orc_search_within_pbc(float cutoff, const float* xyz1, const uint64_t* ids1, size_t n1,
                                 const float* xyz2, const uint64_t* ids2, size_t n2,
                                 const OrcBox* box, uint8_t pbc_dims, int nthreads); /* :560-598 */
/* distance_search_single_pbc (:928-954) reduced to {count, sum, xor} of mix64((min(i,j)<<32)|max(i,j)) over every
   emitted pair — the hash of mb_pairs_checksum — without materialising the list (full-size parity of config 3). */
void orc_search_single_pbc_checksum(float cutoff, const float* xyz, const uint64_t* ids, size_t n,
                                    const OrcBox* box, uint8_t pbc_dims, int nthreads, uint64_t* out3,
                                    uint64_t* dims3 /* may be NULL */);
/* van der Waals search (distance_search.rs:767-879): per-pair cutoff vdw1[i]+vdw2[j]+EPSILON, grid cutoff
   max(vdw1)+max(vdw2)+EPSILON; vdw arrays are per SELECTED atom and the returned indices are LOCAL
   (position within the selection), as in the reference.  box==NULL or pbc_dims==0: non-periodic. */
OrcResult* orc_search_double_vdw(const float* xyz1, const uint64_t* ids1, size_t n1, const float* vdw1,
                                 const float* xyz2, const uint64_t* ids2, size_t n2, const float* vdw2,
                                 const OrcBox* box, uint8_t pbc_dims, int nthreads);
/* Modify::unwrap_connectivity_dim (modify.rs:72-131): contact graph at `cutoff` (periodic search, all dims), walk
   from the lowest unused atom, every reached atom moved to its closest image (`dims`) next to the atom it was
   reached from.  xyz is modified in place; roots_out[k] = position (within the selection) of the start atom of
   k's walk.  Returns the number of start atoms.  See the .cpp for what is and is not defined by the reference. */
int64_t orc_unwrap_connectivity(float cutoff, float* xyz, const uint64_t* ids, size_t n, const OrcBox* box,
                                uint8_t dims, int nthreads, int64_t* roots_out);
/* Measure::min_max (measure.rs:22-36) followed by the +-cutoff+EPS padding the `within`
   AST node applies (selection/ast.rs:598-600). */
void orc_within_bounds(float cutoff, const float* xyz, const uint64_t* ids, size_t n,
                       float* lower3, float* upper3);

/* ---- Measure / Modify (molar/src/measure.rs, modify.rs) ------------------- */
/* *_f32: default build (Float=f32).  *_f64: feature "f64" (aliases.rs:12-13) evaluated on
   the same f32 inputs promoted to f64 — "the reference f64 path".
   Return codes: 0 ok, 1 ZeroMass, 2 Sizes, 3 Svd. */
int orc_center_of_mass_f32(const float* xyz, const float* masses, const uint64_t* ids, size_t n, float* out3);
int orc_center_of_mass_f64(const float* xyz, const float* masses, const uint64_t* ids, size_t n, double* out3);
int orc_gyration_f32(const float* xyz, const float* masses, const uint64_t* ids, size_t n, float* out);
int orc_gyration_f64(const float* xyz, const float* masses, const uint64_t* ids, size_t n, double* out);
int orc_rmsd_f32(const float* xyz1, const uint64_t* ids1, size_t n1,
                 const float* xyz2, const uint64_t* ids2, size_t n2, float* out);
int orc_rmsd_f64(const float* xyz1, const uint64_t* ids1, size_t n1,
                 const float* xyz2, const uint64_t* ids2, size_t n2, double* out);
int orc_rmsd_mw_f32(const float* xyz1, const float* masses1, const uint64_t* ids1, size_t n1,
                    const float* xyz2, const uint64_t* ids2, size_t n2, float* out);
int orc_rmsd_mw_f64(const float* xyz1, const float* masses1, const uint64_t* ids1, size_t n1,
                    const float* xyz2, const uint64_t* ids2, size_t n2, double* out);
/* fit sel1 ONTO sel2 (measure.rs:507-535).  R9 column-major, t3: p' = R p + t. */
int orc_fit_transform_f32(const float* xyz1, const float* masses1, const uint64_t* ids1,
                          const float* xyz2, const float* masses2, const uint64_t* ids2,
                          size_t n, int at_origin, float* R9, float* t3);
int orc_fit_transform_f64(const float* xyz1, const float* masses1, const uint64_t* ids1,
                          const float* xyz2, const float* masses2, const uint64_t* ids2,
                          size_t n, int at_origin, double* R9, double* t3);
/* in place on xyz (modify.rs:32-36) */
void orc_apply_transform_f32(float* xyz, const uint64_t* ids, size_t n, const float* R9, const float* t3);
/* f64 transform applied in f64 to f32 inputs; out is n x 3 f64 (not in place) */
void orc_apply_transform_f64(const float* xyz, const uint64_t* ids, size_t n, const double* R9,
                             const double* t3, double* out);

/* ---- periodic variants, inertia (measure.rs:37-45,88-108,142-252,573-610) ----------------
   prec: 0 = the default f32 build, 1 = the f64 build, 2 = f32 per-atom image arithmetic with f64 sums
   (what the CUDA path is checked against bit-for-bit in its choice of images).  masses == NULL in
   orc_center_pbc selects center_of_geometry_pbc[_dims].  PARITY UNPINNED by the reference (no test
   exercises these); pinned against numpy in tests/test_oracle_measure_pbc.py. */
int orc_center_pbc(const float* xyz, const float* masses, const uint64_t* ids, size_t n, const OrcBox* box,
                   uint8_t dims, int prec, double* out3);
void orc_center_of_geometry(const float* xyz, const uint64_t* ids, size_t n, int prec, double* out3);
int orc_gyration_pbc(const float* xyz, const float* masses, const uint64_t* ids, size_t n, const OrcBox* box,
                     int prec, double* out);
/* box == NULL: inertia, else inertia_pbc.  tensor9 row-major; moments ascending; axes9 column-major
   (col2 = col0 x col1; eigenvector signs are the solver's); centre3 = the centre of mass used. */
int orc_inertia(const float* xyz, const float* masses, const uint64_t* ids, size_t n, const OrcBox* box, int prec,
                double* tensor9, double* moments3, double* axes9, double* centre3);

/* ---- synthetic frames — shared definition of the bench input ---- */
/* pos = M * s, s_axis = float(splitmix64(seed ^ (frame<<32) ^ (atom*3+axis)) >> 40) * 2^-24,
   unfused, a4 order.  masses = 1 + 15*s'. stray_permille: that many per 1000 atoms are
   displaced by +-1 box vector. */
void orc_synth_frame(uint64_t seed, uint64_t frame, size_t n_atoms, const float* m9_colmajor,
                     int stray_permille, float* xyz_out);
void orc_synth_masses(uint64_t seed, size_t n_atoms, float* masses_out);

#ifdef __cplusplus
}
#endif
#endif
