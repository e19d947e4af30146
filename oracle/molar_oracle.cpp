// molar_oracle.cpp — CPU restatement of MolAR's hot-path algorithms.
//
// TEST INFRASTRUCTURE ONLY (see molar_oracle.h).  Build with -ffp-contract=off: Rust never
// contracts a*b+c into an FMA, so every product and sum below must round separately.
//
// Every function cites the reference file:line (relative to /root/reference/) it restates.
// Third-party arithmetic that is NOT under /root/reference: nalgebra 0.34 (Cargo.toml:31,
// semver range, no Cargo.lock => patch version unpinned).  Its published algorithms are
// restated here:
//   * Matrix3 * Vector3  -> gemv column accumulation: ((M_i0*v0) + M_i1*v1) + M_i2*v2
//   * Vector3::norm_squared / dot -> (x*x + y*y) + z*z
//   * Matrix3::try_inverse -> cofactors / determinant (3x3 special case)
//   * SVD::new -> singular values sorted descending; here a one-sided Jacobi SVD (values agree
//     to rounding for non-degenerate covariance; the rotation is unique there).
#include "molar_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// Small linear algebra in the reference's evaluation order (molar/src/aliases.rs:10-27)
// ---------------------------------------------------------------------------------------------
template <class T>
struct V3 {
    T v[3];
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
};
template <class T>
struct M3 {
    T m[3][3];  // m[row][col]
};

template <class T>
inline V3<T> sub(const V3<T>& a, const V3<T>& b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
template <class T>
inline V3<T> add(const V3<T>& a, const V3<T>& b) { return {{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
template <class T>
inline V3<T> scale(T s, const V3<T>& a) { return {{s * a[0], s * a[1], s * a[2]}}; }
// nalgebra dot for 3-vectors: a + b + c with a=x*x, b=y*y, c=z*z  => (a+b)+c
template <class T>
inline T norm_squared(const V3<T>& a) {
    T x = a[0] * a[0], y = a[1] * a[1], z = a[2] * a[2];
    return (x + y) + z;
}
template <class T>
inline T norm(const V3<T>& a) { return std::sqrt(norm_squared(a)); }
// nalgebra gemv: y = col0*v0; y += col1*v1; y += col2*v2
template <class T>
inline V3<T> matvec(const M3<T>& M, const V3<T>& v) {
    V3<T> r;
    for (int i = 0; i < 3; ++i) {
        T acc = M.m[i][0] * v[0];
        acc = acc + M.m[i][1] * v[1];
        acc = acc + M.m[i][2] * v[2];
        r[i] = acc;
    }
    return r;
}
template <class T>
inline M3<T> matmul(const M3<T>& A, const M3<T>& B) {
    M3<T> R;
    for (int j = 0; j < 3; ++j) {
        V3<T> col = {{B.m[0][j], B.m[1][j], B.m[2][j]}};
        V3<T> r = matvec(A, col);
        for (int i = 0; i < 3; ++i) R.m[i][j] = r[i];
    }
    return R;
}
template <class T>
inline T det3(const M3<T>& A) {
    // nalgebra determinant() 3x3: e-minor expansion along first row
    T m11 = A.m[0][0], m12 = A.m[0][1], m13 = A.m[0][2];
    T m21 = A.m[1][0], m22 = A.m[1][1], m23 = A.m[1][2];
    T m31 = A.m[2][0], m32 = A.m[2][1], m33 = A.m[2][2];
    T minor_m12_m23 = m22 * m33 - m32 * m23;
    T minor_m11_m23 = m21 * m33 - m31 * m23;
    T minor_m11_m22 = m21 * m32 - m31 * m22;
    return m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
}
// nalgebra try_inverse, 3x3 special case (cofactors / determinant)
template <class T>
inline bool try_inverse(const M3<T>& A, M3<T>& out) {
    T m11 = A.m[0][0], m12 = A.m[0][1], m13 = A.m[0][2];
    T m21 = A.m[1][0], m22 = A.m[1][1], m23 = A.m[1][2];
    T m31 = A.m[2][0], m32 = A.m[2][1], m33 = A.m[2][2];
    T minor_m12_m23 = m22 * m33 - m32 * m23;
    T minor_m11_m23 = m21 * m33 - m31 * m23;
    T minor_m11_m22 = m21 * m32 - m31 * m22;
    T determinant = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    if (determinant == T(0)) return false;
    out.m[0][0] = minor_m12_m23 / determinant;
    out.m[0][1] = (m13 * m32 - m33 * m12) / determinant;
    out.m[0][2] = (m12 * m23 - m22 * m13) / determinant;
    out.m[1][0] = -minor_m11_m23 / determinant;
    out.m[1][1] = (m11 * m33 - m31 * m13) / determinant;
    out.m[1][2] = (m13 * m21 - m23 * m11) / determinant;
    out.m[2][0] = minor_m11_m22 / determinant;
    out.m[2][1] = (m12 * m31 - m32 * m11) / determinant;
    out.m[2][2] = (m11 * m22 - m21 * m12) / determinant;
    return true;
}

// Rust f32::round / f64::round: half away from zero == C roundf/round.
inline float rround(float x) { return ::roundf(x); }
inline double rround(double x) { return ::round(x); }
// Rust fract(): self - self.trunc()
inline float rfract(float x) { return x - ::truncf(x); }

constexpr uint8_t PBC_FULL = 7, PBC_NONE = 0;
inline bool get_dim(uint8_t p, int d) { return (p >> d) & 1; }

// ---------------------------------------------------------------------------------------------
// PeriodicBox — molar/src/periodic_box.rs
// ---------------------------------------------------------------------------------------------
template <class T>
struct BoxT {
    M3<T> matrix;
    M3<T> inv;
    std::vector<V3<T>> tric_corrections;
};

// periodic_box.rs:25-66
template <class T>
std::vector<V3<T>> build_tric_corrections(const M3<T>& m) {
    if (m.m[0][1] == 0 && m.m[0][2] == 0 && m.m[1][0] == 0 && m.m[1][2] == 0 && m.m[2][0] == 0 &&
        m.m[2][1] == 0)
        return {};
    V3<T> a = {{m.m[0][0], m.m[1][0], m.m[2][0]}};
    V3<T> b = {{m.m[0][1], m.m[1][1], m.m[2][1]}};
    V3<T> c = {{m.m[0][2], m.m[1][2], m.m[2][2]}};
    V3<T> na = scale(T(-1), a);  // -a (unary neg, exact)
    T n1 = norm(add(add(a, b), c));
    T n2 = norm(sub(add(a, b), c));
    T n3 = norm(add(sub(a, b), c));
    T n4 = norm(add(add(na, b), c));
    T half_diag = T(0.5) * std::max(std::max(std::max(n1, n2), n3), n4);
    T two_hd = T(2.0) * half_diag;
    T bound2 = two_hd * two_hd;  // powi(2)
    std::vector<V3<T>> out;
    out.reserve(26);
    for (int i = -1; i <= 1; ++i)
        for (int j = -1; j <= 1; ++j)
            for (int k = -1; k <= 1; ++k) {
                if (i == 0 && j == 0 && k == 0) continue;
                V3<T> s = add(add(scale(T(i), a), scale(T(j), b)), scale(T(k), c));
                if (norm_squared(s) < bound2) out.push_back(s);
            }
    return out;
}

// periodic_box.rs:156-176
template <class T>
bool box_from_matrix(const M3<T>& m, BoxT<T>& out) {
    for (int c = 0; c < 3; ++c) {
        V3<T> col = {{m.m[0][c], m.m[1][c], m.m[2][c]}};
        if (norm(col) == T(0)) return false;  // ZeroLengthVector
    }
    out.matrix = m;
    if (!try_inverse(m, out.inv)) return false;  // InverseFailed
    out.tric_corrections = build_tric_corrections(m);
    return true;
}

// periodic_box.rs:188-235
bool box_from_vectors_angles(float a, float b, float c, float alpha, float beta, float gamma,
                             BoxT<float>& out) {
    M3<float> m;
    std::memset(&m, 0, sizeof(m));
    if (a == 0.0f || b == 0.0f || c == 0.0f) return false;
    if (alpha < 60.0f || beta < 60.0f || gamma < 60.0f) return false;
    m.m[0][0] = a;
    if (alpha != 90.0f || beta != 90.0f || gamma != 90.0f) {
        // Rust to_radians(): self * (PI / 180)
        const float rads_per_deg = 3.14159265358979323846f / 180.0f;
        float cosa = alpha != 90.0f ? std::cos(alpha * rads_per_deg) : 0.0f;
        float cosb = beta != 90.0f ? std::cos(beta * rads_per_deg) : 0.0f;
        float sing = 1.0f, cosg = 0.0f;
        if (gamma != 90.0f) {
            sing = std::sin(gamma * rads_per_deg);
            cosg = std::cos(gamma * rads_per_deg);
        }
        m.m[0][1] = b * cosg;
        m.m[1][1] = b * sing;
        m.m[0][2] = c * cosb;
        m.m[1][2] = c * (cosa - cosb * cosg) / sing;
        m.m[2][2] = std::sqrt(c * c - std::pow(m.m[0][2], 2.0f) - std::pow(m.m[1][2], 2.0f));
    } else {
        m.m[1][1] = b;
        m.m[2][2] = c;
    }
    return box_from_matrix(m, out);
}

// periodic_box.rs:286-318
template <class T>
inline V3<T> shortest_vector_dims(const BoxT<T>& bx, const V3<T>& vec, uint8_t pbc_dims) {
    V3<T> box_vec = matvec(bx.inv, vec);
    for (int i = 0; i < 3; ++i)
        if (get_dim(pbc_dims, i)) box_vec[i] -= rround(box_vec[i]);
    V3<T> start = matvec(bx.matrix, box_vec);
    if (bx.tric_corrections.empty() || pbc_dims != PBC_FULL) return start;
    V3<T> best = start;
    T best2 = norm_squared(start);
    for (const auto& s : bx.tric_corrections) {
        V3<T> cand = add(start, s);
        T n2 = norm_squared(cand);
        if (n2 < best2) {
            best2 = n2;
            best = cand;
        }
    }
    return best;
}
// periodic_box.rs:379-381
template <class T>
inline T distance_squared(const BoxT<T>& bx, const V3<T>& p1, const V3<T>& p2, uint8_t pbc_dims) {
    return norm_squared(shortest_vector_dims(bx, sub(p2, p1), pbc_dims));
}
// periodic_box.rs:369-375
template <class T>
inline V3<T> lab_extents(const BoxT<T>& bx) {
    const auto& m = bx.matrix.m;
    return {{m[0][0] + m[0][1] + m[0][2], m[1][0] + m[1][1] + m[1][2], m[2][0] + m[2][1] + m[2][2]}};
}

// ---------------------------------------------------------------------------------------------
// Grid — molar/src/distance_search.rs:33-215
// ---------------------------------------------------------------------------------------------
using Vf = V3<float>;
using Boxf = BoxT<float>;

struct Entry {
    size_t id;
    const Vf* pos;
};

// Rust `x as usize` for a float: saturating, NaN -> 0
inline size_t f2usize(float x) {
    if (!(x > 0.0f)) return 0;  // also NaN
    if (x >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)x;
}
inline long f2isize(float x) {
    if (x != x) return 0;
    if (x >= 9223372036854775808.0f) return std::numeric_limits<long>::max();
    if (x <= -9223372036854775808.0f) return std::numeric_limits<long>::min();
    return (long)x;
}

struct Grid {
    std::vector<std::vector<Entry>> cells;
    size_t dims[3];
    std::vector<Vf> wrapped_pos;

    void init(const size_t d[3]) {
        for (int i = 0; i < 3; ++i) dims[i] = d[i];
        cells.assign(d[0] * d[1] * d[2], {});
        wrapped_pos.clear();
    }
    size_t loc_to_ind(const size_t loc[3]) const {  // :85-87
        return loc[0] + loc[1] * dims[0] + loc[2] * dims[0] * dims[1];
    }
    // :103-110
    static void dims_from_cutoff_and_extents(float cutoff, const Vf& extents, size_t sz[3]) {
        for (int d = 0; d < 3; ++d) {
            size_t v = f2usize(std::floor(extents[d] / cutoff));
            sz[d] = v < 1 ? 1 : v;
        }
    }
    // :120-142
    // local_ids: entries carry the LOCAL index k (the vdW searches iterate `0..vdw.len()`,
    // distance_search.rs:791-792,852-853) instead of the global atom index
    void populate(const float* xyz, const uint64_t* ids, size_t n, const Vf& lower, const Vf& upper,
                  bool local_ids = false) {
        Vf dim_sz = sub(upper, lower);
        for (size_t k = 0; k < n; ++k) {
            size_t gid = ids ? (size_t)ids[k] : k;
            size_t id = local_ids ? k : gid;
            const Vf* pos = reinterpret_cast<const Vf*>(xyz) + gid;
            size_t loc[3] = {0, 0, 0};
            bool skip = false;
            for (int d = 0; d < 3; ++d) {
                long nn = f2isize(std::floor((float)dims[d] * ((*pos)[d] - lower[d]) / dim_sz[d]));
                if (nn < 0 || nn >= (long)dims[d]) {
                    skip = true;
                    break;
                }
                loc[d] = (size_t)nn;
            }
            if (skip) continue;
            cells[loc_to_ind(loc)].push_back({id, pos});
        }
    }
    // :144-210
    void populate_pbc(const float* xyz, const uint64_t* ids, size_t n, const Boxf& bx, uint8_t pbc_dims,
                      bool local_ids = false) {
        std::vector<std::pair<size_t, size_t>> wrapped_ind;
        wrapped_pos.clear();
        wrapped_pos.reserve(64);
        for (size_t k = 0; k < n; ++k) {
            size_t gid = ids ? (size_t)ids[k] : k;
            size_t id = local_ids ? k : gid;
            const Vf* pos = reinterpret_cast<const Vf*>(xyz) + gid;
            Vf rel = matvec(bx.inv, *pos);
            size_t loc[3] = {0, 0, 0};
            bool correct = true, skip = false;
            for (int d = 0; d < 3; ++d) {
                if (rel[d] < 0.0f || rel[d] >= 1.0f) {
                    if (!get_dim(pbc_dims, d)) {
                        skip = true;  // continue 'outer
                        break;
                    } else {
                        correct = false;
                        break;
                    }
                }
            }
            if (skip) continue;
            if (correct) {
                for (int d = 0; d < 3; ++d) {
                    size_t v = f2usize(std::floor(rel[d] * (float)dims[d]));
                    loc[d] = std::min(v, dims[d] - 1);
                }
                cells[loc_to_ind(loc)].push_back({id, pos});
            } else {
                for (int d = 0; d < 3; ++d) {
                    if (get_dim(pbc_dims, d)) {
                        rel[d] = rfract(rel[d]);
                        if (rel[d] < 0.0f) rel[d] = 1.0f + rel[d];
                    }
                    size_t v = f2usize(std::floor(rel[d] * (float)dims[d]));
                    loc[d] = std::min(v, dims[d] - 1);
                }
                Vf wp = matvec(bx.matrix, rel);
                wrapped_pos.push_back(wp);
                wrapped_ind.push_back({loc_to_ind(loc), id});
            }
        }
        // wrapped_pos no longer grows: pointers are stable now (:203-209)
        for (size_t i = 0; i < wrapped_ind.size(); ++i)
            cells[wrapped_ind[i].first].push_back({wrapped_ind[i].second, &wrapped_pos[i]});
    }
};

// :39-60
const size_t MASK[14][2][3] = {
    {{0, 0, 0}, {0, 0, 0}}, {{0, 0, 0}, {1, 0, 0}}, {{0, 0, 0}, {0, 1, 0}}, {{0, 0, 0}, {0, 0, 1}},
    {{0, 0, 0}, {1, 1, 0}}, {{0, 0, 0}, {1, 0, 1}}, {{0, 0, 0}, {0, 1, 1}}, {{0, 0, 0}, {1, 1, 1}},
    {{1, 0, 0}, {0, 1, 0}}, {{1, 0, 0}, {0, 0, 1}}, {{0, 1, 0}, {0, 0, 1}}, {{1, 1, 0}, {0, 0, 1}},
    {{1, 0, 1}, {0, 1, 0}}, {{0, 1, 1}, {1, 0, 0}},
};

struct PlanEntry {
    size_t c1, c2;
    uint8_t wrapped;
};

// :217-269
std::vector<PlanEntry> search_plan(const Grid& grid1, const Grid* grid2, uint8_t pbc_dims) {
    std::vector<PlanEntry> plan;
    plan.reserve(14 * grid1.dims[0] * grid1.dims[1] * grid1.dims[2]);
    for (size_t x = 0; x < grid1.dims[0]; ++x)
        for (size_t y = 0; y < grid1.dims[1]; ++y)
            for (size_t z = 0; z < grid1.dims[2]; ++z)
                for (int mi = 0; mi < 14; ++mi) {
                    size_t c[2][3] = {
                        {x + MASK[mi][0][0], y + MASK[mi][0][1], z + MASK[mi][0][2]},
                        {x + MASK[mi][1][0], y + MASK[mi][1][1], z + MASK[mi][1][2]},
                    };
                    uint8_t wrapped = PBC_NONE;
                    bool drop = false;
                    for (int i = 0; i <= 1 && !drop; ++i)
                        for (int d = 0; d < 3; ++d)
                            if (c[i][d] == grid1.dims[d]) {
                                if (get_dim(pbc_dims, d)) {
                                    c[i][d] = 0;
                                    wrapped |= (uint8_t)(1u << d);
                                } else {
                                    drop = true;
                                    break;
                                }
                            }
                    if (drop) continue;
                    size_t i1 = grid1.loc_to_ind(c[0]);
                    size_t i2 = grid1.loc_to_ind(c[1]);
                    if (grid2) {
                        if ((!grid1.cells[i1].empty() && !grid2->cells[i2].empty()) ||
                            (!grid2->cells[i1].empty() && !grid1.cells[i2].empty()))
                            plan.push_back({i1, i2, wrapped});
                    } else if (!grid1.cells[i1].empty() && !grid1.cells[i2].empty()) {
                        plan.push_back({i1, i2, wrapped});
                    }
                }
    return plan;
}

struct Triple {
    size_t i, j;
    float d;
};

// :271-322 (pbc == nullptr => non-periodic variant :271-293)
void search_cell_pair_within(float cutoff2, const Grid& g1, const Grid& g2, size_t c1, size_t c2,
                             uint8_t wrapped, const Boxf* pbox, std::vector<size_t>& found) {
    const auto& a = g1.cells[c1];
    const auto& b = g2.cells[c2];
    for (size_t i = 0; i < a.size(); ++i) {
        for (size_t j = 0; j < b.size(); ++j) {
            float d2 = (pbox && wrapped != 0) ? distance_squared(*pbox, *a[i].pos, *b[j].pos, wrapped)
                                              : norm_squared(sub(*b[j].pos, *a[i].pos));
            if (d2 <= cutoff2) {
                found.push_back(a[i].id);
                break;
            }
        }
    }
}

// :324-373
void search_cell_pair_double(float cutoff2, const Grid& g1, const Grid& g2, size_t c1, size_t c2,
                             uint8_t wrapped, const Boxf* pbox, std::vector<Triple>& found) {
    const auto& a = g1.cells[c1];
    const auto& b = g2.cells[c2];
    for (size_t i = 0; i < a.size(); ++i)
        for (size_t j = 0; j < b.size(); ++j) {
            float d2 = (pbox && wrapped != 0) ? distance_squared(*pbox, *a[i].pos, *b[j].pos, wrapped)
                                              : norm_squared(sub(*b[j].pos, *a[i].pos));
            if (d2 <= cutoff2) found.push_back({a[i].id, b[j].id, std::sqrt(d2)});
        }
}

// :375-430  (vdw1/vdw2 indexed by the local ids stored in the grids)
void search_cell_pair_double_vdw(const Grid& g1, const Grid& g2, size_t c1, size_t c2, uint8_t wrapped,
                                 const float* vdw1, const float* vdw2, const Boxf* pbox, std::vector<Triple>& found) {
    const auto& a = g1.cells[c1];
    const auto& b = g2.cells[c2];
    for (size_t i = 0; i < a.size(); ++i)
        for (size_t j = 0; j < b.size(); ++j) {
            float d2 = (pbox && wrapped != 0) ? distance_squared(*pbox, *a[i].pos, *b[j].pos, wrapped)
                                              : norm_squared(sub(*b[j].pos, *a[i].pos));
            float cutoff = vdw1[a[i].id] + vdw2[b[j].id] + std::numeric_limits<float>::epsilon();
            if (d2 <= cutoff * cutoff) found.push_back({a[i].id, b[j].id, std::sqrt(d2)});
        }
}

// :432-517
void search_cell_pair_single(float cutoff2, const Grid& g, size_t c1, size_t c2, uint8_t wrapped,
                             const Boxf* pbox, std::vector<Triple>& found) {
    if (c1 == c2) {
        const auto& a = g.cells[c1];
        size_t n = a.size();
        for (size_t i = 0; i + 1 < n; ++i)
            for (size_t j = i + 1; j < n; ++j) {
                float d2 = (pbox && wrapped != 0)
                               ? distance_squared(*pbox, *a[i].pos, *a[j].pos, wrapped)
                               : norm_squared(sub(*a[j].pos, *a[i].pos));
                if (d2 <= cutoff2) found.push_back({a[i].id, a[j].id, std::sqrt(d2)});
            }
    } else {
        search_cell_pair_double(cutoff2, g, g, c1, c2, wrapped, pbox, found);
    }
}

// rayon `plan.into_par_iter().with_min_len(3).map(..).flatten().collect()` (:949-953):
// chunks of >=3 plan entries, per-task vectors, concatenated in plan order.
template <class Out, class Fn>
std::vector<Out> run_plan(const std::vector<PlanEntry>& plan, int nthreads, Fn&& fn) {
    std::vector<Out> all;
    if (nthreads <= 1 || plan.size() < 6) {
        for (const auto& p : plan) fn(p, all);
        return all;
    }
    const size_t chunk = std::max<size_t>(3, (plan.size() + (size_t)nthreads * 16 - 1) / ((size_t)nthreads * 16));
    const size_t nchunks = (plan.size() + chunk - 1) / chunk;
    std::vector<std::vector<Out>> parts(nchunks);
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            size_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            size_t b = c * chunk, e = std::min(plan.size(), b + chunk);
            for (size_t k = b; k < e; ++k) fn(plan[k], parts[c]);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    all.reserve(total);
    for (auto& p : parts) all.insert(all.end(), p.begin(), p.end());
    return all;
}

// :602-646
void compute_min_max(const float* xyz, const uint64_t* ids, size_t n, Vf& lower, Vf& upper) {
    lower = {{0, 0, 0}};
    upper = {{0, 0, 0}};
    for (size_t k = 0; k < n; ++k) {
        size_t id = ids ? (size_t)ids[k] : k;
        const float* p = xyz + 3 * id;
        for (int d = 0; d < 3; ++d) {
            if (p[d] < lower[d]) lower[d] = p[d];
            if (p[d] > upper[d]) upper[d] = p[d];
        }
    }
}
constexpr float F_EPS = std::numeric_limits<float>::epsilon();
void pad_bounds(float cutoff, Vf& l, Vf& u) {
    float dl = -cutoff - F_EPS, du = cutoff + F_EPS;
    for (int d = 0; d < 3; ++d) {
        l[d] = l[d] + dl;
        u[d] = u[d] + du;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
struct OrcBox {
    Boxf b;
};
struct OrcResult {
    std::vector<Triple> triples;
    std::vector<size_t> ids;
    size_t dims[3];
};

static M3<float> m3_from_colmajor(const float* m9) {
    M3<float> m;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) m.m[r][c] = m9[c * 3 + r];
    return m;
}

extern "C" {

OrcBox* orc_box_from_matrix(const float* m9) {
    auto* ob = new OrcBox;
    if (!box_from_matrix(m3_from_colmajor(m9), ob->b)) {
        delete ob;
        return nullptr;
    }
    return ob;
}
OrcBox* orc_box_from_vectors_angles(float a, float b, float c, float alpha, float beta, float gamma) {
    auto* ob = new OrcBox;
    if (!box_from_vectors_angles(a, b, c, alpha, beta, gamma, ob->b)) {
        delete ob;
        return nullptr;
    }
    return ob;
}
void orc_box_free(OrcBox* b) { delete b; }
void orc_box_get(const OrcBox* b, float* m9, float* i9) {
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            if (m9) m9[c * 3 + r] = b->b.matrix.m[r][c];
            if (i9) i9[c * 3 + r] = b->b.inv.m[r][c];
        }
}
int orc_box_corrections(const OrcBox* b, float* out) {
    if (out)
        for (size_t i = 0; i < b->b.tric_corrections.size(); ++i)
            for (int d = 0; d < 3; ++d) out[3 * i + d] = b->b.tric_corrections[i][d];
    return (int)b->b.tric_corrections.size();
}
void orc_box_lab_extents(const OrcBox* b, float* out3) {
    Vf e = lab_extents(b->b);
    for (int d = 0; d < 3; ++d) out3[d] = e[d];
}
void orc_box_shortest_vector(const OrcBox* b, const float* v3, uint8_t dims, float* out3) {
    Vf r = shortest_vector_dims(b->b, Vf{{v3[0], v3[1], v3[2]}}, dims);
    for (int d = 0; d < 3; ++d) out3[d] = r[d];
}
float orc_box_distance_squared(const OrcBox* b, const float* p1, const float* p2, uint8_t dims) {
    return distance_squared(b->b, Vf{{p1[0], p1[1], p1[2]}}, Vf{{p2[0], p2[1], p2[2]}}, dims);
}

size_t orc_result_len(const OrcResult* r) { return r->triples.empty() ? r->ids.size() : r->triples.size(); }
void orc_result_fill(const OrcResult* r, uint64_t* ij, float* d) {
    for (size_t k = 0; k < r->triples.size(); ++k) {
        if (ij) {
            ij[2 * k] = r->triples[k].i;
            ij[2 * k + 1] = r->triples[k].j;
        }
        if (d) d[k] = r->triples[k].d;
    }
}
void orc_result_fill_ids(const OrcResult* r, uint64_t* ids) {
    for (size_t k = 0; k < r->ids.size(); ++k) ids[k] = r->ids[k];
}
void orc_result_grid_dims(const OrcResult* r, uint64_t* dims3) {
    for (int d = 0; d < 3; ++d) dims3[d] = r->dims[d];
}
void orc_result_free(OrcResult* r) { delete r; }

// distance_search.rs:892-915
OrcResult* orc_search_single(float cutoff, const float* xyz, const uint64_t* ids, size_t n, int nthreads) {
    auto* res = new OrcResult;
    Vf lower, upper;
    compute_min_max(xyz, ids, n, lower, upper);
    pad_bounds(cutoff, lower, upper);
    Grid grid;
    size_t sz[3];
    Grid::dims_from_cutoff_and_extents(cutoff, sub(upper, lower), sz);
    grid.init(sz);
    grid.populate(xyz, ids, n, lower, upper);
    auto plan = search_plan(grid, nullptr, PBC_NONE);
    float c2 = cutoff * cutoff;
    res->triples = run_plan<Triple>(plan, nthreads, [&](const PlanEntry& p, std::vector<Triple>& out) {
        search_cell_pair_single(c2, grid, p.c1, p.c2, p.wrapped, nullptr, out);
    });
    for (int d = 0; d < 3; ++d) res->dims[d] = sz[d];
    return res;
}

// distance_search.rs:928-954
OrcResult* orc_search_single_pbc(float cutoff, const float* xyz, const uint64_t* ids, size_t n,
                                 const OrcBox* box, uint8_t pbc_dims, int nthreads) {
    auto* res = new OrcResult;
    Grid grid;
    size_t sz[3];
    Grid::dims_from_cutoff_and_extents(cutoff, lab_extents(box->b), sz);
    grid.init(sz);
    grid.populate_pbc(xyz, ids, n, box->b, pbc_dims);
    auto plan = search_plan(grid, nullptr, pbc_dims);
    float c2 = cutoff * cutoff;
    res->triples = run_plan<Triple>(plan, nthreads, [&](const PlanEntry& p, std::vector<Triple>& out) {
        search_cell_pair_single(c2, grid, p.c1, p.c2, p.wrapped, &box->b, out);
    });
    for (int d = 0; d < 3; ++d) res->dims[d] = sz[d];
    return res;
}

// Count + order-independent checksum of distance_search_single_pbc's output (distance_search.rs:928-954), computed
// inside the worker pool WITHOUT materialising the pair list, so that the 3.6e8 pairs of a 1M-atom frame can be
// compared with the CUDA path (mb_pairs_checksum / mb_batch_search checksums2: the same hash).  Every emitted
// triple is hashed as mix64((min(i,j) << 32) | max(i,j)); out3 = {count, sum of hashes mod 2^64, xor of hashes}.
static inline uint64_t orc_mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
void orc_search_single_pbc_checksum(float cutoff, const float* xyz, const uint64_t* ids, size_t n, const OrcBox* box,
                                    uint8_t pbc_dims, int nthreads, uint64_t* out3, uint64_t* dims3) {
    Grid grid;
    size_t sz[3];
    Grid::dims_from_cutoff_and_extents(cutoff, lab_extents(box->b), sz);
    grid.init(sz);
    grid.populate_pbc(xyz, ids, n, box->b, pbc_dims);
    auto plan = search_plan(grid, nullptr, pbc_dims);
    const float c2 = cutoff * cutoff;
    if (nthreads < 1) nthreads = 1;
    std::vector<uint64_t> acc((size_t)nthreads * 3, 0);
    std::atomic<size_t> next{0};
    auto worker = [&](int t) {
        uint64_t cnt = 0, sum = 0, x = 0;
        std::vector<Triple> buf;
        for (;;) {
            size_t b = next.fetch_add(3);  // chunks of 3 plan entries (with_min_len(3), :950)
            if (b >= plan.size()) break;
            size_t e = std::min(plan.size(), b + 3);
            for (size_t k = b; k < e; ++k) {
                buf.clear();
                search_cell_pair_single(c2, grid, plan[k].c1, plan[k].c2, plan[k].wrapped, &box->b, buf);
                for (const Triple& tr : buf) {
                    uint64_t lo = std::min(tr.i, tr.j), hi = std::max(tr.i, tr.j);
                    uint64_t h = orc_mix64((lo << 32) | hi);
                    ++cnt;
                    sum += h;
                    x ^= h;
                }
            }
        }
        acc[3 * t] = cnt;
        acc[3 * t + 1] = sum;
        acc[3 * t + 2] = x;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
    out3[0] = out3[1] = out3[2] = 0;
    for (int t = 0; t < nthreads; ++t) {
        out3[0] += acc[3 * t];
        out3[1] += acc[3 * t + 1];
        out3[2] ^= acc[3 * t + 2];
    }
    if (dims3)
        for (int d = 0; d < 3; ++d) dims3[d] = sz[d];
}

// Modify::unwrap_connectivity_dim (modify.rs:72-131) with SearchConnectivity::from_iter (connectivity.rs:18-37).
// xyz: whole frame, modified in place for the selected atoms.  roots_out[k] (k = position in the selection) = the
// position of the atom the walk that reached k started from (its own position for a start atom).  The reference
// returns one selection per start atom holding every atom reached from it — NOT the start atom itself
// (`sel_vec` only receives discovered atoms) and nothing for isolated atoms.  Returns the number of start atoms.
// The neighbour order inside an atom's list is the pair order of the search (serial plan order here; whatever
// rayon produced in the reference), so the spanning tree — and with it the last bits of the unwrapped coordinates —
// is not a defined quantity of the reference; the components are.
int64_t orc_unwrap_connectivity(float cutoff, float* xyz, const uint64_t* ids, size_t n, const OrcBox* box,
                                uint8_t dims, int nthreads, int64_t* roots_out) {
    std::vector<float> pos(3 * n);
    for (size_t k = 0; k < n; ++k) {
        size_t id = ids ? (size_t)ids[k] : k;
        for (int d = 0; d < 3; ++d) pos[3 * k + d] = xyz[3 * id + d];
    }
    OrcResult* r = orc_search_single_pbc(cutoff, pos.data(), nullptr, n, box, PBC_FULL, nthreads);
    std::vector<std::vector<size_t>> conn(n);
    for (const Triple& t : r->triples) {
        conn[t.i].push_back(t.j);
        conn[t.j].push_back(t.i);
    }
    orc_result_free(r);
    std::vector<char> used(n, 0);
    std::vector<size_t> todo;
    int64_t nstart = 0;
    size_t next_unused = 0;
    while (true) {
        while (next_unused < n && used[next_unused]) ++next_unused;  // find_position(|el| !el): lowest unused index
        if (next_unused >= n) break;
        const size_t root = next_unused;
        todo.push_back(root);
        used[root] = 1;
        roots_out[root] = (int64_t)root;
        ++nstart;
        while (!todo.empty()) {
            const size_t c = todo.back();
            todo.pop_back();
            const Vf p0 = {{pos[3 * c], pos[3 * c + 1], pos[3 * c + 2]}};
            for (size_t ind : conn[c]) {
                if (used[ind]) continue;
                const Vf p = {{pos[3 * ind], pos[3 * ind + 1], pos[3 * ind + 2]}};
                const Vf im = add(p0, shortest_vector_dims(box->b, sub(p, p0), dims));  // closest_image_dims
                for (int d = 0; d < 3; ++d) pos[3 * ind + d] = im[d];
                todo.push_back(ind);
                used[ind] = 1;
                roots_out[ind] = (int64_t)root;
            }
        }
    }
    for (size_t k = 0; k < n; ++k) {
        size_t id = ids ? (size_t)ids[k] : k;
        for (int d = 0; d < 3; ++d) xyz[3 * id + d] = pos[3 * k + d];
    }
    return nstart;
}

// distance_search.rs:659-698
OrcResult* orc_search_double(float cutoff, const float* xyz1, const uint64_t* ids1, size_t n1,
                             const float* xyz2, const uint64_t* ids2, size_t n2, int nthreads) {
    auto* res = new OrcResult;
    Vf l1, u1, l2, u2, l, u;
    compute_min_max(xyz1, ids1, n1, l1, u1);
    compute_min_max(xyz2, ids2, n2, l2, u2);
    for (int d = 0; d < 3; ++d) {
        l[d] = std::min(l1[d], l2[d]);
        u[d] = std::max(u1[d], u2[d]);
    }
    pad_bounds(cutoff, l, u);
    Grid g1, g2;
    size_t sz[3];
    Grid::dims_from_cutoff_and_extents(cutoff, sub(u, l), sz);
    g1.init(sz);
    g2.init(sz);
    g1.populate(xyz1, ids1, n1, l, u);
    g2.populate(xyz2, ids2, n2, l, u);
    auto plan = search_plan(g1, &g2, PBC_NONE);
    float c2 = cutoff * cutoff;
    res->triples = run_plan<Triple>(plan, nthreads, [&](const PlanEntry& p, std::vector<Triple>& out) {
        search_cell_pair_double(c2, g1, g2, p.c1, p.c2, p.wrapped, nullptr, out);
        search_cell_pair_double(c2, g1, g2, p.c2, p.c1, p.wrapped, nullptr, out);
    });
    for (int d = 0; d < 3; ++d) res->dims[d] = sz[d];
    return res;
}

// distance_search.rs:713-754
OrcResult* orc_search_double_pbc(float cutoff, const float* xyz1, const uint64_t* ids1, size_t n1,
                                 const float* xyz2, const uint64_t* ids2, size_t n2,
                                 const OrcBox* box, uint8_t pbc_dims, int nthreads) {
    auto* res = new OrcResult;
    Grid g1, g2;
    size_t sz[3];
    Grid::dims_from_cutoff_and_extents(cutoff, lab_extents(box->b), sz);
    g1.init(sz);
    g2.init(sz);
    g1.populate_pbc(xyz1, ids1, n1, box->b, pbc_dims);
    g2.populate_pbc(xyz2, ids2, n2, box->b, pbc_dims);
    auto plan = search_plan(g1, &g2, pbc_dims);
    float c2 = cutoff * cutoff;
    res->triples = run_plan<Triple>(plan, nthreads, [&](const PlanEntry& p, std::vector<Triple>& out) {
        search_cell_pair_double(c2, g1, g2, p.c1, p.c2, p.wrapped, &box->b, out);
        search_cell_pair_double(c2, g1, g2, p.c2, p.c1, p.wrapped, &box->b, out);
    });
    for (int d = 0; d < 3; ++d) res->dims[d] = sz[d];
    return res;
}

// distance_search.rs:519-558
OrcResult* orc_search_within(float cutoff, const float* xyz1, const uint64_t* ids1, size_t n1,
                             const float* xyz2, const uint64_t* ids2, size_t n2,
                             const float* lower3, const float* upper3, int nthreads) {
    auto* res = new OrcResult;
    Vf l = {{lower3[0], lower3[1], lower3[2]}}, u = {{upper3[0], upper3[1], upper3[2]}};
    Grid g1, g2;
    size_t sz[3];
    Grid::dims_from_cutoff_and_extents(cutoff, sub(u, l), sz);
    g1.init(sz);
    g2.init(sz);
    g1.populate(xyz1, ids1, n1, l, u);
    g2.populate(xyz2, ids2, n2, l, u);
    auto plan = search_plan(g1, &g2, PBC_NONE);
    float c2 = cutoff * cutoff;
    res->ids = run_plan<size_t>(plan, nthreads, [&](const PlanEntry& p, std::vector<size_t>& out) {
        search_cell_pair_within(c2, g1, g2, p.c1, p.c2, p.wrapped, nullptr, out);
        search_cell_pair_within(c2, g1, g2, p.c2, p.c1, p.wrapped, nullptr, out);
    });
    for (int d = 0; d < 3; ++d) res->dims[d] = sz[d];
    return res;
}

// distance_search.rs:560-598
OrcResult* orc_search_within_pbc(float cutoff, const float* xyz1, const uint64_t* ids1, size_t n1,
                                 const float* xyz2, const uint64_t* ids2, size_t n2,
                                 const OrcBox* box, uint8_t pbc_dims, int nthreads) {
    auto* res = new OrcResult;
    Grid g1, g2;
    size_t sz[3];
    Grid::dims_from_cutoff_and_extents(cutoff, lab_extents(box->b), sz);
    g1.init(sz);
    g2.init(sz);
    g1.populate_pbc(xyz1, ids1, n1, box->b, pbc_dims);
    g2.populate_pbc(xyz2, ids2, n2, box->b, pbc_dims);
    auto plan = search_plan(g1, &g2, pbc_dims);
    float c2 = cutoff * cutoff;
    res->ids = run_plan<size_t>(plan, nthreads, [&](const PlanEntry& p, std::vector<size_t>& out) {
        search_cell_pair_within(c2, g1, g2, p.c1, p.c2, p.wrapped, &box->b, out);
        search_cell_pair_within(c2, g1, g2, p.c2, p.c1, p.wrapped, &box->b, out);
    });
    for (int d = 0; d < 3; ++d) res->dims[d] = sz[d];
    return res;
}

// distance_search.rs:767-879 (box == NULL or pbc_dims == 0: the non-periodic variant :767-814)
OrcResult* orc_search_double_vdw(const float* xyz1, const uint64_t* ids1, size_t n1, const float* vdw1,
                                 const float* xyz2, const uint64_t* ids2, size_t n2, const float* vdw2,
                                 const OrcBox* box, uint8_t pbc_dims, int nthreads) {
    auto* res = new OrcResult;
    float m1 = vdw1[0], m2 = vdw2[0];
    for (size_t k = 1; k < n1; ++k) m1 = std::fmax(m1, vdw1[k]);  // iter().cloned().reduce(Float::max)
    for (size_t k = 1; k < n2; ++k) m2 = std::fmax(m2, vdw2[k]);
    const float cutoff = m1 + m2 + F_EPS;
    Grid g1, g2;
    size_t sz[3];
    const bool periodic = box && pbc_dims;
    Vf l, u;
    if (periodic) {
        Grid::dims_from_cutoff_and_extents(cutoff, lab_extents(box->b), sz);
    } else {
        Vf l1, u1, l2, u2;
        compute_min_max(xyz1, ids1, n1, l1, u1);
        compute_min_max(xyz2, ids2, n2, l2, u2);
        for (int d = 0; d < 3; ++d) {
            l[d] = std::min(l1[d], l2[d]);
            u[d] = std::max(u1[d], u2[d]);
        }
        pad_bounds(cutoff, l, u);
        Grid::dims_from_cutoff_and_extents(cutoff, sub(u, l), sz);
    }
    g1.init(sz);
    g2.init(sz);
    if (periodic) {
        g1.populate_pbc(xyz1, ids1, n1, box->b, pbc_dims, true);
        g2.populate_pbc(xyz2, ids2, n2, box->b, pbc_dims, true);
    } else {
        g1.populate(xyz1, ids1, n1, l, u, true);
        g2.populate(xyz2, ids2, n2, l, u, true);
    }
    auto plan = search_plan(g1, &g2, periodic ? pbc_dims : PBC_NONE);
    const Boxf* pb = periodic ? &box->b : nullptr;
    res->triples = run_plan<Triple>(plan, nthreads, [&](const PlanEntry& p, std::vector<Triple>& out) {
        search_cell_pair_double_vdw(g1, g2, p.c1, p.c2, p.wrapped, vdw1, vdw2, pb, out);
        search_cell_pair_double_vdw(g1, g2, p.c2, p.c1, p.wrapped, vdw1, vdw2, pb, out);
    });
    for (int d = 0; d < 3; ++d) res->dims[d] = sz[d];
    return res;
}

// measure.rs:22-36 + selection/ast.rs:598-600
void orc_within_bounds(float cutoff, const float* xyz, const uint64_t* ids, size_t n, float* lower3,
                       float* upper3) {
    Vf lower = {{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(),
                 std::numeric_limits<float>::max()}};
    // nalgebra Point::min_value() = Bounded::min_value = f32::MIN (most negative finite)
    Vf upper = {{std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(),
                 std::numeric_limits<float>::lowest()}};
    for (size_t k = 0; k < n; ++k) {
        size_t id = ids ? (size_t)ids[k] : k;
        const float* p = xyz + 3 * id;
        for (int d = 0; d < 3; ++d) {
            if (p[d] < lower[d]) lower[d] = p[d];
            if (p[d] > upper[d]) upper[d] = p[d];
        }
    }
    pad_bounds(cutoff, lower, upper);
    for (int d = 0; d < 3; ++d) {
        lower3[d] = lower[d];
        upper3[d] = upper[d];
    }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Measure / Modify — molar/src/measure.rs:60-87,485-643, modify.rs:32-36
// ---------------------------------------------------------------------------------------------
namespace {

template <class T>
inline V3<T> load_pos(const float* xyz, const uint64_t* ids, size_t k) {
    size_t id = ids ? (size_t)ids[k] : k;
    return {{(T)xyz[3 * id], (T)xyz[3 * id + 1], (T)xyz[3 * id + 2]}};
}
template <class T>
inline T load_mass(const float* masses, const uint64_t* ids, size_t k) {
    size_t id = ids ? (size_t)ids[k] : k;
    return (T)masses[id];
}

// measure.rs:60-75
template <class T>
int center_of_mass(const float* xyz, const float* masses, const uint64_t* ids, size_t n, V3<T>& out) {
    V3<T> cm = {{0, 0, 0}};
    T mass = 0;
    for (size_t k = 0; k < n; ++k) {
        V3<T> c = load_pos<T>(xyz, ids, k);
        T m = load_mass<T>(masses, ids, k);
        for (int d = 0; d < 3; ++d) cm[d] += c[d] * m;
        mass += m;
    }
    if (mass == T(0)) return 1;
    for (int d = 0; d < 3; ++d) out[d] = cm[d] / mass;
    return 0;
}

// measure.rs:78-87 + 561-570
template <class T>
int gyration(const float* xyz, const float* masses, const uint64_t* ids, size_t n, T& out) {
    V3<T> c;
    int rc = center_of_mass<T>(xyz, masses, ids, n, c);
    if (rc) return rc;
    T sd = 0, sm = 0;
    for (size_t k = 0; k < n; ++k) {
        V3<T> d = sub(load_pos<T>(xyz, ids, k), c);
        T m = load_mass<T>(masses, ids, k);
        sd += norm_squared(d) * m;
        sm += m;
    }
    out = std::sqrt(sd / sm);
    return 0;
}

// ---- periodic variants, inertia, principal axes ------------------------------------------------
// TI: precision in which the per-atom image is chosen (the reference's Float), TA: accumulator.
// (float,float) and (double,double) restate the two builds of the reference; (float,double) is the
// f32 build's per-atom arithmetic with noise-free sums — the statement the CUDA path is checked against.
template <class TI>
inline BoxT<TI> box_as(const Boxf& b) {
    M3<TI> m;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m.m[i][j] = (TI)b.matrix.m[i][j];
    BoxT<TI> out;
    box_from_matrix(m, out);
    return out;
}

// center_of_mass_pbc[_dims] (measure.rs:172-214) / center_of_geometry_pbc[_dims] (:142-168); masses == NULL: geometry.
// Keeps the reference's quirk: `let mut cm = p0.coords` — the first atom is not multiplied by its mass.
template <class TI, class TA>
int center_pbc(const float* xyz, const float* masses, const uint64_t* ids, size_t n, const Boxf& box, uint8_t dims,
               V3<TA>& out) {
    const BoxT<TI> b = box_as<TI>(box);
    const V3<TI> p0 = load_pos<TI>(xyz, ids, 0);
    V3<TA> cm = {{(TA)p0[0], (TA)p0[1], (TA)p0[2]}};
    TA mass = masses ? load_mass<TA>(masses, ids, 0) : TA(1);
    for (size_t k = 1; k < n; ++k) {
        const V3<TI> c = load_pos<TI>(xyz, ids, k);
        const V3<TI> im = add(p0, shortest_vector_dims(b, sub(c, p0), dims));  // closest_image_dims (periodic_box.rs:327-330)
        const TA m = masses ? load_mass<TA>(masses, ids, k) : TA(1);
        for (int d = 0; d < 3; ++d) cm[d] += (TA)im[d] * m;
        mass += m;
    }
    if (masses && mass == TA(0)) return 1;
    if (!masses) mass = (TA)n;
    for (int d = 0; d < 3; ++d) out[d] = cm[d] / mass;
    return 0;
}

// center_of_geometry (measure.rs:37-45)
template <class T>
void center_of_geometry(const float* xyz, const uint64_t* ids, size_t n, V3<T>& out) {
    V3<T> cog = {{0, 0, 0}};
    for (size_t k = 0; k < n; ++k) {
        V3<T> c = load_pos<T>(xyz, ids, k);
        for (int d = 0; d < 3; ++d) cog[d] += c[d];
    }
    for (int d = 0; d < 3; ++d) out[d] = cog[d] / (T)n;
}

// distances fed to do_gyration / do_inertia: pos - c, or shortest_vector(pos - c) for the _pbc variants;
// c is a Pos, i.e. it is rounded to the reference's Float (TI) before use
template <class TI, class TA, class F>
void for_each_dist(const float* xyz, const float* masses, const uint64_t* ids, size_t n, const Boxf* box,
                   const V3<TA>& centre, F&& f) {
    const V3<TI> c = {{(TI)centre[0], (TI)centre[1], (TI)centre[2]}};
    BoxT<TI> b;
    if (box) b = box_as<TI>(*box);
    for (size_t k = 0; k < n; ++k) {
        V3<TI> d = sub(load_pos<TI>(xyz, ids, k), c);
        if (box) d = shortest_vector_dims(b, d, PBC_FULL);
        f(V3<TA>{{(TA)d[0], (TA)d[1], (TA)d[2]}}, load_mass<TA>(masses, ids, k));
    }
}

// gyration_pbc (measure.rs:216-226)
template <class TI, class TA>
int gyration_pbc(const float* xyz, const float* masses, const uint64_t* ids, size_t n, const Boxf& box, TA& out) {
    V3<TA> c;
    int rc = center_pbc<TI, TA>(xyz, masses, ids, n, box, PBC_FULL, c);
    if (rc) return rc;
    TA sd = 0, sm = 0;
    for_each_dist<TI, TA>(xyz, masses, ids, n, &box, c, [&](const V3<TA>& d, TA m) {
        sd += norm_squared(d) * m;
        sm += m;
    });
    out = std::sqrt(sd / sm);
    return 0;
}

// cyclic Jacobi for a symmetric 3x3 (stands in for nalgebra::SymmetricEigen, measure.rs:590): eigenvalues
// on the diagonal of A, eigenvectors in the columns of V
template <class T>
void sym_eigen3(T A[3][3], T V[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = i == j ? T(1) : T(0);
    for (int sweep = 0; sweep < 64; ++sweep) {
        T off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
        T diag = std::fabs(A[0][0]) + std::fabs(A[1][1]) + std::fabs(A[2][2]);
        if (off == T(0) || off <= std::numeric_limits<T>::epsilon() * T(0.01) * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == T(0)) continue;
                T theta = (A[q][q] - A[p][p]) / (T(2) * A[p][q]);
                T t = (theta >= 0 ? T(1) : T(-1)) / (std::fabs(theta) + std::sqrt(theta * theta + T(1)));
                T cs = T(1) / std::sqrt(t * t + T(1)), sn = t * cs;
                for (int k = 0; k < 3; ++k) {
                    T akp = A[k][p], akq = A[k][q];
                    A[k][p] = cs * akp - sn * akq;
                    A[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    T apk = A[p][k], aqk = A[q][k];
                    A[p][k] = cs * apk - sn * aqk;
                    A[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    T vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = cs * vkp - sn * vkq;
                    V[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
}

// inertia / inertia_pbc (measure.rs:88-98,228-238) + do_inertia (:573-610).  tensor9 row-major, moments ascending,
// axes9 column-major with col2 = col0 x col1; the sign of col0/col1 is whatever the eigen-solver returns.
template <class TI, class TA>
int inertia(const float* xyz, const float* masses, const uint64_t* ids, size_t n, const Boxf* box, TA* tensor9,
            TA* moments3, TA* axes9, TA* centre3) {
    V3<TA> c;
    int rc = box ? center_pbc<TI, TA>(xyz, masses, ids, n, *box, PBC_FULL, c) : center_of_mass<TA>(xyz, masses, ids, n, c);
    if (rc) return rc;
    TA tens[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for_each_dist<TI, TA>(xyz, masses, ids, n, box, c, [&](const V3<TA>& d, TA m) {
        tens[0][0] += m * (d[1] * d[1] + d[2] * d[2]);
        tens[1][1] += m * (d[0] * d[0] + d[2] * d[2]);
        tens[2][2] += m * (d[0] * d[0] + d[1] * d[1]);
        tens[0][1] -= m * d[0] * d[1];
        tens[0][2] -= m * d[0] * d[2];
        tens[1][2] -= m * d[1] * d[2];
    });
    tens[1][0] = tens[0][1];
    tens[2][0] = tens[0][2];
    tens[2][1] = tens[1][2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) tensor9[3 * i + j] = tens[i][j];
    TA V[3][3];
    sym_eigen3(tens, V);
    int ord[3] = {0, 1, 2};
    std::sort(ord, ord + 3, [&](int a, int b) { return tens[a][a] < tens[b][b]; });
    for (int j = 0; j < 3; ++j) moments3[j] = tens[ord[j]][ord[j]];
    V3<TA> col0 = {{V[0][ord[0]], V[1][ord[0]], V[2][ord[0]]}}, col1 = {{V[0][ord[1]], V[1][ord[1]], V[2][ord[1]]}};
    TA n0 = norm(col0), n1 = norm(col1);
    for (int d = 0; d < 3; ++d) {
        col0[d] /= n0;
        col1[d] /= n1;
    }
    V3<TA> col2 = {{col0[1] * col1[2] - col0[2] * col1[1], col0[2] * col1[0] - col0[0] * col1[2],
                    col0[0] * col1[1] - col0[1] * col1[0]}};
    for (int d = 0; d < 3; ++d) {
        axes9[d] = col0[d];
        axes9[3 + d] = col1[d];
        axes9[6 + d] = col2[d];
        centre3[d] = c[d];
    }
    return 0;
}

// measure.rs:485-504
template <class T>
int rmsd(const float* xyz1, const uint64_t* ids1, size_t n1, const float* xyz2, const uint64_t* ids2,
         size_t n2, T& out) {
    if (n1 != n2) return 2;
    T res = 0;
    for (size_t k = 0; k < n1; ++k)
        res += norm_squared(sub(load_pos<T>(xyz2, ids2, k), load_pos<T>(xyz1, ids1, k)));
    out = std::sqrt(res / (T)n1);
    return 0;
}

// measure.rs:538-558
template <class T>
int rmsd_mw(const float* xyz1, const float* masses1, const uint64_t* ids1, size_t n1, const float* xyz2,
            const uint64_t* ids2, size_t n2, T& out) {
    if (n1 != n2) return 2;
    T res = 0, m_tot = 0;
    for (size_t k = 0; k < n1; ++k) {
        T m = load_mass<T>(masses1, ids1, k);
        res += norm_squared(sub(load_pos<T>(xyz2, ids2, k), load_pos<T>(xyz1, ids1, k))) * m;
        m_tot += m;
    }
    if (m_tot == T(0)) return 1;
    out = std::sqrt(res / m_tot);
    return 0;
}

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix: A = U * diag(s) * V^T, s sorted descending
// (nalgebra SVD::new sorts its singular values in descending order).
template <class T>
bool svd3(const M3<T>& A, M3<T>& U, T s[3], M3<T>& V) {
    T a[3][3], v[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            a[i][j] = A.m[i][j];
            v[i][j] = (i == j) ? T(1) : T(0);
            if (!std::isfinite(a[i][j])) return false;
        }
    const T eps = std::numeric_limits<T>::epsilon();
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                T alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) {
                    alpha += a[i][p] * a[i][p];
                    beta += a[i][q] * a[i][q];
                    gamma += a[i][p] * a[i][q];
                }
                if (gamma == T(0) || std::fabs(gamma) <= eps * std::sqrt(alpha * beta)) continue;
                rotated = true;
                T zeta = (beta - alpha) / (T(2) * gamma);
                T t = (zeta >= 0 ? T(1) : T(-1)) / (std::fabs(zeta) + std::sqrt(T(1) + zeta * zeta));
                T c = T(1) / std::sqrt(T(1) + t * t), sn = c * t;
                for (int i = 0; i < 3; ++i) {
                    T x = a[i][p], y = a[i][q];
                    a[i][p] = c * x - sn * y;
                    a[i][q] = sn * x + c * y;
                    x = v[i][p];
                    y = v[i][q];
                    v[i][p] = c * x - sn * y;
                    v[i][q] = sn * x + c * y;
                }
            }
        if (!rotated) break;
    }
    T sv[3];
    for (int j = 0; j < 3; ++j) sv[j] = std::sqrt(a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j]);
    int order[3] = {0, 1, 2};
    std::sort(order, order + 3, [&](int x, int y) { return sv[x] > sv[y]; });
    T u[3][3];
    for (int jj = 0; jj < 3; ++jj) {
        int j = order[jj];
        s[jj] = sv[j];
        for (int i = 0; i < 3; ++i) {
            V.m[i][jj] = v[i][j];
            u[i][jj] = sv[j] > 0 ? a[i][j] / sv[j] : T(0);
        }
    }
    // complete U to an orthonormal basis when trailing singular values vanish
    const T tiny = s[0] * eps * T(8);
    if (!(s[0] > 0)) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) u[i][j] = (i == j) ? T(1) : T(0);
    } else {
        if (s[1] <= tiny) {
            // any unit vector orthogonal to u0
            int k = 0;
            for (int i = 1; i < 3; ++i)
                if (std::fabs(u[i][0]) < std::fabs(u[k][0])) k = i;
            T e[3] = {0, 0, 0};
            e[k] = 1;
            T dot = u[k][0];
            T w[3], nn = 0;
            for (int i = 0; i < 3; ++i) {
                w[i] = e[i] - dot * u[i][0];
                nn += w[i] * w[i];
            }
            nn = std::sqrt(nn);
            for (int i = 0; i < 3; ++i) u[i][1] = w[i] / nn;
        }
        if (s[2] <= tiny) {
            // u2 = +-(u0 x u1); sign chosen so det(U) = det(V) => det(U V^T) = +1
            T cx = u[1][0] * u[2][1] - u[2][0] * u[1][1];
            T cy = u[2][0] * u[0][1] - u[0][0] * u[2][1];
            T cz = u[0][0] * u[1][1] - u[1][0] * u[0][1];
            M3<T> Vt = V;
            T dv = det3(Vt);
            T sg = dv < 0 ? T(-1) : T(1);
            u[0][2] = sg * cx;
            u[1][2] = sg * cy;
            u[2][2] = sg * cz;
        }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) U.m[i][j] = u[i][j];
    return true;
}

// measure.rs:613-643
template <class T>
int rot_transform(const float* xyz1, const uint64_t* ids1, const V3<T>& c1, const float* xyz2,
                  const uint64_t* ids2, const V3<T>& c2, const float* masses1, size_t n, M3<T>& R) {
    M3<T> cov;
    std::memset(&cov, 0, sizeof(cov));
    for (size_t k = 0; k < n; ++k) {
        V3<T> p1 = sub(load_pos<T>(xyz1, ids1, k), c1);
        V3<T> p2 = sub(load_pos<T>(xyz2, ids2, k), c2);
        T m = load_mass<T>(masses1, ids1, k);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) cov.m[i][j] += (p2[i] * p1[j]) * m;
    }
    M3<T> U, V;
    T s[3];
    if (!svd3(cov, U, s, V)) return 3;
    M3<T> Vt;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Vt.m[i][j] = V.m[j][i];
    T d = det3(matmul(U, Vt)) < T(0) ? T(-1) : T(1);
    M3<T> D;
    std::memset(&D, 0, sizeof(D));
    D.m[0][0] = 1;
    D.m[1][1] = 1;
    D.m[2][2] = d;
    R = matmul(matmul(U, D), Vt);
    return 0;
}

// measure.rs:507-535
template <class T>
int fit_transform(const float* xyz1, const float* masses1, const uint64_t* ids1, const float* xyz2,
                  const float* masses2, const uint64_t* ids2, size_t n, int at_origin, T* R9, T* t3) {
    V3<T> cm1 = {{0, 0, 0}}, cm2 = {{0, 0, 0}};
    if (!at_origin) {
        int rc = center_of_mass<T>(xyz1, masses1, ids1, n, cm1);
        if (rc) return rc;
        rc = center_of_mass<T>(xyz2, masses2, ids2, n, cm2);
        if (rc) return rc;
    }
    M3<T> R;
    int rc = rot_transform<T>(xyz1, ids1, cm1, xyz2, ids2, cm2, masses1, n, R);
    if (rc) return rc;
    // Translation(cm2) * R * Translation(-cm1): t = cm2 + R*(-cm1)
    V3<T> ncm1 = {{-cm1[0], -cm1[1], -cm1[2]}};
    V3<T> rt = matvec(R, ncm1);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) R9[c * 3 + r] = R.m[r][c];
    for (int d = 0; d < 3; ++d) t3[d] = at_origin ? T(0) : rt[d] + cm2[d];
    return 0;
}

inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline float unit_float(uint64_t u) { return (float)(u >> 40) * 5.9604644775390625e-08f; }  // 2^-24

}  // namespace

extern "C" {

int orc_center_of_mass_f32(const float* xyz, const float* masses, const uint64_t* ids, size_t n, float* out3) {
    V3<float> c;
    int rc = center_of_mass<float>(xyz, masses, ids, n, c);
    if (!rc) for (int d = 0; d < 3; ++d) out3[d] = c[d];
    return rc;
}
int orc_center_of_mass_f64(const float* xyz, const float* masses, const uint64_t* ids, size_t n, double* out3) {
    V3<double> c;
    int rc = center_of_mass<double>(xyz, masses, ids, n, c);
    if (!rc) for (int d = 0; d < 3; ++d) out3[d] = c[d];
    return rc;
}
// prec: 0 = (f32 images, f32 sums), 1 = (f64, f64), 2 = (f32 images, f64 sums).  Outputs are doubles.
int orc_center_pbc(const float* xyz, const float* masses, const uint64_t* ids, size_t n, const OrcBox* box,
                   uint8_t dims, int prec, double* out3) {
    int rc;
    if (prec == 0) {
        V3<float> c;
        rc = center_pbc<float, float>(xyz, masses, ids, n, box->b, dims, c);
        if (!rc) for (int d = 0; d < 3; ++d) out3[d] = c[d];
    } else if (prec == 1) {
        V3<double> c;
        rc = center_pbc<double, double>(xyz, masses, ids, n, box->b, dims, c);
        if (!rc) for (int d = 0; d < 3; ++d) out3[d] = c[d];
    } else {
        V3<double> c;
        rc = center_pbc<float, double>(xyz, masses, ids, n, box->b, dims, c);
        if (!rc) for (int d = 0; d < 3; ++d) out3[d] = c[d];
    }
    return rc;
}
void orc_center_of_geometry(const float* xyz, const uint64_t* ids, size_t n, int prec, double* out3) {
    if (prec == 0) {
        V3<float> c;
        center_of_geometry<float>(xyz, ids, n, c);
        for (int d = 0; d < 3; ++d) out3[d] = c[d];
    } else {
        V3<double> c;
        center_of_geometry<double>(xyz, ids, n, c);
        for (int d = 0; d < 3; ++d) out3[d] = c[d];
    }
}
int orc_gyration_pbc(const float* xyz, const float* masses, const uint64_t* ids, size_t n, const OrcBox* box, int prec,
                     double* out) {
    int rc;
    if (prec == 0) {
        float r = 0;
        rc = gyration_pbc<float, float>(xyz, masses, ids, n, box->b, r);
        *out = r;
    } else if (prec == 1) {
        rc = gyration_pbc<double, double>(xyz, masses, ids, n, box->b, *out);
    } else {
        rc = gyration_pbc<float, double>(xyz, masses, ids, n, box->b, *out);
    }
    return rc;
}
// box == NULL: inertia; else inertia_pbc.  tensor9 row-major, axes9 column-major, centre3 = the centre used.
int orc_inertia(const float* xyz, const float* masses, const uint64_t* ids, size_t n, const OrcBox* box, int prec,
                double* tensor9, double* moments3, double* axes9, double* centre3) {
    const Boxf* b = box ? &box->b : nullptr;
    if (prec == 0) {
        float t[9], m[3], a[9], c[3];
        int rc = inertia<float, float>(xyz, masses, ids, n, b, t, m, a, c);
        if (rc) return rc;
        for (int i = 0; i < 9; ++i) { tensor9[i] = t[i]; axes9[i] = a[i]; }
        for (int i = 0; i < 3; ++i) { moments3[i] = m[i]; centre3[i] = c[i]; }
        return 0;
    }
    if (prec == 1) return inertia<double, double>(xyz, masses, ids, n, b, tensor9, moments3, axes9, centre3);
    return inertia<float, double>(xyz, masses, ids, n, b, tensor9, moments3, axes9, centre3);
}
int orc_gyration_f32(const float* xyz, const float* masses, const uint64_t* ids, size_t n, float* out) {
    return gyration<float>(xyz, masses, ids, n, *out);
}
int orc_gyration_f64(const float* xyz, const float* masses, const uint64_t* ids, size_t n, double* out) {
    return gyration<double>(xyz, masses, ids, n, *out);
}
int orc_rmsd_f32(const float* xyz1, const uint64_t* ids1, size_t n1, const float* xyz2, const uint64_t* ids2,
                 size_t n2, float* out) {
    return rmsd<float>(xyz1, ids1, n1, xyz2, ids2, n2, *out);
}
int orc_rmsd_f64(const float* xyz1, const uint64_t* ids1, size_t n1, const float* xyz2, const uint64_t* ids2,
                 size_t n2, double* out) {
    return rmsd<double>(xyz1, ids1, n1, xyz2, ids2, n2, *out);
}
int orc_rmsd_mw_f32(const float* xyz1, const float* masses1, const uint64_t* ids1, size_t n1, const float* xyz2,
                    const uint64_t* ids2, size_t n2, float* out) {
    return rmsd_mw<float>(xyz1, masses1, ids1, n1, xyz2, ids2, n2, *out);
}
int orc_rmsd_mw_f64(const float* xyz1, const float* masses1, const uint64_t* ids1, size_t n1, const float* xyz2,
                    const uint64_t* ids2, size_t n2, double* out) {
    return rmsd_mw<double>(xyz1, masses1, ids1, n1, xyz2, ids2, n2, *out);
}
int orc_fit_transform_f32(const float* xyz1, const float* masses1, const uint64_t* ids1, const float* xyz2,
                          const float* masses2, const uint64_t* ids2, size_t n, int at_origin, float* R9,
                          float* t3) {
    return fit_transform<float>(xyz1, masses1, ids1, xyz2, masses2, ids2, n, at_origin, R9, t3);
}
int orc_fit_transform_f64(const float* xyz1, const float* masses1, const uint64_t* ids1, const float* xyz2,
                          const float* masses2, const uint64_t* ids2, size_t n, int at_origin, double* R9,
                          double* t3) {
    return fit_transform<double>(xyz1, masses1, ids1, xyz2, masses2, ids2, n, at_origin, R9, t3);
}
// modify.rs:32-36: p = tr * p = R*p + t
void orc_apply_transform_f32(float* xyz, const uint64_t* ids, size_t n, const float* R9, const float* t3) {
    M3<float> R = m3_from_colmajor(R9);
    for (size_t k = 0; k < n; ++k) {
        size_t id = ids ? (size_t)ids[k] : k;
        V3<float> p = {{xyz[3 * id], xyz[3 * id + 1], xyz[3 * id + 2]}};
        V3<float> r = matvec(R, p);
        for (int d = 0; d < 3; ++d) xyz[3 * id + d] = r[d] + t3[d];
    }
}
void orc_apply_transform_f64(const float* xyz, const uint64_t* ids, size_t n, const double* R9,
                             const double* t3, double* out) {
    M3<double> R;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) R.m[r][c] = R9[c * 3 + r];
    for (size_t k = 0; k < n; ++k) {
        V3<double> p = load_pos<double>(xyz, ids, k);
        V3<double> r = matvec(R, p);
        for (int d = 0; d < 3; ++d) out[3 * k + d] = r[d] + t3[d];
    }
}

void orc_synth_frame(uint64_t seed, uint64_t frame, size_t n_atoms, const float* m9, int stray_permille,
                     float* xyz_out) {
    M3<float> M = m3_from_colmajor(m9);
    for (size_t a = 0; a < n_atoms; ++a) {
        V3<float> s;
        for (int ax = 0; ax < 3; ++ax)
            s[ax] = unit_float(splitmix64(seed ^ (frame << 32) ^ (uint64_t)(a * 3 + ax)));
        if (stray_permille > 0) {
            uint64_t h = splitmix64(seed ^ (frame << 32) ^ 0x5BD1E995C0FFEEULL ^ ((uint64_t)a << 2));
            if ((int)(h % 1000) < stray_permille) {
                int dim = (int)((h >> 20) % 3);
                float sh = ((h >> 40) & 1) ? 1.0f : -1.0f;
                s[dim] = s[dim] + sh;
            }
        }
        V3<float> p = matvec(M, s);
        for (int d = 0; d < 3; ++d) xyz_out[3 * a + d] = p[d];
    }
}
void orc_synth_masses(uint64_t seed, size_t n_atoms, float* masses_out) {
    for (size_t a = 0; a < n_atoms; ++a) {
        float s = unit_float(splitmix64(seed ^ 0xA5A5A5A5DEADBEEFULL ^ (uint64_t)a));
        masses_out[a] = 1.0f + 15.0f * s;
    }
}

}  // extern "C"
