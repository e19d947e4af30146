// mb_search_lanes.cuh — "home atom per lane" pair enumeration (included by mb_search.cu, namespace mb).
//
// Same tiles, run table and candidate stream as search_cells_kernel, with the roles in the distance
// loop transposed:
//   * every LANE owns one home atom of the tile (position packed twice in registers);
//   * the 64 candidates of a step are staged in shared memory pair-packed, {x0,x1,y0,y1}{z0,z1,id0,id1},
//     so two LDS.128 broadcasts deliver the FADD2/FMUL2 operands of two distance tests per lane;
//   * a hit costs ONE predicated STS.32 and ONE predicated pointer bump: the candidate's id goes into the
//     lane's private column of a [row][lane] queue (bank = lane, conflict-free at any fill level) —
//     no hit masks, no bit extraction, no prefix scan per step;
//   * queue rows that are complete across the active lanes are written to the pair list as they are:
//     lane l stores (its home id, its k-th hit) at base + k*nh + l — one coalesced row per instruction.
//     The ragged remainder (lanes with more hits than the shortest column) is moved down and drained by
//     ballot compaction when the tile is finished.
// The decision taken for every pair is the one of search_cells_kernel (and of the reference): unfused f32
// (dx*dx + dy*dy) + dz*dz against cutoff^2; wrapped cell pairs on the shifted image with the exact
// PeriodicBox::distance_squared re-evaluation inside the rounding band; self tile by sorted index order.
#pragma once

constexpr int LANE_WARPS = 6;
constexpr int Q_ROWS = 64;          // queue rows per warp (u32 per lane and row)
constexpr int LANE_MAX_RUNS = 160;  // run table entries per pass

struct __align__(16) LaneShared {
    float4 cand[64];   // pair q: [2q] = {x0,x1,y0,y1}, [2q+1] = {z0,z1,bits(e0),bits(e1)}  (e = queue entry of the candidate)
    float4 orig[64];   // mixed steps: un-shifted position of candidate (q,half) at [2q+half]
    uint4 aux[32];     // mixed steps: {sorted index of a self-run candidate | ~0, same for the 2nd, wrap dims, wrap dims}
    unsigned rstart[LANE_MAX_RUNS];
    unsigned rpos[LANE_MAX_RUNS + 4];
    unsigned char rflag[LANE_MAX_RUNS];
    unsigned queue[Q_ROWS * 32];
};

__device__ __forceinline__ void sts128f(unsigned addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128u(unsigned addr, unsigned a, unsigned b, unsigned c, unsigned d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts32u(unsigned addr, unsigned a) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void lds128q(unsigned addr, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ uint4 lds128u(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds128f(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void upk2u(unsigned long long u, unsigned& lo, unsigned& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(u));
}

// Write the queue of one warp to the pair list.
//   rows [0, lmin) are complete across the nh active lanes: row k goes to base + k*nh + lane as it is;
//   the rest is either moved down (all == false and it still leaves room for 32 more hits per lane) or
//   drained row by row with ballot compaction.
// MODE 1 entries are sorted-array indices (| wrap dims << 28): the id and the distance are looked up here.
template <int MODE>
__device__ __noinline__ unsigned lane_drain(const SearchParams& P, unsigned colbase, unsigned qptr, unsigned hid, float hx,
                                            float hy, float hz, int nh, unsigned lane, bool all) {
    __syncwarp();
    const unsigned len = (qptr - colbase) >> 7;
    const bool act = lane < (unsigned)nh;
    const unsigned lmin = __reduce_min_sync(0xffffffffu, act ? len : 0xffffu);
    const unsigned lmax = __reduce_max_sync(0xffffffffu, len);
    auto emit = [&](unsigned long long slot, unsigned e) {
        if (MODE == 1) {
            const float4 cnd = __ldg(&P.sortedB[e & 0x0fffffffu]);
            const unsigned w = e >> 28;
            const float d2 = w ? d2_pbc_call(P.g.box, hx, hy, hz, cnd.x, cnd.y, cnd.z, w) : d2_direct(hx, hy, hz, cnd.x, cnd.y, cnd.z);
            P.pairs[slot] = make_uint2(hid, __float_as_uint(cnd.w));
            P.dists[slot] = __fsqrt_rn(d2);
        } else {
            P.pairs[slot] = make_uint2(hid, e);
        }
    };
    if (lmin) {
        const unsigned long long n = (unsigned long long)lmin * (unsigned)nh;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.counter, n);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (act && base + n <= P.pair_cap) {
            unsigned long long slot = base + lane;
            unsigned sa = colbase;
            unsigned k = 0;
            for (; k + 4 <= lmin; k += 4) {
                const unsigned e0 = lds32(sa), e1 = lds32(sa + 128u), e2 = lds32(sa + 256u), e3 = lds32(sa + 384u);
                emit(slot, e0);
                emit(slot + (unsigned)nh, e1);
                emit(slot + 2u * (unsigned)nh, e2);
                emit(slot + 3u * (unsigned)nh, e3);
                slot += 4u * (unsigned)nh;
                sa += 512u;
            }
            for (; k < lmin; ++k) {
                emit(slot, lds32(sa));
                slot += (unsigned)nh;
                sa += 128u;
            }
        }
    }
    const unsigned rest = len - (act ? lmin : 0u);  // inactive lanes never hold entries
    if (!all && lmax - lmin <= (unsigned)(Q_ROWS - 32)) {
        // move the ragged remainder to the top of the column (each lane its own bank)
        if (lmin) {
            unsigned src = colbase + lmin * 128u, dst = colbase;
            for (unsigned k = 0; k < rest; ++k) {
                sts32u(dst, lds32(src));
                src += 128u;
                dst += 128u;
            }
            qptr = colbase + rest * 128u;
        }
        __syncwarp();
        return qptr;
    }
    // ragged drain: one ballot-compacted row at a time
    const unsigned tot = __reduce_add_sync(0xffffffffu, rest);
    if (tot) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.counter, (unsigned long long)tot);
        base = __shfl_sync(0xffffffffu, base, 0);
        const bool room = base + tot <= P.pair_cap;
        unsigned sa = colbase + lmin * 128u;
        const unsigned nrow = lmax - lmin;
        for (unsigned k = 0; k < nrow; ++k) {
            const bool v = k < rest;
            const unsigned m = __ballot_sync(0xffffffffu, v);
            if (v && room) emit(base + __popc(m & ((1u << lane) - 1u)), lds32(sa));
            base += __popc(m);
            sa += 128u;
        }
    }
    __syncwarp();
    return colbase;
}

// MODE: 0 pairs, 1 pairs + distances, 2 count only, 3 `within` flags (two sets)
template <int MODE>
__global__ void __launch_bounds__(LANE_WARPS * 32, 3) search_lanes_kernel(const __grid_constant__ SearchParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    LaneShared& ws = reinterpret_cast<LaneShared*>(smem_raw)[wid];
    const unsigned cand_sa = smem_addr(ws.cand), orig_sa = smem_addr(ws.orig), aux_sa = smem_addr(ws.aux);
    const unsigned colbase = smem_addr(ws.queue) + lane * 4u;
    unsigned qptr = colbase;
    unsigned long long count = 0;
    const GridSpec& g = P.g;
    const int fdx = g.fd[0], fdy = g.fd[1], fdz = g.fd[2];
    const int hx = g.hx, tdx = fdx / hx;
    const unsigned ntiles = (unsigned)(tdx * fdy * fdz);
    const float rc2 = P.rc2;
    const float qnan = __int_as_float(0x7fc00000);
    const float finf = __int_as_float(0x7f800000);
    const float loW = P.fast_pbc ? P.rc2_lo : -1.0f, hiW = P.fast_pbc ? P.rc2_hi : finf;

    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = (unsigned)atomicAdd(P.counter + 1, 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) break;
        const int fx = (int)(tile % (unsigned)tdx) * hx, fy = (int)((tile / (unsigned)tdx) % (unsigned)fdy),
                  fz = (int)(tile / (unsigned)(tdx * fdy));
        const unsigned cell0 = (unsigned)(fx + fdx * (fy + fdy * fz));
        const unsigned hs = P.cell_start[cell0], he = P.cell_start[cell0 + hx];
        if (hs == he) continue;
        const int cx = div_k(fx, g.k[0], g.kmagic[0]), cy = div_k(fy, g.k[1], g.kmagic[1]),
                  cz = div_k(fz, g.k[2], g.kmagic[2]);

        int row0 = 0;
        bool first = true;
#pragma unroll 1
        do {
            // ---------------- Phase A: run table (shared with search_cells_kernel) ----------------
            unsigned nr, T;
            fill_run_table<LANE_MAX_RUNS>(P, ws.rstart, ws.rpos, ws.rflag, fx, fy, fz, cx, cy, cz, hs, he, first, lane, row0, nr,
                                          T);

            // ---------------- Phase B: every lane tests its home atom against the stream ----------------
#pragma unroll 1
            for (unsigned hb = hs; hb < he; hb += 32) {
                const int nh = min(32u, he - hb);
                const float4 h = (lane < (unsigned)nh) ? __ldg(&P.sorted[hb + lane]) : make_float4(qnan, qnan, qnan, 0.f);
                const unsigned hid = __float_as_uint(h.w);
                const unsigned myidx = hb + lane;
                const unsigned long long hx2 = pk2_once(h.x, h.x), hy2 = pk2_once(h.y, h.y), hz2 = pk2_once(h.z, h.z);
                unsigned cnt32 = 0;  // MODE 2
                bool found = false;  // MODE 3
                unsigned cur0 = 0, cur1 = 0;
                float4 n0, n1;
                unsigned f0, f1, a0i, a1i;
                auto fetch = [&](unsigned c0) {
                    const unsigned p0 = c0 + lane, p1 = p0 + 32;
                    n0 = make_float4(qnan, qnan, qnan, 0.f);
                    n1 = n0;
                    f0 = f1 = a0i = a1i = 0;
                    if (p0 < T) {
                        while (ws.rpos[cur0 + 1] <= p0) ++cur0;
                        a0i = ws.rstart[cur0] + (p0 - ws.rpos[cur0]);
                        f0 = ws.rflag[cur0];
                        n0 = __ldg(&P.sortedB[a0i]);
                    }
                    if (p1 < T) {
                        cur1 = max(cur1, cur0);
                        while (ws.rpos[cur1 + 1] <= p1) ++cur1;
                        a1i = ws.rstart[cur1] + (p1 - ws.rpos[cur1]);
                        f1 = ws.rflag[cur1];
                        n1 = __ldg(&P.sortedB[a1i]);
                    }
                };
                if (T) fetch(0);
#pragma unroll 1
                for (unsigned c0 = 0; c0 < T; c0 += 64) {
                    // ---- stage this step's candidates (pair q = this lane's two candidates) ----
                    __syncwarp();
                    const bool mixed = __any_sync(0xffffffffu, (f0 | f1) != 0u);
                    // queue entry of a candidate: its id, or (MODE 1) its sorted index | wrap dims << 28
                    const unsigned e0 = MODE == 1 ? (a0i | ((f0 & 7u) << 28)) : __float_as_uint(n0.w);
                    const unsigned e1 = MODE == 1 ? (a1i | ((f1 & 7u) << 28)) : __float_as_uint(n1.w);
                    if (!mixed) {
                        sts128f(cand_sa + lane * 32u, n0.x, n1.x, n0.y, n1.y);
                        sts128f(cand_sa + lane * 32u + 16u, n0.z, n1.z, __uint_as_float(e0), __uint_as_float(e1));
                    } else {
                        // wrapped runs are tested on the lattice-shifted image (see search_cells_kernel)
                        float sx0 = 0.f, sy0 = 0.f, sz0 = 0.f, sx1 = 0.f, sy1 = 0.f, sz1 = 0.f;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const float bx = P.g.box.m[d], by = P.g.box.m[3 + d], bz = P.g.box.m[6 + d];
                            if ((f0 >> d) & 1u) {
                                const float sg = ((f0 >> (3 + d)) & 1u) ? -1.0f : 1.0f;
                                sx0 += sg * bx; sy0 += sg * by; sz0 += sg * bz;
                            }
                            if ((f1 >> d) & 1u) {
                                const float sg = ((f1 >> (3 + d)) & 1u) ? -1.0f : 1.0f;
                                sx1 += sg * bx; sy1 += sg * by; sz1 += sg * bz;
                            }
                        }
                        sts128f(cand_sa + lane * 32u, xadd(n0.x, sx0), xadd(n1.x, sx1), xadd(n0.y, sy0), xadd(n1.y, sy1));
                        sts128f(cand_sa + lane * 32u + 16u, xadd(n0.z, sz0), xadd(n1.z, sz1), __uint_as_float(e0),
                                __uint_as_float(e1));
                        sts128f(orig_sa + lane * 32u, n0.x, n0.y, n0.z, 0.f);
                        sts128f(orig_sa + lane * 32u + 16u, n1.x, n1.y, n1.z, 0.f);
                        sts128u(aux_sa + lane * 16u, (f0 & RUN_SELF) ? a0i : 0xffffffffu, (f1 & RUN_SELF) ? a1i : 0xffffffffu,
                                f0 & 7u, f1 & 7u);
                    }
                    // prefetch the next step while this one is tested
                    if (c0 + 64 < T) fetch(c0 + 64);
                    __syncwarp();
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        if (MODE == 0 || MODE == 1) {
                            if (__reduce_max_sync(0xffffffffu, qptr - colbase) > (unsigned)((Q_ROWS - 32) * 128))
                                qptr = lane_drain<MODE>(P, colbase, qptr, hid, h.x, h.y, h.z, nh, lane, false);
                        }
                        unsigned sa = cand_sa + (unsigned)half * 512u;
                        if (!mixed) {
                            // software pipeline: operands of pair q+2 are requested before pair q is tested
                            unsigned long long A0, B0, C0, D0, A1, B1, C1, D1;
                            lds128q(sa, A0, B0);
                            lds128q(sa + 16u, C0, D0);
                            lds128q(sa + 32u, A1, B1);
                            lds128q(sa + 48u, C1, D1);
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                unsigned long long A2 = 0, B2 = 0, C2 = 0, D2 = 0;
                                if (q + 2 < 16) {
                                    lds128q(sa + (unsigned)(q + 2) * 32u, A2, B2);
                                    lds128q(sa + (unsigned)(q + 2) * 32u + 16u, C2, D2);
                                }
                                const unsigned long long dx = sub2(A0, hx2), dy = sub2(B0, hy2), dz = sub2(C0, hz2);
                                const unsigned long long xx = mul2(dx, dx), yy = mul2(dy, dy), zz = mul2(dz, dz);
                                float x0, x1, y0, y1, z0, z1;
                                upk2(xx, x0, x1);
                                upk2(yy, y0, y1);
                                upk2(zz, z0, z1);
                                const float d0 = xadd(xadd(x0, y0), z0), d1 = xadd(xadd(x1, y1), z1);
                                unsigned id0, id1;
                                upk2u(D0, id0, id1);
                                if (MODE == 0 || MODE == 1) {
                                    if (d0 <= rc2) { sts32u(qptr, id0); qptr += 128u; }
                                    if (d1 <= rc2) { sts32u(qptr, id1); qptr += 128u; }
                                } else if (MODE == 2) {
                                    cnt32 += (d0 <= rc2 ? 1u : 0u) + (d1 <= rc2 ? 1u : 0u);
                                } else {
                                    found |= (d0 <= rc2) | (d1 <= rc2);
                                }
                                A0 = A1; B0 = B1; C0 = C1; D0 = D1;
                                A1 = A2; B1 = B2; C1 = C2; D1 = D2;
                            }
                        } else {
                            // mixed step: self run (index-order filter) and/or wrapped runs (band + exact re-evaluation)
#pragma unroll 1
                            for (int q = 0; q < 16; ++q) {
                                const unsigned pq = (unsigned)(half * 16 + q);
                                unsigned long long A0, B0, C0, D0;
                                lds128q(sa + (unsigned)q * 32u, A0, B0);
                                lds128q(sa + (unsigned)q * 32u + 16u, C0, D0);
                                const uint4 ax = lds128u(aux_sa + pq * 16u);
                                const unsigned long long dx = sub2(A0, hx2), dy = sub2(B0, hy2), dz = sub2(C0, hz2);
                                const unsigned long long xx = mul2(dx, dx), yy = mul2(dy, dy), zz = mul2(dz, dz);
                                float x0, x1, y0, y1, z0, z1;
                                upk2(xx, x0, x1);
                                upk2(yy, y0, y1);
                                upk2(zz, z0, z1);
                                const float d0 = xadd(xadd(x0, y0), z0), d1 = xadd(xadd(x1, y1), z1);
                                bool hit0 = d0 <= (ax.z ? loW : rc2), hit1 = d1 <= (ax.w ? loW : rc2);
                                if (!hit0 && d0 <= (ax.z ? hiW : rc2)) {
                                    const float4 o = lds128f(orig_sa + pq * 32u);
                                    hit0 = d2_pbc_call(P.g.box, h.x, h.y, h.z, o.x, o.y, o.z, ax.z) <= rc2;
                                }
                                if (!hit1 && d1 <= (ax.w ? hiW : rc2)) {
                                    const float4 o = lds128f(orig_sa + pq * 32u + 16u);
                                    hit1 = d2_pbc_call(P.g.box, h.x, h.y, h.z, o.x, o.y, o.z, ax.w) <= rc2;
                                }
                                hit0 = hit0 && ax.x > myidx;  // home tile against itself: each pair once
                                hit1 = hit1 && ax.y > myidx;
                                unsigned id0, id1;
                                upk2u(D0, id0, id1);
                                if (MODE == 0 || MODE == 1) {
                                    if (hit0) { sts32u(qptr, id0); qptr += 128u; }
                                    if (hit1) { sts32u(qptr, id1); qptr += 128u; }
                                } else if (MODE == 2) {
                                    cnt32 += (hit0 ? 1u : 0u) + (hit1 ? 1u : 0u);
                                } else {
                                    found |= hit0 | hit1;
                                }
                            }
                        }
                    }
                }
                if (MODE == 0 || MODE == 1) {
                    if (__any_sync(0xffffffffu, qptr != colbase))
                        qptr = lane_drain<MODE>(P, colbase, qptr, hid, h.x, h.y, h.z, nh, lane, true);
                } else if (MODE == 2) {
                    count += cnt32;
                } else {
                    if (found && lane < (unsigned)nh) P.flags[hid] = 1;
                }
            }
            first = false;
        } while (row0 < P.nrows);
    }
    if (MODE == 2) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
        if (lane == 0 && count) atomicAdd(P.counter, count);
    }
}
