// sphrt_coop.cuh -- warp-cooperative walker: 32 lanes share ONE ray at a time.
//
// The per-thread walkers (sphrt_walker.cuh / sphrt_fast.cuh) merge three lazy event streams
// inside one thread: serial float64 dependency chains, three generators that diverge
// between the lanes of a warp, and a consumer that spends more instructions on queue
// bookkeeping than on the segment itself (ncu, profiles/README.md).  For grids with
// hundreds of crossings per ray the work is turned by 90 degrees here:
//
//   phase 0  (one lane per ray, 32 rays per warp)  ray constants, classification, and the
//            index RANGES of the shells / cones / half-planes the ray can meet between its
//            entry into and exit from the outermost shell; a 76-byte record per ray in
//            shared memory.
//   phase 1  (32 lanes per ray, one ray after the other)
//            a. every lane evaluates one candidate boundary with the reference's exact
//               arithmetic (the same radial / cone / plane primitives as the other walkers);
//               valid roots inside the window are compacted by ballot into three SORTED
//               per-warp lists in shared memory.  Sorted by construction: shells are met
//               outermost -> innermost -> outermost, the colatitude along a line is
//               unimodal (cones are listed from the turning point outwards, first crossings
//               backwards and second crossings forwards from the middle of the buffer),
//               the azimuth is monotone;
//            b. the lists are VERIFIED (ascending t, clean guard candidates at both ends of
//               every range); a ray that fails any check is handed to the per-thread walker
//               through a bitmask -- the structural shortcuts are never trusted blindly;
//            c. the lists are merged with two merge-path passes (elevation + azimuth, then
//               radial + that): every lane co-ranks the start of its equal share of the
//               merged sequence and merges it sequentially, carrying the voxel state; every
//               merged event starts one segment that ends at the next event -- the
//               reference's sort + forward-fill + diff (raytracer.py:131-151).  Gather /
//               scatter run lane-parallel, one share of the ray per lane.
//
// Equal distances in different families are ordered radial < elevation < azimuth; the
// reference's order of exact ties is unspecified (unstable sort) and
// irrelevant for the result (zero-length slots).
//
// Device only (warp votes and shuffles).  Handles rays that start outside the outermost
// shell and not on a cone / half-plane, on grids whose azimuth covers the full circle; every
// other ray takes the per-thread walker.
#pragma once
#include "sphrt_fast.cuh"

namespace sphrt {

#define SPHRT_FULL 0xffffffffu

// Shared-memory layout of one warp (byte offsets from the warp's base), filled in by the host.
struct CoopPlan {
    int capR, capE, capA;     // event capacity of the three lists
    int off_tR, off_tE, off_tA;   // double arrays, capacity + 1 (sentinel)
    int off_rE, off_rA;           // int16 arrays, capacity + 1 (leading slot = state at entry)
    int capEA, off_tEA, off_sEA;  // merged elevation + azimuth list: t, int32 packed state
    int off_rec;              // per-ray records: 6 double fields, [field][lane]
    int warp_bytes;
    int warps;                // warps per block
    int table_doubles;
};

enum { COOP_MISS = 0, COOP_RAY = 1, COOP_LISTED = 2, COOP_TRACKED = 3 };

struct CoopRecord {           // what phase 1 needs to know about a ray (warp-uniform there)
    double d0, d1, d2, tc, dperp2, t_turn;
    int imin;                 // innermost shell reached
    int e_first, e_count, e_step;     // cone candidates e_first + k * e_step, k < e_count
    int e_mid;                // slot of the elevation buffer where the two runs meet
    int a_first, a_count, a_dir;      // plane candidates, in the direction of rotation
    int se0, sa0;             // elevation / azimuth state at the entry point
};

// The kernel is sensitive to its code size (the per-ray loop is about as large as the 32 KB
// instruction cache): atan2 is only used by the planning phase and kept out of line, and the loops
// marked "rolled" below are kept from being unrolled.
static __device__ __noinline__ double coop_atan2(double y, double x) { return atan2(y, x); }

// geometric bins and reference states of a point, as start_voxel()
__device__ __forceinline__ void coop_locate_point(const GridView& g, double p0, double p1, double p2, int bins[3],
                                  int st[3], double sph[3])
{
    const double xy = p0 * p0 + p1 * p1;
    sph[0] = sqrt(xy + p2 * p2);
    sph[1] = coop_atan2(sqrt(xy), p2);
    sph[2] = coop_atan2(p1, p0);
    bins[0] = locate(g.r_b(), g.nr, sph[0], &st[0]);
    bins[1] = locate(g.e_b(), g.ne, sph[1], &st[1]);
    bins[2] = locate(g.a_b(), g.na, sph[2], &st[2]);
}

// ---- phase 0: classify one ray and plan its candidate ranges ---------------------------------
// elevation bin of a point from its cosine: last cone with cos e_i >= cos e (the table descends)
__device__ __forceinline__ int coop_bin_cos(const GridView& g, double cosv)
{
    const double* ce = g.cos_e();
    int lo = 0, hi = g.ne + 1;                     // first index with ce[idx] < cosv
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (ce[mid] >= cosv) lo = mid + 1; else hi = mid; }
    return lo - 1;
}

// Start points are shared by whole observations: their classification is cached per lane.
struct CoopStartCache {
    int64_t index; bool inside, on_boundary;       // inside the outermost shell; on a cone / half-plane
    int se, sa, a_bin;                             // reference states and azimuth bin of the start
};

// kMode selects which kinds of ray this instantiation takes (the other kind is handed over as
// COOP_LISTED).  Large problems run the kernel twice, once per kind, so that the rare all-cones
// code stays out of the main kernel's instruction stream (inlined into one kernel it cost the
// config-5 forward pass 11 %); small ones run it once for both (a second launch costs them more).
// kMode: COOP_ORDINARY (hand the all-cones rays over), COOP_ALLCONES (only those), COOP_BOTH.
enum { COOP_ORDINARY = 0, COOP_ALLCONES = 1, COOP_BOTH = 2 };

// kAllPlanes (second-generation walker only): a ray that passes (nearly) through the Z axis has no
// monotone azimuth sweep -- every half-plane has its root at the axis, kept or dropped by the last
// bits of the half-plane test, as in the reference.  Such a ray is walked in all-cones mode with
// EVERY half-plane evaluated and rank-sorted as well (a_count == 0 in its record) instead of being
// left to one thread of pass B.
template <int kMode, bool kAllPlanes = false>
__device__ __forceinline__ int coop_plan_ray(const GridView& g, const RayConsts& c, const double x[3],
                                             const CoopPlan& P, int64_t start_index,
                                             const int32_t* cells, CoopStartCache& cache, CoopRecord& rec)
{
    // a ray that never reaches the outermost shell has no segment, whatever else holds for it
    if (c.dperp2 != c.dperp2) return COOP_MISS;                   // no radial root at all
    const double rN = g.r_b()[g.nr];
    const double argN = rN * rN - c.dperp2;
    if (!(argN >= 0)) return COOP_MISS;                           // passes outside the grid
    const double sN = sqrt(argN);
    const double t_in = c.tc - sN, t_out = c.tc + sN;
    if (t_out < 0) return COOP_MISS;                              // the grid lies behind the start
    // Rays whose cone roots lose adjacency (snapped discriminant or leading coefficient,
    // ray_needs_cone_list) and rays that start on a boundary are walked from t = 0 with EVERY cone
    // evaluated and the roots sorted explicitly ("all-cones mode", e_mid == 0)
    const bool listed = ray_needs_cone_list(c, g.c2_sorted(), g.n_c2);
    // (the all-cones run only sees flagged rays and has one way out: pass B, which takes any ray)
    const int hand_over = (kMode == COOP_ALLCONES || listed) ? COOP_LISTED : COOP_TRACKED;
    if (start_index != cache.index) {
        int bins[3], st[3];
        bool on_b;
        if (cells != nullptr) {
            // the reference's own binning of the start point, evaluated on the host (sphrt.h)
            const int32_t* p = cells + 8 * start_index;
            bins[0] = __ldg(p); bins[1] = __ldg(p + 1); bins[2] = __ldg(p + 2);
            st[0] = __ldg(p + 3); st[1] = __ldg(p + 4); st[2] = __ldg(p + 5);
            on_b = __ldg(p + 6) != 0;
        } else {
            double sph[3];
            coop_locate_point(g, x[0], x[1], x[2], bins, st, sph);
            on_b = near_boundary(g.e_b(), g.ne, bins[1], sph[1], 1e-9) ||
                   near_boundary(g.a_b(), g.na, bins[2], sph[2], 1e-9);
        }
        cache.index = start_index;
        cache.inside = bins[0] != g.nr;
        cache.on_boundary = on_b;
        cache.se = st[1]; cache.sa = st[2]; cache.a_bin = bins[2];
    }
    // partial azimuth range (gap sector logic): per-thread walkers; the bin tracker only when the
    // start is not on a boundary
    if (!g.wrap) return cache.on_boundary ? COOP_LISTED : hand_over;
    // starts inside the grid (quirk 1, no entry event): per-thread walkers; the bin tracker only
    // when the start is not on a boundary
    if (cache.inside) return cache.on_boundary ? COOP_LISTED : hand_over;
    if (!(t_in >= 0) || !(t_out < SPHRT_INF)) return hand_over;
    // a start ON a cone or half-plane makes the reference keep or lose the t ~ 0 crossing by
    // rounding: no shortcut from the geometry of the entry point then, but the all-cones walk from
    // t = 0 applies the reference's own rule (t >= 0) to the reference's own roots
    const double cross = c.x0 * c.d1 - c.x1 * c.d0;
    const bool near_axis = !(fabs(cross) > 1e-9 * sqrt(c.nx2));                         // (nearly) through the Z axis
    const bool all_cones = listed || cache.on_boundary || (kAllPlanes && near_axis);
    if (kMode != COOP_BOTH && all_cones != (kMode == COOP_ALLCONES)) return COOP_LISTED;   // the other run's ray
    // innermost shell with a root: first i with r_i^2 - dperp2 >= 0 (same expression as the walk)
    int lo = 0, hi = g.nr;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double r = g.r_b()[mid];
        if (r * r - c.dperp2 >= 0) hi = mid; else lo = mid + 1;
    }
    rec.imin = lo;
    rec.d0 = c.d0; rec.d1 = c.d1; rec.d2 = c.d2;
    rec.tc = c.tc; rec.dperp2 = c.dperp2;

    const double q[3] = {c.d0 * t_out + c.x0, c.d1 * t_out + c.x1, c.d2 * t_out + c.x2};
    int qs_a;
    const int qb_a = locate(g.a_b(), g.na, (double)atan2f((float)q[1], (float)q[0]), &qs_a);
    int a_bin_in;
    if (kMode != COOP_ORDINARY && all_cones) {
        // walk from the start point: its states, all cones, planes from its sector
        rec.se0 = cache.se; rec.sa0 = cache.sa; a_bin_in = cache.a_bin;
        rec.e_first = 0; rec.e_count = g.ne + 1; rec.e_step = 1; rec.t_turn = -SPHRT_INF; rec.e_mid = 0;
    } else {
    // voxel state at the entry point (as walk_ray_fast's fast-forward): exact, float64
    const double p[3] = {c.d0 * t_in + c.x0, c.d1 * t_in + c.x1, c.d2 * t_in + c.x2};
    int pb[3], ps[3];
    double psph[3];
    coop_locate_point(g, p[0], p[1], p[2], pb, ps, psph);
    rec.se0 = ps[1]; rec.sa0 = ps[2]; a_bin_in = pb[2];
    // the bins of the exit and turning points only size the candidate ranges, whose two ends are
    // guards: cosines and a float atan2 are accurate enough (an off-by-one trips a guard)
    const double cos_p = p[2] / sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    const double cos_q = q[2] / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const int qb_e = coop_bin_cos(g, cos_q);

    // ---- elevation: colatitude is unimodal along the chord; turning point at t*
    const double den = c.d2 * c.dx - c.x2;
    const double tstar = (c.x2 * c.dx - c.d2 * c.nx2) / den;
    int first, count, step, mid = 1, up = 0;
    double t_turn = -SPHRT_INF;
    if (tstar > t_in && tstar < t_out) {
        const double m[3] = {c.d0 * tstar + c.x0, c.d1 * tstar + c.x1, c.d2 * tstar + c.x2};
        const double cos_m = m[2] / sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
        const int mb = coop_bin_cos(g, cos_m);
        t_turn = tstar;
        const int top = pb[1] > qb_e ? pb[1] : qb_e, bot = pb[1] < qb_e ? pb[1] : qb_e;
        if (cos_m > cos_p) { step = +1; first = mb; count = top + 2 - mb; }       // dips towards +Z
        else               { step = -1; first = mb + 1; count = mb + 2 - bot; }   // rises towards -Z
        // first crossings (entry -> turn) fill the buffer downwards from `mid`, second crossings
        // (turn -> exit) upwards; a chord's colatitude varies by less than pi in total
        mid = (pb[1] > mb ? pb[1] - mb : mb - pb[1]) + 2;
        up = (qb_e > mb ? qb_e - mb : mb - qb_e) + 2;
    } else if (cos_q <= cos_p) { step = +1; first = pb[1]; count = qb_e + 2 - pb[1]; up = count; }
    else                       { step = -1; first = pb[1] + 1; count = pb[1] + 2 - qb_e; up = count; }
    if (count < 2 || mid + up + 1 > P.capE) return COOP_TRACKED;
    rec.e_first = first; rec.e_count = count; rec.e_step = step; rec.t_turn = t_turn; rec.e_mid = mid;
    }

    // ---- azimuth: monotone rotation, direction = sign of (x x d)_z
    if (near_axis) {
        if (!kAllPlanes) return hand_over;
        rec.a_dir = 1; rec.a_first = 0; rec.a_count = 0;             // every half-plane, rank-sorted
        return COOP_RAY;
    }
    const int dir = cross > 0 ? +1 : -1;
    const int s_in = FastAzimuth::to_sector(g, a_bin_in), s_out = FastAzimuth::to_sector(g, qb_a);
    int sweep = (s_out - s_in) * dir;
    if (sweep < 0) sweep += g.na;
    if (sweep > g.na / 2 + 1) sweep = 0;          // rounding put the exit a sector behind the entry
    rec.a_dir = dir;
    rec.a_first = dir > 0 ? s_in + 1 : s_in;
    rec.a_count = sweep + 1;                      // the last candidate is the guard
    if (rec.a_count > P.capA) return hand_over;
    return COOP_RAY;
}

// ---- phase 1 ------------------------------------------------------------------------------------
struct CoopLists {
    double *tR, *tE, *tA;     // ascending distances; [n] = +inf
    int16_t *rE, *rA;         // region after the event; [-1] = state at entry
    int nR, nE, nA;
    int m, imin, nr;          // radial events: k < m first root of shell nr - k, else second root
                              // of shell imin + k - m
    // radial index after radial event k (k = -1: before the entry)
    __device__ __forceinline__ int radial_region(int k) const
    {
        const int reg = k < m ? nr - 1 - k : imin + (k - m);
        return reg == nr ? -1 : reg;
    }
};

// Build the three sorted lists of ray `rec`.  Returns false if a structural assumption did not
// hold (the ray then goes to the per-thread walker).  All lanes of the warp call this.
template <int kMode>
__device__ __forceinline__ bool coop_build(const GridView& g, const CoopPlan& P, char* wbase,
                                           const CoopRecord& rec, const double x[3], int lane,
                                           CoopLists& L)
{
    RayConsts c;
    c.x0 = x[0]; c.x1 = x[1]; c.x2 = x[2];
    c.d0 = rec.d0; c.d1 = rec.d1; c.d2 = rec.d2;
    c.tc = rec.tc; c.dperp2 = rec.dperp2;
    { const double nx = norm3(c.x0, c.x1, c.x2); c.nx2 = nx * nx; }      // as ray_setup
    c.dz2 = c.d2 * c.d2;
    c.dzxz = c.d2 * c.x2;
    c.dx = dot3(c.d0, c.d1, c.d2, c.x0, c.x1, c.x2);
    c.xz2 = c.x2 * c.x2;
    c.r0 = c.r1 = c.r2 = 0.0;                      // only needed by the tangent test (bails out)
    bool bad = false;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // ---- radial: first roots of shells nr .. imin, then second roots of imin .. nr
    double* tR = reinterpret_cast<double*>(wbase + P.off_tR);
    const int m = g.nr - rec.imin + 1;
    const int nR = 2 * m;
    const double tang = 1e-17 * (1.0 + c.nx2);
    for (int k = lane; k < m; k += 32) {           // shell nr - k: both roots share the square root
        const double r = g.r_b()[g.nr - k];
        const double r2 = r * r;
        const double arg = r2 - c.dperp2;
        const double t1c = sqrt(arg);
        // inward roots enter shell i - 1, outward roots shell i (CoopLists::radial_region); when
        // tangent within rounding the reference decides from the sign of d.p: bail out
        if (!(arg >= tang + 1e-17 * r2)) bad = true;
        tR[k] = c.tc - t1c;
        tR[nR - 1 - k] = c.tc + t1c;
    }
    if (lane == 0) tR[nR] = SPHRT_INF;
    __syncwarp();
    const bool all_cones = kMode == COOP_BOTH ? rec.e_mid == 0 : kMode == COOP_ALLCONES;
    // events are kept from the entry point on, or from the start point in all-cones mode
    const double t_in = all_cones ? 0.0 : tR[0], t_out = tR[nR - 1];

    // ---- elevation: candidates from the turning point outwards
    double* tEbuf = reinterpret_cast<double*>(wbase + P.off_tE);
    int16_t* rEbuf = reinterpret_cast<int16_t*>(wbase + P.off_rE) + 1;
    int nDown = 0, nUp = 0;
    bool unchanged = false;
    if (kMode != COOP_ORDINARY && all_cones) {
        // every cone, both roots, unsorted into the (still unused) merge buffers ...
        double* tmpT = reinterpret_cast<double*>(wbase + P.off_tEA);
        int32_t* tmpR = reinterpret_cast<int32_t*>(wbase + P.off_sEA) + 1;
        int n = 0;
        for (int k0 = 0; k0 <= g.ne; k0 += 32) {
            const int i = k0 + lane;
            double t1 = SPHRT_INF, t2 = SPHRT_INF;
            bool v1 = false, v2 = false;
            if (i <= g.ne) {
                const ConeQuad q = cone_quadratic(g, c, i);
                cone_roots(q, fabs(q.delta) < SPHRT_TAU ? 0.0 : q.delta, t1, t2);
                v1 = t1 >= t_in && t1 <= t_out && !cone_shadow(g, c, i, t1);
                v2 = t2 >= t_in && t2 <= t_out && !cone_shadow(g, c, i, t2);
            }
            const uint32_t m1 = __ballot_sync(SPHRT_FULL, v1), m2 = __ballot_sync(SPHRT_FULL, v2);
            const int s1 = n + __popc(m1 & lt_mask), s2 = n + __popc(m1) + __popc(m2 & lt_mask);
            int geo;
            if (v1) { if (s1 < P.capE) { tmpT[s1] = t1; tmpR[s1] = cone_region(g, c, i, t1, &geo); } else bad = true; }
            if (v2) { if (s2 < P.capE) { tmpT[s2] = t2; tmpR[s2] = cone_region(g, c, i, t2, &geo); } else bad = true; }
            n += __popc(m1) + __popc(m2);
        }
        __syncwarp();
        // ... then to their rank (the reference's sort, :131; ties in evaluation order)
        if (n <= P.capE) {
            for (int e = lane; e < n; e += 32) {
                const double te = tmpT[e];
                int rank = 0;
                for (int j = 0; j < n; j++) {
                    const double tj = tmpT[j];
                    rank += (tj < te || (tj == te && j < e)) ? 1 : 0;
                }
                const int reg = tmpR[e];
                tEbuf[rank] = te; rEbuf[rank] = (int16_t)reg;
                unchanged |= reg == -2;
            }
        }
        nUp = n;
    } else if (kMode != COOP_ALLCONES)
    for (int k0 = 0; k0 < rec.e_count; k0 += 32) {
        const int k = k0 + lane;
        const int i = rec.e_first + k * rec.e_step;
        double ta = SPHRT_INF, tb = SPHRT_INF;
        int ra = 0, rb = 0;
        if (k < rec.e_count && i >= 0 && i <= g.ne) {
            const ConeQuad q = cone_quadratic(g, c, i);
            double t1, t2;
            cone_roots(q, fabs(q.delta) < SPHRT_TAU ? 0.0 : q.delta, t1, t2);
            // keep the roots inside the chord that lie on the cone's own nappe (cone_event);
            // the region is only worked out for those
            const bool v1 = t1 >= t_in && t1 <= t_out && !cone_shadow(g, c, i, t1);
            const bool v2 = t2 >= t_in && t2 <= t_out && !cone_shadow(g, c, i, t2);
            if (v1) ta = t1;
            if (v2) { if (v1) tb = t2; else ta = t2; }
            if (tb < ta) { const double tt = ta; ta = tb; tb = tt; }
            if (tb == ta) tb = SPHRT_INF;                      // merged double root: one event
            int geo;
            if (ta < SPHRT_INF) ra = cone_region(g, c, i, ta, &geo);
            // the candidates at both ends of the range are guards: they must stay silent
            if ((k == 0 || k == rec.e_count - 1) && ta < SPHRT_INF) bad = true;
        }
        if (__any_sync(SPHRT_FULL, tb < SPHRT_INF)) {           // cones crossed twice inside the chord
            int geo;
            if (tb < SPHRT_INF) rb = cone_region(g, c, i, tb, &geo);
        }
        const bool a_down = ta < rec.t_turn;                    // first crossing (towards the turn)
        const bool a_up = ta < SPHRT_INF && !a_down;
        const bool b_up = tb < SPHRT_INF;
        if (a_up && b_up) bad = true;                           // two crossings after the turn
        const uint32_t md = __ballot_sync(SPHRT_FULL, a_down);
        const uint32_t mu = __ballot_sync(SPHRT_FULL, a_up || b_up);
        if (a_down) {
            const int slot = rec.e_mid - 1 - (nDown + __popc(md & lt_mask));
            if (slot >= 0) { tEbuf[slot] = ta; rEbuf[slot] = (int16_t)ra; } else bad = true;
            unchanged |= ra == -2;
        }
        if (a_up || b_up) {
            const int slot = rec.e_mid + nUp + __popc(mu & lt_mask);
            if (slot < P.capE) { tEbuf[slot] = a_up ? ta : tb; rEbuf[slot] = (int16_t)(a_up ? ra : rb); }
            else bad = true;
            unchanged |= (a_up ? ra : rb) == -2;
        }
        nDown += __popc(md); nUp += __popc(mu);
    }
    if (__any_sync(SPHRT_FULL, bad)) return false;        // (also keeps the offsets below in range)
    double* tE = tEbuf + (rec.e_mid - nDown);
    int16_t* rE = rEbuf + (rec.e_mid - nDown);
    const int nE = nDown + nUp;
    if (lane == 0) { tE[nE] = SPHRT_INF; rE[-1] = (int16_t)rec.se0; }

    // ---- azimuth: planes in the direction of rotation
    double* tA = reinterpret_cast<double*>(wbase + P.off_tA);
    int16_t* rA = reinterpret_cast<int16_t*>(wbase + P.off_rA) + 1;
    int nA = 0;
    for (int k0 = 0; k0 < rec.a_count; k0 += 32) {
        const int k = k0 + lane;
        double t = SPHRT_INF;
        int geo = 0;
        if (k < rec.a_count) {
            int pl = rec.a_first + k * rec.a_dir;              // seam: plane na going up, 0 going down
            if (pl > g.na) pl -= g.na;
            if (pl < 0) pl += g.na;
            if (pl == 0 || pl == g.na) {
                // the seam: one half-plane the reference evaluates twice, earlier root wins
                FastAzimuth::eval(g, c, pl, t_in, false, t, geo);
                if (!(t <= t_out)) t = SPHRT_INF;
            } else {
                // plane_event() with the validity kept in a predicate
                const double ca = g.cos_a()[pl], sa = g.sin_a()[pl];
                const double num = fma_rn(0.0, c.x2, fma_rn(ca, c.x1, (-sa) * c.x0));
                const double den = fma_rn(0.0, c.d2, fma_rn(ca, c.d1, (-sa) * c.d0));
                const double tt = -num / den;
                const double cross = fma_rn(ca, c.d1, -(sa * c.d0));
                geo = pl - (cross < 0 ? 1 : 0);
                const double px = tt * c.d0 + c.x0, py = tt * c.d1 + c.x1;
                const bool valid = !(fabs(cross) <= SPHRT_A_PARALLEL) && !(ca * px + sa * py < 0) &&
                                   tt >= t_in && tt <= t_out;
                if (valid) t = tt;
            }
            if (k == rec.a_count - 1 && t < SPHRT_INF) bad = true;     // guard
        }
        const uint32_t mv = __ballot_sync(SPHRT_FULL, t < SPHRT_INF);
        if (t < SPHRT_INF) {
            const int slot = nA + __popc(mv & lt_mask);
            tA[slot] = t; rA[slot] = (int16_t)FastAzimuth::to_region(g, geo);
        }
        nA += __popc(mv);
    }
    if (lane == 0) { tA[nA] = SPHRT_INF; rA[-1] = (int16_t)rec.sa0; }
    __syncwarp();

    // (the ascending order of both lists is verified while they are merged, coop_merge_ea)
    if (nE + nA > P.capEA) bad = true;
    if (__any_sync(SPHRT_FULL, bad)) return false;
    // glancing crossings leave the elevation index unchanged (region -2): forward-fill
    if (__any_sync(SPHRT_FULL, unchanged)) {
        if (lane == 0) {
#pragma unroll 1                                   // rolled (rare path)
            for (int k = 0; k < nE; k++) if (rE[k] == -2) rE[k] = rE[k - 1];
        }
        __syncwarp();
    }
    L.tR = tR; L.tE = tE; L.tA = tA; L.rE = rE; L.rA = rA;
    L.nR = nR; L.nE = nE; L.nA = nA;
    L.m = m; L.imin = rec.imin; L.nr = g.nr;
    return true;
}

// ---- merging -----------------------------------------------------------------------------------
// Merge path: lane L owns the merged ranks [L*c, (L+1)*c) of two ascending lists and finds where
// that range starts in each list with one co-rank search; the c-step sequential merge that
// follows has the same trip count in every lane.  X wins ties (radial < elevation < azimuth).

// number of X entries among the first g entries of merge(X, Y); `steps` >= log2(nX) + 1, uniform
__device__ __forceinline__ int coop_corank(const double* X, int nX, const double* Y, int nY, int g,
                                           int steps)
{
    int lo = g - nY > 0 ? g - nY : 0, hi = g < nX ? g : nX;
#pragma unroll 1                                   // rolled
    for (int s = 0; s < steps; s++) {
        const int mid = (lo + hi) >> 1;
        const bool open = lo < hi;
        const bool take = X[mid] <= Y[g - mid - 1];      // harmless garbage when !open
        lo = (open && take) ? mid + 1 : lo;
        hi = (open && !take) ? mid : hi;
    }
    return lo;
}

__device__ __forceinline__ int coop_log2_steps(int n)     // iterations that bisect [0, n] down to one
{
    return 32 - __clz(n > 0 ? n : 1);
}

// stage 1: elevation + azimuth -> one list of (t, elevation | azimuth << 16 after the event).
// Returns the merged length, or -1 if one of the two lists is not ascending (every entry is
// compared with its successor here: inside a lane's share when the successor is loaded, across
// shares when a lane starts).
__device__ __forceinline__ int coop_merge_ea(const CoopPlan& P, char* wbase, const CoopLists& L,
                                             int lane, double*& tEA, int32_t*& sEA)
{
    tEA = reinterpret_cast<double*>(wbase + P.off_tEA);
    sEA = reinterpret_cast<int32_t*>(wbase + P.off_sEA) + 1;
    const int n = L.nE + L.nA;
    const int c = (n + 31) >> 5;
    const int g0 = lane * c;
    const int gs = g0 < n ? g0 : n;
    const int i = coop_corank(L.tE, L.nE, L.tA, L.nA, gs, coop_log2_steps(L.nE));
    const int j = gs - i;
    const double* pE = L.tE + i;
    const double* pA = L.tA + j;
    const int16_t* qE = L.rE + i;
    const int16_t* qA = L.rA + j;
    int se = qE[-1], sa = qA[-1];
    double xv = *pE, yv = *pA;
    bool viol = (i > 0 && i < L.nE && pE[-1] > xv) || (j > 0 && j < L.nA && pA[-1] > yv);
    bool tx = xv <= yv;
    for (int g = g0; g < g0 + c; g++) {             // same trip count in every lane: no divergence
        const double t = tx ? xv : yv;
        const bool live = g < n;                    // beyond the end the heads are stale memory
        if (tx) { se = *qE++; xv = *++pE; viol |= live && xv < t; }
        else    { sa = *qA++; yv = *++pA; viol |= live && yv < t; }
        tx = xv <= yv;
        if (live) { tEA[g] = t; sEA[g] = (se & 0xffff) | (sa << 16); }
    }
    if (lane == 0) { tEA[n] = SPHRT_INF; sEA[-1] = (L.rE[-1] & 0xffff) | ((int)L.rA[-1] << 16); }
    if (__any_sync(SPHRT_FULL, viol)) return -1;
    __syncwarp();
    return n;
}

// stage 2: radial + (elevation, azimuth); every merged event starts one segment that ends at the
// next event: the reference's sort + forward-fill + diff (raytracer.py:131-151).
// The loop is unrolled by kCoopDepth so that a sink can keep that many gathers in flight
// (Sink::segment_at<slot>): the volume sits in L2, a few hundred cycles away.
#ifndef SPHRT_COOP_DEPTH
#define SPHRT_COOP_DEPTH 1
#endif
constexpr int kCoopDepth = SPHRT_COOP_DEPTH;      // unroll of gathering sinks (Sink::kDepth): 1, 2 or 4

template <bool kEvents, class Sink>
__device__ __forceinline__ void coop_segments(const CoopLists& L, const double* tEA, const int32_t* sEA,
                                              int nEA, int lane, Sink& sink)
{
    const int n = L.nR + nEA;
    const int c = (n + 31) >> 5;
    const int g0 = lane * c;
    const int gs = g0 < n ? g0 : n;
    int i = coop_corank(L.tR, L.nR, tEA, nEA, gs, coop_log2_steps(L.nR));
    const int j = gs - i;
    const double* pR = L.tR + i;
    const double* pE = tEA + j;
    const int32_t* qE = sEA + j;
    int sr = L.radial_region(i - 1), sea = qE[-1];
    double xv = *pR, yv = *pE;
    bool tx = xv <= yv;
    double t = tx ? xv : yv;
    int g = g0;
    // one merge step + one segment; the trip count is the same in every lane (no divergence)
#define SPHRT_COOP_STEP(SLOT)                                                                     \
    {                                                                                             \
        if (tx) { sr = L.radial_region(i); i++; xv = *++pR; }                                     \
        else    { sea = *qE++; yv = *++pE; }                                                      \
        tx = xv <= yv;                                                                            \
        const double tn = tx ? xv : yv;                                                           \
        const double len = tn - t;                                                                \
        /* after the exit event the radial index is -1: the +inf sentinel never yields a length */\
        const bool ok = g < n && (sr | sea) >= 0 && (sea & 0x8000) == 0 && len > 0;               \
        if constexpr (kEvents) {                                                                  \
            if (g < n) sink.event(g, sr, (int)(int16_t)(sea & 0xffff), sea >> 16,                 \
                                  ok && tn < SPHRT_INF ? len : 0.0);                              \
        } else {                                                                                  \
            sink.template segment_at<SLOT>(ok, sr, sea & 0xffff, sea >> 16, len);                 \
        }                                                                                         \
        t = tn; g++;                                                                              \
    }
    const int g_end = g0 + c;
    constexpr int kDepth = Sink::kDepth;           // 1 for sinks without loads to hide
    // (left to the compiler's 4x unroll: rolled, the forward gathers lose their overlap, +9 %)
    while (g + kDepth <= g_end) {
        SPHRT_COOP_STEP(0)
        if constexpr (kDepth > 1) SPHRT_COOP_STEP(1)
        if constexpr (kDepth > 2) { SPHRT_COOP_STEP(2) SPHRT_COOP_STEP(3) }
    }
    if constexpr (kDepth > 1) { if (g < g_end) SPHRT_COOP_STEP(0) }
    if constexpr (kDepth > 2) { if (g < g_end) SPHRT_COOP_STEP(1) if (g < g_end) SPHRT_COOP_STEP(2) }
#undef SPHRT_COOP_STEP
}

}  // namespace sphrt
