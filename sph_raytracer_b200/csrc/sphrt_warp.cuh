// sphrt_warp.cuh -- warp-cooperative walker, second generation: the kernel that walks the ordinary
// rays of large launches (every ray that starts outside the grid, off the cones / half-planes, and
// has no snapped cone: 98-99 % of the rays that hit the grid at config 5).
//
// Same algorithm and the same exact arithmetic as sphrt_coop.cuh (three ascending event lists per
// ray, evaluated by 32 lanes with the reference's formulas, verified, merged with two merge-path
// stages -- the reference's sort + forward-fill + diff, raytracer.py:131-151), reorganised around
// what ncu showed of that kernel (profiles/README.md, round 1d: 227 thread-instructions per segment
// of which only ~20 % are float64 arithmetic, 16 warps per SM, instruction footprint twice the
// instruction cache):
//
//   * PLANNING IS ITS OWN KERNEL (plan_warp_kernel, one thread per ray): classification, entry /
//     exit / turning-point bins and candidate ranges go into an 80-byte record per ray that is
//     walked; rays that never reach the grid produce nothing, the other kinds are flagged for the
//     existing passes (all-cones run, per-thread walker, pass B).  The walking kernel is handed a
//     dense list of records: no miss handling, no atan2, no binary searches in its instruction
//     stream, and every warp-iteration is a ray with work.
//   * SHARED MEMORY PER WARP IS 9 KB INSTEAD OF 13 KB at config 5: the radial list is built AFTER the
//     elevation and azimuth lists have been merged and overlays them.  With <= 80 registers that is
//     24 warps per SM instead of 16.
//   * THE MERGE LOOPS CARRY PACKED STATE: elevation and azimuth regions are stored pre-multiplied
//     (e * na, a) so that the merged list holds e * na + a, the radial state is the running offset
//     r * ne * na, and ONE unsigned comparison of their sum against the voxel count replaces the
//     three sign tests (an index outside the grid is a large negative number in either part); the
//     radial list is padded with +inf so that the loop needs no end-of-sequence test; the gather is
//     issued under its predicate and consumed one or two steps later under the same predicate.
//
// Rays whose lists fail verification are flagged for the per-thread walker exactly as before.
#pragma once
#include "sphrt_coop.cuh"
#ifdef SPHRT_DEBUG_FAIL
#include <cstdio>
#endif

namespace sphrt {

// Stage 2 loads only the head it consumed again (two selects more, one shared-memory load less per
// event: the walking kernels keep the LSU data pipe 85-91 % busy, two thirds of it shared-memory
// wavefronts at 2-3x their conflict-free count -- per-lane list positions are as good as random
// banks; forward -2.6 %, adjoint -2.4 %).  The same in stage 1 (SPHRT_ONE_HEAD1) pays for its six
// selects with all it saves.
#ifndef SPHRT_ONE_HEAD
#define SPHRT_ONE_HEAD 1
#endif
#ifndef SPHRT_SKEW1
#define SPHRT_SKEW1 1
#endif
#ifndef SPHRT_ONE_HEAD1
#define SPHRT_ONE_HEAD1 0
#endif
#ifndef SPHRT_WARP_DEPTH
#define SPHRT_WARP_DEPTH 4
#endif
constexpr int kWarpDepth = SPHRT_WARP_DEPTH;         // gathers in flight per lane in the segment loop: 2 or 4

constexpr int kInvalidState = -(1 << 30);            // packed state of an index outside the grid
constexpr int kUnchangedState = kInvalidState - 1;   // glancing cone crossing: keep the previous one

// Shared-memory layout of one warp (byte offsets from the warp's base), filled in by the host.
struct WarpPlan {
    int capR, padR, capE, capA, capEA;
    int off_tE, off_rE, off_tA, off_rA;      // phase A: elevation / azimuth lists
    int off_tR;                              // phase B: radial list, overlays phase A
    int off_tEA, off_sEA;                    // merged elevation + azimuth list
    int off_flag;                            // two flag words
    int warp_bytes, warps, table_doubles;
};

// What the walking kernel needs to know about a ray (written by plan_warp_kernel).
struct alignas(16) PlanRec {
    double d0, d1, d2;        // direction as e_torch / a_torch see it (normalised three times)
    double tc, dperp2;        // radial constants (r_torch, :288-289)
    double t_turn;            // colatitude turning point (-inf: monotone along the chord)
    double nx2;               // |x|^2 as e_torch computes it (norm()**2, :375)
    int32_t ray, start;       // ray index relative to the chunk, row of `starts`
    int16_t imin, e_mid;      // innermost shell reached; slot of the cone buffer where the runs meet
    int16_t e_first, e_cs;    // cone candidates e_first + k * sign(e_cs), k < |e_cs|
    int16_t a_first, a_cd;    // plane candidates a_first + k * sign(a_cd), k < |a_cd|
    int16_t se0, sa0;         // elevation / azimuth state at the entry point
};
static_assert(sizeof(PlanRec) == 80, "PlanRec is loaded as five 16-byte words");

__device__ __forceinline__ PlanRec load_plan_rec(const PlanRec* p)
{
    union { PlanRec r; int4 w[5]; } u;
    const int4* s = reinterpret_cast<const int4*>(p);
#pragma unroll
    for (int k = 0; k < 5; k++) u.w[k] = __ldg(s + k);
    return u.r;
}

__device__ __forceinline__ void store_plan_rec(PlanRec* p, const PlanRec& r)
{
    union { PlanRec r; int4 w[5]; } u;
    u.r = r;
    int4* d = reinterpret_cast<int4*>(p);
#pragma unroll
    for (int k = 0; k < 5; k++) d[k] = u.w[k];
}

// ---- shared memory behind plain 32-bit addresses ---------------------------------------------------
// The lists and tables are addressed with 32-bit shared-window addresses and explicit ld/st.shared:
// the compiler then keeps one address register per list (with generic pointers it re-derived the
// window base -- S2UR SR_CgaCtaId -- inside the loops) and cannot turn the arithmetic list updates of
// the merge loops back into pairs of predicated moves.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_s32(uint32_t a, int v)
{
    asm volatile("st.shared.s32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}

__device__ __forceinline__ void lds_f64x2(uint32_t a, double& v0, double& v1)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a));
}

// Tables of the walking kernel in shared memory, derived once per block from the grid's boundary
// tables (same products as the per-ray code would form, so nothing changes numerically):
//   r2   (nr + 1, padded to an even count)  r_b^2
//   cone (ne + 1 pairs)   { cos^2 e_b, sign: 0 for cones exempt from the shadow test by
//                           |e_b - pi/2| < tau (raytracer.py:449-461), else +1 / -1 for cos e_b >= 0 / < 0 }
//   plane (na + 1 pairs)  { cos a_b, sin a_b }
// One base register; the other tables sit at offsets that only depend on the grid shape.
struct Tab32 {
    uint32_t r2, cone, plane;
    int nr, ne, na;
    __host__ __device__ static int doubles(int nr, int ne, int na)
    {
        return ((nr + 2) & ~1) + 2 * (ne + 1) + 2 * (na + 1);
    }
    __device__ __forceinline__ Tab32(uint32_t base, int nr_, int ne_, int na_)
    {
        nr = nr_; ne = ne_; na = na_;
        r2 = base;
        cone = base + 8u * (uint32_t)((nr + 2) & ~1);
        plane = cone + 16u * (uint32_t)(ne + 1);
    }
};

// v if lo <= v <= hi and !(side < 0) and !(mag <= eps), else +inf
__device__ __forceinline__ double select_in_range(double v, double lo, double hi, double side, double mag, double eps)
{
    double r;
    asm("{\n .reg .pred p;\n"
        " setp.ge.f64 p, %1, %2;\n"
        " setp.le.and.f64 p, %1, %3, p;\n"
        " setp.geu.and.f64 p, %4, 0d0000000000000000, p;\n"
        " setp.gtu.and.f64 p, %5, %6, p;\n"
        " selp.f64 %0, %1, 0d7FF0000000000000, p;\n}"
        : "=d"(r) : "d"(v), "d"(lo), "d"(hi), "d"(side), "d"(mag), "d"(eps));
    return r;
}

// number of X entries among the first g entries of merge(X, Y), X winning ties; X / Y are the
// shared addresses of the lists; `steps` bisections (>= log2(min(nX, nY)) + 1, warp-uniform)
__device__ __forceinline__ int warp_corank(uint32_t X, int nX, uint32_t Y, int nY, int g, int steps)
{
    int lo = g - nY > 0 ? g - nY : 0, hi = g < nX ? g : nX;
    const uint32_t Yg = Y + 8u * (uint32_t)(g - 1);
#pragma unroll 1
    for (int s = 0; s < steps; s++) {
        const int mid = (lo + hi) >> 1;
        const bool open = lo < hi;
        const bool take = lds_f64(X + 8u * mid) <= lds_f64(Yg - 8u * mid);   // stale values when !open
        lo = (open && take) ? mid + 1 : lo;
        hi = (open && !take) ? mid : hi;
    }
    return lo;
}

// Walk one planned ray with all 32 lanes.  Returns the number of merged events, or -1 if one of
// the structural assumptions did not hold (nothing has been given to the sink then).
// the seam of a full-circle grid: planes 0 and na are one half-plane that the reference evaluates
// twice (with table values that differ in the last bits); the earlier root not behind t_now wins
// (FastAzimuth::eval of sphrt_fast.cuh on the walking kernel's own tables)
__device__ __forceinline__ void warp_seam_plane(const Tab32& T, const RayConsts& c, int i, double t_now,
                                              double& t, int& geo)
{
    double ca, sa;
    lds_f64x2(T.plane + 16u * i, ca, sa);
    Event ev = plane_event_v(ca, sa, c, i, T.na, 1);
    lds_f64x2(T.plane + 16u * (T.na - i), ca, sa);
    const Event tw = plane_event_v(ca, sa, c, T.na - i, T.na, 1);
    const bool ev_ok = !(ev.t < t_now), tw_ok = !(tw.t < t_now);
    if (tw_ok && (!ev_ok || tw.t < ev.t)) ev = tw;
    t = ev.t < t_now ? SPHRT_INF : ev.t;
    geo = ev.geo;
}

// kAllCones: the rays whose cone roots lose adjacency (snapped discriminant or leading coefficient,
// ray_needs_cone_list) or that start ON a cone / half-plane: walked from the START point (t = 0)
// with EVERY cone evaluated and the valid roots put in order by an explicit rank sort -- the
// reference's algorithm (:131) restricted to the elevation family -- instead of the candidate range
// around the turning point.  Everything after the elevation list is the same code.
template <bool kEvents, bool kAllCones, class Sink>
__device__ __forceinline__ int warp_walk_ray(const Tab32& T, const WarpPlan& P, uint32_t wbase,
                                             const PlanRec& R, const double* xs, int lane, uint32_t lt_mask,
                                             Sink& sink)
{
    RayConsts c;
    c.x0 = __ldg(xs); c.x1 = __ldg(xs + 1); c.x2 = __ldg(xs + 2);
    c.d0 = R.d0; c.d1 = R.d1; c.d2 = R.d2;
    c.tc = R.tc; c.dperp2 = R.dperp2; c.nx2 = R.nx2;
    c.dz2 = c.d2 * c.d2;
    c.dzxz = c.d2 * c.x2;
    c.dx = dot3(c.d0, c.d1, c.d2, c.x0, c.x1, c.x2);
    c.xz2 = c.x2 * c.x2;
    c.r0 = c.r1 = c.r2 = 0.0;
    // Two per-warp flag words in shared memory instead of per-lane booleans carried through the
    // loops (the compiler packed those into bytes of one register and spent a dozen instructions
    // per candidate pass on keeping them): [0] a structural check failed, [1] a glancing crossing
    // (region "unchanged") is in the elevation list.  Set by predicated stores, read after the loops.
    const uint32_t flags = wbase + P.off_flag;
    sts_s32(flags, 0); sts_s32(flags + 4u, 0);
    __syncwarp();
#ifdef SPHRT_DEBUG_FAIL
#define SPHRT_BAD_IF(cond) do { if (cond) { sts_s32(flags, 1); printf("bad ray %d line %d lane %d\n", R.ray, __LINE__, lane); } } while (0)
#else
#define SPHRT_BAD_IF(cond) do { if (cond) sts_s32(flags, 1); } while (0)
#endif
#define SPHRT_FAILED() (__syncwarp(), __any_sync(SPHRT_FULL, lds_s32(flags) != 0))
    const int na = T.na;

    // the chord inside the outermost shell: same expressions as the radial list below
    double t_in, t_out;
    {
        const double sN = sqrt(lds_f64(T.r2 + 8u * T.nr) - c.dperp2);
        t_in = kAllCones ? 0.0 : c.tc - sN; t_out = c.tc + sN;
    }

    // ---- elevation: candidates from the turning point outwards (as coop_build)
    const uint32_t tEbuf = wbase + P.off_tE, rEbuf = wbase + P.off_rE + 4u;
    const int e_count = R.e_cs < 0 ? -R.e_cs : R.e_cs, e_step = R.e_cs < 0 ? -1 : 1;
    const int e_mid = R.e_mid;
    int nDown = 0, nUp = 0;
#define SPHRT_UNCHANGED_IF(cond) do { if (cond) sts_s32(flags + 4u, 1); } while (0)
    if constexpr (kAllCones) {
        // every cone, both roots, unsorted into the (still unused) merge buffers ...
        const uint32_t tmpT = wbase + P.off_tEA, tmpR = wbase + P.off_sEA + 4u;
        int n = 0;
#pragma unroll 1
        for (int i = lane; __any_sync(SPHRT_FULL, i <= T.ne); i += 32) {
            double t1 = SPHRT_INF, t2 = SPHRT_INF;
            bool v1 = false, v2 = false;
            if (i <= T.ne) {
                double c2, sgn;
                lds_f64x2(T.cone + 16u * i, c2, sgn);
                const ConeQuad q = cone_quadratic_c2(c, c2);
                cone_roots_guarded(q, fabs(q.delta) < SPHRT_TAU ? 0.0 : q.delta, t1, t2);
                v1 = t1 >= t_in && t1 <= t_out && !cone_shadow_sgn(c, sgn, t1);
                v2 = t2 >= t_in && t2 <= t_out && !cone_shadow_sgn(c, sgn, t2);
            }
            const uint32_t m1 = __ballot_sync(SPHRT_FULL, v1), m2 = __ballot_sync(SPHRT_FULL, v2);
            const int s1 = n + __popc(m1 & lt_mask), s2 = n + __popc(m1) + __popc(m2 & lt_mask);
            int geo;
            if (v1) {
                if (s1 < P.capE) { sts_f64(tmpT + 8u * s1, t1); sts_s32(tmpR + 4u * s1, cone_region_ne(T.ne, c, i, t1, &geo)); }
                else SPHRT_BAD_IF(true);
            }
            if (v2) {
                if (s2 < P.capE) { sts_f64(tmpT + 8u * s2, t2); sts_s32(tmpR + 4u * s2, cone_region_ne(T.ne, c, i, t2, &geo)); }
                else SPHRT_BAD_IF(true);
            }
            n += __popc(m1) + __popc(m2);
        }
        __syncwarp();
        // ... then to their rank (the reference's sort, :131; ties in evaluation order)
        if (n <= P.capE) {
#pragma unroll 1
            for (int e = lane; e < n; e += 32) {
                const double te = lds_f64(tmpT + 8u * e);
                int rank = 0;
#pragma unroll 1
                for (int j = 0; j < n; j++) {
                    const double tj = lds_f64(tmpT + 8u * j);
                    rank += (tj < te || (tj == te && j < e)) ? 1 : 0;
                }
                const int reg = lds_s32(tmpR + 4u * e);
                sts_f64(tEbuf + 8u * rank, te);
                sts_s32(rEbuf + 4u * rank, reg >= 0 ? reg * na : (reg == -2 ? kUnchangedState : kInvalidState));
                SPHRT_UNCHANGED_IF(reg == -2);
            }
        } else SPHRT_BAD_IF(true);
        nUp = n;
    } else {
        int k = lane, i = R.e_first + lane * e_step;
        const int i_step = 32 * e_step;
#pragma unroll 1
        for (; __any_sync(SPHRT_FULL, k < e_count); k += 32, i += i_step) {
            double ta = SPHRT_INF, tb = SPHRT_INF;
            int ra = 0, rb = 0;
            if (k < e_count && i >= 0 && i <= T.ne) {
                double c2, sgn;
                lds_f64x2(T.cone + 16u * i, c2, sgn);
                const ConeQuad q = cone_quadratic_c2(c, c2);
                double t1, t2;
                cone_roots_guarded(q, fabs(q.delta) < SPHRT_TAU ? 0.0 : q.delta, t1, t2);
                const bool v1 = t1 >= t_in && t1 <= t_out && !cone_shadow_sgn(c, sgn, t1);
                const bool v2 = t2 >= t_in && t2 <= t_out && !cone_shadow_sgn(c, sgn, t2);
                if (v1) ta = t1;
                if (v2) { if (v1) tb = t2; else ta = t2; }
                if (tb < ta) { const double tt = ta; ta = tb; tb = tt; }
                if (tb == ta) tb = SPHRT_INF;                      // merged double root: one event
                int geo;
                if (ta < SPHRT_INF) ra = cone_region_ne(T.ne, c, i, ta, &geo);
                // the candidates at both ends of the range are guards: they must stay silent
                SPHRT_BAD_IF((k == 0 || k == e_count - 1) && ta < SPHRT_INF);
            }
            if (__any_sync(SPHRT_FULL, tb < SPHRT_INF)) {           // cones crossed twice inside the chord
                int geo;
                if (tb < SPHRT_INF) rb = cone_region_ne(T.ne, c, i, tb, &geo);
            }
            const bool a_down = ta < R.t_turn;                      // first crossing (towards the turn)
            const bool a_up = ta < SPHRT_INF && !a_down;
            const bool b_up = tb < SPHRT_INF;
            SPHRT_BAD_IF(a_up && b_up);                             // two crossings after the turn
            const uint32_t md = __ballot_sync(SPHRT_FULL, a_down);
            const uint32_t mu = __ballot_sync(SPHRT_FULL, a_up || b_up);
            if (a_down) {
                const int slot = e_mid - 1 - (nDown + __popc(md & lt_mask));
                if (slot >= 0) {
                    sts_f64(tEbuf + 8u * slot, ta);
                    sts_s32(rEbuf + 4u * slot, ra >= 0 ? ra * na : (ra == -2 ? kUnchangedState : kInvalidState));
                } else SPHRT_BAD_IF(true);
                SPHRT_UNCHANGED_IF(ra == -2);
            }
            if (a_up || b_up) {
                const int slot = e_mid + nUp + __popc(mu & lt_mask);
                const int rr = a_up ? ra : rb;
                if (slot < P.capE) {
                    sts_f64(tEbuf + 8u * slot, a_up ? ta : tb);
                    sts_s32(rEbuf + 4u * slot, rr >= 0 ? rr * na : (rr == -2 ? kUnchangedState : kInvalidState));
                } else SPHRT_BAD_IF(true);
                SPHRT_UNCHANGED_IF(rr == -2);
            }
            nDown += __popc(md); nUp += __popc(mu);
        }
    }
    if (SPHRT_FAILED()) return -1;                        // (also keeps the offsets below in range)
    const uint32_t tE = tEbuf + 8u * (e_mid - nDown), rE = rEbuf + 4u * (e_mid - nDown);
    const int nE = nDown + nUp;
    if (lane == 0) {
        sts_f64(tE + 8u * nE, SPHRT_INF);
        sts_s32(rE - 4u, R.se0 >= 0 ? R.se0 * na : kInvalidState);
    }

    // ---- azimuth: planes in the direction of rotation (as coop_build)
    const uint32_t tA = wbase + P.off_tA, rA = wbase + P.off_rA + 4u;
    const int a_count = R.a_cd < 0 ? -R.a_cd : R.a_cd, a_dir = R.a_cd < 0 ? -1 : 1;
    int nA = 0;
    bool all_planes = false;
    if constexpr (kAllCones) all_planes = a_count == 0;
    if (all_planes) {
        // every half-plane (both seam planes, as the reference evaluates them), the valid roots
        // unsorted into the still unused merge buffers, then to their rank (:131; ties in
        // evaluation order)
        const uint32_t tmpT = wbase + P.off_tEA, tmpR = wbase + P.off_sEA + 4u;
        int n = 0;
#pragma unroll 1
        for (int i = lane; __any_sync(SPHRT_FULL, i <= na); i += 32) {
            double t = SPHRT_INF;
            int reg = 0;
            if (i <= na) {
                double ca, sa;
                lds_f64x2(T.plane + 16u * i, ca, sa);
                const Event ev = plane_event_v(ca, sa, c, i, na, 1);
                if (ev.t >= t_in && ev.t <= t_out) t = ev.t;
                reg = ev.region;
            }
            const bool hit = t < SPHRT_INF;
            const uint32_t mv = __ballot_sync(SPHRT_FULL, hit);
            const int slot = n + __popc(mv & lt_mask);
            if (hit) {
                if (slot < P.capEA) { sts_f64(tmpT + 8u * slot, t); sts_s32(tmpR + 4u * slot, reg); }
                else SPHRT_BAD_IF(true);
            }
            n += __popc(mv);
        }
        __syncwarp();
        if (n <= P.capA && n <= P.capEA) {
#pragma unroll 1
            for (int e = lane; e < n; e += 32) {
                const double te = lds_f64(tmpT + 8u * e);
                int rank = 0;
#pragma unroll 1
                for (int j = 0; j < n; j++) {
                    const double tj = lds_f64(tmpT + 8u * j);
                    rank += (tj < te || (tj == te && j < e)) ? 1 : 0;
                }
                sts_f64(tA + 8u * rank, te);
                sts_s32(rA + 4u * rank, lds_s32(tmpR + 4u * e));
            }
        } else SPHRT_BAD_IF(true);
        nA = n;
        __syncwarp();
    } else {
        int k = lane, pl0 = R.a_first + lane * a_dir;
        const int pl_step = 32 * a_dir;
        uint32_t oT = tA, oR = rA;
#pragma unroll 1
        for (; __any_sync(SPHRT_FULL, k < a_count); k += 32, pl0 += pl_step) {
            double t = SPHRT_INF;
            int geo = 0;
            if (k < a_count) {
                int pl = pl0;                                      // seam: plane na going up, 0 going down
                if (pl > na) pl -= na;
                if (pl < 0) pl += na;
                if (pl == 0 || pl == na) {
                    warp_seam_plane(T, c, pl, t_in, t, geo);
                    if (!(t <= t_out)) t = SPHRT_INF;
                    geo = geo < 0 ? geo + na : (geo >= na ? geo - na : geo);   // full circle
                } else {
                    // plane_event() with the validity kept in a predicate.  Two exact identities
                    // are used: fma(0, z, y) == y for finite z (the reference's einsum carries the
                    // zero z-component of the plane normal, :508-513), and cross == den bit for bit
                    // ((-sa) * d0 == -(sa * d0))
                    double ca, sa;
                    lds_f64x2(T.plane + 16u * pl, ca, sa);
                    const double num = fma_rn(ca, c.x1, (-sa) * c.x0);
                    const double den = fma_rn(ca, c.d1, (-sa) * c.d0);
                    const double tt = -num / den;
                    geo = pl - (den < 0 ? 1 : 0);                    // 1 <= pl < na: already a sector
                    const double px = tt * c.d0 + c.x0, py = tt * c.d1 + c.x1;
                    // valid = !(|den| <= 1e-15) && !(ca px + sa py < 0) && t_in <= tt <= t_out, as ONE
                    // predicate chain (written as C++ the four tests become four pairs of selects)
                    t = select_in_range(tt, t_in, t_out, ca * px + sa * py, fabs(den), SPHRT_A_PARALLEL);
                }
                SPHRT_BAD_IF(k == a_count - 1 && t < SPHRT_INF);       // guard
            }
            const bool hit = t < SPHRT_INF;
            const uint32_t mv = __ballot_sync(SPHRT_FULL, hit);
            if (hit) {
                const int slot = __popc(mv & lt_mask);
                sts_f64(oT + 8u * slot, t);
                sts_s32(oR + 4u * slot, geo);
            }
            const int cnt = __popc(mv);
            nA += cnt; oT += 8u * cnt; oR += 4u * cnt;
        }
    }
    if (lane == 0) {
        sts_f64(tA + 8u * nA, SPHRT_INF);
        sts_s32(rA - 4u, R.sa0 >= 0 ? R.sa0 : kInvalidState);
    }
    const int nEA = nE + nA;
    if (nEA > P.capEA) {
#ifdef SPHRT_DEBUG_FAIL
        if (lane == 0) printf("capEA ray %d nE %d nA %d\n", R.ray, nE, nA);
#endif
        return -1;
    }
    if (SPHRT_FAILED()) return -1;
    // glancing crossings leave the elevation index unchanged: forward-fill (rare)
    if (__any_sync(SPHRT_FULL, lds_s32(flags + 4u) != 0)) {
        if (lane == 0) {
#pragma unroll 1
            for (int k = 0; k < nE; k++)
                if (lds_s32(rE + 4u * k) == kUnchangedState) sts_s32(rE + 4u * k, lds_s32(rE + 4u * k - 4u));
        }
        __syncwarp();
    }

    // ---- stage 1: elevation + azimuth -> (t, packed state after the event); both lists are
    // verified to ascend on the way (every entry is compared with its successor).  Written without
    // data-dependent control flow: ONE counter (elevation events taken) gives all four addresses,
    // and both heads are (re)loaded after every step.  The packed state is
    // e * na + a + (imin - 1) * ne * na, so that stage 2 only adds |radial steps from the turning
    // point| * ne * na to get the voxel offset.
    const uint32_t tEA = wbase + P.off_tEA, sEA = wbase + P.off_sEA + 4u;
    const int nena = T.ne * na;
    const int base_r = (R.imin - 1) * nena;
    {
#if SPHRT_SKEW1
        // shares of c1 or c1 - 1 entries (every fourth lane): with equal shares of an even size the
        // lanes' output slots fall on a few banks.  A lane with the shorter share writes the first
        // entry of the next one as well -- the same values.
        const int c1 = (nEA + 38) >> 5;
        const int g0 = lane * c1 - (lane >> 2);
#else
        const int c1 = (nEA + 31) >> 5;
        const int g0 = lane * c1;
#endif
        const int gs = g0 < nEA ? g0 : nEA;
        const int nmin = nE < nA ? nE : nA;
        int sE = warp_corank(tE, nE, tA, nA, gs, 32 - __clz(nmin > 0 ? nmin : 1));   // E events taken
        const int j0 = gs - sE;
        double xv = lds_f64(tE + 8u * sE), yv = lds_f64(tA + 8u * j0);
        bool viol = (sE > 0 && sE < nE && lds_f64(tE + 8u * sE - 8u) > xv) ||
                    (j0 > 0 && j0 < nA && lds_f64(tA + 8u * j0 - 8u) > yv);
        // head of A after k steps: tA[gs + k - sE]; state of A: rA[gs + k - sE - 1]
        uint32_t cAt = tA + 8u * gs, cAr = rA + 4u * gs - 4u;
        uint32_t oT = tEA + 8u * g0, oS = sEA + 4u * g0;
        const uint32_t rE1 = rE - 4u;
        int left = nEA - g0;                           // live entries of this share (may be <= 0)
        bool tx = xv <= yv;
        int se = 0, sa = 0;
        if (SPHRT_ONE_HEAD1) { se = lds_s32(rE1 + 4u * sE); sa = lds_s32(cAr - 4u * sE); }
#define SPHRT_STAGE1_STEP(K)                                                                      \
        {                                                                                         \
            const double t = tx ? xv : yv;                                                        \
            sE += tx ? 1 : 0;                                                                     \
            if (SPHRT_ONE_HEAD1) {                 /* only the consumed head / state is loaded again */ \
                const double v = lds_f64(tx ? tE + 8u * sE : cAt + 8u * ((K) + 1) - 8u * sE);     \
                const int ns = lds_s32(tx ? rE1 + 4u * sE : cAr + 4u * ((K) + 1) - 4u * sE);      \
                xv = tx ? v : xv; yv = tx ? yv : v;                                               \
                se = tx ? ns : se; sa = tx ? sa : ns;                                             \
            } else {                                                                              \
                xv = lds_f64(tE + 8u * sE);                                                       \
                yv = lds_f64(cAt + 8u * ((K) + 1) - 8u * sE);                                     \
                se = lds_s32(rE1 + 4u * sE); sa = lds_s32(cAr + 4u * ((K) + 1) - 4u * sE);        \
            }                                                                                     \
            const int st = se + sa + base_r;                                                      \
            const bool live = (K) < left;          /* beyond the end the heads are stale memory */ \
            viol |= live && (xv < t || yv < t);                                                   \
            if (live) { sts_f64(oT + 8u * (K), t); sts_s32(oS + 4u * (K), st); }                  \
            tx = xv <= yv;                                                                        \
        }
        int q = 0;
        for (; q + 2 <= c1; q += 2) {                  // same trip count in every lane
            SPHRT_STAGE1_STEP(0)
            SPHRT_STAGE1_STEP(1)
            cAt += 16u; cAr += 8u; oT += 16u; oS += 8u; left -= 2;
        }
        if (q < c1) SPHRT_STAGE1_STEP(0)
#undef SPHRT_STAGE1_STEP
        if (lane == 0) {
            sts_f64(tEA + 8u * nEA, SPHRT_INF);
            sts_s32(sEA - 4u, lds_s32(rE - 4u) + lds_s32(rA - 4u) + base_r);
        }
        if (__any_sync(SPHRT_FULL, viol)) {
#ifdef SPHRT_DEBUG_FAIL
            if (viol) printf("viol ray %d lane %d nE %d nA %d allp %d\n", R.ray, lane, nE, nA, (int)all_planes);
#endif
            return -1;
        }
    }
    __syncwarp();                                      // phase A lists are dead from here on

    // ---- radial list (overlays phase A): first roots of shells nr .. imin, then second roots of
    // imin .. nr; padded with +inf so that the merge below may run past its end
    const uint32_t tR = wbase + P.off_tR;
    const int m = T.nr - R.imin + 1;
    const int nR = 2 * m;
    {
        const double tang = 1e-17 * (1.0 + c.nx2);
        uint32_t src = T.r2 + 8u * (T.nr - lane), lo = tR + 8u * lane, hi = tR + 8u * (nR - 1 - lane);
#pragma unroll 1
        for (int k = lane; k < m; k += 32, src -= 256u, lo += 256u, hi -= 256u) {   // both roots share the square root
            const double r2 = lds_f64(src);
            const double arg = r2 - c.dperp2;
            const double t1c = sqrt(arg);
            // inward roots enter shell i - 1, outward roots shell i; when tangent within rounding
            // the reference decides from the sign of d.p: bail out.  (Not for an innermost "shell" of
            // radius 0 met by a ray through the origin: both roots are tc, the reference gives both
            // region 0, here the ray is below the grid for a length of exactly zero in between.)
            SPHRT_BAD_IF(!(arg >= tang + 1e-17 * r2) && r2 != 0.0);
            sts_f64(lo, c.tc - t1c);
            sts_f64(hi, c.tc + t1c);
        }
        for (int k = lane; k < P.padR; k += 32) sts_f64(tR + 8u * (nR + k), SPHRT_INF);
    }
    if (SPHRT_FAILED()) return -1;

    // ---- stage 2: radial + (elevation, azimuth); every merged event starts one segment that ends
    // at the next event (the reference's sort + forward-fill + diff, raytracer.py:131-151).
    // ONE signed counter u = (radial events taken) - m gives everything: the radial head is
    // tR[m + u], the other head is tEA[rank - m - u], and the radial index after the events taken
    // is imin - 1 + |u| (inward events enter shells nr-1 .. imin-1, outward events imin .. nr), so
    // the voxel offset is |u| * ne * na + packed state.  Before the entry and after the exit that
    // is r = nr, below the innermost shell r = -1, and an (e, a) index outside the grid is a large
    // negative number: one unsigned comparison with the voxel count tests all of it.  A length is
    // never negative (the lists were verified); zero lengths (ties) add nothing.
    const int n = nR + nEA;
    {
        const unsigned nvox = (unsigned)(T.nr * nena);
        const int c2 = (n + 31) >> 5;
        const int g0 = lane * c2;
        const int gs = g0 < n ? g0 : n;
        const int nmin = nR < nEA ? nR : nEA;
        const int i0 = warp_corank(tR, nR, tEA, nEA, gs, 32 - __clz(nmin > 0 ? nmin : 1));
        // One sweep over this lane's share of the merged sequence.  A sink may ask for a second
        // one (Sink::kTwoSweeps: the fused normal-equations step scatters A^T of a weight that
        // depends on the ray's own integral): the lists are still in shared memory, only the
        // segment loop is repeated.
        auto sweep = [&](auto& sink) {
        int u = i0 - m;
        const uint32_t cR = tR + 8u * m;               // radial head: cR + 8 u
        uint32_t cE = tEA + 8u * (gs - m);             // other head after k steps: cE + 8 k - 8 u
        uint32_t cS = sEA + 4u * (gs - m) - 4u;        // its state:               cS + 4 k - 4 u
        double xv = lds_f64(cR + 8u * u), yv = lds_f64(cE - 8u * u);
        bool tx = xv <= yv;
        double t = tx ? xv : yv;
#define SPHRT_WARP_STEP(SLOT, K)                                                                  \
        {                                                                                         \
            u += tx ? 1 : 0;                                                                      \
            if (SPHRT_ONE_HEAD) {                  /* only the consumed head is loaded again */    \
                const double v = lds_f64(tx ? cR + 8u * u : cE + 8u * ((K) + 1) - 8u * u);        \
                xv = tx ? v : xv; yv = tx ? yv : v;                                               \
            } else {                                                                              \
                xv = lds_f64(cR + 8u * u);                                                        \
                yv = lds_f64(cE + 8u * ((K) + 1) - 8u * u);                                       \
            }                                                                                     \
            const int sea = lds_s32(cS + 4u * ((K) + 1) - 4u * u);                                \
            const unsigned idx = (unsigned)((u < 0 ? -u : u) * nena + sea);                       \
            tx = xv <= yv;                                                                        \
            const double tn = tx ? xv : yv;                                                       \
            const double len = tn - t;                                                            \
            const bool ok = idx < nvox;                                                           \
            if constexpr (kEvents) {                                                              \
                if (g0 + q + (K) < n) sink.event(g0 + q + (K), R.imin - 1 + (u < 0 ? -u : u), sea - base_r, na, \
                                                 ok && len > 0 && tn < SPHRT_INF ? len : 0.0);    \
            } else {                                                                              \
                sink.template seg<SLOT>(ok, idx, len);                                            \
            }                                                                                     \
            t = tn;                                                                               \
        }
        int q = 0;
        if constexpr (kWarpDepth == 4) {
            for (; q + 4 <= c2; q += 4) {
                SPHRT_WARP_STEP(0, 0)
                SPHRT_WARP_STEP(1, 1)
                SPHRT_WARP_STEP(2, 2)
                SPHRT_WARP_STEP(3, 3)
                cE += 32u; cS += 16u;
            }
            if (q + 2 <= c2) {
                SPHRT_WARP_STEP(0, 0)
                SPHRT_WARP_STEP(1, 1)
                cE += 16u; cS += 8u; q += 2;
            }
            if (q < c2) SPHRT_WARP_STEP(2, 0)
        } else {
            for (; q + 2 <= c2; q += 2) {
                SPHRT_WARP_STEP(0, 0)
                SPHRT_WARP_STEP(1, 1)
                cE += 16u; cS += 8u;
            }
            if (q < c2) SPHRT_WARP_STEP(0, 0)
        }
#undef SPHRT_WARP_STEP
        };
        sweep(sink);
        if constexpr (Sink::kTwoSweeps) {
            auto second = sink.second(lane);
            if (second.active) sweep(second);
        }
    }
#undef SPHRT_BAD_IF
#undef SPHRT_UNCHANGED_IF
#undef SPHRT_FAILED
    return n;
}

}  // namespace sphrt
