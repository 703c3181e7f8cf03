// sphrt_walker.cuh -- per-ray crossing generator for a spherical voxel grid.
//
// One ray is walked by one thread.  Instead of materialising and sorting every
// candidate intersection like the reference's trace_indices()
// (sph_raytracer/raytracer.py:48-230), three monotone event streams are generated
// lazily in registers and merged on the fly:
//
//   radial    : shells are met outermost->innermost (first roots) and then
//               innermost->outermost (second roots); both runs are monotone in t
//               because sqrt and +- are monotone                    (r_torch, :248-325)
//   elevation : the colatitude along a line is unimodal, so the cones crossed next are
//               always the two that bound the current elevation bin; each cone is
//               evaluated with the reference's exact quadratic incl. its snapping
//               thresholds                                          (e_torch, :328-468)
//   azimuth   : azimuth is monotone along a line; the next half-plane is one of the two
//               that bound the current sector                       (a_torch, :471-552)
//
// Every event carries the reference's *absolute* region value (new index of that
// coordinate after the crossing, -1 outside, -2 "leave unchanged"), so the state
// machine below is the reference's sort + forward_fill (:131-146) restated as a merge.
// Arithmetic is IEEE double with the fused operations placed exactly where the
// reference's ATen kernels fuse; translation units including
// this header must be compiled with -fmad=false (nvcc) / -ffp-contract=off (host test
// build) so that no other contraction happens.
//
// The header is __host__ __device__ clean: tests/host_walker.cpp compiles it with g++
// for CPU-side algorithm tests.  The product only ever runs it on the GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SPHRT_HD __host__ __device__ __forceinline__
#else
#define SPHRT_HD inline
#endif

namespace sphrt {

// isclose() threshold of the reference: finfo(float64).resolution ** (1/3)  (:233-246)
#define SPHRT_TAU 0x1.4f8b588e368f4p-17
// tr.isclose(cross, 0, atol=finfo.resolution)                               (:521)
#define SPHRT_A_PARALLEL 1e-15
#define SPHRT_HALF_PI 0x1.921fb54442d18p+0
#define SPHRT_PI 0x1.921fb54442d18p+1
#define SPHRT_INF (__builtin_huge_val())

// All boundary tables live in ONE packed array (one base register instead of seven):
//   [ r_b (nr+1) | e_b (ne+1) | cos_e (ne+1) | a_b (na+1) | cos_a (na+1) | sin_a (na+1) | c2_sorted ]
struct GridView {
    int nr, ne, na;        // voxel counts (boundary tables have n+1 entries)
    int wrap;              // azimuth covers [-pi, pi]: planes 0 and na coincide (:528)
    int n_c2;
    const double* tab;
    SPHRT_HD const double* r_b() const { return tab; }                               // shell radii
    SPHRT_HD const double* e_b() const { return tab + (nr + 1); }                    // cone angles
    SPHRT_HD const double* cos_e() const { return tab + (nr + 1) + (ne + 1); }       // torch.cos(e_b)
    SPHRT_HD const double* a_b() const { return tab + (nr + 1) + 2 * (ne + 1); }     // plane angles
    SPHRT_HD const double* cos_a() const { return a_b() + (na + 1); }                // torch.cos(a_b)
    SPHRT_HD const double* sin_a() const { return a_b() + 2 * (na + 1); }            // torch.sin(a_b)
    // ascending cos^2(e_b) of the cones with |e - pi/2| >= tau
    SPHRT_HD const double* c2_sorted() const { return a_b() + 3 * (na + 1); }
    SPHRT_HD static int table_doubles(int nr, int ne, int na, int n_c2)
    {
        return (nr + 1) + 2 * (ne + 1) + 3 * (na + 1) + n_c2;
    }
};

SPHRT_HD double fma_rn(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
// einsum('...j,...j->...'): separately rounded products, left-to-right sum (Appendix A)
SPHRT_HD double dot3(double a0, double a1, double a2, double b0, double b1, double b2)
{
    return (a0 * b0 + a1 * b1) + a2 * b2;
}
// tr.linalg.norm over a 3-vector: fused sum of squares (Appendix A)
SPHRT_HD double norm3(double v0, double v1, double v2)
{
    return sqrt(fma_rn(v2, v2, fma_rn(v1, v1, v0 * v0)));
}

// searchsorted(b, v, right=True) - 1 with the fix-ups of find_starts (:630-642).
// Returns the *geometric* bin in [-1, n]; `state` receives the reference's region
// (n -> -1, v == b[n] -> n-1).
SPHRT_HD int locate(const double* b, int n, double v, int* state)
{
    int lo = 0, hi = n + 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (b[mid] > v) hi = mid; else lo = mid + 1;
    }
    int bin = lo - 1;
    if (v == b[n]) bin = n - 1;
    *state = (bin == n) ? -1 : bin;
    return bin;
}

struct Event {
    double t;      // distance along the ray (+inf: no event)
    int region;    // reference region value: >=0 new index, -1 outside, -2 unchanged
    int geo;       // index - negative_crossing before any snapping (geometric bin entered)
};

// ---------------------------------------------------------------------------------------
// Per-ray constants shared by the three families.
struct RayConsts {
    double x0, x1, x2;        // start
    double r0, r1, r2;        // direction used by r_torch  (normalised once more, :281)
    double d0, d1, d2;        // direction used by e_torch / a_torch (normalised again, :365)
    double tc, dperp2;        // radial: closest-approach parameter and d_perp**2 (:288-289)
    double dz2, dzxz, dx, xz2, nx2;   // elevation quadratic ingredients (:373-375)
};

SPHRT_HD void ray_setup(RayConsts& c, const double x[3], const double d_in[3])
{
    c.x0 = x[0]; c.x1 = x[1]; c.x2 = x[2];
    double n = norm3(d_in[0], d_in[1], d_in[2]);
    c.r0 = d_in[0] / n; c.r1 = d_in[1] / n; c.r2 = d_in[2] / n;
    n = norm3(c.r0, c.r1, c.r2);
    c.d0 = c.r0 / n; c.d1 = c.r1 / n; c.d2 = c.r2 / n;
    c.tc = dot3(-c.x0, -c.x1, -c.x2, c.r0, c.r1, c.r2);
    const double dperp = sqrt(dot3(c.x0, c.x1, c.x2, c.x0, c.x1, c.x2) - c.tc * c.tc);
    c.dperp2 = dperp * dperp;
    c.dz2 = c.d2 * c.d2;
    c.dzxz = c.d2 * c.x2;
    c.dx = dot3(c.d0, c.d1, c.d2, c.x0, c.x1, c.x2);
    c.xz2 = c.x2 * c.x2;
    const double nx = norm3(c.x0, c.x1, c.x2);
    c.nx2 = nx * nx;
}

// ---- radial family (r_torch :293-323) -------------------------------------------------
// root of shell i: second == false -> tc - sqrt(r_i^2 - dperp^2), true -> tc + sqrt(...)
SPHRT_HD Event radial_event(const GridView& g, const RayConsts& c, int i, bool second)
{
    Event ev;
    const double r = g.r_b()[i];
    const double t1c = sqrt(r * r - c.dperp2);
    const double t = second ? c.tc + t1c : c.tc - t1c;
    const double p0 = c.r0 * t + c.x0, p1 = c.r1 * t + c.x1, p2 = c.r2 * t + c.x2;
    const int neg = dot3(c.r0, c.r1, c.r2, p0, p1, p2) < 0;
    ev.geo = i - neg;
    ev.region = (ev.geo == g.nr) ? -1 : ev.geo;
    ev.t = (t != t) ? SPHRT_INF : t;
    return ev;
}

// ---- elevation family (e_torch :373-466) ---------------------------------------------------
// Quadratic aa t^2 + bb t + cc of cone i along the ray (:373-375); aa carries the
// |aa| < tau snap (:376), delta is still unsnapped.
struct ConeQuad { double aa, bb, cc, delta; };

// (the same with the cone's cosine passed in: kernels that keep the tables behind plain
// shared-memory addresses load it themselves)
SPHRT_HD ConeQuad cone_quadratic_ce(const RayConsts& c, double ce);

SPHRT_HD ConeQuad cone_quadratic(const GridView& g, const RayConsts& c, int i)
{
    return cone_quadratic_ce(c, g.cos_e()[i]);
}

SPHRT_HD ConeQuad cone_quadratic_c2(const RayConsts& c, double c2);
SPHRT_HD ConeQuad cone_quadratic_ce(const RayConsts& c, double ce)
{
    return cone_quadratic_c2(c, ce * ce);
}
// (... or its square: tables of cos^2 hold the same product)
SPHRT_HD ConeQuad cone_quadratic_c2(const RayConsts& c, double c2)
{
    ConeQuad q;
    q.aa = c.dz2 - c2;
    q.bb = 2 * (c.dzxz - c.dx * c2);
    q.cc = c.xz2 - c.nx2 * c2;
    if (fabs(q.aa) < SPHRT_TAU) q.aa = 0;
    q.delta = q.bb * q.bb - (4 * q.aa) * q.cc;
    return q;
}

// the two roots for a given discriminant, with the parallel / degenerate branches (:385-410)
SPHRT_HD void cone_roots(const ConeQuad& q, double delta, double& t1, double& t2)
{
    const double sq = sqrt(delta);
    t1 = (-q.bb + sq) / (2 * q.aa);
    t2 = (-q.bb - sq) / (2 * q.aa);
    if (q.aa == 0 && !(fabs(q.bb) < SPHRT_TAU)) { t1 = -q.cc / q.bb; t2 = SPHRT_INF; }
    if (q.aa == 0 && q.bb == 0 && q.cc == 0) { t1 = SPHRT_INF; t2 = SPHRT_INF; }
}

// cone_roots() for the GPU's warp walkers: a negative (or NaN) discriminant does not go through
// sqrt and the two divisions -- their NaN slow paths cost a warp ~100 instructions for one lane, and
// the guard candidates at the ends of every cone range have exactly that -- the roots are NaN as in
// the reference (sqrt of a negative number, :385-386).  Same results as cone_roots().
SPHRT_HD void cone_roots_guarded(const ConeQuad& q, double delta, double& t1, double& t2)
{
    const bool none = !(delta >= 0);
    // (sqrt(0) itself takes the slow path too: 1.0 stands in for both and the zero is put back.
    // The empty asm keeps the select in FRONT of the square root: the compiler otherwise takes
    // sqrt(delta) first and selects afterwards, which is the same value and the slow path again --
    // 0.75 calls per config-5 ray, one lane each.)
    double arg = (none || delta == 0) ? 1.0 : delta;
#ifdef __CUDA_ARCH__
    asm("" : "+d"(arg));
#endif
    double sq = sqrt(arg);
    if (delta == 0) sq = 0.0;
    t1 = (-q.bb + sq) / (2 * q.aa);
    t2 = (-q.bb - sq) / (2 * q.aa);
    if (none) { t1 = __builtin_nan(""); t2 = __builtin_nan(""); }
    if (q.aa == 0 && !(fabs(q.bb) < SPHRT_TAU)) { t1 = -q.cc / q.bb; t2 = SPHRT_INF; }
    if (q.aa == 0 && q.bb == 0 && q.cc == 0) { t1 = SPHRT_INF; t2 = SPHRT_INF; }
}

// region a ray enters through cone i at distance t (:418-446, 464): sign of d . (p x phi_hat),
// glancing crossings leave the index unchanged (-2)
SPHRT_HD int cone_region_ne(int ne, const RayConsts& c, int i, double t, int* geo);
SPHRT_HD int cone_region(const GridView& g, const RayConsts& c, int i, double t, int* geo)
{
    return cone_region_ne(g.ne, c, i, t, geo);
}
SPHRT_HD int cone_region_ne(int ne, const RayConsts& c, int i, double t, int* geo)
{
    const double p0 = c.d0 * t + c.x0, p1 = c.d1 * t + c.x1, p2 = c.d2 * t + c.x2;
    const double n0 = -(p2 * p0), n1 = -(p2 * p1), n2 = fma_rn(p0, p0, p1 * p1);
    const double prod = dot3(c.d0, c.d1, c.d2, n0, n1, n2);
    const int neg = prod > 0;
    *geo = i - neg;
    int region = i - neg;
    if (fabs(prod) < SPHRT_TAU) region = -2;
    if (region == ne) region = -1;
    return region;
}

// true if the root at distance t lies on the mirror nappe of cone i (:449-461)
SPHRT_HD bool cone_shadow_v(const RayConsts& c, double ce, double eb, double t)
{
    const double p2 = c.d2 * t + c.x2;
    bool shadow = (p2 >= 0) != (ce >= 0);
    if (fabs(SPHRT_HALF_PI - eb) < SPHRT_TAU) shadow = false;
    return shadow;
}
// the same from a per-cone sign: 0 for the cones exempt by |e - pi/2| < tau, else +1 / -1 for
// cos e >= 0 / < 0
SPHRT_HD bool cone_shadow_sgn(const RayConsts& c, double sgn, double t)
{
    const double p2 = c.d2 * t + c.x2;
    return (p2 >= 0) ? (sgn < 0) : (sgn > 0);
}
SPHRT_HD double cone_shadow_sign(double ce, double eb)
{
    if (fabs(SPHRT_HALF_PI - eb) < SPHRT_TAU) return 0.0;
    return ce >= 0 ? 1.0 : -1.0;
}
SPHRT_HD bool cone_shadow(const GridView& g, const RayConsts& c, int i, double t)
{
    return cone_shadow_v(c, g.cos_e()[i], g.e_b()[i], t);
}

// region / shadow / NaN handling of one cone root (:418-466)
SPHRT_HD Event cone_event(const GridView& g, const RayConsts& c, int i, double t)
{
    Event ev;
    ev.region = cone_region(g, c, i, t, &ev.geo);
    if (cone_shadow(g, c, i, t)) t = SPHRT_INF;
    ev.t = (t != t) ? SPHRT_INF : t;
    return ev;
}

// ---- azimuth family (a_torch :505-550): the single root of half-plane i -------------------
SPHRT_HD Event plane_event_v(double ca, double sa, const RayConsts& c, int i, int na, int wrap);
SPHRT_HD Event plane_event(const GridView& g, const RayConsts& c, int i)
{
    return plane_event_v(g.cos_a()[i], g.sin_a()[i], c, i, g.na, g.wrap);
}
SPHRT_HD Event plane_event_v(double ca, double sa, const RayConsts& c, int i, int na, int wrap)
{
    Event ev;
    const double num = fma_rn(0.0, c.x2, fma_rn(ca, c.x1, (-sa) * c.x0));
    const double den = fma_rn(0.0, c.d2, fma_rn(ca, c.d1, (-sa) * c.d0));
    double t = -num / den;
    const double cross = fma_rn(ca, c.d1, -(sa * c.d0));
    if (fabs(cross) <= SPHRT_A_PARALLEL) t = SPHRT_INF;
    const int neg = cross < 0;
    ev.geo = i - neg;                      // -1 .. na ; -1 and na are the gap sector
    int region = ev.geo;
    if (wrap) { region %= na; if (region < 0) region += na; }
    else if (region == na) region = -1;
    const double px = t * c.d0 + c.x0, py = t * c.d1 + c.x1;
    if (ca * px + sa * py < 0) t = SPHRT_INF;
    ev.region = region;
    ev.t = (t != t) ? SPHRT_INF : t;
    return ev;
}

// ---------------------------------------------------------------------------------------
// Stream states.  Each yields its events with t >= 0 in ascending t.

struct RadialStream {
    int i;          // next shell index
    int second;     // 0: inward run (first roots), 1: outward run (second roots), 2: done
    Event head;

    SPHRT_HD void advance(const GridView& g, const RayConsts& c)
    {
        // skip events behind the start (t < 0): they only matter for quirk 1 (see walk)
        for (;;) {
            if (second == 0) {
                if (i < 0) { second = 1; i = 0; continue; }
                Event ev = radial_event(g, c, i, false);
                if (ev.t == SPHRT_INF) { second = 1; i = i + 1; continue; }   // shell not reached
                i -= 1;
                if (ev.t < 0) continue;
                head = ev; return;
            } else if (second == 1) {
                if (i > g.nr) { second = 2; continue; }
                Event ev = radial_event(g, c, i, true);
                i += 1;
                if (ev.t < 0 || ev.t == SPHRT_INF) continue;
                head = ev; return;
            } else {
                head.t = SPHRT_INF; head.region = -2; head.geo = -2; return;
            }
        }
    }
    SPHRT_HD void init(const GridView& g, const RayConsts& c)
    {
        i = g.nr; second = 0;
        // NaN dperp2 (ray numerically through the origin): every root is NaN -> no events
        if (c.dperp2 != c.dperp2) second = 2;
        advance(g, c);
    }
};

// two roots of one bounding cone that are still ahead of the ray
struct ConeSlot {
    double ta, tb;     // ta <= tb ; +inf = none
    int ra, rb;        // packed region/geo of each
};
SPHRT_HD int pack_rg(const Event& e) { return (e.region & 0xffff) | (e.geo << 16); }
SPHRT_HD int unpack_region(int p) { return (int)(int16_t)(p & 0xffff); }
SPHRT_HD int unpack_geo(int p) { return p >> 16; }

SPHRT_HD bool behind(double t, double t_now, bool strict)
{
    return t < t_now || (strict && t == t_now);
}

// reference roots of cone i (incl. the |delta| < tau snap, :383) as events
SPHRT_HD void cone_events(const GridView& g, const RayConsts& c, int i, Event& e1, Event& e2)
{
    const ConeQuad q = cone_quadratic(g, c, i);
    double t1, t2;
    cone_roots(q, fabs(q.delta) < SPHRT_TAU ? 0.0 : q.delta, t1, t2);
    e1 = cone_event(g, c, i, t1);
    e2 = cone_event(g, c, i, t2);
}

// roots of cone i that are still ahead of the ray (t >= t_now, or t > t_now if strict)
SPHRT_HD ConeSlot fresh_cone(const GridView& g, const RayConsts& c, int i, double t_now,
                             bool strict)
{
    ConeSlot s; s.ta = SPHRT_INF; s.tb = SPHRT_INF; s.ra = 0; s.rb = 0;
    if (i < 0 || i > g.ne) return s;
    Event e1, e2;
    cone_events(g, c, i, e1, e2);
    if (behind(e1.t, t_now, strict)) e1.t = SPHRT_INF;
    if (behind(e2.t, t_now, strict)) e2.t = SPHRT_INF;
    if (e2.t < e1.t) { Event tmp = e1; e1 = e2; e2 = tmp; }
    s.ta = e1.t; s.ra = pack_rg(e1);
    s.tb = e2.t; s.rb = pack_rg(e2);
    return s;
}

// ---------------------------------------------------------------------------------------
// The snap |delta| < tau (:383) is an absolute threshold on delta = 4 c2 (A - c2 B), with
// c2 = cos^2(e_i), A = |dz x - xz d|^2 and B = |x|^2 - (d.x)^2.  For an ordinary ray it only
// fires on the plane-like cone e = pi/2 (c2 ~ 0), where it is harmless.  For rays that run
// in the equatorial plane or pass close to the origin it fires on whole ranges of cones;
// the reference then reports their crossings at the midpoint of the real and the
// shadow-nappe root (or not at all), i.e. at distances unrelated to where the ray meets
// the cone, and adjacency between consecutive events is lost.  Such rays are detected up
// front (conservatively) and walked from an explicitly sorted list of every cone root
// instead of the bin tracker, which keeps them identical to the reference.
//
// c2_sorted: ascending cos^2(e_i) of all cones that are not exempt by |e_i - pi/2| < tau.
SPHRT_HD bool table_hits(const double* tab, int n, double lo, double hi)
{
    int a = 0, b = n;                      // first index with tab[idx] >= lo
    while (a < b) { const int m = (a + b) >> 1; if (tab[m] < lo) a = m + 1; else b = m; }
    return a < n && tab[a] < hi;
}
// The snap |aa| < tau (:376) turns a cone's quadratic into the linear equation bb t + cc = 0 when
// the ray's inclination matches the cone's (dz^2 ~ cos^2 e_i).  For the plane-like cone
// e = pi/2 that root lies at HALF the distance of the real crossing (the quadratic there is the
// perfect square (dz t + xz)^2), so for nearly horizontal rays the reference changes the
// elevation index somewhere else than the geometry does.  The walkers reproduce that by
// evaluating the same formulas; but a displaced root also breaks the adjacency the bin tracker
// relies on (sequences like 71 -> 73 -> 72 in the reference), and the geometric elevation bin of
// a point is no longer the reference's state there.  Such rays (0.1-2 % at config 5) are walked
// from the sorted list of all cone roots, like the rays with snapped discriminants.
SPHRT_HD bool ray_has_snapped_aa(const RayConsts& c, const double* c2_sorted, int n_c2)
{
    const double w = SPHRT_TAU * 1.0001;
    if (!(c.dz2 >= w + 1e-9)) return true;         // cones within tau of pi/2 (not in the table)
    return table_hits(c2_sorted, n_c2, c.dz2 - w, c.dz2 + w);
}

SPHRT_HD bool ray_needs_cone_list(const RayConsts& c, const double* c2_sorted, int n_c2)
{
    if (ray_has_snapped_aa(c, c2_sorted, n_c2)) return true;
    const double v0 = c.d2 * c.x0 - c.x2 * c.d0, v1 = c.d2 * c.x1 - c.x2 * c.d1;
    const double A = v0 * v0 + v1 * v1;
    double B = c.nx2 - c.dx * c.dx;
    if (!(B > 0)) B = 0;
    // tau/4 with a margin for the rounding noise of bb^2 - 4 aa cc
    const double eps = 0.25 * SPHRT_TAU * 1.001 + 1e-12 * (1.0 + c.nx2);
    if (!(A == A)) return true;
    // { c >= 0 : c |A - c B| < eps } = [0, cm) u (cp, c3)   (one interval if A^2 < 4 B eps)
    const double disc = A * A - 4 * B * eps;
    const double c3 = B > 0 ? (A + sqrt(A * A + 4 * B * eps)) / (2 * B) : SPHRT_INF;
    if (disc < 0) return table_hits(c2_sorted, n_c2, 0.0, c3);
    const double sq = sqrt(disc);
    const double cm = 2 * eps / (A + sq);
    if (table_hits(c2_sorted, n_c2, 0.0, cm)) return true;
    if (B > 0) {
        const double cp = (A + sq) / (2 * B);
        if (table_hits(c2_sorted, n_c2, cp, c3)) return true;
    }
    return false;
}

// Per-thread scratch for the sorted cone-root list (element k at [k * stride]).
struct ConeList {
    double* t; int* r; int64_t stride; int capacity;
    SPHRT_HD double& T(int k) const { return t[(int64_t)k * stride]; }
    SPHRT_HD int& R(int k) const { return r[(int64_t)k * stride]; }
};

// Elevation events in ascending t.
//   tracker mode: the next cone crossed is always one of the two that bound the current
//                 elevation bin; each is evaluated when it becomes a bounding cone.
//   list mode   : all roots >= 0 of all cones, heap-sorted in scratch memory.
struct ElevationStream {
    int bin;            // tracker: geometric elevation bin in [-1, ne]; cones bin, bin+1
    ConeSlot lo, hi;
    int probed;         // tracker: neighbours of an exhausted bin were already probed
    int list_n, list_i; // list mode: number of events, next index (list_n < 0: tracker mode)
    Event head;

    SPHRT_HD void peek()
    {
        const int p = (lo.ta <= hi.ta) ? lo.ra : hi.ra;
        head.t = fmin(lo.ta, hi.ta); head.region = unpack_region(p); head.geo = unpack_geo(p);
    }
    // Both bounding cones are exhausted.  Normally the ray stays in this bin for good; the
    // neighbours are probed once in case a rounding-level decision (shadow sign) hid a
    // crossing.  Strict comparison keeps coincident apex roots from re-entering.
    SPHRT_HD void recover(const GridView& g, const RayConsts& c, double t_now)
    {
        if (probed || lo.ta != SPHRT_INF || hi.ta != SPHRT_INF) return;
        probed = 1;
        const ConeSlot dn = fresh_cone(g, c, bin - 1, t_now, true);
        const ConeSlot up = fresh_cone(g, c, bin + 2, t_now, true);
        if (dn.ta == SPHRT_INF && up.ta == SPHRT_INF) return;
        if (dn.ta <= up.ta) { hi = lo; lo = dn; bin -= 1; }
        else                { lo = hi; hi = up; bin += 1; }
    }
    SPHRT_HD void init_tracker(const GridView& g, const RayConsts& c, int start_bin)
    {
        list_n = -1; list_i = 0;
        bin = start_bin; probed = 0;
        lo = fresh_cone(g, c, bin, 0.0, false);
        hi = fresh_cone(g, c, bin + 1, 0.0, false);
        recover(g, c, 0.0);
        peek();
    }
    SPHRT_HD void list_head(const ConeList& L)
    {
        if (list_i < list_n) { head.t = L.T(list_i); head.region = L.R(list_i); }
        else { head.t = SPHRT_INF; head.region = -2; }
        head.geo = 0;
    }
    SPHRT_HD void init_list(const GridView& g, const RayConsts& c, const ConeList& L)
    {
        int n = 0;
        for (int i = 0; i <= g.ne; i++) {
            Event e1, e2;
            cone_events(g, c, i, e1, e2);
            if (e1.t >= 0 && e1.t < SPHRT_INF && n < L.capacity) { L.T(n) = e1.t; L.R(n) = e1.region; n++; }
            if (e2.t >= 0 && e2.t < SPHRT_INF && n < L.capacity) { L.T(n) = e2.t; L.R(n) = e2.region; n++; }
        }
        // heap sort, ascending t
        for (int start = n / 2 - 1, end = n; ; ) {
            int root;
            double tv; int rv;
            if (start >= 0) { root = start; tv = L.T(root); rv = L.R(root); start--; }
            else {
                if (--end <= 0) break;
                tv = L.T(end); rv = L.R(end); L.T(end) = L.T(0); L.R(end) = L.R(0); root = 0;
            }
            for (;;) {
                int child = 2 * root + 1;
                if (child >= end) break;
                if (child + 1 < end && L.T(child + 1) > L.T(child)) child++;
                if (!(L.T(child) > tv)) break;
                L.T(root) = L.T(child); L.R(root) = L.R(child); root = child;
            }
            L.T(root) = tv; L.R(root) = rv;
        }
        list_n = n; list_i = 0;
        lo.ta = lo.tb = hi.ta = hi.tb = SPHRT_INF; lo.ra = lo.rb = hi.ra = hi.rb = 0;
        bin = 0; probed = 1;
        list_head(L);
    }
    // consume the head event and bring up the next one
    template <bool kList>
    SPHRT_HD void advance(const GridView& g, const RayConsts& c, const ConeList& L)
    {
        if (kList) { list_i++; list_head(L); return; }
        const double t_now = head.t;
        const bool from_lo = lo.ta <= hi.ta;
        ConeSlot& s = from_lo ? lo : hi;
        const int nb = head.geo;
        if (s.tb == s.ta) { s.ta = SPHRT_INF; s.tb = SPHRT_INF; }   // merged double root
        else { s.ta = s.tb; s.ra = s.rb; s.tb = SPHRT_INF; }
        if (from_lo && nb == bin - 1) {            // went down through the lower cone
            hi = lo; bin = nb; lo = fresh_cone(g, c, bin, t_now, false); probed = 0;
        } else if (!from_lo && nb == bin + 1) {    // went up through the upper cone
            lo = hi; bin = nb; hi = fresh_cone(g, c, bin + 1, t_now, false); probed = 0;
        }
        recover(g, c, t_now);
        peek();
    }
};

struct AzimuthStream {
    int sector;         // 0..na-1 voxel sectors; na = gap sector (zero width when wrap)
    double t_lo, t_hi;  // pending root of the lower / upper bounding plane
    int r_lo, r_hi;
    Event head;

    SPHRT_HD int n_sectors(const GridView& g) const { return g.wrap ? g.na : g.na + 1; }
    SPHRT_HD int lower_plane(const GridView&, int s) const { return s; }
    SPHRT_HD int upper_plane(const GridView& g, int s) const
    {
        if (g.wrap) return s + 1;                 // s == na-1 -> plane na (== plane 0, the seam)
        return s == g.na ? 0 : s + 1;
    }
    // root of plane i ahead of the ray; on a wrapped grid planes 0 and na are the same
    // half-plane evaluated twice by the reference -- the earlier of the two is the boundary
    SPHRT_HD Event eval(const GridView& g, const RayConsts& c, int i, double t_now) const
    {
        Event ev = plane_event(g, c, i);
        if (ev.t < t_now) ev.t = SPHRT_INF;
        if (g.wrap && (i == 0 || i == g.na)) {
            Event tw = plane_event(g, c, g.na - i);
            if (tw.t < t_now) tw.t = SPHRT_INF;
            if (tw.t < ev.t) ev = tw;
        }
        return ev;
    }
    SPHRT_HD int geo_to_sector(const GridView& g, int geo) const
    {
        if (g.wrap) { int s = geo % g.na; return s < 0 ? s + g.na : s; }
        return (geo < 0 || geo >= g.na) ? g.na : geo;
    }
    SPHRT_HD void peek()
    {
        if (t_lo <= t_hi) { head.t = t_lo; head.region = unpack_region(r_lo); head.geo = unpack_geo(r_lo); }
        else              { head.t = t_hi; head.region = unpack_region(r_hi); head.geo = unpack_geo(r_hi); }
    }
    SPHRT_HD void init(const GridView& g, const RayConsts& c, int start_bin)
    {
        sector = geo_to_sector(g, start_bin);
        Event a = eval(g, c, lower_plane(g, sector), 0.0);
        Event b = eval(g, c, upper_plane(g, sector), 0.0);
        t_lo = a.t; r_lo = pack_rg(a);
        t_hi = b.t; r_hi = pack_rg(b);
        if (t_lo == SPHRT_INF && t_hi == SPHRT_INF) {
            // No root ahead in either bounding plane: a start ON a half-plane whose t ~ 0 root came
            // out negative is geometrically in the neighbouring sector already (the reference
            // ignores that root and meets the next plane with the start bin still in place).
            const int nsec = n_sectors(g);
            const int sd = sector == 0 ? nsec - 1 : sector - 1, su = sector == nsec - 1 ? 0 : sector + 1;
            a = eval(g, c, lower_plane(g, sd), 0.0);
            b = eval(g, c, upper_plane(g, su), 0.0);
            if (a.t <= b.t) { if (a.t < SPHRT_INF) { sector = sd; t_lo = a.t; r_lo = pack_rg(a); } }
            else            { sector = su; t_hi = b.t; r_hi = pack_rg(b); }
        }
        peek();
    }
    SPHRT_HD void advance(const GridView& g, const RayConsts& c)
    {
        const double t_now = head.t;
        const bool from_lo = t_lo <= t_hi;
        const int ns = geo_to_sector(g, head.geo);
        const int nsec = n_sectors(g);
        if (from_lo) t_lo = SPHRT_INF; else t_hi = SPHRT_INF;
        const int down = sector == 0 ? nsec - 1 : sector - 1;
        const int up = sector == nsec - 1 ? 0 : sector + 1;
        if (from_lo && ns == down && ns != sector) {
            // crossed the lower plane: it becomes the upper plane of the new sector
            sector = ns; t_hi = SPHRT_INF;
            Event a = eval(g, c, lower_plane(g, sector), t_now);
            t_lo = a.t; r_lo = pack_rg(a);
        } else if (!from_lo && ns == up && ns != sector) {
            sector = ns; t_lo = SPHRT_INF;
            Event b = eval(g, c, upper_plane(g, sector), t_now);
            t_hi = b.t; r_hi = pack_rg(b);
        }
        peek();
    }
};

// ---------------------------------------------------------------------------------------
// Quirk 1 of the reference: candidates behind the start (t < 0) are kept in
// the sorted list with the start voxel, so a ray that starts *inside* the grid integrates
// the start voxel from the most negative finite candidate on.  Brute force over every
// boundary; only taken for rays whose start voxel is valid.
SPHRT_HD double most_negative_candidate(const GridView& g, const RayConsts& c)
{
    double tmin = 0.0;
    for (int i = 0; i <= g.nr; i++) {
        const double a = radial_event(g, c, i, false).t, b = radial_event(g, c, i, true).t;
        if (a < tmin && a > -SPHRT_INF) tmin = a;
        if (b < tmin && b > -SPHRT_INF) tmin = b;
    }
    for (int i = 0; i <= g.ne; i++) {
        Event e1, e2;
        cone_events(g, c, i, e1, e2);
        const double a = e1.t, b = e2.t;
        if (a < tmin && a > -SPHRT_INF) tmin = a;
        if (b < tmin && b > -SPHRT_INF) tmin = b;
    }
    for (int i = 0; i <= g.na; i++) {
        const double a = plane_event(g, c, i).t;
        if (a < tmin && a > -SPHRT_INF) tmin = a;
    }
    return tmin;
}

// start voxel (find_starts + cart2sph, :555-578, 605-644).  bins: geometric bins,
// state: reference region values.
SPHRT_HD void start_voxel(const GridView& g, const double x[3], int bins[3], int state[3],
                          double* sph = nullptr)
{
    const double xy = x[0] * x[0] + x[1] * x[1];
    const double r = sqrt(xy + x[2] * x[2]);
    const double e = atan2(sqrt(xy), x[2]);
    const double a = atan2(x[1], x[0]);
    bins[0] = locate(g.r_b(), g.nr, r, &state[0]);
    bins[1] = locate(g.e_b(), g.ne, e, &state[1]);
    bins[2] = locate(g.a_b(), g.na, a, &state[2]);
    if (sph) { sph[0] = r; sph[1] = e; sph[2] = a; }
}

// Start voxel of a ray with the classification the kernels need: geometric bins, reference
// states, spherical coordinates, and whether the point sits on a cone or half-plane (the
// reference then keeps or loses the t ~ 0 crossing by rounding).
struct StartCell {
    int bins[3], st[3];
    double sph[3];
    bool on_boundary;
};

// true if v lies within `tol` of one of the boundaries of its bin
SPHRT_HD bool near_boundary(const double* b, int n, int bin, double v, double tol)
{
    const int lo = bin < 0 ? 0 : (bin > n ? n : bin);
    const int hi = bin + 1 > n ? n : (bin + 1 < 0 ? 0 : bin + 1);
    return fabs(v - b[lo]) < tol || fabs(v - b[hi]) < tol;
}

SPHRT_HD StartCell start_cell(const GridView& g, const double x[3])
{
    StartCell s;
    start_voxel(g, x, s.bins, s.st, s.sph);
    s.on_boundary = near_boundary(g.e_b(), g.ne, s.bins[1], s.sph[1], 1e-9) ||
                    near_boundary(g.a_b(), g.na, s.bins[2], s.sph[2], 1e-9);
    return s;
}

// ---------------------------------------------------------------------------------------
// Walk one ray.  sink.segment(r, e, a, len) is called for every interval between two
// consecutive events whose voxel is valid and whose length is finite and > 0
// (trace_indices :150-170), in order of increasing distance.
//
// kList selects the elevation event source and must equal
// ray_needs_cone_list(c, g.c2_sorted(), g.n_c2) for the ray; `c` comes from ray_setup().
template <bool kList, class Sink>
SPHRT_HD void walk_ray(const GridView& g, const RayConsts& c, const double x[3],
                       const ConeList& scratch, Sink& sink, const StartCell* pre = nullptr)
{
    int bins[3], st[3];
    if (pre != nullptr) { for (int k = 0; k < 3; k++) { bins[k] = pre->bins[k]; st[k] = pre->st[k]; } }
    else start_voxel(g, x, bins, st);

    RadialStream rs;
    rs.init(g, c);
    // A ray that starts outside the outermost shell and never reaches it keeps r = -1
    // for good: every length is zeroed (:168).
    if (st[0] < 0 && rs.head.t == SPHRT_INF) return;

    ElevationStream es;
    if (kList) es.init_list(g, c, scratch);
    else es.init_tracker(g, c, bins[1]);
    AzimuthStream as;
    as.init(g, c, bins[2]);

    double t_prev = 0.0;
    if (st[0] >= 0 && st[1] >= 0 && st[2] >= 0) t_prev = most_negative_candidate(g, c);

    // every boundary root is consumed at most once (plus re-evaluated cone roots on the way
    // back); the cap only exists so that a numerical surprise can never hang a GPU thread
    int budget = 4 * (g.nr + g.ne + g.na) + 64;
    for (; budget > 0; --budget) {
        const double tr = rs.head.t, te = es.head.t, ta = as.head.t;
        const double tn = fmin(tr, fmin(te, ta));
        if (st[0] >= 0 && st[1] >= 0 && st[2] >= 0) {
            const double len = tn - t_prev;
            if (len > 0 && len < SPHRT_INF) sink.segment(st[0], st[1], st[2], len);
        }
        if (tn == SPHRT_INF) break;
        if (tr == tn) {
            if (rs.head.region != -2) st[0] = rs.head.region;
            rs.advance(g, c);
            // left through the outermost shell (or the run is exhausted outside): nothing
            // further along the ray can have a valid radial index
            if (st[0] < 0 && rs.head.t == SPHRT_INF) break;
        } else if (te == tn) {
            if (es.head.region != -2) st[1] = es.head.region;
            es.template advance<kList>(g, c, scratch);
        } else {
            if (as.head.region != -2) st[2] = as.head.region;
            as.advance(g, c);
        }
        if (tn > t_prev) t_prev = tn;
    }
}

// ---------------------------------------------------------------------------------------
// Warp-cooperative variant of walk_ray<false>.
//
// In walk_ray every loop iteration consumes one event and immediately regenerates the
// stream it came from; the 32 rays of a warp sit in different streams at any moment, so
// the three (expensive, float64) generators execute serially with a third of the lanes
// each.  Here every stream produces AHEAD into a small per-thread ring buffer, and a
// buffer is refilled for the whole warp at once when any lane runs dry: the generators
// run converged, and the consumer loop between refills is a few cheap instructions.
// The event sequence of each ray -- hence every result -- is identical to walk_ray's.
//
// Queue storage is caller-provided: element k of this thread at [k * stride].
template <int Q, int STRIDE>
struct EventQueue {
    double* t; int* r;
    int head, count, done;          // done: generator exhausted, nothing more will be pushed
    static constexpr int stride = STRIDE;
    SPHRT_HD void reset(double* tt, int* rr) { t = tt; r = rr; head = 0; count = 0; done = 0; }
    SPHRT_HD void push(double tt, int rr)
    {
        const int k = (head + count) & (Q - 1);
        t[k * stride] = tt; r[k * stride] = rr; count++;
    }
    SPHRT_HD void pop() { head = (head + 1) & (Q - 1); count--; }
    SPHRT_HD bool ready() const { return count > 0 || done; }
    SPHRT_HD double front_t() const { return count > 0 ? t[head * stride] : SPHRT_INF; }
    SPHRT_HD int front_r() const { return count > 0 ? r[head * stride] : -2; }
};

// `any(pred)` must be a warp-wide vote in which ALL lanes of the warp take part (lanes
// without a ray pass active = false); on the host it is the identity.
template <int Q, int STRIDE, class Sink, class Any>
SPHRT_HD void walk_ray_queued(const GridView& g, const RayConsts& c, const double x[3], bool active,
                              double* qt, int* qr, Sink& sink, Any any)
{
    constexpr int qstride = STRIDE;
    int bins[3] = {0, 0, 0}, st[3] = {-1, -1, -1};
    RadialStream rs; ElevationStream es; AzimuthStream as;
    EventQueue<Q, STRIDE> qR, qE, qA;
    qR.reset(qt, qr);
    qE.reset(qt + Q * qstride, qr + Q * qstride);
    qA.reset(qt + 2 * Q * qstride, qr + 2 * Q * qstride);
    const ConeList none{nullptr, nullptr, 0, 0};
    double t_prev = 0.0;

    if (active) {
        start_voxel(g, x, bins, st);
        rs.init(g, c);
        if (st[0] < 0 && rs.head.t == SPHRT_INF) active = false;      // never enters the grid
    }
    if (active) {
        es.init_tracker(g, c, bins[1]);
        as.init(g, c, bins[2]);
        if (st[0] >= 0 && st[1] >= 0 && st[2] >= 0) t_prev = most_negative_candidate(g, c);
    } else {
        qR.done = qE.done = qA.done = 1;
        rs.head.t = es.head.t = as.head.t = SPHRT_INF;
    }
    int budget = 4 * (g.nr + g.ne + g.na) + 64;

    while (any(active)) {
        // ---- refill: whole-warp, one stream at a time ------------------------------------
        if (any(active && !qR.ready())) {
            for (int k = 0; k < Q; k++) {
                const bool need = active && !qR.done && qR.count < Q;
                if (!any(need)) break;
                if (need) {
                    if (rs.head.t == SPHRT_INF) qR.done = 1;
                    else { qR.push(rs.head.t, rs.head.region); rs.advance(g, c); }
                }
            }
        }
        if (any(active && !qE.ready())) {
            for (int k = 0; k < Q; k++) {
                const bool need = active && !qE.done && qE.count < Q;
                if (!any(need)) break;
                if (need) {
                    if (es.head.t == SPHRT_INF) qE.done = 1;
                    else { qE.push(es.head.t, es.head.region); es.template advance<false>(g, c, none); }
                }
            }
        }
        if (any(active && !qA.ready())) {
            for (int k = 0; k < Q; k++) {
                const bool need = active && !qA.done && qA.count < Q;
                if (!any(need)) break;
                if (need) {
                    if (as.head.t == SPHRT_INF) qA.done = 1;
                    else { qA.push(as.head.t, as.head.region); as.advance(g, c); }
                }
            }
        }
        // ---- consume in lockstep until some lane needs a refill ---------------------------
        for (;;) {
            const bool can = active && qR.ready() && qE.ready() && qA.ready();
            if (any(active && !can)) break;
            if (!any(active)) break;
            if (can) {
                const double tr = qR.front_t(), te = qE.front_t(), ta = qA.front_t();
                const double tn = fmin(tr, fmin(te, ta));
                if (st[0] >= 0 && st[1] >= 0 && st[2] >= 0) {
                    const double len = tn - t_prev;
                    if (len > 0 && len < SPHRT_INF) sink.segment(st[0], st[1], st[2], len);
                }
                if (tn == SPHRT_INF || --budget <= 0) active = false;
                else if (tr == tn) {
                    const int reg = qR.front_r();
                    if (reg != -2) st[0] = reg;
                    qR.pop();
                    // left through the outermost shell: nothing further can be valid
                    if (st[0] < 0 && qR.count == 0 && (qR.done || rs.head.t == SPHRT_INF)) active = false;
                } else if (te == tn) {
                    const int reg = qE.front_r();
                    if (reg != -2) st[1] = reg;
                    qE.pop();
                } else {
                    const int reg = qA.front_r();
                    if (reg != -2) st[2] = reg;
                    qA.pop();
                }
                if (tn > t_prev) t_prev = tn;
            }
        }
    }
}

}  // namespace sphrt
