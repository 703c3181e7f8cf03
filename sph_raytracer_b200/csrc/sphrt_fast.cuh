// sphrt_fast.cuh -- lean, warp-cooperative walker used by the production kernels (pass A).
//
// Same event sequences as walk_ray<false> in sphrt_walker.cuh (the straightforward merge,
// kept as the readable reference and used for the sorted-list rays of pass B), organised
// for instruction economy on the GPU:
//
//   * every stream produces AHEAD into a per-thread ring buffer in shared memory and is
//     refilled for the whole warp at once, so the float64 generators run converged
//     instead of serially with a third of the lanes each;
//   * the consumer keeps the three queue fronts in registers; an exhausted stream parks a
//     +inf sentinel in its queue, so the hot loop has no emptiness tests and one warp vote;
//   * rays that start outside the outermost shell are fast-forwarded to their entry point:
//     the elevation / azimuth trackers start from the bins of the entry point instead of
//     replaying the crossings in empty space (for rays without snapped cones -- the only
//     ones that reach this walker -- the reference's state at entry is the geometric bin);
//   * radial regions are inferred structurally (inward root -> i-1, outward root -> i)
//     unless the ray is tangent to the shell within rounding, where the reference's own
//     dot-product test is evaluated;
//   * cone roots already behind the ray are discarded before their region is evaluated.
//
// Host/device clean like sphrt_walker.cuh; tests/host_walker.cpp runs it on the CPU.
#pragma once
#include "sphrt_walker.cuh"

namespace sphrt {

// ---- radial ------------------------------------------------------------------------------
struct FastRadial {
    int i, phase;        // phase 0: inward run, 1: outward run, 2: exhausted

    SPHRT_HD void init(const GridView& g, const RayConsts& c)
    {
        i = g.nr; phase = (c.dperp2 != c.dperp2) ? 2 : 0;
    }
    // next event with t >= 0; false when exhausted
    SPHRT_HD bool next(const GridView& g, const RayConsts& c, double& t, int& region)
    {
        for (;;) {
            if (phase == 2) return false;
            if (phase == 0 && i < 0) { phase = 1; i = 0; }
            if (phase == 1 && i > g.nr) { phase = 2; return false; }
            const int k = i;
            const double r = g.r_b()[k];
            const double arg = r * r - c.dperp2;
            if (!(arg >= 0)) {                     // shell not reached (sqrt would be NaN)
                if (phase == 0) { phase = 1; i = k + 1; } else i = k + 1;
                continue;
            }
            const double t1c = sqrt(arg);
            const bool out = phase == 1;
            const double tt = out ? c.tc + t1c : c.tc - t1c;
            i = out ? k + 1 : k - 1;
            if (tt < 0) continue;                  // behind the start
            int reg = out ? k : k - 1;
            if (arg < 1e-17 * (1.0 + c.nx2 + r * r)) {
                // tangent within rounding: the sign of d.p is decided by noise; ask the
                // reference formula (r_torch :307-318)
                reg = radial_event(g, c, k, out).geo;
            }
            region = (reg == g.nr) ? -1 : reg;
            t = tt;
            return true;
        }
    }
};

// ---- azimuth ------------------------------------------------------------------------------
struct FastAzimuth {
    int sector, dir;     // current sector; dir +1 / -1 once known, 0 = exhausted
    double t_next;       // pending event
    int geo_next;
    // A start that lies exactly ON a half-plane makes BOTH planes of the start sector report a
    // root ahead (the one at t = 0 does not leave the sector); the later one is parked here.
    double t_alt;
    int geo_alt;

    SPHRT_HD static int n_sectors(const GridView& g) { return g.wrap ? g.na : g.na + 1; }
    SPHRT_HD static int to_sector(const GridView& g, int geo)
    {
        if (g.wrap) return geo < 0 ? geo + g.na : (geo >= g.na ? geo - g.na : geo);
        return (geo < 0 || geo >= g.na) ? g.na : geo;
    }
    SPHRT_HD static int to_region(const GridView& g, int geo)
    {
        if (g.wrap) return geo < 0 ? geo + g.na : (geo >= g.na ? geo - g.na : geo);
        return geo >= g.na ? -1 : geo;             // geo == -1 stays -1
    }
    // plane bounding `s` on the side `d` (wrap: plane na stands for the seam 0/na)
    SPHRT_HD static int plane_of(const GridView& g, int s, int d)
    {
        if (d < 0) return s;
        if (g.wrap) return s + 1;
        return s == g.na ? 0 : s + 1;
    }
    // root of plane i not behind t_now (strictly ahead if `strict`); seam planes of a
    // wrapped grid are one half-plane evaluated twice by the reference: earlier root wins
    SPHRT_HD static void eval(const GridView& g, const RayConsts& c, int i, double t_now, bool strict,
                              double& t, int& geo)
    {
        Event ev = plane_event(g, c, i);
        if (g.wrap && (i == 0 || i == g.na)) {
            const Event tw = plane_event(g, c, g.na - i);
            const bool ev_ok = !behind(ev.t, t_now, strict), tw_ok = !behind(tw.t, t_now, strict);
            if (tw_ok && (!ev_ok || tw.t < ev.t)) ev = tw;
        }
        t = behind(ev.t, t_now, strict) ? SPHRT_INF : ev.t;
        geo = ev.geo;
    }
    SPHRT_HD void init(const GridView& g, const RayConsts& c, int start_bin, double t_now)
    {
        sector = to_sector(g, start_bin);
        double ta, tb; int ga, gb;
        eval(g, c, plane_of(g, sector, -1), t_now, false, ta, ga);
        eval(g, c, plane_of(g, sector, +1), t_now, false, tb, gb);
        if (ta <= tb) { t_next = ta; geo_next = ga; dir = -1; t_alt = tb; geo_alt = gb; }
        else          { t_next = tb; geo_next = gb; dir = +1; t_alt = ta; geo_alt = ga; }
        if (t_next == SPHRT_INF) {
            // No root ahead in either bounding plane.  Usually the ray never leaves this sector;
            // but a start ON a half-plane whose t ~ 0 root came out negative is geometrically in
            // the neighbouring sector already (the reference ignores that root and meets the
            // next plane with the start bin still in place): probe the planes beyond.
            const int nsec = n_sectors(g);
            const int sd = sector == 0 ? nsec - 1 : sector - 1, su = sector == nsec - 1 ? 0 : sector + 1;
            eval(g, c, plane_of(g, sd, -1), t_now, false, ta, ga);
            eval(g, c, plane_of(g, su, +1), t_now, false, tb, gb);
            if (ta <= tb) { t_next = ta; geo_next = ga; dir = -1; sector = sd; }
            else          { t_next = tb; geo_next = gb; dir = +1; sector = su; }
            t_alt = SPHRT_INF;
            if (t_next == SPHRT_INF) dir = 0;
        }
    }
    SPHRT_HD bool next(const GridView& g, const RayConsts& c, double& t, int& region)
    {
        if (dir == 0) return false;
        t = t_next; region = to_region(g, geo_next);
        // move into the neighbouring sector and look at its far plane
        const int nsec = n_sectors(g);
        const int expect = dir > 0 ? (sector == nsec - 1 ? 0 : sector + 1)
                                   : (sector == 0 ? nsec - 1 : sector - 1);
        const int ns = to_sector(g, geo_next);
        if (ns != expect || ns == sector) {
            // no move through this plane: continue with the other bounding plane, if any
            t_next = t_alt; geo_next = geo_alt; t_alt = SPHRT_INF;
            dir = t_next == SPHRT_INF ? 0 : -dir;
            return true;
        }
        t_alt = SPHRT_INF;
        sector = ns;
        eval(g, c, plane_of(g, sector, dir), t, true, t_next, geo_next);
        if (t_next == SPHRT_INF) dir = 0;
        return true;
    }
};

// ---- elevation (bin tracker) ------------------------------------------------------------------
SPHRT_HD ConeSlot fresh_cone_fast(const GridView& g, const RayConsts& c, int i, double t_now, bool strict)
{
    ConeSlot s; s.ta = SPHRT_INF; s.tb = SPHRT_INF; s.ra = 0; s.rb = 0;
    if (i < 0 || i > g.ne) return s;
    const ConeQuad q = cone_quadratic(g, c, i);
    double t1, t2;
    cone_roots(q, fabs(q.delta) < SPHRT_TAU ? 0.0 : q.delta, t1, t2);
    // roots behind the ray (or NaN / inf) need no region
    const bool k1 = t1 < SPHRT_INF && !behind(t1, t_now, strict);
    const bool k2 = t2 < SPHRT_INF && !behind(t2, t_now, strict);
    if (k1) { const Event e = cone_event(g, c, i, t1); s.ta = e.t; s.ra = pack_rg(e); }
    if (k2) { const Event e = cone_event(g, c, i, t2); s.tb = e.t; s.rb = pack_rg(e); }
    if (s.tb < s.ta) { const double tt = s.ta; s.ta = s.tb; s.tb = tt; const int rr = s.ra; s.ra = s.rb; s.rb = rr; }
    return s;
}

struct FastElevation {
    int bin, probed;
    ConeSlot lo, hi;

    SPHRT_HD void recover(const GridView& g, const RayConsts& c, double t_now)
    {
        if (probed || lo.ta != SPHRT_INF || hi.ta != SPHRT_INF) return;
        probed = 1;
        const ConeSlot dn = fresh_cone_fast(g, c, bin - 1, t_now, true);
        const ConeSlot up = fresh_cone_fast(g, c, bin + 2, t_now, true);
        if (dn.ta == SPHRT_INF && up.ta == SPHRT_INF) return;
        if (dn.ta <= up.ta) { hi = lo; lo = dn; bin -= 1; }
        else                { lo = hi; hi = up; bin += 1; }
    }
    SPHRT_HD void init(const GridView& g, const RayConsts& c, int start_bin, double t_now)
    {
        bin = start_bin; probed = 0;
        lo = fresh_cone_fast(g, c, bin, t_now, false);
        hi = fresh_cone_fast(g, c, bin + 1, t_now, false);
        recover(g, c, t_now);
    }
    SPHRT_HD bool next(const GridView& g, const RayConsts& c, double& t, int& region)
    {
        const bool from_lo = lo.ta <= hi.ta;
        const double tt = from_lo ? lo.ta : hi.ta;
        if (tt == SPHRT_INF) return false;
        const int p = from_lo ? lo.ra : hi.ra;
        t = tt; region = unpack_region(p);
        const int nb = unpack_geo(p);
        if (from_lo) {
            if (lo.tb == lo.ta) { lo.ta = SPHRT_INF; lo.tb = SPHRT_INF; }
            else { lo.ta = lo.tb; lo.ra = lo.rb; lo.tb = SPHRT_INF; }
            if (nb == bin - 1) { hi = lo; bin = nb; lo = fresh_cone_fast(g, c, bin, tt, false); probed = 0; }
        } else {
            if (hi.tb == hi.ta) { hi.ta = SPHRT_INF; hi.tb = SPHRT_INF; }
            else { hi.ta = hi.tb; hi.ra = hi.rb; hi.tb = SPHRT_INF; }
            if (nb == bin + 1) { lo = hi; bin = nb; hi = fresh_cone_fast(g, c, bin + 1, tt, false); probed = 0; }
        }
        recover(g, c, tt);
        return true;
    }
};

// ---- per-thread event ring buffers ---------------------------------------------------------------
// Element k of a queue lives at [k * STRIDE] from the thread's base (shared memory,
// [slot][thread] layout: conflict-free).  count == 0 only transiently; an exhausted stream
// leaves a +inf sentinel at the tail that is never popped.
template <int Q, int STRIDE>
struct Ring {
    double* t; int* r;
    int head, count;
    SPHRT_HD void reset(double* tt, int* rr) { t = tt; r = rr; head = 0; count = 0; }
    SPHRT_HD void push(double tt, int rr)
    {
        const int k = (head + count) & (Q - 1);
        t[k * STRIDE] = tt; r[k * STRIDE] = rr; count++;
    }
    SPHRT_HD void pop() { head = (head + 1) & (Q - 1); count--; }
    SPHRT_HD void front(double& tt, int& rr) const { tt = t[head * STRIDE]; rr = r[head * STRIDE]; }
};

// Walk one ray (pass A).  `any(pred)` is a warp-wide vote in which ALL lanes of the warp take
// part (lanes without a ray pass active = false); on the host it is the identity.
template <int Q, int STRIDE, class Sink, class Any>
SPHRT_HD void walk_ray_fast(const GridView& g, const RayConsts& c, const double x[3], bool active,
                            double* qt, int* qr, Sink& sink, Any any, const StartCell* pre = nullptr)
{
    FastRadial rs; FastElevation es; FastAzimuth as;
    Ring<Q, STRIDE> qR, qE, qA;
    qR.reset(qt, qr);
    qE.reset(qt + Q * STRIDE, qr + Q * STRIDE);
    qA.reset(qt + 2 * Q * STRIDE, qr + 2 * Q * STRIDE);
    bool doneR = true, doneE = true, doneA = true;   // generator exhausted (sentinel pushed)
    int sr = -1, se = -1, sa = -1;
    double t_prev = 0.0;
    double tR = SPHRT_INF, tE = SPHRT_INF, tA = SPHRT_INF;
    int rR = -2, rE = -2, rA = -2;

    if (active) {
        int bins[3], st[3];
        double sph[3];
        if (pre != nullptr) {
            for (int k = 0; k < 3; k++) { bins[k] = pre->bins[k]; st[k] = pre->st[k]; sph[k] = pre->sph[k]; }
        } else start_voxel(g, x, bins, st, sph);
        rs.init(g, c);
        double t0; int r0;
        const bool has_r = rs.next(g, c, t0, r0);
        if (st[0] < 0 && !has_r) active = false;              // never enters the grid
        else {
            sr = st[0]; se = st[1]; sa = st[2];
            double t_now = 0.0;
            if (has_r) { qR.push(t0, r0); doneR = false; tR = t0; rR = r0; }
            else { qR.push(SPHRT_INF, -2); }
            // A start that sits ON a cone or half-plane makes the reference lose or keep the
            // t ~ 0 crossing by rounding (e.g. single_vantage.py: pos (5,0,0) lies on the
            // equator cone and on the plane a = 0); replay from t = 0 then.  Likewise when a
            // cone's quadratic is snapped to a linear one (ray_has_snapped_aa).
            const bool on_boundary = near_boundary(g.e_b(), g.ne, bins[1], sph[1], 1e-9) ||
                                     near_boundary(g.a_b(), g.na, bins[2], sph[2], 1e-9);
            if (bins[0] == g.nr && has_r && !on_boundary && !ray_has_snapped_aa(c, g.c2_sorted(), g.n_c2)) {
                // starts outside the outermost shell: jump to the entry point.  Until then
                // r = -1 and every length is zero; the elevation / azimuth state at entry
                // is the bin of the entry point.
                t_now = t0;
                const double p[3] = {c.d0 * t0 + c.x0, c.d1 * t0 + c.x1, c.d2 * t0 + c.x2};
                int pb[3], ps[3];
                start_voxel(g, p, pb, ps);
                bins[1] = pb[1]; bins[2] = pb[2]; se = ps[1]; sa = ps[2];
                t_prev = t0;
            } else if (sr >= 0 && se >= 0 && sa >= 0) {
                t_prev = most_negative_candidate(g, c);        // quirk 1, see sphrt_walker.cuh
            }
            es.init(g, c, bins[1], t_now);
            as.init(g, c, bins[2], t_now);
            doneE = false; doneA = false;
        }
    }
    int budget = 4 * (g.nr + g.ne + g.na) + 64;

    while (any(active)) {
        // ---- refill: whole warp, one stream at a time, each lane tops its queue up ---------
        if (any(active && qR.count == 0)) {
            for (int k = 0; k < Q; k++) {
                const bool need = active && !doneR && qR.count < Q;
                if (!any(need)) break;
                if (need) {
                    double t; int reg;
                    if (rs.next(g, c, t, reg)) qR.push(t, reg);
                    else { qR.push(SPHRT_INF, -2); doneR = true; }
                }
            }
            if (active) {
                qR.front(tR, rR);
                if (sr < 0 && tR == SPHRT_INF) active = false;   // outside for good
            }
        }
        if (any(active && qE.count == 0)) {
            for (int k = 0; k < Q; k++) {
                const bool need = active && !doneE && qE.count < Q;
                if (!any(need)) break;
                if (need) {
                    double t; int reg;
                    if (es.next(g, c, t, reg)) qE.push(t, reg);
                    else { qE.push(SPHRT_INF, -2); doneE = true; }
                }
            }
            if (active) qE.front(tE, rE);
        }
        if (any(active && qA.count == 0)) {
            for (int k = 0; k < Q; k++) {
                const bool need = active && !doneA && qA.count < Q;
                if (!any(need)) break;
                if (need) {
                    double t; int reg;
                    if (as.next(g, c, t, reg)) qA.push(t, reg);
                    else { qA.push(SPHRT_INF, -2); doneA = true; }
                }
            }
            if (active) qA.front(tA, rA);
        }
        // ---- consume in lockstep until a lane runs dry or finishes ------------------------------
        if (!any(active)) break;           // the refill above can retire the last lanes
        for (;;) {
            bool stop = false;         // this lane needs a refill or has just finished
            if (active) {
                const double tn = fmin(tR, fmin(tE, tA));
                if ((sr | se | sa) >= 0) {
                    const double len = tn - t_prev;
                    if (len > 0 && tn < SPHRT_INF) sink.segment(sr, se, sa, len);
                }
                if (tn == SPHRT_INF || --budget <= 0) { active = false; }
                else if (tR == tn) {
                    if (rR != -2) sr = rR;
                    qR.pop();
                    if (qR.count > 0) {
                        qR.front(tR, rR);
                        // left through the outermost shell: nothing further can be valid
                        if (sr < 0 && tR == SPHRT_INF) active = false;
                    }
                } else if (tE == tn) {
                    if (rE != -2) se = rE;
                    qE.pop();
                    if (qE.count > 0) qE.front(tE, rE);
                } else {
                    if (rA != -2) sa = rA;
                    qA.pop();
                    if (qA.count > 0) qA.front(tA, rA);
                }
                if (tn > t_prev) t_prev = tn;
                stop = !active || qR.count == 0 || qE.count == 0 || qA.count == 0;
            }
            if (any(stop)) break;
        }
    }
}

}  // namespace sphrt
