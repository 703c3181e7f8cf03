// TEST-ONLY host build of the per-ray walker (sph_raytracer_b200/csrc/sphrt_walker.cuh).
// Lets the crossing-merge algorithm be checked against the oracle and the golden
// fixtures on a machine without a GPU.  The product never loads this library: the
// shipped path is the CUDA extension (libsphrt.so) and nothing else.
//
// Build (tests/test_walker_host.py does it): g++ -O2 -ffp-contract=off -shared -fPIC
#include "../sph_raytracer_b200/csrc/sphrt_fast.cuh"
#include <cstdint>
#include <vector>

using namespace sphrt;

namespace {
struct CanonicalSink {
    int32_t* vox; double* len; int64_t max_seg; int64_t count = 0;
    int pr = -9, pe = -9, pa = -9;
    void segment(int r, int e, int a, double l)
    {
        if (count > 0 && r == pr && e == pe && a == pa) {
            if (count <= max_seg) len[count - 1] += l;
            return;
        }
        if (count < max_seg) {
            vox[3 * count] = r; vox[3 * count + 1] = e; vox[3 * count + 2] = a; len[count] = l;
        }
        pr = r; pe = e; pa = a; count++;
    }
};
}  // namespace

static int g_queued = 0;
extern "C" void hw_set_queued(int q) { g_queued = q; }

extern "C" void hw_trace_segments(int nr, int ne, int na, const double* r_b, const double* e_b,
                                  const double* cos_e, const double* a_b, const double* cos_a,
                                  const double* sin_a, const double* c2_sorted, int n_c2,
                                  int64_t n, const double* xs,
                                  const double* dirs, int64_t max_seg, int32_t* vox, double* len,
                                  int32_t* count)
{
    std::vector<double> tab;
    auto app = [&](const double* p, int n) { tab.insert(tab.end(), p, p + n); };
    app(r_b, nr + 1); app(e_b, ne + 1); app(cos_e, ne + 1); app(a_b, na + 1); app(cos_a, na + 1);
    app(sin_a, na + 1); app(c2_sorted, n_c2);
    GridView g{nr, ne, na, 0, n_c2, tab.data()};
    g.wrap = (-a_b[0] == a_b[na]) && (a_b[na] == SPHRT_PI);
    std::vector<double> lt(2 * (ne + 1)); std::vector<int> lr(2 * (ne + 1));
    ConeList scratch{lt.data(), lr.data(), 1, 2 * (ne + 1)};
    for (int64_t i = 0; i < n; i++) {
        CanonicalSink s{vox + 3 * i * max_seg, len + i * max_seg, max_seg};
        RayConsts c;
        ray_setup(c, xs + 3 * i, dirs + 3 * i);
        if (ray_needs_cone_list(c, g.c2_sorted(), g.n_c2)) walk_ray<true>(g, c, xs + 3 * i, scratch, s);
        else if (!g_queued) walk_ray<false>(g, c, xs + 3 * i, scratch, s);
        else {
            double qt[3 * 8]; int qr[3 * 8];
            if (g_queued == 1) walk_ray_queued<8, 1>(g, c, xs + 3 * i, true, qt, qr, s, [](bool p) { return p; });
            else walk_ray_fast<8, 1>(g, c, xs + 3 * i, true, qt, qr, s, [](bool p) { return p; });
        }
        count[i] = (int32_t)s.count;
    }
}

// number of rays that take the sorted-cone-list path (pass B) -- diagnostics
extern "C" int64_t hw_count_listed(int ne, const double* e_b, const double* cos_e,
                                   const double* c2_sorted, int n_c2, int64_t n, const double* xs,
                                   const double* dirs)
{
    (void)ne; (void)e_b; (void)cos_e;
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++) {
        RayConsts c;
        ray_setup(c, xs + 3 * i, dirs + 3 * i);
        k += ray_needs_cone_list(c, c2_sorted, n_c2);
    }
    return k;
}
