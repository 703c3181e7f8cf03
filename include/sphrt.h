/*
 * sphrt.h -- C ABI of the B200 line-integration kernels (libsphrt.so).
 *
 * Drop-in boundary for the forward / adjoint hot path of Evidlo/sph_raytracer
 * (sph_raytracer/raytracer.py).  The reference has no native code and no FFI; its
 * boundary is the Python class `Operator` (raytracer.py:647-755).  Each entry point
 * below replaces the reference code cited next to it and is what the Python mirror
 * (sph_raytracer_b200/raytracer.py) binds through ctypes -- see INTEGRATION.md for the
 * stub a maintainer would add to the reference itself.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host
 *   - all arrays are dense, C-contiguous; volumes are (n_vol, nr, ne, na), azimuth fastest
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - the library is stateless: it never allocates, frees or synchronises
 *   - return value: 0 on success, a negative SPHRT_E* code otherwise
 *     (sphrt_error_string() describes it; CUDA launch errors are SPHRT_ECUDA and the
 *     cudaError_t is available through sphrt_last_cuda_error())
 */
#ifndef SPHRT_H
#define SPHRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHRT_VERSION 7

enum {
    SPHRT_OK = 0,
    SPHRT_EINVAL = -1,     /* null pointer, negative size, unknown dtype ...            */
    SPHRT_ETOOBIG = -2,    /* a grid axis exceeds SPHRT_MAX_AXIS                        */
    SPHRT_ESCRATCH = -3,   /* scratch buffer smaller than sphrt_scratch_bytes()         */
    SPHRT_ECUDA = -4       /* kernel launch failed; see sphrt_last_cuda_error()         */
};

#define SPHRT_MAX_AXIS 32000   /* voxels per axis (region indices are packed in 16 bits) */

enum { SPHRT_F32 = 0, SPHRT_F64 = 1 };

/* Spherical grid: what Operator reads from SphericalGrid (geometry.py:107-183:
 * r_b, e_b, a_b, shape) plus trigonometric tables the caller computes ON THE HOST
 * with the same library the reference uses (torch.cos / torch.sin), so that no libm
 * difference can move a crossing (raytracer.py:373-375, 452, 505-506). */
typedef struct sphrt_grid {
    int32_t nr, ne, na;        /* voxel counts; the boundary vectors hold n+1 values    */
    int32_t n_c2;              /* entries of c2_sorted                                   */
    const double *r_b;         /* (nr+1) shell radii, ascending                          */
    const double *e_b;         /* (ne+1) cone angles from +Z, ascending                  */
    const double *cos_e;       /* (ne+1) cos(e_b)                                        */
    const double *a_b;         /* (na+1) half-plane angles from +X, ascending            */
    const double *cos_a;       /* (na+1) cos(a_b)                                        */
    const double *sin_a;       /* (na+1) sin(a_b)                                        */
    const double *c2_sorted;   /* ascending cos(e_b)^2 of the cones with |e-pi/2|>=1e-5  */
} sphrt_grid;

/* Lines of sight: what Operator reads from ViewGeom (geometry.py:284-288, 389-395,
 * 494-513): ray i starts at starts[i / rays_per_start] and points along
 * dirs[i / rays_per_dir] (float64 xyz triples).  rays_per_start == n_rays expresses one
 * pinhole shared by all pixels, == H*W one pinhole per observation, == 1 explicit
 * per-ray starts; likewise for directions (parallel beams share one direction).
 * Directions are used exactly as trace_indices receives them (raytracer.py:48-80). */
typedef struct sphrt_rays {
    int64_t n_rays;
    const double *starts;
    int64_t rays_per_start;
    const double *dirs;
    int64_t rays_per_dir;
    const int32_t *start_cells;/* optional (may be NULL): for each row of `starts`, 8 int32 --
                                * the reference's find_starts() of that point evaluated ON THE HOST
                                * with the reference's own library (torch: cart2sph + searchsorted,
                                * raytracer.py:555-578, 605-644): geometric bins r,e,a (searchsorted
                                * - 1, a value equal to the last boundary counted into the last bin),
                                * states r,e,a (bin n -> -1), 1 if the point lies within 1e-9 of a
                                * cone or half-plane, 0 (pad).  A start that sits on a boundary to
                                * the last bit is binned by the last bit of atan2, which differs
                                * between libms; with this table the kernels bin it like the
                                * reference does.  NULL: binned in the kernel (CUDA atan2). */
    int64_t walker;            /* which kernel walks the ordinary rays: 0 = library's choice,
                                * 1 = warp-cooperative, 2 = per-thread.  Same results either way;
                                * the caller may time both once and keep the faster one (the Python
                                * Operator does).  SPHRT_WALKER in the environment overrides. */
    /* Optional analytic ray generation for pinhole detectors: instead of
     * `dirs` (24 bytes per ray, 6.3 GB at config 5) the directions are formed inside the kernels
     * from a 96-byte frame per observation and float32 pixel-coordinate tables, with exactly the
     * arithmetic of the reference's detector classes (geometry.py:494-508 rectangular,
     * :573-582 polar), so that they are bit-identical to geom.rays:
     *     d = (lookdir + A * u) + B * v ;  d /= sqrt(fma(d2, d2, fma(d1, d1, d0 * d0)))
     *     gen_kind 1 (ConeRectGeom): A = ta[row], B = tb[col]
     *     gen_kind 2 (ConeCircGeom): A = ta[row] * tb[col], B = ta[row] * tc[col]  (float32 products)
     * Ray i belongs to observation i / (gen_h * gen_w), row (i / gen_w) % gen_h, column i % gen_w;
     * `starts` holds the pinhole of every observation (rays_per_start == gen_h * gen_w).
     * gen_kind 0: `dirs` is used. */
    int32_t gen_kind, gen_h, gen_w, gen_per_obs_tables;
    const double *gen_frames;  /* (n_obs, 12): pos, lookdir, u = lookdir x updir, v = updir      */
    const float *gen_ta;       /* (gen_h) rows, or (n_obs, gen_h) if gen_per_obs_tables          */
    const float *gen_tb;       /* (gen_w) columns, or (n_obs, gen_w)                             */
    const float *gen_tc;       /* polar detectors only                                           */
    /* Optional plan cache (sphrt_forward / sphrt_adjoint / sphrt_normal, warp-cooperative walker
     * only): the per-ray records the planning kernel writes for the walking kernel -- direction,
     * radial constants, candidate ranges, 80 bytes for every ray that reaches the grid -- and its
     * two bits per ray are the same in every call on the same rays.  With a caller-owned buffer of
     * sphrt_plan_cache_bytes() bytes the first call (plan_cache_state 1) keeps them and the
     * following ones (plan_cache_state 2) skip the planning kernel (~5 % of a pass).  NULL / 0:
     * planned in every call, in the scratch buffer.  The buffer belongs to one (grid, rays) pair
     * and to one value of SPHRT_WARP_CHUNK. */
    void *plan_cache;
    int64_t plan_cache_bytes;
    int64_t plan_cache_state;
} sphrt_rays;

/* Bytes of a plan cache for these rays (0: this grid is not served by the kernels that use one). */
int64_t sphrt_plan_cache_bytes(const sphrt_grid *grid, int64_t n_rays);

/* Bytes of caller-owned scratch the tracing entry points need for this grid and ray count:
 * two bits per ray (the rays that run in the equatorial plane or pass close to the origin
 * and are walked from a sorted cone-root list in a later pass; the rays the warp-cooperative
 * walker hands to the per-thread walker) plus, only for grids with more than ~500 cones,
 * global-memory space for those lists. */
int64_t sphrt_scratch_bytes(const sphrt_grid *grid, int64_t n_rays);

/* Operator.__call__ (raytracer.py:692-713) fused with trace_indices (:48-230):
 *   rays_per_vol == 0 : out[v, i] = sum_seg density[v, vox(seg)] * len(seg)   v < n_vol
 *                       (static volume, or n_vol channels sharing the geometry)
 *   rays_per_vol  > 0 : out[i]    = sum_seg density[(i / rays_per_vol) % n_vol, vox] * len
 *                       (dynamic volume: observation o reads time-slice o, :705-706)
 * density / out have element type `dtype`; lengths and the accumulation are float64. */
int sphrt_forward(const sphrt_grid *grid, const sphrt_rays *rays,
                  const void *density, int dtype, int64_t n_vol, int64_t rays_per_vol,
                  void *out, void *scratch, int64_t scratch_bytes, void *stream);

/* Adjoint: autograd backward of __call__ and Operator.T (raytracer.py:715-748):
 *   grad[v, vox(seg)] += y[v, i] * len(seg)        (rays_per_vol == 0)
 *   grad[(i / rays_per_vol) % n_vol, vox] += y[i] * len(seg)   (rays_per_vol > 0)
 * `grad` is accumulated into (the caller zero-fills it), element type `dtype`. */
int sphrt_adjoint(const sphrt_grid *grid, const sphrt_rays *rays,
                  const void *y, int dtype, int64_t n_vol, int64_t rays_per_vol,
                  void *grad, void *scratch, int64_t scratch_bytes, void *stream);

/* Per-ray number of canonical segments (non-zero length, consecutive identical voxels
 * merged) = what (Operator.lens > 0) holds after merging; `total` (one uint64, may be
 * NULL) accumulates their sum.  Used for the ray-segments/s metric. */
int sphrt_count_segments(const sphrt_grid *grid, const sphrt_rays *rays,
                         int32_t *count, unsigned long long *total,
                         void *scratch, int64_t scratch_bytes, void *stream);

/* Debug / parity export of Operator.regs and Operator.lens (raytracer.py:675-679) in
 * canonical form: vox (n_rays, max_seg, 3) int32 r,e,a; seg_len (n_rays, max_seg) f64;
 * count (n_rays) true segment count (segments beyond max_seg are dropped). */
int sphrt_trace_segments(const sphrt_grid *grid, const sphrt_rays *rays, int64_t max_seg,
                         int32_t *vox, double *seg_len, int32_t *count,
                         void *scratch, int64_t scratch_bytes, void *stream);

/* Fused normal-equations step (extension for retrieval loops; the reference spends one forward and
 * one adjoint trace per iteration on SquareLoss + autograd, loss.py:86-90, retrieval.py:99-116):
 *     out  = A x                                  (n_rays)
 *     grad += A^T (scale * mask * (A x - b))      (n_vol, nr, ne, na); the caller zero-fills grad
 * in ONE walk of the rays.  b and mask are (n_rays) arrays of the volume's dtype and may be NULL
 * (b = 0, mask = 1).  One volume per ray: n_vol = 1, or rays_per_vol > 0 (time-indexed volume, as in
 * sphrt_forward).  Returns SPHRT_EINVAL for grids the warp-cooperative walker does not serve
 * (sphrt_walker_kind); the caller then runs sphrt_forward and sphrt_adjoint. */
int sphrt_normal(const sphrt_grid *grid, const sphrt_rays *rays, const void *density, int dtype,
                 int64_t n_vol, int64_t rays_per_vol, void *out, const void *b, const void *mask,
                 double scale, void *grad, void *scratch, int64_t scratch_bytes, void *stream);

/* Stored operator (optional).  The reference precomputes and stores its operator
 * (Operator.regs / .lens, raytracer.py:675-679: 32 bytes for each of K padded slots per ray) and
 * applies it with a gather (:710-713).  For problems that apply one operator many times and have
 * the memory, the canonical segments can be stored here too, 8 bytes per segment (12 with
 * float64 lengths): offsets (n_rays + 1) int64 = exclusive prefix sum of sphrt_count_segments,
 * vox[offsets[i] + k] = flat voxel offset (r*ne + e)*na + a of segment k of ray i, seg_len its
 * length (len_dtype SPHRT_F32 / SPHRT_F64).  sphrt_forward_stored / sphrt_adjoint_stored then
 * apply it exactly like sphrt_forward / sphrt_adjoint (same n_vol / rays_per_vol rules) without
 * tracing; they stream the lists once (HBM-bound). */
int sphrt_store_segments(const sphrt_grid *grid, const sphrt_rays *rays, const int64_t *offsets,
                         int32_t *vox, void *seg_len, int len_dtype,
                         void *scratch, int64_t scratch_bytes, void *stream);
int sphrt_forward_stored(const int64_t *offsets, const int32_t *vox, const void *seg_len, int len_dtype,
                         int64_t n_rays, const void *density, int dtype, int64_t n_vol,
                         int64_t rays_per_vol, int64_t nvox, void *out, void *stream);
int sphrt_adjoint_stored(const int64_t *offsets, const int32_t *vox, const void *seg_len, int len_dtype,
                         int64_t n_rays, const void *y, int dtype, int64_t n_vol,
                         int64_t rays_per_vol, int64_t nvox, void *grad, void *stream);

/* Packed stored operator: 4 bytes per segment, for the observations that fit in device memory
 * when the operator is applied many times (a retrieval loop; the reference stores its operator
 * at 32 bytes per padded slot, raytracer.py:66-72, 675-679, and applies it with a gather,
 * :710-713).  One 32-bit word per canonical segment: bits 31..29 a step code relative to the
 * previous voxel of the ray (0/1 = +-1 azimuth, 2/3 = +-na elevation, 4/5 = +-ne*na radial, in
 * flat voxel offsets (r*ne + e)*na + a; 6 = same voxel), bits 28..0 the length as a float32
 * rounded to 29 bits (bits << 2 is the float; relative error 2.4e-7).  Code 7 is an escape: bits
 * 28..0 are an absolute voxel offset, no length (first voxel of a ray, azimuth seam, holes,
 * diagonal moves); the next word is a code-6 word with the length.  Needs nr*ne*na < 2^29.
 * sphrt_pack_count: words per ray (count, n_rays int32) and their sum (total, may be NULL;
 * caller zero-fills); offsets (n_rays + 1) int64 = exclusive prefix sum of count;
 * sphrt_pack_segments writes the words; sphrt_forward_packed / sphrt_adjoint_packed apply them
 * like sphrt_forward / sphrt_adjoint (same n_vol / rays_per_vol rules, float32 or float64
 * volumes).  A caller may pack a PREFIX of its rays (whole observations) and trace the rest with
 * sphrt_forward / sphrt_adjoint on a sphrt_rays that starts at the first unpacked observation. */
int sphrt_pack_count(const sphrt_grid *grid, const sphrt_rays *rays, int32_t *count,
                     unsigned long long *total, void *scratch, int64_t scratch_bytes, void *stream);
int sphrt_pack_segments(const sphrt_grid *grid, const sphrt_rays *rays, const int64_t *offsets,
                        uint32_t *words, void *scratch, int64_t scratch_bytes, void *stream);
int sphrt_forward_packed(const int64_t *offsets, const uint32_t *words, int64_t n_rays, int ne, int na,
                         const void *density, int dtype, int64_t n_vol, int64_t rays_per_vol,
                         int64_t nvox, void *out, void *stream);
int sphrt_adjoint_packed(const int64_t *offsets, const uint32_t *words, int64_t n_rays, int ne, int na,
                         const void *y, int dtype, int64_t n_vol, int64_t rays_per_vol,
                         int64_t nvox, void *grad, void *stream);

/* Parity export of the warp-cooperative walker (the kernel that serves large grids): for
 * every ray it handles itself, ALL its events in merged order -- the reference's sorted
 * candidate list (raytracer.py:131-170) restricted to the chord inside the outermost shell:
 * vox (n_rays, max_ev, 3) int32 voxel after the event, ev_len (n_rays, max_ev) f64 distance
 * to the next event (0 where the reference zeroes the slot), count (n_rays) number of events.
 * count[i] is left untouched for rays this walker does not handle (the caller presets -1).
 * stats (4 x uint64, may be NULL, caller zero-fills): rays walked cooperatively, rays needing
 * the sorted cone list, rays handed to the per-thread walker, of those: failed verification. */
int sphrt_trace_events(const sphrt_grid *grid, const sphrt_rays *rays, int64_t max_ev,
                       int32_t *vox, double *ev_len, int32_t *count,
                       unsigned long long *stats,
                       void *scratch, int64_t scratch_bytes, void *stream);

/* r_torch / e_torch / a_torch (raytracer.py:248-552) as trace_indices calls them (:83-88): the
 * unsorted candidate table of every ray, K1 = 2(nr+1) + 2(ne+1) + (na+1) entries in the order
 * [first radial roots | second radial roots | first cone roots | second cone roots | planes]:
 * t (n_rays, K1) f64 distance (+inf: none), region (n_rays, K1) the reference's region value
 * (-1 outside, -2 unchanged), geo (n_rays, K1) boundary index minus negative_crossing.
 * Debug / parity export (the Python mirror's r_torch, e_torch, a_torch). */
int sphrt_candidates(const sphrt_grid *grid, const sphrt_rays *rays, double *t, int32_t *region,
                     int32_t *geo, void *stream);

/* Which kernel walks the ordinary rays of this grid in sphrt_forward / sphrt_adjoint:
 * 1 = warp-cooperative walker (csrc/sphrt_coop.cuh; large grids), 0 = per-thread walker.
 * Both produce the same segments; the choice (grid size, overridable with the environment
 * variable SPHRT_WALKER = auto | coop | tracked) only affects speed.  Negative: error code. */
int sphrt_walker_kind(const sphrt_grid *grid, int64_t n_rays);

/* find_starts (raytracer.py:605-644): voxel of n points; out is (3, n) int32 with -1
 * for "outside" exactly like the reference. */
int sphrt_find_starts(const sphrt_grid *grid, int64_t n, const double *xs, int32_t *out,
                      void *stream);

/* The direction of every ray as the kernels see it, (n_rays, 3) float64: `dirs` broadcast, or the
 * analytically generated directions (parity export: they must equal geom.rays bit for bit). */
int sphrt_export_rays(const sphrt_rays *rays, double *dirs_out, void *stream);

int sphrt_version(void);
const char *sphrt_error_string(int code);
int sphrt_last_cuda_error(void);
/* Number of CUDA kernels this library has launched in this process so far (every entry point
 * counts its own launches; benchmarks report the difference over their timed region). */
unsigned long long sphrt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SPHRT_H */
