#pragma once
// sphrt_kernels.cuh -- sm_100a kernels and their host launchers (C ABI: sphrt_forward.cu,
// sphrt_adjoint.cu, sphrt_api.cu; include/sphrt.h).  Three translation units include this header
// so that nvcc compiles the forward, adjoint and export kernels in parallel.
//
// One thread walks one ray with the register-resident crossing merge of
// sphrt_walker.cuh; nothing per ray-slot is ever materialised (the reference stores
// 32 B for each of K = 2(Nr+1)+2(Ne+1)+(Na+1)+1 slots per ray, raytracer.py:66-72,230).
//
//   forward  : gather-multiply-sum fused with the walk           (Operator.__call__, :692-713)
//   adjoint  : scatter-add with red.global.add fused with the walk   (Operator.T, :715-748 and
//              the autograd backward of __call__)
//   count / trace : parity and metric support (canonical segment lists)
//
// Every entry point runs two launches over the same rays:
//   pass A  one thread per ray, elevation events from the bin tracker; rays that need the
//           sorted cone-root list (ray_needs_cone_list) are skipped;
//   pass B  a small persistent grid re-scans the rays and walks only those, each thread
//           owning a slice of the caller-provided scratch buffer.
// Boundary tables are staged in shared memory once per block (threads index them
// divergently, which shared memory serves at full rate and constant memory would not).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo ...
// (-fmad=false: the walker's arithmetic must not be contracted, see sphrt_walker.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sphrt.h"
#include <stdlib.h>
#include <string.h>

#include "sphrt_warp.cuh"

using namespace sphrt;

namespace {

#ifndef SPHRT_BLOCK
#define SPHRT_BLOCK 128
#endif
constexpr int kBlock = SPHRT_BLOCK;    // threads per block, both passes
constexpr int kListBlock = 128;         // pass B: max threads per block
constexpr int kMaxSmemBytes = 96 * 1024;   // boundary tables must fit in shared memory
#ifndef SPHRT_QUEUE
#define SPHRT_QUEUE 8
#endif
constexpr int kQueue = SPHRT_QUEUE;    // pass A: events buffered per stream and thread
constexpr int kQueueBytes = 3 * kQueue * kBlock * (int)(sizeof(double) + sizeof(int));

// one slot for the whole library (defined in sphrt_api.cu): the last cudaError_t a launch returned
}  // namespace
extern int sphrt_g_last_cuda_error;
extern unsigned long long sphrt_g_launches;      // kernels this library has launched (sphrt_launch_count)
namespace {

struct KRays {
    int64_t n;
    const double* starts; int64_t rps;
    const double* dirs; int64_t rpd;
    const int32_t* cells;          // host-evaluated find_starts of every start row (may be null)
    // analytic ray generation (sphrt_rays::gen_*): 0 = directions come from `dirs`
    int gen_kind, gen_h, gen_w, gen_per_obs;
    const double* frames; const float* ta; const float* tb; const float* tc;
};

inline KRays make_krays(const sphrt_rays* r)
{
    return KRays{r->n_rays, r->starts, r->rays_per_start, r->dirs, r->rays_per_dir, r->start_cells,
                 r->gen_kind, r->gen_h, r->gen_w, r->gen_per_obs_tables,
                 r->gen_frames, r->gen_ta, r->gen_tb, r->gen_tc};
}

// Start voxel of ray i: from the caller's table (the reference's own binning) or, without one,
// evaluated here.
__device__ __forceinline__ StartCell ray_start_cell(const KRays& kr, int64_t i, const GridView& g,
                                                    const double x[3])
{
    if (kr.cells == nullptr) return start_cell(g, x);
    const int32_t* p = kr.cells + 8 * (i / kr.rps);
    StartCell s;
    s.bins[0] = __ldg(p); s.bins[1] = __ldg(p + 1); s.bins[2] = __ldg(p + 2);
    s.st[0] = __ldg(p + 3); s.st[1] = __ldg(p + 4); s.st[2] = __ldg(p + 5);
    s.on_boundary = __ldg(p + 6) != 0;
    s.sph[0] = s.sph[1] = s.sph[2] = 0.0;          // only used for on_boundary, which is given
    return s;
}

__host__ __device__ inline int64_t table_doubles(const sphrt_grid& kg)
{
    return (int64_t)(kg.nr + 1) + 2 * (int64_t)(kg.ne + 1) + 3 * (int64_t)(kg.na + 1) + kg.n_c2;
}

// Copy the boundary tables into shared memory, packed in GridView order, and build the
// walker's view of the grid.
__device__ __forceinline__ GridView stage_tables(const sphrt_grid& kg, double* sm)
{
    GridView g;
    g.nr = kg.nr; g.ne = kg.ne; g.na = kg.na; g.n_c2 = kg.n_c2;
    g.tab = sm;
    const double* src[7] = {kg.r_b, kg.e_b, kg.cos_e, kg.a_b, kg.cos_a, kg.sin_a, kg.c2_sorted};
    const int cnt[7] = {kg.nr + 1, kg.ne + 1, kg.ne + 1, kg.na + 1, kg.na + 1, kg.na + 1, kg.n_c2};
    double* p = sm;
#pragma unroll
    for (int k = 0; k < 7; k++) {
        for (int i = threadIdx.x; i < cnt[k]; i += blockDim.x) p[i] = __ldg(src[k] + i);
        p += cnt[k];
    }
    __syncthreads();
    // full azimuth range: planes 0 and na are the same half-plane (raytracer.py:528)
    g.wrap = (-g.a_b()[0] == g.a_b()[g.na]) && (g.a_b()[g.na] == SPHRT_PI);
    return g;
}

// Direction of pixel (row, col) of a pinhole detector exactly as the reference's detector classes
// form it (geometry.py:494-508, 573-582): float32 pixel coordinates promoted on multiply, two
// products and two sums in float64 left to right, then the 3-vector normalisation of
// torch.linalg.norm (fused sum of squares).
__device__ __forceinline__ void generate_ray(const KRays& kr, int64_t i, double d[3])
{
    const int64_t hw = (int64_t)kr.gen_h * kr.gen_w;
    const int64_t obs = i / hw;
    const int rem = (int)(i - obs * hw);
    const int row = rem / kr.gen_w, col = rem - row * kr.gen_w;
    const double* f = kr.frames + 12 * obs;
    const int64_t ot = kr.gen_per_obs ? obs : 0;
    const float a_row = __ldg(kr.ta + ot * kr.gen_h + row);
    float A, B;
    if (kr.gen_kind == 2) {
        A = __fmul_rn(a_row, __ldg(kr.tb + ot * kr.gen_w + col));
        B = __fmul_rn(a_row, __ldg(kr.tc + ot * kr.gen_w + col));
    } else {
        A = a_row;
        B = __ldg(kr.tb + ot * kr.gen_w + col);
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
        d[k] = (__ldg(f + 3 + k) + __ldg(f + 6 + k) * (double)A) + __ldg(f + 9 + k) * (double)B;
    const double n = norm3(d[0], d[1], d[2]);
    d[0] /= n; d[1] /= n; d[2] /= n;
}

__device__ __forceinline__ void load_ray(const KRays& kr, int64_t i, double x[3], double d[3])
{
    const double* xs = kr.starts + 3 * (i / kr.rps);
    x[0] = __ldg(xs); x[1] = __ldg(xs + 1); x[2] = __ldg(xs + 2);
    if (kr.gen_kind != 0) { generate_ray(kr, i, d); return; }
    const double* ds = kr.dirs + 3 * (i / kr.rpd);
    d[0] = __ldg(ds); d[1] = __ldg(ds + 1); d[2] = __ldg(ds + 2);
}

__global__ void export_rays_kernel(KRays kr, double* out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kr.n) return;
    double x[3], d[3];
    load_ray(kr, i, x, d);
    out[3 * i] = d[0]; out[3 * i + 1] = d[1]; out[3 * i + 2] = d[2];
}

// ---- per-ray operations ------------------------------------------------------------------

// volume a ray reads / writes: channel block (static) or time-slice of its observation
struct VolumeMap {
    int64_t n_rays, nvox, n_vol, rays_per_vol;
    int ne, na;
    __device__ __forceinline__ int64_t offset(int r, int e, int a) const
    {
        return ((int64_t)r * ne + e) * na + a;
    }
    __device__ __forceinline__ int64_t first_volume(int64_t ray, int ch0) const
    {
        return rays_per_vol > 0 ? (ray / rays_per_vol) % n_vol : ch0;
    }
};

// predicated accumulate / gather of the warp walker's forward sink (see ForwardOp::WSink)
#ifndef SPHRT_WARP_WARPS
#define SPHRT_WARP_WARPS 24
#endif
#ifndef SPHRT_FWD_WARPS
#define SPHRT_FWD_WARPS 20
#endif
#ifndef SPHRT_FWD_OPAQUE
#define SPHRT_FWD_OPAQUE true
#endif
constexpr int kWarpMaxWarps = SPHRT_WARP_WARPS;      // warps per block of the second-generation warp walker

__device__ __forceinline__ void fma_if(float& acc, float a, float b, unsigned p)
{
    asm("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q fma.rn.f32 %0, %1, %2, %0;\n}"
        : "+f"(acc) : "f"(a), "f"(b), "r"(p));
}
__device__ __forceinline__ void fma_if(double& acc, double a, double b, unsigned p)
{
    asm("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q fma.rn.f64 %0, %1, %2, %0;\n}"
        : "+d"(acc) : "d"(a), "d"(b), "r"(p));
}
__device__ __forceinline__ void ldg_if(float& v, const float* base, unsigned idx, unsigned p)
{
    asm volatile("{\n .reg .pred q;\n .reg .u64 a;\n setp.ne.u32 q, %3, 0;\n mad.wide.u32 a, %2, 4, %1;\n"
                 " @q ld.global.nc.f32 %0, [a];\n}" : "+f"(v) : "l"(base), "r"(idx), "r"(p));
}
__device__ __forceinline__ void ldg_if(double& v, const double* base, unsigned idx, unsigned p)
{
    asm volatile("{\n .reg .pred q;\n .reg .u64 a;\n setp.ne.u32 q, %3, 0;\n mad.wide.u32 a, %2, 8, %1;\n"
                 " @q ld.global.nc.f64 %0, [a];\n}" : "+d"(v) : "l"(base), "r"(idx), "r"(p));
}

__device__ __forceinline__ float fmaf_or_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fmaf_or_fma(double a, double b, double c) { return __fma_rn(a, b, c); }

template <typename T, int NCH>
struct ForwardOp {
    const T* density; T* out; VolumeMap m; int ch0;
    // The gather of segment k is issued when the walker reports it and consumed when it
    // reports segment k+1, so the L2 latency overlaps one iteration of the event loop.
    struct Sink {
        const T* vol; int nvox; int ne, na; int nch;
        double acc[NCH], plen;
        T pval[NCH];
        // cooperative walker: kCoopDepth gathers in flight per lane (slot k is consumed when the
        // k-th next segment arrives); one channel only, wider batches take segment()
        static constexpr int kDepth = NCH == 1 ? kCoopDepth : 1;
        T rval[kCoopDepth]; double rlen[kCoopDepth];
        template <int SLOT>
        __device__ __forceinline__ void segment_at(bool ok, int r, int e, int a, double len)
        {
            if constexpr (NCH == 1) {
                acc[0] = fma_rn((double)rval[SLOT], rlen[SLOT], acc[0]);
                const int idx = ok ? (r * ne + e) * na + a : 0;
                rval[SLOT] = __ldg(vol + idx);
                rlen[SLOT] = ok ? len : 0.0;
            } else {
                if (ok) segment(r, e, a, len);
            }
        }
        __device__ __forceinline__ void flush()
        {
#pragma unroll
            for (int c = 0; c < NCH; c++) acc[c] = fma_rn((double)pval[c], plen, acc[c]);
            if constexpr (NCH == 1) {
#pragma unroll
                for (int k = 0; k < kCoopDepth; k++) acc[0] = fma_rn((double)rval[k], rlen[k], acc[0]);
            }
        }
        __device__ __forceinline__ void segment(int r, int e, int a, double len)
        {
            const int idx = (r * ne + e) * na + a;
            flush();
#pragma unroll
            for (int c = 0; c < NCH; c++)
                pval[c] = (NCH == 1 || c < nch) ? __ldg(vol + (size_t)c * nvox + idx) : T(0);
            plen = len;
        }
    };
    __device__ __forceinline__ Sink begin(int64_t ray) const
    {
        Sink s;
        s.vol = density + m.first_volume(ray, ch0) * m.nvox;
        s.nvox = (int)m.nvox; s.ne = m.ne; s.na = m.na;
        s.nch = m.rays_per_vol > 0 ? 1 : (int)min((int64_t)NCH, m.n_vol - ch0);
        s.plen = 0.0;
#pragma unroll
        for (int c = 0; c < NCH; c++) { s.acc[c] = 0.0; s.pval[c] = T(0); }
#pragma unroll
        for (int k = 0; k < kCoopDepth; k++) { s.rval[k] = T(0); s.rlen[k] = 0.0; }
        return s;
    }
    __device__ __forceinline__ void end(int64_t ray, Sink& s) const
    {
        s.flush();
        if (m.rays_per_vol > 0) { out[ray] = (T)s.acc[0]; return; }
#pragma unroll
        for (int c = 0; c < NCH; c++)
            if (c < s.nch) out[(int64_t)(ch0 + c) * m.n_rays + ray] = (T)s.acc[c];
    }
    // second-generation warp walker (sphrt_warp.cuh): segments arrive as (flat voxel offset,
    // length).  One channel: the gather of a segment is issued under its predicate and consumed two
    // segments later under the same predicate (predicated inside inline PTX: written as C++ `if`s the
    // compiler selects between old and new sums instead); float32 volumes accumulate value * length
    // in float32 per lane (ten-odd terms; the reference itself rounds every product to float32 and
    // sums in float32, raytracer.py:710-713), the 32 partial sums are added in float64.  Wider
    // batches accumulate directly in float64.
    struct WSink {
        static constexpr bool kTwoSweeps = false;
        const T* vol; int nvox; int nch;
        double acc[NCH];
        T facc, pv[kWarpDepth], pl[kWarpDepth]; unsigned pk[kWarpDepth];
        template <int SLOT>
        __device__ __forceinline__ void seg(bool ok, unsigned idx, double len)
        {
            if constexpr (NCH == 1) {
                fma_if(facc, pv[SLOT], pl[SLOT], pk[SLOT]);
                pk[SLOT] = ok ? 1u : 0u; pl[SLOT] = (T)len;
                ldg_if(pv[SLOT], vol, idx, pk[SLOT]);
            } else {
                if (ok) {
#pragma unroll
                    for (int c = 0; c < NCH; c++)
                        if (c < nch) acc[c] = fma_rn((double)__ldg(vol + (size_t)c * nvox + idx), len, acc[c]);
                }
            }
        }
    };
    __device__ __forceinline__ WSink wbegin(int64_t ray) const
    {
        WSink s;
        s.vol = density + m.first_volume(ray, ch0) * m.nvox;
        s.nvox = (int)m.nvox;
        s.nch = m.rays_per_vol > 0 ? 1 : (int)min((int64_t)NCH, m.n_vol - ch0);
#pragma unroll
        for (int c = 0; c < NCH; c++) s.acc[c] = 0.0;
        s.facc = T(0);
#pragma unroll
        for (int k = 0; k < kWarpDepth; k++) { s.pk[k] = 0u; s.pl[k] = T(0); s.pv[k] = T(0); }
        return s;
    }
    __device__ __forceinline__ void wend(int64_t ray, WSink& s, int lane, int) const
    {
        if constexpr (NCH == 1) {
#pragma unroll
            for (int k = 0; k < kWarpDepth; k++) fma_if(s.facc, s.pv[k], s.pl[k], s.pk[k]);
            s.acc[0] = (double)s.facc;
        }
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            double a = s.acc[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            s.acc[c] = a;
        }
        if (lane != 0) return;
        if (m.rays_per_vol > 0) { out[ray] = (T)s.acc[0]; return; }
#pragma unroll
        for (int c = 0; c < NCH; c++)
            if (c < s.nch) out[(int64_t)(ch0 + c) * m.n_rays + ray] = (T)s.acc[c];
    }
    // rays that never reach the grid are not walked at all: the result is zero-filled up front
    int prefill(cudaStream_t st) const
    {
        if (m.rays_per_vol > 0) return (int)cudaMemsetAsync(out, 0, sizeof(T) * (size_t)m.n_rays, st);
        const int64_t nch = min((int64_t)NCH, m.n_vol - ch0);
        return (int)cudaMemsetAsync(out + (int64_t)ch0 * m.n_rays, 0, sizeof(T) * (size_t)(nch * m.n_rays), st);
    }
    int32_t* miss_count() const { return nullptr; }
    // warp-cooperative walker: every lane holds a partial sum of the same ray
    static constexpr bool kCoop = true, kEvents = false, kOpaqueBases = SPHRT_FWD_OPAQUE;
    static constexpr int kWarpWarps = SPHRT_FWD_WARPS;
    __device__ __forceinline__ bool void_ray(int64_t) const { return false; }
    __device__ __forceinline__ void coop_end(int64_t ray, Sink& s, int lane, int) const
    {
        s.flush();
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            double a = s.acc[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            s.acc[c] = a;
        }
        if (lane != 0) return;
        if (m.rays_per_vol > 0) { out[ray] = (T)s.acc[0]; return; }
#pragma unroll
        for (int c = 0; c < NCH; c++)
            if (c < s.nch) out[(int64_t)(ch0 + c) * m.n_rays + ray] = (T)s.acc[c];
    }
};

// red.global.add on a pointer the compiler can no longer see the address space of
// (predicated inside the asm: a C++ `if` around it becomes a branch per segment)
__device__ __forceinline__ void red_add_global_if(bool ok, float* p, float v)
{
    asm volatile("" :: "l"(p), "f"(v));          // value and address are formed outside the guard
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %2, 0;\n @q red.global.add.f32 [%0], %1;\n}"
                 :: "l"(p), "f"(v), "r"((unsigned)ok) : "memory");
}
__device__ __forceinline__ void red_add_global_if(bool ok, double* p, double v)
{
    asm volatile("" :: "l"(p), "d"(v));
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %2, 0;\n @q red.global.add.f64 [%0], %1;\n}"
                 :: "l"(p), "d"(v), "r"((unsigned)ok) : "memory");
}

template <typename T, int NCH>
struct AdjointOp {
    const T* y; T* grad; VolumeMap m; int ch0;
    struct Sink {
        static constexpr int kDepth = 1;
        T* vol; int nvox; int ne, na; int nch;
        double w[NCH];
        template <int SLOT>
        __device__ __forceinline__ void segment_at(bool ok, int r, int e, int a, double len)
        {
            if (ok) segment(r, e, a, len);
        }
        __device__ __forceinline__ void segment(int r, int e, int a, double len)
        {
            const int idx = (r * ne + e) * na + a;
#pragma unroll
            for (int c = 0; c < NCH; c++)
                if ((NCH == 1 || c < nch) && w[c] != 0.0) atomicAdd(vol + (size_t)c * nvox + idx, (T)(w[c] * len));
        }
    };
    __device__ __forceinline__ Sink begin(int64_t ray) const
    {
        Sink s;
        s.vol = grad + m.first_volume(ray, ch0) * m.nvox;
        s.nvox = (int)m.nvox; s.ne = m.ne; s.na = m.na;
        if (m.rays_per_vol > 0) {
            s.nch = 1;
            s.w[0] = (double)__ldg(y + ray);
        } else {
            s.nch = (int)min((int64_t)NCH, m.n_vol - ch0);
#pragma unroll
            for (int c = 0; c < NCH; c++)
                s.w[c] = c < s.nch ? (double)__ldg(y + (int64_t)(ch0 + c) * m.n_rays + ray) : 0.0;
        }
        return s;
    }
    struct WSink {
        static constexpr bool kTwoSweeps = false;
        T* vol; int nvox; int nch;
        double w[NCH];
        template <int SLOT>
        __device__ __forceinline__ void seg(bool ok, unsigned idx, double len)
        {
#pragma unroll
            for (int c = 0; c < NCH; c++)
                red_add_global_if(ok && (NCH == 1 || (c < nch && w[c] != 0.0)),
                                  vol + (size_t)c * nvox + idx, (T)(w[c] * len));
        }
    };
    __device__ __forceinline__ WSink wbegin(int64_t ray) const
    {
        WSink s;
        s.vol = grad + m.first_volume(ray, ch0) * m.nvox;
        asm volatile("" : "+l"(s.vol));
        s.nvox = (int)m.nvox;
        if (m.rays_per_vol > 0) {
            s.nch = 1;
            s.w[0] = (double)__ldg(y + ray);
        } else {
            s.nch = (int)min((int64_t)NCH, m.n_vol - ch0);
#pragma unroll
            for (int c = 0; c < NCH; c++)
                s.w[c] = c < s.nch ? (double)__ldg(y + (int64_t)(ch0 + c) * m.n_rays + ray) : 0.0;
        }
        return s;
    }
    __device__ __forceinline__ void wend(int64_t, WSink&, int, int) const {}
    int prefill(cudaStream_t) const { return 0; }
    int32_t* miss_count() const { return nullptr; }
    __device__ __forceinline__ void end(int64_t, Sink&) const {}
    static constexpr bool kCoop = true, kEvents = false, kOpaqueBases = false;
    static constexpr int kWarpWarps = kWarpMaxWarps;
    __device__ __forceinline__ void coop_end(int64_t, Sink&, int, int) const {}
    // a ray whose weights are all zero (masked measurements: loss.py projection_mask) adds nothing
    __device__ __forceinline__ bool void_ray(int64_t ray) const
    {
        if (m.rays_per_vol > 0) return __ldg(y + ray) == T(0);
        const int nch = (int)min((int64_t)NCH, m.n_vol - ch0);
        bool zero = true;
#pragma unroll
        for (int c = 0; c < NCH; c++)
            if (c < nch) zero &= __ldg(y + (int64_t)(ch0 + c) * m.n_rays + ray) == T(0);
        return zero;
    }
};

// Fused normal-equations step for retrieval loops (an extension, not a path of the reference:
// SquareLoss + autograd there cost one forward and one adjoint walk per iteration,
// loss.py:86-90, retrieval.py:99-116): out = A x and grad += A^T (scale * mask * (A x - b)) in ONE
// walk of the warp walker.  The weight of a ray's back-projection depends only on that ray's own
// integral, and after the forward sweep its merged lists are still in shared memory: the second
// sweep repeats the segment loop only (~15 % of a walk).  One volume per ray (static, or the
// time-slice of its observation).
template <typename T>
struct NormalOp {
    const T* density; T* out; const T* b; const T* mask; double scale; T* grad; VolumeMap m;
    // second sweep: the adjoint sink with the weight just computed
    struct WSink2 {
        static constexpr bool kTwoSweeps = false;
        bool active; T* vol; double w;
        template <int SLOT>
        __device__ __forceinline__ void seg(bool ok, unsigned idx, double len)
        {
            red_add_global_if(ok, vol + idx, (T)(w * len));
        }
    };
    struct WSink {
        static constexpr bool kTwoSweeps = true;
        T* out; const T* b; const T* mask; double scale; T* grad; int64_t ray, vol_index, nvox;
        const T* vol;
        T facc, pv[kWarpDepth], pl[kWarpDepth]; unsigned pk[kWarpDepth];
        template <int SLOT>
        __device__ __forceinline__ void seg(bool ok, unsigned idx, double len)
        {
            fma_if(facc, pv[SLOT], pl[SLOT], pk[SLOT]);
            pk[SLOT] = ok ? 1u : 0u; pl[SLOT] = (T)len;
            ldg_if(pv[SLOT], vol, idx, pk[SLOT]);
        }
        __device__ __forceinline__ WSink2 second(int lane)
        {
#pragma unroll
            for (int k = 0; k < kWarpDepth; k++) fma_if(facc, pv[k], pl[k], pk[k]);
            double a = (double)facc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            const T y = (T)a;                          // the integral as the forward pass stores it
            if (lane == 0) out[ray] = y;
            double w = scale * ((double)y - (b != nullptr ? (double)__ldg(b + ray) : 0.0));
            if (mask != nullptr) w *= (double)__ldg(mask + ray);
            return WSink2{w != 0.0, grad + vol_index * nvox, w};
        }
    };
    __device__ __forceinline__ WSink wbegin(int64_t ray) const
    {
        WSink s;
        s.out = out; s.b = b; s.mask = mask; s.scale = scale; s.grad = grad;
        s.ray = ray; s.vol_index = m.first_volume(ray, 0); s.nvox = m.nvox;
        s.vol = density + s.vol_index * m.nvox;
        s.facc = T(0);
#pragma unroll
        for (int k = 0; k < kWarpDepth; k++) { s.pk[k] = 0u; s.pl[k] = T(0); s.pv[k] = T(0); }
        return s;
    }
    __device__ __forceinline__ void wend(int64_t, WSink&, int, int) const {}
    // rays that never reach the grid: integral 0 (their weight scatters nowhere)
    int prefill(cudaStream_t st) const { return (int)cudaMemsetAsync(out, 0, sizeof(T) * (size_t)m.n_rays, st); }
    int32_t* miss_count() const { return nullptr; }
    static constexpr bool kCoop = true, kEvents = false, kOpaqueBases = true;
    static constexpr int kWarpWarps = SPHRT_FWD_WARPS;
    __device__ __forceinline__ bool void_ray(int64_t) const { return false; }
    // (the per-thread fallback passes run as a forward and an adjoint launch: launch_normal)
    struct Sink { static constexpr int kDepth = 1; };
};

// adjoint of the fused step for the rays the fallback passes walk: the weight comes from the
// integral the forward fallback launch has just stored
template <typename T>
struct NormalAdjointOp {
    const T* out; const T* b; const T* mask; double scale; T* grad; VolumeMap m;
    using Sink = typename AdjointOp<T, 1>::Sink;
    __device__ __forceinline__ double weight(int64_t ray) const
    {
        double w = scale * ((double)__ldg(out + ray) - (b != nullptr ? (double)__ldg(b + ray) : 0.0));
        if (mask != nullptr) w *= (double)__ldg(mask + ray);
        return w;
    }
    __device__ __forceinline__ Sink begin(int64_t ray) const
    {
        Sink s;
        s.vol = grad + m.first_volume(ray, 0) * m.nvox;
        s.nvox = (int)m.nvox; s.ne = m.ne; s.na = m.na; s.nch = 1;
        s.w[0] = weight(ray);
        return s;
    }
    __device__ __forceinline__ void end(int64_t, Sink&) const {}
    static constexpr bool kCoop = false, kEvents = false;
    __device__ __forceinline__ bool void_ray(int64_t ray) const { return weight(ray) == 0.0; }
    __device__ __forceinline__ void coop_end(int64_t, Sink&, int, int) const {}
};

// canonical segment bookkeeping shared by count / trace: consecutive identical voxels merge
struct CountOp {
    int32_t* count; unsigned long long* total;
    struct Sink {
        static constexpr int kDepth = 1;
        int pr, pe, pa; int n;
        __device__ __forceinline__ void segment(int r, int e, int a, double)
        {
            if (n > 0 && r == pr && e == pe && a == pa) return;
            pr = r; pe = e; pa = a; n++;
        }
    };
    __device__ __forceinline__ Sink begin(int64_t) const { Sink s; s.pr = s.pe = s.pa = -9; s.n = 0; return s; }
    __device__ __forceinline__ void end(int64_t ray, Sink& s) const
    {
        count[ray] = s.n;
        if (total != nullptr && s.n > 0) atomicAdd(total, (unsigned long long)s.n);
    }
    // consecutive identical voxels merge: needs the segments in order, per-thread walker only
    static constexpr bool kCoop = false, kEvents = false;
    __device__ __forceinline__ void coop_end(int64_t, Sink&, int, int) const {}
};

struct TraceOp {
    int32_t* vox; double* seg_len; int32_t* count; int64_t max_seg;
    struct Sink {
        static constexpr int kDepth = 1;
        int32_t* vox; double* len; int64_t max_seg;
        int pr, pe, pa; int64_t n;
        __device__ __forceinline__ void segment(int r, int e, int a, double l)
        {
            if (n > 0 && r == pr && e == pe && a == pa) {
                if (n <= max_seg) len[n - 1] += l;
                return;
            }
            if (n < max_seg) { vox[3 * n] = r; vox[3 * n + 1] = e; vox[3 * n + 2] = a; len[n] = l; }
            pr = r; pe = e; pa = a; n++;
        }
    };
    __device__ __forceinline__ Sink begin(int64_t ray) const
    {
        Sink s;
        s.vox = vox + 3 * ray * max_seg; s.len = seg_len + ray * max_seg; s.max_seg = max_seg;
        s.pr = s.pe = s.pa = -9; s.n = 0;
        return s;
    }
    __device__ __forceinline__ void end(int64_t ray, Sink& s) const { count[ray] = (int32_t)s.n; }
    static constexpr bool kCoop = false, kEvents = false;
    __device__ __forceinline__ void coop_end(int64_t, Sink&, int, int) const {}
};

// Stored operator: the canonical segments of every ray written once as a
// compact CSR list -- flat voxel offset (int32) + length (float32 or float64) per segment,
// 8 or 12 bytes against the reference's 32 bytes for each of its K padded slots per ray
// (raytracer.py:66-72, 675-679) -- for problems that apply the same operator many times.
template <typename L>
struct StoreOp {
    const int64_t* offsets; int32_t* vox; L* seg_len; int ne, na;
    struct Sink {
        static constexpr int kDepth = 1;
        int32_t* vox; L* len; int ne, na;
        int cur; double acc; int64_t n;          // pending segment (consecutive identical voxels merge)
        __device__ __forceinline__ void segment(int r, int e, int a, double l)
        {
            const int idx = (r * ne + e) * na + a;
            if (idx == cur) { acc += l; return; }
            if (cur >= 0) { vox[n] = cur; len[n] = (L)acc; n++; }
            cur = idx; acc = l;
        }
    };
    __device__ __forceinline__ Sink begin(int64_t ray) const
    {
        const int64_t o = __ldg(offsets + ray);
        return Sink{vox + o, seg_len + o, ne, na, -1, 0.0, 0};
    }
    __device__ __forceinline__ void end(int64_t, Sink& s) const
    {
        if (s.cur >= 0) { s.vox[s.n] = s.cur; s.len[s.n] = (L)s.acc; }
    }
    static constexpr bool kCoop = false, kEvents = false;
    __device__ __forceinline__ void coop_end(int64_t, Sink&, int, int) const {}
};

// Raw event export of the warp-cooperative walker (parity tests): every event of a ray at its
// rank in the merged order, with the voxel that follows it and the length up to the next
// event (0 where the reference zeroes the slot).  count stays untouched for rays that the
// cooperative walker hands to the per-thread walker.
struct EventTraceOp {
    int32_t* vox; double* ev_len; int32_t* count; int64_t max_ev;
    struct Sink {
        static constexpr int kDepth = 1;
        int32_t* vox; double* len; int64_t max_ev;
        __device__ __forceinline__ void segment(int, int, int, double) {}
        template <int SLOT>
        __device__ __forceinline__ void segment_at(bool, int, int, int, double) {}
        __device__ __forceinline__ void event(int rank, int r, int e, int a, double l)
        {
            if (rank >= max_ev) return;
            vox[3 * rank] = r; vox[3 * rank + 1] = e; vox[3 * rank + 2] = a; len[rank] = l;
        }
    };
    __device__ __forceinline__ Sink begin(int64_t ray) const
    {
        return Sink{vox + 3 * ray * max_ev, ev_len + ray * max_ev, max_ev};
    }
    struct WSink {
        static constexpr bool kTwoSweeps = false;
        int32_t* vox; double* len; int64_t max_ev; int nr;
        template <int SLOT>
        __device__ __forceinline__ void seg(bool, unsigned, double) {}
        // state of the second-generation walker: radial index, e * na + a (negative: outside)
        __device__ __forceinline__ void event(int rank, int r, int sea, int na, double l)
        {
            if (rank >= max_ev) return;
            if (r >= nr) r = -1;
            const int e = sea < 0 ? -1 : sea / na, a = sea < 0 ? -1 : sea % na;
            vox[3 * rank] = r; vox[3 * rank + 1] = e; vox[3 * rank + 2] = a; len[rank] = l;
        }
    };
    int nr_for_wsink;
    __device__ __forceinline__ WSink wbegin(int64_t ray) const
    {
        return WSink{vox + 3 * ray * max_ev, ev_len + ray * max_ev, max_ev, nr_for_wsink};
    }
    __device__ __forceinline__ void wend(int64_t ray, WSink&, int lane, int n_events) const
    {
        if (lane == 0) count[ray] = n_events;
    }
    int prefill(cudaStream_t) const { return 0; }
    int32_t* miss_count() const { return count; }
    __device__ __forceinline__ void end(int64_t ray, Sink&) const { count[ray] = 0; }   // a miss
    static constexpr bool kCoop = true, kEvents = true, kOpaqueBases = false;
    static constexpr int kWarpWarps = kWarpMaxWarps;
    __device__ __forceinline__ bool void_ray(int64_t) const { return false; }
    __device__ __forceinline__ void coop_end(int64_t ray, Sink&, int lane, int n_events) const
    {
        if (lane == 0) count[ray] = n_events;
    }
};

// ---- the two passes ------------------------------------------------------------------------

// pass A: one thread per ray, bin-tracker elevation stream, warp-cooperative refills
struct WarpAny {
    __device__ __forceinline__ bool operator()(bool p) const { return __any_sync(0xffffffffu, p); }
};

#ifndef SPHRT_TRACKED_MIN_BLOCKS
#define SPHRT_TRACKED_MIN_BLOCKS (512 / SPHRT_BLOCK)      /* 16 warps per SM: <= 128 registers */
#endif
template <class Op>
__global__ void __launch_bounds__(kBlock, SPHRT_TRACKED_MIN_BLOCKS)
walk_tracked_kernel(sphrt_grid kg, KRays kr, Op op, int table_doubles_smem, uint32_t* listed_mask)
{
    extern __shared__ double sm_tables[];
    const GridView g = stage_tables(kg, sm_tables);
    // event ring buffers behind the tables: [stream][slot][thread]
    double* qt = sm_tables + table_doubles_smem + threadIdx.x;
    int* qr = reinterpret_cast<int*>(sm_tables + table_doubles_smem + 3 * kQueue * kBlock) + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = i < kr.n;
    double x[3] = {0, 0, 0}, d[3] = {0, 0, 1};
    RayConsts c;
    StartCell cell;
    if (active) {
        load_ray(kr, i, x, d);
        ray_setup(c, x, d);
        active = !ray_needs_cone_list(c, g.c2_sorted(), g.n_c2);     // pass B takes those
        if (active) {
            // ... and the rays that start ON a cone or half-plane: the bin trackers are not exact
            // there (tools/fuzz.py), the sorted-list walkers apply the reference's own rule
            cell = ray_start_cell(kr, i, g, x);
            active = !cell.on_boundary;
        }
    }
    // one bit per ray: the rays pass B has to walk (ray index and warp are 32-aligned)
    const uint32_t listed = __ballot_sync(0xffffffffu, i < kr.n && !active);
    if ((threadIdx.x & 31) == 0 && i < kr.n) listed_mask[i >> 5] = listed;
    const bool mine = active;
    typename Op::Sink sink = op.begin(mine ? i : 0);
    walk_ray_fast<kQueue, kBlock>(g, c, x, active, qt, qr, sink, WarpAny(), &cell);
    if (mine) op.end(i, sink);
}

// Flagged-ray iterator of the fallback passes.  Flagged rays are rare (about 1 %) and scattered,
// so walking them where they sit in their 32-ray mask word leaves most lanes idle for the whole
// (long, serial) walk.  Instead a warp scans mask words in chunks handed out by an atomic
// counter, compacts the flagged ray indices into a small shared-memory queue and calls
// `walk(ray, active)` with 32 of them at a time; the remainder is flushed at the end.
// Flagged rays come in dense runs (whole detector rows near the equatorial plane), so the 32
// words of one hand-out are taken from 32 far-apart, mutually rotated places of the mask: every
// hand-out then holds about the same number of flagged rays.
constexpr int kChunkWords = 32;       // mask words (1024 rays) per hand-out
constexpr int kFlagQueue = 64;        // ray indices buffered per warp

template <class Walk>
__device__ __forceinline__ void for_each_flagged(const uint32_t* mask, int64_t n_words,
                                                 unsigned long long* chunk_counter,
                                                 int64_t* queue /* this warp's kFlagQueue slots */,
                                                 Walk walk)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int head = 0, count = 0;          // ring buffer state, warp-uniform
    const int64_t n_chunks = (n_words + kChunkWords - 1) / kChunkWords;
    for (;;) {
        unsigned long long next = 0;
        if (lane == 0) next = atomicAdd(chunk_counter, 1ull);
        const int64_t chunk = (int64_t)__shfl_sync(0xffffffffu, next, 0);
        if (chunk >= n_chunks) break;
        // lane k reads word k * n_chunks + (chunk + k * R) mod n_chunks: a bijection onto the mask
        const int64_t my_word = lane * n_chunks + (chunk + lane * (int64_t)0x9E3779B1u) % n_chunks;
        const uint32_t mine = my_word < n_words ? __ldg(mask + my_word) : 0u;
        uint32_t nonzero = __ballot_sync(0xffffffffu, mine != 0u);
        for (; nonzero; nonzero &= nonzero - 1) {
            const int src = __ffs(nonzero) - 1;
            const uint32_t m = __shfl_sync(0xffffffffu, mine, src);
            const int64_t word = __shfl_sync(0xffffffffu, my_word, src);
            if ((m >> lane) & 1u)
                queue[(head + count + __popc(m & lt_mask)) & (kFlagQueue - 1)] = (word << 5) + lane;
            count += __popc(m);
            __syncwarp();
            if (count >= 32) {
                const int64_t ray = queue[(head + lane) & (kFlagQueue - 1)];
                head = (head + 32) & (kFlagQueue - 1); count -= 32;
                __syncwarp();
                walk(ray, true);
                __syncwarp();
            }
        }
    }
    if (count > 0) {
        const bool active = lane < count;
        const int64_t ray = active ? queue[(head + lane) & (kFlagQueue - 1)] : 0;
        walk(ray, active);
    }
}

// pass A', masked: the per-thread walker for the rays the cooperative walker handed over
// (starts inside the grid or on a boundary, failed verification).  Persistent warps, one
// 32-ray mask word at a time.
template <class Op>
__global__ void __launch_bounds__(kBlock, SPHRT_TRACKED_MIN_BLOCKS)
walk_tracked_masked_kernel(sphrt_grid kg, KRays kr, Op op, int table_doubles_smem,
                           const uint32_t* mask, int64_t n_words, unsigned long long* chunk_counter,
                           const unsigned* n_work /* may be null: number of flagged rays, 0 = nothing to do */)
{
    if (n_work != nullptr && *n_work == 0) return;
    extern __shared__ double sm_tables[];
    __shared__ int64_t flag_queue[(kBlock / 32) * kFlagQueue];
    const GridView g = stage_tables(kg, sm_tables);
    double* qt = sm_tables + table_doubles_smem + threadIdx.x;
    int* qr = reinterpret_cast<int*>(sm_tables + table_doubles_smem + 3 * kQueue * kBlock) + threadIdx.x;
    for_each_flagged(mask, n_words, chunk_counter, flag_queue + (threadIdx.x >> 5) * kFlagQueue,
                     [&](int64_t i, bool active) {
        double x[3] = {0, 0, 0}, d[3] = {0, 0, 1};
        RayConsts c;
        StartCell cell;
        if (active) { load_ray(kr, i, x, d); ray_setup(c, x, d); cell = ray_start_cell(kr, i, g, x); }
        typename Op::Sink sink = op.begin(active ? i : 0);
        walk_ray_fast<kQueue, kBlock>(g, c, x, active, qt, qr, sink, WarpAny(), &cell);
        if (active) op.end(i, sink);
    });
}

// Warp-cooperative walker (sphrt_coop.cuh): persistent warps, 32 rays per tile; phase 0 plans
// one ray per lane, phase 1 walks the tile's rays one after the other with all 32 lanes.
#ifndef SPHRT_COOP_WARPS
#define SPHRT_COOP_WARPS 16
#endif
constexpr int kCoopMaxWarps = SPHRT_COOP_WARPS;
constexpr int kCoopRecDoubles = 6;
constexpr int64_t kCoopSplitRays = 4 << 20;    // from here on the all-cones rays get their own launch
// SPHRT_COOP_SPLIT_RAYS in the environment overrides the threshold (tests force the two-launch
// path on small cases; 0 = always split)
inline int64_t coop_split_rays()
{
    const char* e = getenv("SPHRT_COOP_SPLIT_RAYS");
    if (e == nullptr || *e == 0) return kCoopSplitRays;
    char* end = nullptr;
    const long long v = strtoll(e, &end, 10);
    return (end == e || v < 0) ? kCoopSplitRays : (int64_t)v;
}

// phase 0 of the cooperative walker for one ray: load, set up, classify, plan; the float64 part of
// the record goes to shared memory ([field][lane]), the integer part comes back in registers
struct Phase0Result { int cls; int ri[6]; int64_t start_index; };

template <int kMode>
__device__ __forceinline__ Phase0Result coop_phase0(const GridView& g, const KRays& kr, const CoopPlan& P,
                                                         int64_t i, int lane, double* recd,
                                                         CoopStartCache* cache, bool table_ok)
{
    Phase0Result r;
    r.ri[0] = r.ri[1] = r.ri[2] = r.ri[3] = r.ri[4] = r.ri[5] = 0;
    double x[3], d[3];
    load_ray(kr, i, x, d);
    r.start_index = i / kr.rps;
    RayConsts c;
    ray_setup(c, x, d);
    CoopRecord rec;
    r.cls = coop_plan_ray<kMode>(g, c, x, P, r.start_index, kr.cells, *cache, rec);
    if (!table_ok && r.cls == COOP_RAY) r.cls = rec.e_mid == 0 ? COOP_LISTED : COOP_TRACKED;
    if (r.cls == COOP_RAY) {
        recd[0 * 32 + lane] = rec.d0; recd[1 * 32 + lane] = rec.d1; recd[2 * 32 + lane] = rec.d2;
        recd[3 * 32 + lane] = rec.tc; recd[4 * 32 + lane] = rec.dperp2;
        recd[5 * 32 + lane] = rec.t_turn;
        r.ri[0] = rec.imin | (rec.e_mid << 16);
        r.ri[1] = rec.e_first; r.ri[2] = rec.e_count * rec.e_step;
        r.ri[3] = rec.a_first; r.ri[4] = rec.a_count * rec.a_dir;
        r.ri[5] = (rec.se0 & 0xffff) | (rec.sa0 << 16);
    }
    return r;
}

template <class Op, int kMode>
__global__ void __launch_bounds__(kCoopMaxWarps * 32, 1)
walk_coop_kernel(sphrt_grid kg, KRays kr, Op op, CoopPlan P, const uint32_t* in_mask,
                 uint32_t* listed_mask, uint32_t* tracked_mask, unsigned long long* tile_counter,
                 unsigned long long* stats)
{
    // COOP_ORDINARY / COOP_BOTH, in_mask == nullptr: every ray; rays left for the all-cones run or
    // pass B are flagged in listed_mask.  COOP_ALLCONES: only the rays flagged in in_mask (by the
    // run before, or by the per-thread walker), walked in all-cones mode; what is left over is
    // flagged in listed_mask again (which may be in_mask itself: one warp owns a word).
    extern __shared__ double sm_tables[];
    const GridView g = stage_tables(kg, sm_tables);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    char* wbase = reinterpret_cast<char*>(sm_tables + P.table_doubles) + (size_t)warp * P.warp_bytes;
    double* recd = reinterpret_cast<double*>(wbase + P.off_rec);
    const int64_t n_tiles = (kr.n + 31) >> 5;
    unsigned long long n_coop = 0, n_handed = 0;
    CoopStartCache start_cache{-1, true, true, 0, 0, 0};
    // the cooperative ranges assume cos(e_b) descends (e_b ascending within [0, pi])
    bool table_ok = true;
    for (int k = threadIdx.x; k < g.ne; k += blockDim.x) table_ok &= g.cos_e()[k] > g.cos_e()[k + 1];
    table_ok = __syncthreads_and(table_ok);
    for (;;) {
        // tiles are handed out dynamically: their cost varies from nothing (misses) to ~1 ms
        unsigned long long next = 0;
        if (lane == 0) next = atomicAdd(tile_counter, 1ull);
        const int64_t tile = (int64_t)__shfl_sync(0xffffffffu, next, 0);
        if (tile >= n_tiles) break;
        // ---- phase 0: one lane per ray
        const int64_t i = (tile << 5) + lane;
        int cls = COOP_MISS;
        int ri0 = 0, ri1 = 0, ri2 = 0, ri3 = 0, ri4 = 0, ri5 = 0;      // integer part of the record
        int64_t start_index = 0;
        // (plain load: in_mask may be listed_mask itself, which this kernel writes)
        const uint32_t sel = in_mask != nullptr ? in_mask[tile] : 0xffffffffu;
        if (sel == 0u) { if (lane == 0) listed_mask[tile] = 0u; continue; }
        if (i < kr.n && ((sel >> lane) & 1u) && !op.void_ray(i)) {
            const Phase0Result r0 = coop_phase0<kMode>(g, kr, P, i, lane, recd, &start_cache, table_ok);
            cls = r0.cls; start_index = r0.start_index;
            ri0 = r0.ri[0]; ri1 = r0.ri[1]; ri2 = r0.ri[2]; ri3 = r0.ri[3]; ri4 = r0.ri[4]; ri5 = r0.ri[5];
        }
        if (cls == COOP_MISS && i < kr.n && ((sel >> lane) & 1u)) {   // never enters the grid: empty result
            typename Op::Sink none = op.begin(i);
            op.end(i, none);
        }
        const uint32_t m_ray = __ballot_sync(0xffffffffu, cls == COOP_RAY);
        uint32_t m_listed = __ballot_sync(0xffffffffu, cls == COOP_LISTED);
        uint32_t m_tracked = __ballot_sync(0xffffffffu, cls == COOP_TRACKED);
        __syncwarp();
        // ---- phase 1: 32 lanes per ray
        for (uint32_t m = m_ray; m; m &= m - 1) {
            const int j = __ffs(m) - 1;
            const int64_t ray = (tile << 5) + j;
            CoopRecord rec;
            rec.d0 = recd[0 * 32 + j]; rec.d1 = recd[1 * 32 + j]; rec.d2 = recd[2 * 32 + j];
            rec.tc = recd[3 * 32 + j]; rec.dperp2 = recd[4 * 32 + j];
            rec.t_turn = recd[5 * 32 + j];
            const int im = __shfl_sync(0xffffffffu, ri0, j);
            rec.imin = im & 0xffff; rec.e_mid = im >> 16;
            rec.e_first = __shfl_sync(0xffffffffu, ri1, j);
            const int ec = __shfl_sync(0xffffffffu, ri2, j), ac = __shfl_sync(0xffffffffu, ri4, j);
            const int stt = __shfl_sync(0xffffffffu, ri5, j);
            rec.e_step = ec < 0 ? -1 : 1; rec.e_count = ec < 0 ? -ec : ec;
            rec.a_first = __shfl_sync(0xffffffffu, ri3, j);
            rec.a_dir = ac < 0 ? -1 : 1; rec.a_count = ac < 0 ? -ac : ac;
            rec.se0 = (int)(int16_t)(stt & 0xffff); rec.sa0 = stt >> 16;
            const double* xs = kr.starts + 3 * __shfl_sync(0xffffffffu, start_index, j);
            const double x[3] = {__ldg(xs), __ldg(xs + 1), __ldg(xs + 2)};
            CoopLists L;
            // a ray that fails a check goes back to the per-thread walker of its kind
            uint32_t& m_back = rec.e_mid == 0 ? m_listed : m_tracked;
            if (!coop_build<kMode>(g, P, wbase, rec, x, lane, L)) { m_back |= 1u << j; n_handed++; continue; }
            double* tEA; int32_t* sEA;
            const int nEA = coop_merge_ea(P, wbase, L, lane, tEA, sEA);
            if (nEA < 0) { m_back |= 1u << j; n_handed++; continue; }        // a list out of order
            typename Op::Sink sink = op.begin(ray);
            coop_segments<Op::kEvents>(L, tEA, sEA, nEA, lane, sink);
            op.coop_end(ray, sink, lane, L.nR + L.nE + L.nA);
            n_coop++;
            __syncwarp();              // the lists are reused by the next ray
        }
        if (lane == 0) { listed_mask[tile] = m_listed; if (in_mask == nullptr) tracked_mask[tile] = m_tracked; }
        if (stats != nullptr && lane == 0) {
            if (kMode != COOP_ORDINARY) atomicAdd(stats + 1, (unsigned long long)__popc(m_listed));    // left for pass B
            if (kMode != COOP_ALLCONES) atomicAdd(stats + 2, (unsigned long long)__popc(m_tracked));
        }
    }
    if (stats != nullptr && lane == 0) { atomicAdd(stats + 0, n_coop); atomicAdd(stats + 3, n_handed); }
}


// ---- second-generation warp walker (sphrt_warp.cuh) ---------------------------------------------
// plan_warp_kernel: one thread per ray of a chunk.  Classifies the ray and, for the rays the warp
// walker takes, appends an 80-byte record to a dense list; the other kinds are flagged in the two
// bitmasks of the fallback passes exactly like walk_coop_kernel<COOP_ORDINARY> does.
constexpr int kPlanBlock = 256;

__global__ void __launch_bounds__(kPlanBlock)
plan_warp_kernel(sphrt_grid kg, KRays kr, int64_t base, int64_t n, CoopPlan P, PlanRec* recs,
                 unsigned* n_recs /* [0] ordinary rays, from the front; [1] all-cones rays, from the back */,
                 unsigned* n_fallback /* [0] rays left for pass B, [1] for the per-thread walker */,
                 int64_t cap, uint32_t* mask_listed, uint32_t* mask_tracked, int32_t* miss_count,
                 unsigned long long* stats)
{
    extern __shared__ double sm_tables[];
    const GridView g = stage_tables(kg, sm_tables);
    bool table_ok = true;                  // the candidate ranges assume cos(e_b) descends
    for (int k = threadIdx.x; k < g.ne; k += blockDim.x) table_ok &= g.cos_e()[k] > g.cos_e()[k + 1];
    table_ok = __syncthreads_and(table_ok);
    const int lane = threadIdx.x & 31;
    const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // index inside the chunk
    const int64_t i = base + li;
    int cls = COOP_MISS;
    CoopRecord rec;
    RayConsts c;
    int64_t start_index = 0;
    if (li < n) {
        double x[3], d[3];
        load_ray(kr, i, x, d);
        start_index = i / kr.rps;
        ray_setup(c, x, d);
        CoopStartCache cache{-1, true, true, 0, 0, 0};
        cls = coop_plan_ray<COOP_BOTH, true>(g, c, x, P, start_index, kr.cells, cache, rec);
        if (!table_ok && cls == COOP_RAY) cls = rec.e_mid == 0 ? COOP_LISTED : COOP_TRACKED;
        if (cls == COOP_MISS && miss_count != nullptr) miss_count[i] = 0;
    }
    const bool allc = cls == COOP_RAY && rec.e_mid == 0;           // walked in all-cones mode
    const uint32_t m_ray = __ballot_sync(0xffffffffu, cls == COOP_RAY && !allc);
    const uint32_t m_all = __ballot_sync(0xffffffffu, allc);
    const uint32_t m_listed = __ballot_sync(0xffffffffu, cls == COOP_LISTED);
    const uint32_t m_tracked = __ballot_sync(0xffffffffu, cls == COOP_TRACKED);
    unsigned slot0 = 0, slot1 = 0;
    if (lane == 0 && li < n) {
        mask_listed[i >> 5] = m_listed;
        mask_tracked[i >> 5] = m_tracked;
        if (m_ray) slot0 = atomicAdd(n_recs, (unsigned)__popc(m_ray));
        if (m_all) slot1 = atomicAdd(n_recs + 1, (unsigned)__popc(m_all));
        if (m_listed) atomicAdd(n_fallback, (unsigned)__popc(m_listed));      // work of the fallback passes
        if (m_tracked) atomicAdd(n_fallback + 1, (unsigned)__popc(m_tracked));
        if (stats != nullptr) {
            if (m_listed) atomicAdd(stats + 1, (unsigned long long)__popc(m_listed));
            if (m_tracked) atomicAdd(stats + 2, (unsigned long long)__popc(m_tracked));
        }
    }
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    slot1 = __shfl_sync(0xffffffffu, slot1, 0);
    if (cls == COOP_RAY) {
        PlanRec r;
        r.d0 = rec.d0; r.d1 = rec.d1; r.d2 = rec.d2;
        r.tc = rec.tc; r.dperp2 = rec.dperp2; r.t_turn = rec.t_turn; r.nx2 = c.nx2;
        r.ray = (int32_t)li; r.start = (int32_t)start_index;
        r.imin = (int16_t)rec.imin; r.e_mid = (int16_t)rec.e_mid;
        r.e_first = (int16_t)rec.e_first; r.e_cs = (int16_t)(rec.e_count * rec.e_step);
        r.a_first = (int16_t)rec.a_first; r.a_cd = (int16_t)(rec.a_count * rec.a_dir);
        r.se0 = (int16_t)rec.se0; r.sa0 = (int16_t)rec.sa0;
        const uint32_t lt = (1u << lane) - 1u;
        if (allc) store_plan_rec(recs + (cap - 1 - (int64_t)(slot1 + __popc(m_all & lt))), r);
        else store_plan_rec(recs + slot0 + __popc(m_ray & lt), r);
    }
}

#ifndef SPHRT_WARP_BATCH
#define SPHRT_WARP_BATCH 8
#endif
constexpr unsigned kWarpBatch = SPHRT_WARP_BATCH;   // records per hand-out

// walk_warp_kernel: persistent warps take batches of planned rays from the dense record list.
// kAllCones: the list of all-cones records (stored from the back of the buffer: `recs` points at its
// first record and is walked downwards); failures go to pass B's mask.
// Op::kWarpWarps: warps per block (= per SM) the kernel is compiled for.  24 warps leave 80
// registers per thread, 20 leave 96: the forward pass is faster with 20 on large grids (its
// gathers evict the spill slots from L1, so it pays for every spilled register with an L2 round
// trip: 9.23 against 9.60 ms per 16 config-5 observations), the adjoint with 24 (8.33 / 8.72).
template <class Op, bool kAllCones>
__global__ void __launch_bounds__(Op::kWarpWarps * 32, 1)
walk_warp_kernel(sphrt_grid kg, const double* __restrict__ starts, Op op, WarpPlan P,
                 const PlanRec* __restrict__ recs, const unsigned* n_recs_p, int64_t ray_base,
                 uint32_t* mask_back, unsigned* counter, unsigned* n_back, unsigned long long* stats)
{
    extern __shared__ __align__(16) double sm_tables[];
    // derived tables (Tab32): r_b^2 | { cos^2 e_b, shadow sign } | { cos a_b, sin a_b }
    {
        const int o_cone = (kg.nr + 2) & ~1, o_plane = o_cone + 2 * (kg.ne + 1);
        for (int i = threadIdx.x; i <= kg.nr; i += blockDim.x) { const double r = __ldg(kg.r_b + i); sm_tables[i] = r * r; }
        for (int i = threadIdx.x; i <= kg.ne; i += blockDim.x) {
            const double ce = __ldg(kg.cos_e + i);
            sm_tables[o_cone + 2 * i] = ce * ce;
            sm_tables[o_cone + 2 * i + 1] = cone_shadow_sign(ce, __ldg(kg.e_b + i));
        }
        for (int i = threadIdx.x; i <= kg.na; i += blockDim.x) {
            sm_tables[o_plane + 2 * i] = __ldg(kg.cos_a + i);
            sm_tables[o_plane + 2 * i + 1] = __ldg(kg.sin_a + i);
        }
        __syncthreads();
    }
    // The lane index, its mask and the two shared-memory bases are passed through a shuffle so that
    // ptxas keeps them in registers: it otherwise re-derives them from SR_TID.X / SR_CgaCtaId
    // inside the loops (S2R has the latency of a shared-memory load).
    int lane = threadIdx.x & 31;
    uint32_t tab32 = smem_u32(sm_tables);
    uint32_t wbase = tab32 + 8u * (uint32_t)P.table_doubles + (uint32_t)((threadIdx.x >> 5) * P.warp_bytes);
    lane = __shfl_sync(0xffffffffu, lane, threadIdx.x);
    tab32 = __shfl_sync(0xffffffffu, tab32, lane);
    wbase = __shfl_sync(0xffffffffu, wbase, lane);
    const uint32_t lt_mask = __shfl_sync(0xffffffffu, (1u << lane) - 1u, lane);
    const unsigned n_recs = *n_recs_p;
    if (n_recs == 0) return;
    unsigned n_ok = 0, n_fail = 0;
    auto rec_at = [&](unsigned k) { return kAllCones ? recs - k : recs + k; };
    // batches are handed out one ahead: the atomic and the record prefetch of the next batch are
    // in flight while the current one is walked
    // (records per hand-out: kWarpBatch when every warp gets many batches; small launches -- a few
    // records per warp, like the all-cones rays of a 50^3 problem -- take them one or two at a time,
    // or most warps would idle while a few walk eight rays in a row)
    unsigned batch = n_recs / (8u * gridDim.x * (blockDim.x >> 5));
    batch = batch < 1u ? 1u : (batch > kWarpBatch ? kWarpBatch : batch);
    auto grab = [&]() {
        unsigned v = 0;
        if (lane == 0) v = atomicAdd(counter, batch);
        v = __shfl_sync(0xffffffffu, v, 0);
        if (v < n_recs && lane < 6) {
            const unsigned e = v + batch < n_recs ? v + batch : n_recs;
            const PlanRec* first = kAllCones ? rec_at(e - 1) : rec_at(v);       // lowest address of the batch
            const char* p = reinterpret_cast<const char*>(first) + 128 * lane;
            if (p < reinterpret_cast<const char*>(first + (e - v))) asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
        }
        return v;
    };
    unsigned b = grab();
    while (b < n_recs) {
        const unsigned e = b + batch < n_recs ? b + batch : n_recs;
        const unsigned b_next = grab();
        for (unsigned k = b; k < e; k++) {
            const PlanRec R = load_plan_rec(rec_at(k));
            const int64_t ray = ray_base + R.ray;
            if (op.void_ray(ray)) continue;
            typename Op::WSink sink = op.wbegin(ray);
            // Forward pass: the two shared-memory bases are made opaque once per ray, so that the ten
            // list / table addresses derived from them are formed where they are used (one add each)
            // instead of being hoisted out of the ray loop into registers that end up spilled -- the
            // gathers leave the spill slots no room in L1 there (87 % of the local loads missed;
            // forward -3.6 %).  The adjoint's spills stay in L1 and it is faster with the hoisting.
            uint32_t wb = wbase, tb = tab32;
            if constexpr (Op::kOpaqueBases) asm volatile("" : "+r"(wb), "+r"(tb));
            const Tab32 Tr(tb, kg.nr, kg.ne, kg.na);
            const int nev = warp_walk_ray<Op::kEvents, kAllCones>(Tr, P, wb, R, starts + 3 * (int64_t)R.start, lane, lt_mask, sink);
            if (nev < 0) {
                // a structural check failed: a per-thread walker takes the ray (sorted-list walker
                // of pass B for the all-cones rays)
                if (lane == 0) { atomicOr(mask_back + (ray >> 5), 1u << (ray & 31)); atomicAdd(n_back, 1u); }
                n_fail++;
            } else {
                op.wend(ray, sink, lane, nev);
                n_ok++;
            }
            __syncwarp();              // the lists are reused by the next ray
        }
        b = b_next;
    }
    if (stats != nullptr && lane == 0) {
        atomicAdd(stats + 0, (unsigned long long)n_ok);
        atomicAdd(stats + (kAllCones ? 1 : 2), (unsigned long long)n_fail);
        atomicAdd(stats + 3, (unsigned long long)n_fail);
    }
}

// pass B: persistent warps walk the rays flagged by pass A from a heap-sorted list of all cone
// roots.  A warp takes one 32-ray mask word at a time (flagged rays come in runs along a
// detector row, so its lanes fill up); each lane owns a [slot][thread] list column in shared
// memory, or -- grids with very many cones -- a contiguous slice of global scratch.
template <class Op>
__global__ void __launch_bounds__(kListBlock)
walk_listed_kernel(sphrt_grid kg, KRays kr, Op op, const uint32_t* listed_mask, int64_t n_words,
                   int table_doubles_smem, double* glist_t, int* glist_r,
                   unsigned long long* chunk_counter, const unsigned* n_work)
{
    if (n_work != nullptr && *n_work == 0) return;
    extern __shared__ double sm_tables[];
    __shared__ int64_t flag_queue[(kListBlock / 32) * kFlagQueue];
    const GridView g = stage_tables(kg, sm_tables);
    const int cap = 2 * (kg.ne + 1);
    ConeList list;
    if (glist_t == nullptr) {
        double* lt = sm_tables + table_doubles_smem;
        int* lr = reinterpret_cast<int*>(lt + (size_t)cap * blockDim.x);
        list = ConeList{lt + threadIdx.x, lr + threadIdx.x, (int64_t)blockDim.x, cap};
    } else {
        const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        list = ConeList{glist_t + tid * cap, glist_r + tid * cap, 1, cap};
    }
    for_each_flagged(listed_mask, n_words, chunk_counter, flag_queue + (threadIdx.x >> 5) * kFlagQueue,
                     [&](int64_t i, bool active) {
        if (!active) return;
        double x[3], d[3];
        load_ray(kr, i, x, d);
        RayConsts c;
        ray_setup(c, x, d);
        typename Op::Sink sink = op.begin(i);
        const StartCell cell = ray_start_cell(kr, i, g, x);
        walk_ray<true>(g, c, x, list, sink, &cell);
        op.end(i, sink);
    });
}

// Unsorted candidate table of the reference (trace_indices :83-122 before the sort): for every
// ray [first radial roots | second radial roots | first cone roots | second cone roots | planes],
// distance, region after the crossing and geometric index (boundary index - negative crossing).
// Debug / parity export behind r_torch / e_torch / a_torch of the Python mirror.
__global__ void candidates_kernel(sphrt_grid kg, KRays kr, double* t_out, int32_t* reg_out, int32_t* geo_out)
{
    extern __shared__ double sm_tables[];
    const GridView g = stage_tables(kg, sm_tables);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kr.n) return;
    double x[3], d[3];
    load_ray(kr, i, x, d);
    RayConsts c;
    ray_setup(c, x, d);
    const int nr1 = g.nr + 1, ne1 = g.ne + 1, na1 = g.na + 1;
    const int64_t K1 = 2 * nr1 + 2 * ne1 + na1;
    double* t = t_out + i * K1;
    int32_t* reg = reg_out + i * K1;
    int32_t* geo = geo_out + i * K1;
    for (int k = 0; k < nr1; k++) {
        const Event a = radial_event(g, c, k, false), b = radial_event(g, c, k, true);
        t[k] = a.t; reg[k] = a.region; geo[k] = a.geo;
        t[nr1 + k] = b.t; reg[nr1 + k] = b.region; geo[nr1 + k] = b.geo;
    }
    for (int k = 0; k < ne1; k++) {
        Event a, b;
        cone_events(g, c, k, a, b);
        const int o = 2 * nr1;
        t[o + k] = a.t; reg[o + k] = a.region; geo[o + k] = a.geo;
        t[o + ne1 + k] = b.t; reg[o + ne1 + k] = b.region; geo[o + ne1 + k] = b.geo;
    }
    for (int k = 0; k < na1; k++) {
        const Event a = plane_event(g, c, k);
        const int o = 2 * nr1 + 2 * ne1;
        t[o + k] = a.t; reg[o + k] = a.region; geo[o + k] = a.geo;
    }
}

__global__ void find_starts_kernel(sphrt_grid kg, int64_t n, const double* xs, int32_t* out)
{
    extern __shared__ double sm_tables[];
    const GridView g = stage_tables(kg, sm_tables);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x[3] = {xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]};
    int bins[3], st[3];
    start_voxel(g, x, bins, st);
    out[i] = st[0]; out[n + i] = st[1]; out[2 * n + i] = st[2];
}

// ---- host side -------------------------------------------------------------------------------

int num_sms()
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        return 148;
    return sms;
}

constexpr int kSmemBudget = 200 * 1024;     // per block, pass B lists + tables

// pass B launch shape: threads per block such that the per-lane cone lists fit in shared
// memory (0 = they do not: global lists)
struct ListPlan { int threads; size_t smem; int64_t blocks; bool global_lists; };

ListPlan plan_lists(const sphrt_grid& g, int64_t n_rays)
{
    ListPlan p;
    const size_t tables = (size_t)table_doubles(g) * sizeof(double);
    const size_t per_thread = (size_t)2 * (g.ne + 1) * (sizeof(double) + sizeof(int));
    int threads = (int)((kSmemBudget - tables) / per_thread) / 32 * 32;
    if (threads > kListBlock) threads = kListBlock;
    p.global_lists = threads < 32;
    p.threads = p.global_lists ? kListBlock : threads;
    p.smem = tables + (p.global_lists ? 0 : per_thread * p.threads);
    int per_sm = (int)((220 * 1024) / (p.smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const int64_t words = (n_rays + 31) / 32;
    const int64_t want = (words * 32 + p.threads - 1) / p.threads;
    p.blocks = (int64_t)num_sms() * per_sm;
    if (p.blocks > want) p.blocks = want;
    if (p.blocks < 1) p.blocks = 1;
    return p;
}

int64_t mask_bytes(int64_t n_rays) { return ((n_rays + 31) / 32 * 4 + 255) / 256 * 256; }

// ---- warp-cooperative walker: shared-memory plan and selection -----------------------------------
struct CoopLaunch { CoopPlan P; size_t smem; int64_t blocks; bool ok; };

CoopLaunch plan_coop(const sphrt_grid& g, int64_t n_rays)
{
    CoopLaunch c;
    CoopPlan& P = c.P;
    P.capR = 2 * (g.nr + 1);
    P.capE = g.ne + 12;
    P.capA = g.na / 2 + 4;
    int off = 0;
    auto take = [&off](int bytes) { const int o = off; off += (bytes + 7) / 8 * 8; return o; };
    P.off_tR = take((P.capR + 1) * 8);
    P.off_tE = take((P.capE + 1) * 8);
    P.off_tA = take((P.capA + 1) * 8);
    P.off_rE = take((P.capE + 2) * 2);
    P.off_rA = take((P.capA + 2) * 2);
    // a chord's colatitude varies by less than pi and its azimuth by at most pi
    P.capEA = g.ne + g.na / 2 + 16;
    P.off_tEA = take((P.capEA + 1) * 8);
    P.off_sEA = take((P.capEA + 2) * 4);
    P.off_rec = take(32 * kCoopRecDoubles * 8);
    P.warp_bytes = off;
    P.table_doubles = (int)table_doubles(g);
    const int64_t tables = table_doubles(g) * (int64_t)sizeof(double);
    int64_t warps = (220 * 1024 - tables) / P.warp_bytes;
    if (warps > kCoopMaxWarps) warps = kCoopMaxWarps;
    c.ok = warps >= 4;
    P.warps = (int)(warps > 0 ? warps : 1);
    c.smem = (size_t)(tables + (int64_t)P.warps * P.warp_bytes);
    const int64_t tiles = (n_rays + 31) / 32;
    c.blocks = num_sms();
    const int64_t want = (tiles + P.warps - 1) / P.warps;
    if (c.blocks > want) c.blocks = want;
    if (c.blocks < 1) c.blocks = 1;
    return c;
}


// ---- second-generation warp walker: shared-memory plan, chunking, selection -----------------------
constexpr int kSmemMaxDynamic = 227 * 1024;
constexpr int64_t kWarpChunkRays = 4 << 20;    // rays planned and walked per pair of launches

struct WarpLaunch { WarpPlan P; CoopPlan caps; size_t smem; bool ok; };

WarpLaunch plan_warp(const sphrt_grid& g)
{
    WarpLaunch w;
    WarpPlan& P = w.P;
    P.capR = 2 * (g.nr + 1);
    P.capE = g.ne + 12;
    P.capA = g.na / 2 + 4;
    // a chord's colatitude varies by less than pi and its azimuth by at most pi
    P.capEA = g.ne + g.na / 2 + 16;
    // (rays through the Z axis, walked with every half-plane evaluated, can meet more of them: on
    // grids where that costs little every half-plane gets a slot; elsewhere such a ray overflows
    // into pass B)
    if (g.na <= 128 && g.ne <= 128) {
        P.capA = g.na + 2;                 // every half-plane once
        P.capE = 2 * (g.ne + 1) + 2;       // every cone twice (such a ray can pass close to the origin)
        P.capEA = P.capE + P.capA;
    }
    P.padR = (P.capR + P.capEA + 31) / 32 + 2;
    int off = 0;
    auto take = [&off](int bytes) { const int o = off; off += (bytes + 7) / 8 * 8; return o; };
    P.off_tE = take((P.capE + 1) * 8);
    P.off_rE = take((P.capE + 2) * 4);
    P.off_tA = take((P.capA + 1) * 8);
    P.off_rA = take((P.capA + 2) * 4);
    const int size_a = off;
    P.off_tR = 0;
    const int size_b = (P.capR + P.padR) * 8;
    off = ((size_a > size_b ? size_a : size_b) + 15) / 16 * 16;
    P.off_tEA = take((P.capEA + 1) * 8);
    P.off_sEA = take((P.capEA + 2) * 4);
    P.off_flag = take(8);
    P.warp_bytes = (off + 15) / 16 * 16;
    P.table_doubles = Tab32::doubles(g.nr, g.ne, g.na);        // the walking kernel's own tables
    const int64_t tables = P.table_doubles * (int64_t)sizeof(double);
    int64_t warps = (kSmemMaxDynamic - tables) / P.warp_bytes;
    if (warps > kWarpMaxWarps) warps = kWarpMaxWarps;
    // packed states need nvox <= 2^30; the records hold 16-bit indices
    // packed states: (nr + padding) * ne * na must stay below 2^30; the records hold 16-bit indices
    w.ok = warps >= 4 && (int64_t)(g.nr + P.padR + 2) * g.ne * g.na <= (1LL << 30) &&
           g.nr < 32000 && g.ne < 32000 && g.na < 32000;
    P.warps = (int)(warps > 0 ? warps : 1);
    w.smem = (size_t)(tables + (int64_t)P.warps * P.warp_bytes);
    memset(&w.caps, 0, sizeof(w.caps));
    w.caps.capE = P.capE; w.caps.capA = P.capA; w.caps.capR = P.capR; w.caps.capEA = P.capEA;
    return w;
}

int64_t warp_chunk_rays()
{
    const char* e = getenv("SPHRT_WARP_CHUNK");
    if (e != nullptr && *e) {
        const long long v = strtoll(e, nullptr, 10);
        if (v >= 32) return (int64_t)(v / 32 * 32);
    }
    return kWarpChunkRays;
}

// SPHRT_COOP_GEN = 2 (default) | 1 : generation of the cooperative walker that takes the ordinary rays
bool want_warp_gen2()
{
    const char* e = getenv("SPHRT_COOP_GEN");
    return !(e != nullptr && strcmp(e, "1") == 0);
}

int64_t warp_scratch_bytes(const sphrt_grid& g, int64_t n_rays)
{
    if (!plan_warp(g).ok) return 0;
    const int64_t chunk = n_rays < warp_chunk_rays() ? n_rays : warp_chunk_rays();
    return 256 + ((chunk + 31) / 32 * 32) * (int64_t)sizeof(PlanRec);
}

// SPHRT_WALKER = auto (default) | coop | tracked : which walker takes the ordinary rays.
// auto: the cooperative walker pays off once a ray meets a few hundred boundaries.
// SPHRT_COOP_LISTED=0 keeps the sorted-cone-list rays of the per-thread path in pass B (tests)
bool no_coop_listed()
{
    const char* e = getenv("SPHRT_COOP_LISTED");
    return e != nullptr && strcmp(e, "0") == 0;
}

bool want_coop(const sphrt_grid& g, int64_t hint = 0)
{
    const char* e = getenv("SPHRT_WALKER");
    if (e != nullptr && strcmp(e, "coop") == 0) return true;
    if (e != nullptr && strcmp(e, "tracked") == 0) return false;
    if (hint == 1) return true;
    if (hint == 2) return false;
    return g.nr + g.ne + g.na >= 300;
}

int check_grid(const sphrt_grid* g)
{
    if (g == nullptr) return SPHRT_EINVAL;
    if (g->nr < 1 || g->ne < 1 || g->na < 1 || g->n_c2 < 0) return SPHRT_EINVAL;
    if (g->nr > SPHRT_MAX_AXIS || g->ne > SPHRT_MAX_AXIS || g->na > SPHRT_MAX_AXIS) return SPHRT_ETOOBIG;
    if ((int64_t)g->nr * g->ne * g->na > 0x7fffffffLL) return SPHRT_ETOOBIG;   // 32-bit voxel offsets
    if (table_doubles(*g) * (int64_t)sizeof(double) > kMaxSmemBytes) return SPHRT_ETOOBIG;   // tables live in smem
    if (!g->r_b || !g->e_b || !g->cos_e || !g->a_b || !g->cos_a || !g->sin_a) return SPHRT_EINVAL;
    if (g->n_c2 > 0 && !g->c2_sorted) return SPHRT_EINVAL;
    return SPHRT_OK;
}

int check_rays(const sphrt_rays* r)
{
    if (r == nullptr || r->n_rays < 0) return SPHRT_EINVAL;
    if (r->n_rays > 0 && (!r->starts || r->rays_per_start < 1)) return SPHRT_EINVAL;
    if (r->gen_kind == 0) {
        if (r->n_rays > 0 && (!r->dirs || r->rays_per_dir < 1)) return SPHRT_EINVAL;
    } else {
        if (r->gen_kind < 0 || r->gen_kind > 2 || r->gen_h < 1 || r->gen_w < 1) return SPHRT_EINVAL;
        if (!r->gen_frames || !r->gen_ta || !r->gen_tb || (r->gen_kind == 2 && !r->gen_tc)) return SPHRT_EINVAL;
        if (r->n_rays % ((int64_t)r->gen_h * r->gen_w) != 0) return SPHRT_EINVAL;
    }
    if (r->walker < 0 || r->walker > 2) return SPHRT_EINVAL;
    return SPHRT_OK;
}

int cuda_status()          // called once after every kernel launch
{
    sphrt_g_launches++;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { sphrt_g_last_cuda_error = (int)e; return SPHRT_ECUDA; }
    return SPHRT_OK;
}


// offset of the warp walker's record buffer inside the scratch area (after masks, counters, lists)
int64_t scratch_bytes_before_recs(const sphrt_grid* grid, int64_t n_rays)
{
    int64_t bytes = 2 * mask_bytes(n_rays) + 256;
    const ListPlan lp = plan_lists(*grid, n_rays);
    if (lp.global_lists)
        bytes += lp.blocks * lp.threads * 2 * (int64_t)(grid->ne + 1) * (int64_t)(sizeof(double) + sizeof(int));
    return (bytes + 255) / 256 * 256;
}

template <class Op>
int launch_listed(const sphrt_grid* grid, const KRays& kr, const Op& op, const uint32_t* mask,
                  void* lists, size_t smem, const ListPlan& lp, unsigned long long* chunk_counter,
                  cudaStream_t st, const unsigned* n_work = nullptr)
{
    double* glist_t = nullptr;
    int* glist_r = nullptr;
    if (lp.global_lists) {
        glist_t = (double*)lists;
        glist_r = (int*)(glist_t + lp.blocks * lp.threads * 2 * (int64_t)(grid->ne + 1));
    }
    if (lp.smem > 40 * 1024)       // (+ the static flag queue)
        cudaFuncSetAttribute(walk_listed_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.smem);
    walk_listed_kernel<Op><<<(unsigned)lp.blocks, lp.threads, lp.smem, st>>>(
        *grid, kr, op, mask, (kr.n + 31) / 32, (int)(smem / sizeof(double)), glist_t, glist_r,
        chunk_counter, n_work);
    return cuda_status();
}

// Everything a pass needs to know about the launch: rays, scratch carved into its parts.
struct PassContext {
    const sphrt_grid* grid;
    KRays kr;
    cudaStream_t st;
    size_t smem_tables;                  // bytes of boundary tables in shared memory
    uint32_t* mask_listed;               // rays for the sorted-list walkers (all-cones run, pass B)
    uint32_t* mask_tracked;              // rays the cooperative walker hands to the per-thread walker
    unsigned long long* counters;        // [0] cooperative tiles  [1] masked per-thread pass
                                         // [2] pass B  [3] cooperative all-cones tiles
    void* lists;                         // global cone lists of pass B (very large grids only)
    int64_t n_words;
    PlanRec* recs;                       // records of the second-generation warp walker (one chunk)
    // plan cache (sphrt_rays.plan_cache): every chunk's records and counts kept between calls
    PlanRec* cache_recs = nullptr;       // chunk c at cache_recs + c * cache_stride
    int64_t cache_stride = 0;
    unsigned* cache_counts = nullptr;    // 8 unsigned per chunk: [0] ordinary, [1] all-cones records
    unsigned* cache_totals = nullptr;    // [0], [1]: rays of the two fallback passes
    int cache_state = 0;                 // 1: fill, 2: use
};

// Layout of a plan cache: [256 B totals][32 B per chunk][listed mask][tracked mask][records]
struct PlanCacheLayout { int64_t n_chunks, off_counts, off_masks, off_recs, stride, bytes; };
inline PlanCacheLayout plan_cache_layout(int64_t n_rays)
{
    PlanCacheLayout L;
    const int64_t chunk = warp_chunk_rays();
    L.n_chunks = (n_rays + chunk - 1) / chunk;
    L.stride = (n_rays < chunk ? n_rays : chunk) + 32;          // records of one chunk
    L.off_counts = 256;
    L.off_masks = (L.off_counts + 32 * L.n_chunks + 255) / 256 * 256;
    L.off_recs = L.off_masks + 2 * mask_bytes(n_rays);
    L.bytes = L.off_recs + L.n_chunks * L.stride * (int64_t)sizeof(PlanRec);          // chunk c at c * stride
    return L;
}

// Point the pass at the caller's plan cache, if there is one that fits (masks, records, counts).
inline void attach_plan_cache(PassContext& cx, const sphrt_rays* rays)
{
    if (rays->plan_cache == nullptr || rays->plan_cache_state < 1 || rays->plan_cache_state > 2) return;
    const PlanCacheLayout L = plan_cache_layout(rays->n_rays);
    if (rays->plan_cache_bytes < L.bytes) return;
    char* c = (char*)rays->plan_cache;
    cx.cache_totals = (unsigned*)c;
    cx.cache_counts = (unsigned*)(c + L.off_counts);
    cx.mask_listed = (uint32_t*)(c + L.off_masks);
    cx.mask_tracked = (uint32_t*)(c + L.off_masks + mask_bytes(rays->n_rays));
    cx.cache_recs = (PlanRec*)(c + L.off_recs);
    cx.cache_stride = L.stride;
    cx.cache_state = (int)rays->plan_cache_state;
}

template <class Op, int kMode>
int launch_coop(const PassContext& cx, const CoopLaunch& cl, const Op& op, const uint32_t* in_mask,
                unsigned long long* tile_counter, unsigned long long* stats)
{
    cudaFuncSetAttribute(walk_coop_kernel<Op, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cl.smem);
    walk_coop_kernel<Op, kMode><<<(unsigned)cl.blocks, cl.P.warps * 32, cl.smem, cx.st>>>(
        *cx.grid, cx.kr, op, cl.P, in_mask, cx.mask_listed, cx.mask_tracked, tile_counter, stats);
    return cuda_status();
}


// second-generation warp walker over every ray, chunk by chunk: plan -> walk (ordinary rays) ->
// walk (all-cones rays).  Counters (unsigned, behind the pass counters): [0] ordinary records
// [1] all-cones records  [2] / [3] hand-out counters of the two walks  [4] rays left for pass B
// [5] rays left for the per-thread walker (accumulated over the chunks).
template <class Op>
int launch_warp(const PassContext& cx, const WarpLaunch& wl, const Op& op, unsigned long long* stats)
{
    int rc = op.prefill(cx.st);
    if (rc != 0) { sphrt_g_last_cuda_error = rc; return SPHRT_ECUDA; }
    const size_t smem_plan = cx.smem_tables;
    if (smem_plan > 48 * 1024)
        cudaFuncSetAttribute(plan_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_plan);
    cudaFuncSetAttribute(walk_warp_kernel<Op, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wl.smem);
    cudaFuncSetAttribute(walk_warp_kernel<Op, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wl.smem);
    unsigned* wc = reinterpret_cast<unsigned*>(cx.counters + 8);
    const int64_t chunk = warp_chunk_rays();
    const int sms = num_sms();
    cudaMemsetAsync(wc, 0, 8 * sizeof(unsigned), cx.st);
    // With a plan cache every chunk has its own record area and counts; the first call (state 1)
    // fills them, the later ones (state 2) skip the planning kernel.  Events are never cached.
    const bool cached = cx.cache_recs != nullptr && !Op::kEvents;
    const bool plan_now = !cached || cx.cache_state == 1;
    if (cached && !plan_now)        // the planner's counts of the rays left to the fallback passes
        cudaMemcpyAsync(wc + 4, cx.cache_totals, 2 * sizeof(unsigned), cudaMemcpyDeviceToDevice, cx.st);
    int64_t c = 0;
    for (int64_t base = 0; base < cx.kr.n; base += chunk, c++) {
        const int64_t n = cx.kr.n - base < chunk ? cx.kr.n - base : chunk;
        const int64_t cap = (n + 31) / 32 * 32;
        PlanRec* recs = cached ? cx.cache_recs + c * cx.cache_stride : cx.recs;
        unsigned* cnt = cached ? cx.cache_counts + 8 * c : wc;          // [0] ordinary, [1] all-cones records
        if (base > 0 || cached) cudaMemsetAsync(wc, 0, 4 * sizeof(unsigned), cx.st);
        if (plan_now) {
            if (cached) cudaMemsetAsync(cnt, 0, 8 * sizeof(unsigned), cx.st);
            const int64_t pb = (n + kPlanBlock - 1) / kPlanBlock;
            plan_warp_kernel<<<(unsigned)pb, kPlanBlock, smem_plan, cx.st>>>(
                *cx.grid, cx.kr, base, n, wl.caps, recs, cnt, wc + 4, cap, cx.mask_listed, cx.mask_tracked,
                op.miss_count(), stats);
            rc = cuda_status();
            if (rc != SPHRT_OK) return rc;
        }
        // (the plan's layout holds for fewer warps; their lists are not allocated, which leaves the
        // SM a larger L1: that, more than the registers, is what the forward pass gains from)
        const int warps = wl.P.warps < Op::kWarpWarps ? wl.P.warps : Op::kWarpWarps;
        const size_t smem = wl.smem - (size_t)(wl.P.warps - warps) * wl.P.warp_bytes;
        int64_t blocks = (n + (int64_t)warps * kWarpBatch - 1) / ((int64_t)warps * kWarpBatch);
        if (blocks > sms) blocks = sms;
        walk_warp_kernel<Op, false><<<(unsigned)blocks, warps * 32, smem, cx.st>>>(
            *cx.grid, cx.kr.starts, op, wl.P, recs, cnt, base, cx.mask_tracked, wc + 2, wc + 5, stats);
        rc = cuda_status();
        if (rc != SPHRT_OK) return rc;
        walk_warp_kernel<Op, true><<<(unsigned)blocks, warps * 32, smem, cx.st>>>(
            *cx.grid, cx.kr.starts, op, wl.P, recs + (cap - 1), cnt + 1, base, cx.mask_listed, wc + 3, wc + 4, stats);
        rc = cuda_status();
        if (rc != SPHRT_OK) return rc;
    }
    if (cached && plan_now)         // (taken after the walks: a superset of the planner's counts)
        cudaMemcpyAsync(cx.cache_totals, wc + 4, 2 * sizeof(unsigned), cudaMemcpyDeviceToDevice, cx.st);
    return SPHRT_OK;
}

template <class Op>
int launch_tracked_masked(const PassContext& cx, const Op& op, const unsigned* n_work = nullptr,
                          unsigned long long* chunk_counter = nullptr)
{
    if (chunk_counter == nullptr) chunk_counter = cx.counters + 1;
    const size_t smem = cx.smem_tables + kQueueBytes;
    if (smem > 40 * 1024)   // (+ the static flag queue)
        cudaFuncSetAttribute(walk_tracked_masked_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (int64_t)num_sms() * SPHRT_TRACKED_MIN_BLOCKS;
    const int64_t want = (cx.n_words + 32 * (kBlock / 32) - 1) / (32 * (kBlock / 32));
    if (blocks > want) blocks = want;
    if (blocks < 1) blocks = 1;
    walk_tracked_masked_kernel<Op><<<(unsigned)blocks, kBlock, smem, cx.st>>>(
        *cx.grid, cx.kr, op, (int)(cx.smem_tables / sizeof(double)), cx.mask_tracked, cx.n_words,
        chunk_counter, n_work);
    return cuda_status();
}

template <class Op>
int launch_tracked_dense(const PassContext& cx, const Op& op)
{
    const size_t smem = cx.smem_tables + kQueueBytes;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(walk_tracked_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t blocks = (cx.kr.n + kBlock - 1) / kBlock;
    if (blocks > 0x7fffffffLL) return SPHRT_ETOOBIG;
    walk_tracked_kernel<Op><<<(unsigned)blocks, kBlock, smem, cx.st>>>(
        *cx.grid, cx.kr, op, (int)(cx.smem_tables / sizeof(double)), cx.mask_listed);
    return cuda_status();
}

// Walk every ray with `op`.  Scratch layout: [listed mask][tracked mask][counters][global cone lists].
//
//   cooperative path : cooperative walker (ordinary rays, then the all-cones rays; one launch for
//                      both below kCoopSplitRays) -> per-thread walker on what it handed over
//                      -> pass B on what is left of the sorted-list rays
//   per-thread path  : per-thread walker -> cooperative all-cones run on its sorted-list rays
//                      (ops that can run cooperatively) -> pass B on what is left
template <class Op>
int launch_passes(const sphrt_grid* grid, const sphrt_rays* rays, const Op& op,
                  void* scratch, int64_t scratch_bytes, void* stream, bool force_coop = false,
                  unsigned long long* stats = nullptr)
{
    if (rays->n_rays == 0) return SPHRT_OK;
    if (scratch_bytes < sphrt_scratch_bytes(grid, rays->n_rays) || scratch == nullptr) return SPHRT_ESCRATCH;
    PassContext cx;
    cx.grid = grid;
    cx.kr = make_krays(rays);
    cx.st = (cudaStream_t)stream;
    cx.smem_tables = (size_t)table_doubles(*grid) * sizeof(double);
    cx.mask_listed = (uint32_t*)scratch;
    cx.mask_tracked = (uint32_t*)((char*)scratch + mask_bytes(rays->n_rays));
    cx.counters = (unsigned long long*)((char*)scratch + 2 * mask_bytes(rays->n_rays));
    cx.lists = (char*)scratch + 2 * mask_bytes(rays->n_rays) + 256;
    cx.n_words = (rays->n_rays + 31) / 32;
    cx.recs = reinterpret_cast<PlanRec*>((char*)scratch + scratch_bytes_before_recs(grid, rays->n_rays));
    cudaMemsetAsync(cx.counters, 0, 4 * sizeof(unsigned long long), cx.st);
    const ListPlan lp = plan_lists(*grid, rays->n_rays);
    int rc;

    if constexpr (Op::kCoop) {
        const CoopLaunch cl = plan_coop(*grid, rays->n_rays);
        if (cl.ok && (force_coop || want_coop(*grid, rays->walker))) {
            const WarpLaunch wl = plan_warp(*grid);
            const unsigned* wc = nullptr;          // gen 2 counts the rays it leaves for the fallback passes
            if (wl.ok && want_warp_gen2()) {
                wc = reinterpret_cast<unsigned*>(cx.counters + 8);
                // ordinary rays: plan + walk kernels of sphrt_warp.cuh; then the all-cones rays,
                // in place on their mask
                if constexpr (!Op::kEvents) attach_plan_cache(cx, rays);
                rc = launch_warp(cx, wl, op, stats);
            } else if (rays->n_rays < coop_split_rays()) {
                rc = launch_coop<Op, COOP_BOTH>(cx, cl, op, nullptr, cx.counters, stats);
            } else {
                rc = launch_coop<Op, COOP_ORDINARY>(cx, cl, op, nullptr, cx.counters, stats);
                // the rays that need every cone evaluated and sorted (about 1 %), in place on their mask
                if (rc == SPHRT_OK)
                    rc = launch_coop<Op, COOP_ALLCONES>(cx, cl, op, cx.mask_listed, cx.counters + 3, stats);
            }
            if (rc != SPHRT_OK) return rc;
            if constexpr (Op::kEvents) {
                return SPHRT_OK;               // raw event export: cooperative walker only
            } else {
                rc = launch_tracked_masked(cx, op, wc ? wc + 5 : nullptr);
                if (rc != SPHRT_OK) return rc;
                return launch_listed(grid, cx.kr, op, cx.mask_listed, cx.lists, cx.smem_tables, lp,
                                     cx.counters + 2, cx.st, wc ? wc + 4 : nullptr);
            }
        }
    }
    if constexpr (Op::kEvents) {
        return SPHRT_ETOOBIG;                  // the event export has no per-thread form
    } else {
        rc = launch_tracked_dense(cx, op);
        if (rc != SPHRT_OK) return rc;
        if constexpr (Op::kCoop) {
            // The sorted-list rays are a few per cent of the rays but, walked by one thread each,
            // more than half of a step on the 50^3 grids: walk them cooperatively in all-cones mode
            // first; pass B keeps what that cannot take (starts inside the grid, rays through the
            // Z axis).  SPHRT_COOP_LISTED=0 leaves them all to pass B (tests).
            const CoopLaunch cl = plan_coop(*grid, rays->n_rays);
            if (cl.ok && !no_coop_listed()) {
                rc = launch_coop<Op, COOP_ALLCONES>(cx, cl, op, cx.mask_listed, cx.counters + 3, nullptr);
                if (rc != SPHRT_OK) return rc;
            }
        }
        return launch_listed(grid, cx.kr, op, cx.mask_listed, cx.lists, cx.smem_tables, lp,
                             cx.counters + 2, cx.st);
    }
}

template <template <typename, int> class OpT, typename T>
int launch_volume_op(const sphrt_grid* grid, const sphrt_rays* rays, const T* in, T* out,
                     int64_t n_vol, int64_t rays_per_vol, bool forward,
                     void* scratch, int64_t scratch_bytes, void* stream)
{
    VolumeMap m;
    m.n_rays = rays->n_rays; m.nvox = (int64_t)grid->nr * grid->ne * grid->na;
    m.n_vol = n_vol; m.rays_per_vol = rays_per_vol; m.ne = grid->ne; m.na = grid->na;
    (void)forward;
    if (rays_per_vol > 0 || n_vol == 1) {
        OpT<T, 1> op{in, out, m, 0};
        return launch_passes(grid, rays, op, scratch, scratch_bytes, stream);
    }
    // channels share the geometry: float32 volumes are walked four at a time, float64 volumes
    // (rare; their 4-wide kernels would be a third of this library's compile time) one by one
    if constexpr (sizeof(T) == 4) {
        for (int64_t ch0 = 0; ch0 < n_vol; ch0 += 4) {
            OpT<T, 4> op{in, out, m, (int)ch0};
            const int rc = launch_passes(grid, rays, op, scratch, scratch_bytes, stream);
            if (rc != SPHRT_OK) return rc;
        }
    } else {
        for (int64_t ch0 = 0; ch0 < n_vol; ch0++) {
            OpT<T, 1> op{in, out, m, (int)ch0};
            const int rc = launch_passes(grid, rays, op, scratch, scratch_bytes, stream);
            if (rc != SPHRT_OK) return rc;
        }
    }
    return SPHRT_OK;
}

// Fused normal-equations step: plan + warp walk with two sweeps for the cooperative rays; the rays
// left to the per-thread walkers get a forward and then an adjoint launch of the two fallback
// passes (both return at once when the planner counted none).  SPHRT_EINVAL if the grid is not
// served by the second-generation warp walker (the caller then runs the two passes apart).
template <typename T>
int launch_normal(const sphrt_grid* grid, const sphrt_rays* rays, const T* density, int64_t n_vol,
                  int64_t rays_per_vol, T* out, const T* b, const T* mask, double scale, T* grad,
                  void* scratch, int64_t scratch_bytes, void* stream)
{
    if (rays->n_rays == 0) return SPHRT_OK;
    if (scratch_bytes < sphrt_scratch_bytes(grid, rays->n_rays) || scratch == nullptr) return SPHRT_ESCRATCH;
    const WarpLaunch wl = plan_warp(*grid);
    if (!wl.ok || !plan_coop(*grid, rays->n_rays).ok) return SPHRT_EINVAL;
    VolumeMap m;
    m.n_rays = rays->n_rays; m.nvox = (int64_t)grid->nr * grid->ne * grid->na;
    m.n_vol = n_vol; m.rays_per_vol = rays_per_vol; m.ne = grid->ne; m.na = grid->na;
    PassContext cx;
    cx.grid = grid;
    cx.kr = make_krays(rays);
    cx.st = (cudaStream_t)stream;
    cx.smem_tables = (size_t)table_doubles(*grid) * sizeof(double);
    cx.mask_listed = (uint32_t*)scratch;
    cx.mask_tracked = (uint32_t*)((char*)scratch + mask_bytes(rays->n_rays));
    cx.counters = (unsigned long long*)((char*)scratch + 2 * mask_bytes(rays->n_rays));
    cx.lists = (char*)scratch + 2 * mask_bytes(rays->n_rays) + 256;
    cx.n_words = (rays->n_rays + 31) / 32;
    cx.recs = reinterpret_cast<PlanRec*>((char*)scratch + scratch_bytes_before_recs(grid, rays->n_rays));
    cudaMemsetAsync(cx.counters, 0, 8 * sizeof(unsigned long long), cx.st);
    const ListPlan lp = plan_lists(*grid, rays->n_rays);
    const unsigned* wc = reinterpret_cast<unsigned*>(cx.counters + 8);
    const NormalOp<T> op{density, out, b, mask, scale, grad, m};
    attach_plan_cache(cx, rays);
    int rc = launch_warp(cx, wl, op, nullptr);
    if (rc != SPHRT_OK) return rc;
    // what the warp walker left: integrals first ...
    const ForwardOp<T, 1> fwd{density, out, m, 0};
    rc = launch_tracked_masked(cx, fwd, wc + 5);
    if (rc != SPHRT_OK) return rc;
    rc = launch_listed(grid, cx.kr, fwd, cx.mask_listed, cx.lists, cx.smem_tables, lp, cx.counters + 2, cx.st, wc + 4);
    if (rc != SPHRT_OK) return rc;
    // ... then their back-projection, weighted by the integrals just stored
    const NormalAdjointOp<T> adj{out, b, mask, scale, grad, m};
    rc = launch_tracked_masked(cx, adj, wc + 5, cx.counters + 4);
    if (rc != SPHRT_OK) return rc;
    return launch_listed(grid, cx.kr, adj, cx.mask_listed, cx.lists, cx.smem_tables, lp, cx.counters + 5, cx.st, wc + 4);
}

// ---- stored operator: apply passes ---------------------------------------------------------------
// One warp per ray streams the ray's (voxel, length) pairs with coalesced loads; the volume is
// gathered from L2.  Bound by HBM: 8 B (12 B for float64 lengths) per segment.
// Packed stored operator: 4 bytes per segment.  Along a ray the canonical segments visit voxels
// that almost always differ from their predecessor by one step along one axis, so a segment is ONE
// 32-bit word: a 3-bit step code (+-1 azimuth, +-na elevation, +-ne*na radial, flat voxel offsets)
// and its length as a float32 rounded to 29 bits (sign bit and two mantissa bits dropped: relative
// 2.4e-7).  Anything else -- the first voxel of a ray, the azimuth seam, a hole between two valid
// voxels (inner shell, gap sector), a diagonal move after an exact tie -- is an escape word that
// carries the absolute voxel offset, followed by a "stay" word with the length.  The reference
// stores 32 bytes for each of its K padded slots per ray (raytracer.py:66-72, 675-679).
constexpr uint32_t kPackLenMask = 0x1fffffffu;
#ifndef SPHRT_PACK_GROUPS
#define SPHRT_PACK_GROUPS 2
#endif
enum { PACK_STAY = 6, PACK_ESC = 7 };

__host__ __device__ __forceinline__ int pack_code(int d, int na, int nena)
{
    if (d == 1) return 0;
    if (d == -1) return 1;
    if (d == na) return 2;
    if (d == -na) return 3;
    if (d == nena) return 4;
    if (d == -nena) return 5;
    return PACK_ESC;
}
__device__ __forceinline__ int pack_delta(uint32_t code, int na, int nena)
{
    const int mag = code < 2u ? 1 : (code < 4u ? na : (code < 6u ? nena : 0));
    return (code & 1u) ? -mag : mag;
}
__device__ __forceinline__ uint32_t pack_len(double l)
{
    return (__float_as_uint((float)l) + 2u) >> 2;
}

// kStore == false: count the words of every ray; true: write them at offsets[ray]
template <bool kStore>
struct PackOp {
    const int64_t* offsets; uint32_t* words; int32_t* count; unsigned long long* total; int ne, na;
    struct Sink {
        static constexpr int kDepth = 1;
        uint32_t* out; int ne, na;
        int cur, prev; double acc; int n;        // pending segment (consecutive identical voxels merge)
        __device__ __forceinline__ void put(uint32_t w) { if constexpr (kStore) out[n] = w; n++; }
        __device__ __forceinline__ void flush()
        {
            if (cur < 0) return;
            const int code = prev < 0 ? (int)PACK_ESC : pack_code(cur - prev, na, ne * na);
            if (code == PACK_ESC) {
                put(((uint32_t)PACK_ESC << 29) | (uint32_t)cur);
                put(((uint32_t)PACK_STAY << 29) | pack_len(acc));
            } else {
                put(((uint32_t)code << 29) | pack_len(acc));
            }
            prev = cur;
        }
        __device__ __forceinline__ void segment(int r, int e, int a, double l)
        {
            const int idx = (r * ne + e) * na + a;
            if (idx == cur) { acc += l; return; }
            flush();
            cur = idx; acc = l;
        }
    };
    __device__ __forceinline__ Sink begin(int64_t ray) const
    {
        return Sink{kStore ? words + __ldg(offsets + ray) : nullptr, ne, na, -1, -1, 0.0, 0};
    }
    __device__ __forceinline__ void end(int64_t ray, Sink& s) const
    {
        s.flush();
        if constexpr (!kStore) {
            count[ray] = s.n;
            if (total != nullptr && s.n > 0) atomicAdd(total, (unsigned long long)s.n);
        }
    }
    static constexpr bool kCoop = false, kEvents = false;
    __device__ __forceinline__ void coop_end(int64_t, Sink&, int, int) const {}
};

constexpr int kStoredBlock = 256;

template <typename T, typename L, bool kAdjoint>
__global__ void __launch_bounds__(kStoredBlock)
stored_apply_kernel(const int64_t* __restrict__ offsets, const int32_t* __restrict__ vox,
                    const L* __restrict__ seg_len, int64_t n_rays, const T* __restrict__ in,
                    T* __restrict__ out, int64_t n_vol, int64_t rays_per_vol, int64_t nvox)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < n_rays; ray += warps) {
        const int64_t b = __ldg(offsets + ray), e = __ldg(offsets + ray + 1);
        if (b == e && kAdjoint) continue;
        const int64_t v0 = rays_per_vol > 0 ? (ray / rays_per_vol) % n_vol : 0;
        const int64_t nv = rays_per_vol > 0 ? 1 : n_vol;          // channels share the list
        for (int64_t v = 0; v < nv; v++) {
            if constexpr (kAdjoint) {
                // in = y (n_vol, n_rays) or (n_rays); out = gradient volumes
                const double w = (double)__ldg(in + (rays_per_vol > 0 ? ray : v * n_rays + ray));
                if (w == 0.0) continue;
                T* vol = out + (v0 + v) * nvox;
                for (int64_t k = b + lane; k < e; k += 32)
                    atomicAdd(vol + __ldcs(vox + k), (T)(w * (double)__ldcs(seg_len + k)));
            } else {
                const T* vol = in + (v0 + v) * nvox;
                double acc = 0.0;
                for (int64_t k = b + lane; k < e; k += 32)
                    acc = fma_rn((double)__ldg(vol + __ldcs(vox + k)), (double)__ldcs(seg_len + k), acc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) out[rays_per_vol > 0 ? ray : v * n_rays + ray] = (T)acc;
            }
        }
    }
}

template <typename T, typename L, bool kAdjoint>
int launch_stored(const int64_t* offsets, const int32_t* vox, const void* seg_len, int64_t n_rays,
                  const void* in, void* out, int64_t n_vol, int64_t rays_per_vol, int64_t nvox,
                  void* stream)
{
    if (n_rays == 0) return SPHRT_OK;
    int64_t blocks = (int64_t)num_sms() * (2048 / kStoredBlock);
    const int64_t want = (n_rays * 32 + kStoredBlock - 1) / kStoredBlock;
    if (blocks > want) blocks = want;
    stored_apply_kernel<T, L, kAdjoint><<<(unsigned)blocks, kStoredBlock, 0, (cudaStream_t)stream>>>(
        offsets, vox, (const L*)seg_len, n_rays, (const T*)in, (T*)out, n_vol, rays_per_vol, nvox);
    return cuda_status();
}

// One warp per ray streams its words (coalesced, read once: ld.global.cs), 32 at a time: the voxel
// offsets are a segmented prefix sum of the step codes (an escape restarts the sum), carried from
// one group of 32 words to the next; gathers / reductions as in stored_apply_kernel.
template <typename T, bool kAdjoint>
__global__ void __launch_bounds__(kStoredBlock)
packed_apply_kernel(const int64_t* __restrict__ offsets, const uint32_t* __restrict__ words,
                    int64_t n_rays, int na, int nena, const T* __restrict__ in, T* __restrict__ out,
                    int64_t n_vol, int64_t rays_per_vol, int64_t nvox)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // (the warp's ray index is broadcast from lane 0: that is what tells the compiler that the loops
    // below are warp-uniform -- without it every shuffle of the scan is wrapped in a convergence
    // barrier and the kernel, L2-bound by design, becomes instruction-bound)
    const int64_t first = __shfl_sync(0xffffffffu, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, 0);
    const bool l1 = lane >= 1, l2 = lane >= 2, l4 = lane >= 4, l8 = lane >= 8, l16 = lane >= 16;
    for (int64_t ray = first; ray < n_rays; ray += warps) {
        const int64_t b = __ldg(offsets + ray);
        const int cnt = (int)(__ldg(offsets + ray + 1) - b);
        if (cnt == 0 && kAdjoint) continue;
        const uint32_t* __restrict__ wp = words + b;
        const int64_t v0 = rays_per_vol > 0 ? (ray / rays_per_vol) % n_vol : 0;
        const int nv = rays_per_vol > 0 ? 1 : (int)n_vol;         // channels share the list
        for (int v = 0; v < nv; v++) {
            float w = 0.0f;
            // per-lane partial sums in the volume's precision (ten-odd terms per lane, as in the
            // walker's forward sink), added over the warp in float64
            T acc = T(0);
            const T* vin = nullptr; T* vout = nullptr;
            double wd = 0.0;
            if constexpr (kAdjoint) {
                wd = (double)__ldg(in + (rays_per_vol > 0 ? ray : (int64_t)v * n_rays + ray));
                if (wd == 0.0) continue;
                w = (float)wd;
                vout = out + (v0 + v) * nvox;
            } else {
                vin = in + (v0 + v) * nvox;
            }
            // The scan carries ONE integer per lane: sum of steps since the last escape, plus 2^30
            // once an escape is part of it (voxel offsets are below 2^29, so the sum of steps stays
            // inside (-2^29, 2^29) and the biased value inside [2^30, 2^30 + 2^29)).  The words of
            // the next groups are loaded before the current groups' gathers are issued.
            constexpr int kBias = 1 << 30, kHalf = 1 << 29;
            constexpr int G = SPHRT_PACK_GROUPS;               // groups of 32 words in flight per warp
            const uint32_t pad = (uint32_t)PACK_STAY << 29;
            int carry = 0;
            uint32_t next[G];
#pragma unroll
            for (int g = 0; g < G; g++) { const int k = 32 * g + lane; next[g] = k < cnt ? __ldcs(wp + k) : pad; }
            for (int k0 = 0; k0 < cnt; k0 += 32 * G) {
                int x[G]; float len[G];
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const uint32_t word = next[g];
                    const int kn = k0 + 32 * (G + g) + lane;
                    next[g] = kn < cnt ? __ldcs(wp + kn) : pad;
                    const uint32_t code = word >> 29, payload = word & kPackLenMask;
                    const bool esc = code == (uint32_t)PACK_ESC;
                    x[g] = esc ? (int)payload + kBias : pack_delta(code, na, nena);
                    len[g] = esc ? 0.0f : __uint_as_float(payload << 2);
                }
#define SPHRT_SCAN_STEP(O, P)                                                                      \
                _Pragma("unroll")                                                                  \
                for (int g = 0; g < G; g++) {                                                      \
                    const int px = __shfl_up_sync(0xffffffffu, x[g], O);                           \
                    if (P && x[g] < kHalf) x[g] += px;                                             \
                }
                SPHRT_SCAN_STEP(1, l1) SPHRT_SCAN_STEP(2, l2) SPHRT_SCAN_STEP(4, l4)
                SPHRT_SCAN_STEP(8, l8) SPHRT_SCAN_STEP(16, l16)
#undef SPHRT_SCAN_STEP
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const int idx = x[g] >= kHalf ? x[g] - kBias : carry + x[g];
                    carry = __shfl_sync(0xffffffffu, idx, 31);
                    if (len[g] != 0.0f) {
                        if constexpr (kAdjoint) {
                            if constexpr (sizeof(T) == 4) atomicAdd(vout + idx, (T)(w * len[g]));
                            else atomicAdd(vout + idx, (T)(wd * (double)len[g]));
                        } else {
                            acc = fmaf_or_fma(__ldg(vin + idx), (T)len[g], acc);
                        }
                    }
                }
            }
            if constexpr (!kAdjoint) {
                double a = (double)acc;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) out[rays_per_vol > 0 ? ray : (int64_t)v * n_rays + ray] = (T)a;
            }
        }
    }
}

template <typename T, bool kAdjoint>
int launch_packed(const int64_t* offsets, const uint32_t* words, int64_t n_rays, int ne, int na,
                  const void* in, void* out, int64_t n_vol, int64_t rays_per_vol, int64_t nvox, void* stream)
{
    if (n_rays == 0) return SPHRT_OK;
    int64_t blocks = (int64_t)num_sms() * (2048 / kStoredBlock);
    const int64_t want = (n_rays * 32 + kStoredBlock - 1) / kStoredBlock;
    if (blocks > want) blocks = want;
    packed_apply_kernel<T, kAdjoint><<<(unsigned)blocks, kStoredBlock, 0, (cudaStream_t)stream>>>(
        offsets, words, n_rays, na, ne * na, (const T*)in, (T*)out, n_vol, rays_per_vol, nvox);
    return cuda_status();
}

template <bool kAdjoint>
int dispatch_packed(const int64_t* offsets, const uint32_t* words, int64_t n_rays, int ne, int na,
                    const void* in, int dtype, void* out, int64_t n_vol, int64_t rays_per_vol,
                    int64_t nvox, void* stream)
{
    if (n_rays < 0 || n_vol < 1 || rays_per_vol < 0 || nvox < 1 || ne < 1 || na < 1) return SPHRT_EINVAL;
    if (nvox > (int64_t)kPackLenMask) return SPHRT_ETOOBIG;              // 29-bit voxel offsets
    if (n_rays > 0 && (!offsets || !words || !in || !out)) return SPHRT_EINVAL;
    if (dtype == SPHRT_F32)
        return launch_packed<float, kAdjoint>(offsets, words, n_rays, ne, na, in, out, n_vol, rays_per_vol, nvox, stream);
    if (dtype == SPHRT_F64)
        return launch_packed<double, kAdjoint>(offsets, words, n_rays, ne, na, in, out, n_vol, rays_per_vol, nvox, stream);
    return SPHRT_EINVAL;
}

template <bool kAdjoint>
int dispatch_stored(const int64_t* offsets, const int32_t* vox, const void* seg_len, int len_dtype,
                    int64_t n_rays, const void* in, int dtype, void* out, int64_t n_vol,
                    int64_t rays_per_vol, int64_t nvox, void* stream)
{
    if (n_rays < 0 || n_vol < 1 || rays_per_vol < 0 || nvox < 1) return SPHRT_EINVAL;
    if (n_rays > 0 && (!offsets || !in || !out)) return SPHRT_EINVAL;
    if (dtype == SPHRT_F32 && len_dtype == SPHRT_F32)
        return launch_stored<float, float, kAdjoint>(offsets, vox, seg_len, n_rays, in, out, n_vol, rays_per_vol, nvox, stream);
    if (dtype == SPHRT_F32 && len_dtype == SPHRT_F64)
        return launch_stored<float, double, kAdjoint>(offsets, vox, seg_len, n_rays, in, out, n_vol, rays_per_vol, nvox, stream);
    if (dtype == SPHRT_F64 && len_dtype == SPHRT_F64)
        return launch_stored<double, double, kAdjoint>(offsets, vox, seg_len, n_rays, in, out, n_vol, rays_per_vol, nvox, stream);
    if (dtype == SPHRT_F64 && len_dtype == SPHRT_F32)
        return launch_stored<double, float, kAdjoint>(offsets, vox, seg_len, n_rays, in, out, n_vol, rays_per_vol, nvox, stream);
    return SPHRT_EINVAL;
}

}  // namespace

