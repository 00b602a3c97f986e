// cp_api.cu — kernels and the C ABI of libcpgpu.so (see include/cp_gpu.h).
//
// Built for sm_100a only, with -fmad=false (no contraction: see cp_device.cuh).  There is no CPU
// path in this library: every compute entry point launches CUDA kernels or fails.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "cp_device.cuh"
#include "cp_internal.h"
#include "cp_launch.h"

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

cp_status cp_fail(cp_status st, const char *msg) {
    g_last_error = msg ? msg : "";
    return st;
}

static cp_status cuda_fail(cudaError_t e, const char *what) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    return cp_fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? CP_ERR_NO_DEVICE : CP_ERR_CUDA, buf);
}

#define CP_CUDA(call)                                         \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);   \
    } while (0)

// .max() over replicas ordered by score (main.rs:253, packed.rs:49-68).  rayon's max() is
// reduce_with(Ord::max) in index order and Ord::max returns its second argument on Equal, so of several
// replicas with the maximal score the one with the HIGHEST index wins.
struct BestKey {
    double score;
    unsigned long long index;
};
__device__ __forceinline__ bool better(const BestKey &a, const BestKey &b) {
    // is a better than b?  NaN (None) never wins
    if (!(a.score == a.score)) return false;
    if (!(b.score == b.score)) return true;
    return a.score > b.score || (a.score == b.score && a.index > b.index);
}
__global__ void __launch_bounds__(256)
best_kernel(const cp_replica_result *res, const BestKey *partial_in, unsigned long long n, BestKey *out,
            cp_best *final_best, unsigned long long replica_begin) {
    __shared__ BestKey sh[256];
    BestKey best{__longlong_as_double(0x7FF8000000000000ll), ~0ull};
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        BestKey c = partial_in ? partial_in[i] : BestKey{res[i].score, i};
        if (better(c, best)) best = c;
    }
    sh[threadIdx.x] = best;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s && better(sh[threadIdx.x + s], sh[threadIdx.x])) sh[threadIdx.x] = sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[blockIdx.x] = sh[0];
        if (final_best) {  // last pass: materialise the winner as a cp_best in device memory
            cp_best b;
            if (sh[0].index != ~0ull) {
                b.result = res[sh[0].index];
                b.replica = replica_begin + sh[0].index;
            } else {
                for (int i = 0; i < CP_MAX_DOF; ++i) b.result.dof[i] = 0.;
                b.result.score = __longlong_as_double(0x7FF8000000000000ll);
                b.result.rejections = b.result.moves = 0;
                b.replica = ~0ull;
            }
            *final_best = b;
        }
    }
}

// FP pipe peaks for the roofline denominator: 8 independent FMA chains per thread, no memory traffic
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        if constexpr (sizeof(T) == 8) {
            x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
        } else {
            x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
            x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
        }
    }
    T r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == (T)123456789) out[blockIdx.x * blockDim.x + threadIdx.x] = r;  // never true; keeps the chains live
}

// ------------------------------------------------------------------------------------------ context
struct cp_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[7] = {};  // h2d 0-1, mc 2-3, reduce end 4, d2h 5-6
    int lanes = 0;  // 0 = automatic
    // prepared job
    bool prepared = false, launched = false;
    KParams params{};
    cp_replica_result *d_out = nullptr;
    double *d_init = nullptr;
    BestKey *d_partial = nullptr;  // [kPartials + 1]
    cp_best *d_best = nullptr;
    int *d_error = nullptr;
    size_t cap_out = 0, cap_init = 0;
    cp_timing timing{};
    bool timed_h2d = false;
};
constexpr int kPartials = 592;  // 148 SMs x 4

static double powi6_host(double x) { double x2 = x * x, x4 = x2 * x2; return x2 * x4; }
static double powi12_host(double x) { double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4; return x4 * x8; }

static cp_status make_kproblem(const cp_problem *p, KProblem *k) {
    if (!p) return cp_fail(CP_ERR_INVALID_ARG, "null problem");
    if (p->kind < CP_POLYGON || p->kind > CP_LJ) return cp_fail(CP_ERR_INVALID_ARG, "problem.kind out of range");
    if (p->n_items < 1 || p->n_items > CP_MAX_ITEMS) return cp_fail(CP_ERR_INVALID_ARG, "problem.n_items out of range");
    if (p->n_syms < 1 || p->n_syms > CP_MAX_SYMS) return cp_fail(CP_ERR_INVALID_ARG, "problem.n_syms out of range");
    if (p->family < CP_MONOCLINIC || p->family > CP_TETRAGONAL) return cp_fail(CP_ERR_INVALID_ARG, "problem.family out of range");
    std::memset(k, 0, sizeof *k);
    k->kind = p->kind;
    k->n_items = p->n_items;
    k->family = p->family;
    k->n_syms = p->n_syms;
    std::memcpy(k->items, p->items, sizeof k->items);
    std::memcpy(k->syms, p->syms, sizeof k->syms);
    std::memcpy(k->dof, p->dof, sizeof k->dof);
    k->area = p->area;
    k->radius = p->enclosing_radius;
    k->lj_reach = INFINITY;
    if (p->kind == CP_LJ) {
        double cmax = 0., rmax = 0.;
        for (int i = 0; i < p->n_items; ++i) {
            const double x = p->items[i][4];
            cmax = std::isnan(x) ? INFINITY : std::fmax(cmax, x);
            rmax = std::fmax(rmax, std::sqrt(p->items[i][0] * p->items[i][0] + p->items[i][1] * p->items[i][1]));
        }
        k->lj_reach = cmax + 2. * rmax;
    }
    // single-precision copies for the certificates (cp_filter.h)
    for (int s = 0; s < p->n_syms; ++s) {
        k->symsf[s][0] = (float)p->syms[s][0];
        k->symsf[s][1] = (float)p->syms[s][1];
        k->symsf[s][2] = (float)p->syms[s][3];
        k->symsf[s][3] = (float)p->syms[s][4];
    }
    // Signed images: the rotation part of every operator is that of operator 0 with the signs of its
    // rows flipped (identity, two-fold axis, mirrors: all seven plane groups of the reference).  Then
    // the rotated shape points of image s are those of image 0 with the same sign flips, bit for bit
    // (cp_thread.cuh: t_place evaluates fmaf(+-x, c, +-y * s)), and one set serves all images.
    k->signed_images = 1;
    for (int s = 0; s < p->n_syms; ++s) {
        bool found = false;
        for (int a = 0; a < 4 && !found; ++a) {
            const float ax = (a & 1) ? -1.0f : 1.0f, ay = (a & 2) ? -1.0f : 1.0f;
            if (k->symsf[s][0] == ax * k->symsf[0][0] && k->symsf[s][1] == ax * k->symsf[0][1] &&
                k->symsf[s][2] == ay * k->symsf[0][2] && k->symsf[s][3] == ay * k->symsf[0][3]) {
                k->sgnf[s][0] = ax;
                k->sgnf[s][1] = ay;
                found = true;
            }
        }
        if (!found) k->signed_images = 0;
    }
    if (!k->signed_images)
        for (int s = 0; s < CP_MAX_SYMS; ++s) k->sgnf[s][0] = k->sgnf[s][1] = 1.0f;
    k->radf = (float)p->enclosing_radius;
    k->chain = 0;
    if (p->kind == CP_POLYGON) {
        // a closed chain: every segment ends where the next one starts (LineShape::from_radial computes
        // the two from different expressions, so they agree to rounding only: line_shape.rs:128-146)
        bool chain = p->n_items >= 3;
        const double tol = 1e-12 * p->enclosing_radius;
        for (int i = 0; i < p->n_items && chain; ++i) {
            const double *e = p->items[i], *s = p->items[(i + 1) % p->n_items];
            chain = std::fabs(e[2] - s[0]) <= tol && std::fabs(e[3] - s[1]) <= tol;
        }
        k->chain = chain ? 1 : 0;
        if (cp_thread_route_ni(*k) > 0) {
            for (int i = 0; i < p->n_items; ++i) {
                k->ptsf[i][0] = (float)p->items[i][0];
                k->ptsf[i][1] = (float)p->items[i][1];
            }
        } else {
            for (int i = 0; i < p->n_items; ++i) {
                k->ptsf[2 * i][0] = (float)p->items[i][0];
                k->ptsf[2 * i][1] = (float)p->items[i][1];
                k->ptsf[2 * i + 1][0] = (float)p->items[i][2];
                k->ptsf[2 * i + 1][1] = (float)p->items[i][3];
            }
        }
    } else if (p->kind == CP_DISCS) {
        for (int i = 0; i < p->n_items; ++i) {
            k->ptsf[i][0] = (float)p->items[i][0];
            k->ptsf[i][1] = (float)p->items[i][1];
            k->discrf[i] = (float)p->items[i][2];
        }
    }
    if (p->kind == CP_LJ)
        for (int i = 0; i < p->n_items; ++i) {
            const double sigma = p->items[i][2], eps = p->items[i][3], x = p->items[i][4];
            // lj2.rs:50-51: 4. * epsilon * ((sigma / x).powi(12) - (sigma / x).powi(6))
            k->lj_shift[i] = std::isnan(x) ? 0. : 4. * eps * (powi12_host(sigma / x) - powi6_host(sigma / x));
        }
    return CP_OK;
}

static int auto_lanes(const KProblem &k) {
    // Hard shapes: one thread per replica whenever the per-thread geometry fits in shared memory
    // (measured 4-5x faster than a warp per replica on B200, profiles/README.md).  LJ and very
    // large shapes: lane groups sized to the parallel work of one move.
    // at least two 64-thread blocks must fit an SM, else the warp-per-replica kernel is faster
    // (12-gon in p2gg: 117 KB per block; measured with tools/lane_sweep.py)
    if (cp_thread_kernel_fits(k) && cp_thread_shared_bytes(k) <= 110 * 1024) return 1;
    const int pair_tests = k.n_syms * (k.n_syms - 1) / 2 * k.n_items * k.n_items;
    const int cands = k.n_syms * k.n_syms * (k.kind == CP_LJ ? 48 : 8);
    const int work = pair_tests > cands ? pair_tests : cands;
    if (work >= 32) return 32;
    if (work >= 16) return 16;
    return 8;
}

static int effective_lanes(const cp_ctx *ctx, const KProblem &k) {
    int lanes = ctx->lanes ? ctx->lanes : auto_lanes(k);
    if (lanes == 1 && !cp_thread_kernel_fits(k)) lanes = 8;  // LJ / very large shapes: group kernel
    return lanes;
}

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char *cp_last_error(void) { return g_last_error.c_str(); }
const char *cp_version(void) { return "crystal-packing_b200 0.1 (sm_100a)"; }

cp_status cp_device_count(int *out) {
    if (!out) return cp_fail(CP_ERR_INVALID_ARG, "cp_device_count: null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *out = 0;
        return cuda_fail(e, "cudaGetDeviceCount");
    }
    *out = n;
    return CP_OK;
}

static cp_status init_members(cp_ctx *ctx, int device) {
    ctx->device = device;
    cudaDeviceProp prop;
    CP_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    CP_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    for (auto &ev : ctx->ev) CP_CUDA(cudaEventCreate(&ev));
    CP_CUDA(cudaMalloc(&ctx->d_partial, sizeof(BestKey) * (kPartials + 1)));
    CP_CUDA(cudaMalloc(&ctx->d_error, sizeof(int)));
    CP_CUDA(cudaMalloc(&ctx->d_best, sizeof(cp_best)));
    return CP_OK;
}

cp_status cp_init(int device, cp_ctx **out) {
    if (!out) return cp_fail(CP_ERR_INVALID_ARG, "cp_init: null output");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    if (n == 0) return cp_fail(CP_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    if (device < 0 || device >= n) return cp_fail(CP_ERR_INVALID_ARG, "cp_init: device index out of range");
    CP_CUDA(cudaSetDevice(device));
    cp_ctx *ctx = new (std::nothrow) cp_ctx();
    if (!ctx) return cp_fail(CP_ERR_CUDA, "out of host memory");
    const cp_status st = init_members(ctx, device);
    if (st != CP_OK) {  // a half-built context is released here: the caller never sees it
        const std::string why = g_last_error;
        cp_destroy(ctx);
        (void)cudaGetLastError();
        g_last_error = why;
        return st;
    }
    *out = ctx;
    return CP_OK;
}

void cp_destroy(cp_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_out);
    cudaFree(ctx->d_init);
    cudaFree(ctx->d_partial);
    cudaFree(ctx->d_error);
    cudaFree(ctx->d_best);
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

cp_status cp_set_stream(cp_ctx *ctx, void *cuda_stream) {
    if (!ctx) return cp_fail(CP_ERR_INVALID_ARG, "cp_set_stream: null ctx");
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return CP_OK;
}

cp_status cp_set_lanes(cp_ctx *ctx, int lanes) {
    if (!ctx) return cp_fail(CP_ERR_INVALID_ARG, "cp_set_lanes: null ctx");
    if (lanes != 0 && lanes != 1 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
        return cp_fail(CP_ERR_INVALID_ARG, "cp_set_lanes: lanes must be 0, 1, 4, 8, 16 or 32");
    ctx->lanes = lanes;
    return CP_OK;
}

cp_status cp_prepare(cp_ctx *ctx, const cp_problem *problem, const cp_schedule *stages, int n_stages,
                     int rng_mode, uint64_t seed_base, uint64_t replica_begin, uint64_t n_replicas,
                     const double *init_dof) {
    if (!ctx) return cp_fail(CP_ERR_INVALID_ARG, "cp_prepare: null ctx");
    if (!stages || n_stages < 1 || n_stages > CP_MAX_STAGES)
        return cp_fail(CP_ERR_INVALID_ARG, "cp_prepare: n_stages must be 1..CP_MAX_STAGES");
    if (rng_mode != CP_RNG_PCG64MCG && rng_mode != CP_RNG_PHILOX)
        return cp_fail(CP_ERR_INVALID_ARG, "cp_prepare: unknown rng_mode");
    if (n_replicas == 0) return cp_fail(CP_ERR_INVALID_ARG, "cp_prepare: n_replicas must be > 0");
    ctx->prepared = false;
    KParams &K = ctx->params;
    cp_status st = make_kproblem(problem, &K.prob);
    if (st != CP_OK) return st;
    for (int s = 0; s < n_stages; ++s) {
        K.stages[s].kt_start = stages[s].kt_start;
        K.stages[s].kt_ratio = stages[s].kt_ratio;
        K.stages[s].max_step = stages[s].max_step_size;
        K.stages[s].steps = stages[s].steps;
        K.stages[s].inner = stages[s].inner_steps;
        K.stages[s].convergence = stages[s].convergence;
    }
    K.n_stages = n_stages;
    K.rng_mode = rng_mode;
    K.seed_base = seed_base;
    K.replica_begin = replica_begin;
    K.n_replicas = n_replicas;
    K.draw_cap = 3;
    K.warp_fill = 0;  // 0: chosen at the launch from the occupancy (cp_thread_inst.cuh)
    if (const char *e = std::getenv("CP_WARP_FILL")) K.warp_fill = std::atoi(e);  // tuning knob
    if (const char *e = std::getenv("CP_DRAW_CAP")) {  // tuning knob, see cp_thread.cuh
        const int c = std::atoi(e);
        if (c >= 1 && c <= 64) K.draw_cap = c;
    }
    CP_CUDA(cudaSetDevice(ctx->device));
    if (ctx->cap_out < n_replicas) {
        cudaFree(ctx->d_out);
        ctx->d_out = nullptr;
        ctx->cap_out = 0;
        CP_CUDA(cudaMalloc(&ctx->d_out, sizeof(cp_replica_result) * n_replicas));
        ctx->cap_out = n_replicas;
    }
    K.out = ctx->d_out;
    K.error_flag = ctx->d_error;
    K.init_dof = nullptr;
    ctx->timed_h2d = false;
    if (init_dof) {
        if (ctx->cap_init < n_replicas) {
            cudaFree(ctx->d_init);
            ctx->d_init = nullptr;
            ctx->cap_init = 0;
            CP_CUDA(cudaMalloc(&ctx->d_init, sizeof(double) * CP_MAX_DOF * n_replicas));
            ctx->cap_init = n_replicas;
        }
        CP_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
        CP_CUDA(cudaMemcpyAsync(ctx->d_init, init_dof, sizeof(double) * CP_MAX_DOF * n_replicas,
                                cudaMemcpyHostToDevice, ctx->stream));
        CP_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
        K.init_dof = ctx->d_init;
        ctx->timed_h2d = true;
    }
    ctx->prepared = true;
    ctx->launched = false;
    return CP_OK;
}

cp_status cp_launch(cp_ctx *ctx) {
    if (!ctx) return cp_fail(CP_ERR_INVALID_ARG, "cp_launch: null ctx");
    if (!ctx->prepared) return cp_fail(CP_ERR_NOT_PREPARED, "cp_launch: call cp_prepare first");
    CP_CUDA(cudaSetDevice(ctx->device));
    const KParams &K = ctx->params;
    const int lanes = effective_lanes(ctx, K.prob);
    CP_CUDA(cudaMemsetAsync(ctx->d_error, 0, sizeof(int), ctx->stream));
    CP_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    CP_CUDA(cp_launch_mc(K, lanes, ctx->stream));
    CP_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    best_kernel<<<kPartials, 256, 0, ctx->stream>>>(ctx->d_out, nullptr, K.n_replicas, ctx->d_partial, nullptr, 0);
    CP_CUDA(cudaGetLastError());
    best_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_out, ctx->d_partial, kPartials, ctx->d_partial + kPartials,
                                            ctx->d_best, K.replica_begin);
    CP_CUDA(cudaGetLastError());
    CP_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
    ctx->timing.launches = 3;
    ctx->timing.lanes_per_replica = (uint32_t)lanes;
    ctx->launched = true;
    return CP_OK;
}

cp_status cp_synchronize(cp_ctx *ctx) {
    if (!ctx) return cp_fail(CP_ERR_INVALID_ARG, "cp_synchronize: null ctx");
    CP_CUDA(cudaStreamSynchronize(ctx->stream));
    return CP_OK;
}

cp_status cp_fetch(cp_ctx *ctx, cp_best *out_best, cp_replica_result *out_all) {
    if (!ctx) return cp_fail(CP_ERR_INVALID_ARG, "cp_fetch: null ctx");
    if (!ctx->launched) return cp_fail(CP_ERR_NOT_PREPARED, "cp_fetch: nothing was launched");
    CP_CUDA(cudaSetDevice(ctx->device));
    const KParams &K = ctx->params;
    cp_best best{};
    int err = 0;
    CP_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
    if (out_all)
        CP_CUDA(cudaMemcpyAsync(out_all, ctx->d_out, sizeof(cp_replica_result) * K.n_replicas,
                                cudaMemcpyDeviceToHost, ctx->stream));
    CP_CUDA(cudaMemcpyAsync(&best, ctx->d_best, sizeof best, cudaMemcpyDeviceToHost, ctx->stream));
    CP_CUDA(cudaMemcpyAsync(&err, ctx->d_error, sizeof err, cudaMemcpyDeviceToHost, ctx->stream));
    CP_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
    CP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_best) *out_best = best;
    float ms = 0.f;
    ctx->timing.h2d_ms = 0.f;
    if (ctx->timed_h2d && cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->timing.h2d_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]) == cudaSuccess) ctx->timing.kernel_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]) == cudaSuccess) ctx->timing.reduce_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]) == cudaSuccess) ctx->timing.d2h_ms = ms;
    if (err) return cp_fail(CP_ERR_INVALID_STATE, "Invalid configuration passed to function, exiting.");
    return CP_OK;
}

cp_status cp_device_buffers(cp_ctx *ctx, void **d_results, void **d_best) {
    if (!ctx) return cp_fail(CP_ERR_INVALID_ARG, "cp_device_buffers: null ctx");
    if (!ctx->prepared) return cp_fail(CP_ERR_NOT_PREPARED, "cp_device_buffers: call cp_prepare first");
    if (d_results) *d_results = ctx->d_out;
    if (d_best) *d_best = ctx->d_best;
    return CP_OK;
}

cp_status cp_last_timing(cp_ctx *ctx, cp_timing *out) {
    if (!ctx || !out) return cp_fail(CP_ERR_INVALID_ARG, "cp_last_timing: null argument");
    if (ctx->launched) {  // valid once the stream has passed the events (after a synchronise)
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]) == cudaSuccess) ctx->timing.kernel_ms = ms;
        if (cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]) == cudaSuccess) ctx->timing.reduce_ms = ms;
        (void)cudaGetLastError();
    }
    *out = ctx->timing;
    return CP_OK;
}

cp_status cp_run(cp_ctx *ctx, const cp_problem *problem, const cp_schedule *stages, int n_stages,
                 int rng_mode, uint64_t seed_base, uint64_t replica_begin, uint64_t n_replicas,
                 const double *init_dof, cp_best *out_best, cp_replica_result *out_all) {
    cp_status st = cp_prepare(ctx, problem, stages, n_stages, rng_mode, seed_base, replica_begin, n_replicas, init_dof);
    if (st == CP_OK) st = cp_launch(ctx);
    if (st == CP_OK) st = cp_fetch(ctx, out_best, out_all);
    return st;
}

// ---- several GPUs behind one handle -------------------------------------------------------------
// analyse_state's replica loop (main.rs:221-253) over a list of devices: the range of replicas is cut
// into contiguous shards, one per device, every device runs its shard on its own stream at the same
// time (prepare + launch are asynchronous; one host thread drives them all), and the per-device
// maxima are reduced on the host.  The seed of a replica is its global index, so the result does not
// depend on the number of devices.
// The reduction is n x 80 bytes that cp_fetch has already copied to the host (the caller wants
// out_best in host memory anyway), so a device-side collective would only add a launch and a second
// synchronisation to every call; processes that keep their results on the device (one process per
// GPU, bench.py) all-gather the cp_best of cp_device_buffers over NCCL instead.
struct cp_multi {
    std::vector<cp_ctx *> ctx;
    std::vector<uint64_t> begin, count;  // shards of the last run (global replica indices)
};

static bool host_better(const cp_best &a, const cp_best &b) {  // the rule of better() above
    if (!(a.result.score == a.result.score)) return false;
    if (!(b.result.score == b.result.score)) return true;
    return a.result.score > b.result.score || (a.result.score == b.result.score && a.replica > b.replica);
}

cp_status cp_init_multi(const int *device_ids, int n_devices, cp_multi **out) {
    if (!out) return cp_fail(CP_ERR_INVALID_ARG, "cp_init_multi: null output");
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > 64)
        return cp_fail(CP_ERR_INVALID_ARG, "cp_init_multi: need 1..64 device ids");
    cp_multi *m = new (std::nothrow) cp_multi();
    if (!m) return cp_fail(CP_ERR_CUDA, "out of host memory");
    for (int k = 0; k < n_devices; ++k) {
        cp_ctx *c = nullptr;
        const cp_status st = cp_init(device_ids[k], &c);
        if (st != CP_OK) {
            const std::string why = g_last_error;
            cp_multi_destroy(m);
            g_last_error = why;
            return st;
        }
        m->ctx.push_back(c);
    }
    m->begin.assign(n_devices, 0);
    m->count.assign(n_devices, 0);
    *out = m;
    return CP_OK;
}

void cp_multi_destroy(cp_multi *m) {
    if (!m) return;
    for (cp_ctx *c : m->ctx) cp_destroy(c);
    delete m;
}

cp_status cp_multi_device_count(const cp_multi *m, int *out) {
    if (!m || !out) return cp_fail(CP_ERR_INVALID_ARG, "cp_multi_device_count: null argument");
    *out = (int)m->ctx.size();
    return CP_OK;
}

cp_status cp_multi_context(cp_multi *m, int k, cp_ctx **out) {
    if (!m || !out || k < 0 || k >= (int)m->ctx.size())
        return cp_fail(CP_ERR_INVALID_ARG, "cp_multi_context: index out of range");
    *out = m->ctx[k];
    return CP_OK;
}

cp_status cp_multi_shard(const cp_multi *m, int k, uint64_t *replica_begin, uint64_t *n_replicas) {
    if (!m || k < 0 || k >= (int)m->ctx.size()) return cp_fail(CP_ERR_INVALID_ARG, "cp_multi_shard: index out of range");
    if (replica_begin) *replica_begin = m->begin[k];
    if (n_replicas) *n_replicas = m->count[k];
    return CP_OK;
}

cp_status cp_multi_run(cp_multi *m, const cp_problem *problem, const cp_schedule *stages, int n_stages, int rng_mode,
                       uint64_t seed_base, uint64_t replica_begin, uint64_t n_replicas, const double *init_dof,
                       cp_best *out_best, cp_replica_result *out_all) {
    if (!m) return cp_fail(CP_ERR_INVALID_ARG, "cp_multi_run: null handle");
    if (n_replicas == 0) return cp_fail(CP_ERR_INVALID_ARG, "cp_multi_run: n_replicas must be > 0");
    const uint64_t nd = m->ctx.size();
    const uint64_t base = n_replicas / nd, extra = n_replicas % nd;
    uint64_t at = 0;
    for (uint64_t k = 0; k < nd; ++k) {  // contiguous shards, the first `extra` of them one replica longer
        m->begin[k] = replica_begin + at;
        m->count[k] = base + (k < extra ? 1 : 0);
        at += m->count[k];
    }
    // every device is started before any is waited for
    for (uint64_t k = 0; k < nd; ++k) {
        if (m->count[k] == 0) continue;
        const uint64_t off = m->begin[k] - replica_begin;
        cp_status st = cp_prepare(m->ctx[k], problem, stages, n_stages, rng_mode, seed_base, m->begin[k], m->count[k],
                                  init_dof ? init_dof + off * CP_MAX_DOF : nullptr);
        if (st == CP_OK) st = cp_launch(m->ctx[k]);
        if (st != CP_OK) {
            const std::string why = g_last_error;
            for (uint64_t j = 0; j < k; ++j)
                if (m->count[j]) cp_synchronize(m->ctx[j]);
            g_last_error = why;
            return st;
        }
    }
    cp_status first_error = CP_OK;
    std::string why;
    cp_best best{};
    best.result.score = std::nan("");
    best.replica = ~0ull;
    for (uint64_t k = 0; k < nd; ++k) {
        if (m->count[k] == 0) continue;
        cp_best b{};
        const uint64_t off = m->begin[k] - replica_begin;
        const cp_status st = cp_fetch(m->ctx[k], &b, out_all ? out_all + off : nullptr);
        if (st != CP_OK && first_error == CP_OK) {
            first_error = st;
            why = g_last_error;
        }
        if (st == CP_OK && host_better(b, best)) best = b;
    }
    if (out_best) *out_best = best;
    if (first_error != CP_OK) g_last_error = why;
    return first_error;
}

cp_status cp_measure_peaks(cp_ctx *ctx, double *fp64_tflops, double *fp32_tflops) {
    if (!ctx || !fp64_tflops || !fp32_tflops) return cp_fail(CP_ERR_INVALID_ARG, "cp_measure_peaks: null argument");
    CP_CUDA(cudaSetDevice(ctx->device));
    void *buf = nullptr;
    CP_CUDA(cudaMalloc(&buf, 8));
    const int blocks = ctx->sm_count * 8, iters = 1 << 14;
    cudaEvent_t a = ctx->ev[0], b = ctx->ev[1];
    double best64 = 0., best32 = 0.;
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 4 && e == cudaSuccess; ++rep) {
        float ms = 0.f;
        cudaEventRecord(a, ctx->stream);
        fma_peak_kernel<double><<<blocks, 256, 0, ctx->stream>>>((double *)buf, iters, 0.999999, 1e-9);
        cudaEventRecord(b, ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) break;
        cudaEventElapsedTime(&ms, a, b);
        double tf = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best64) best64 = tf;
        cudaEventRecord(a, ctx->stream);
        fma_peak_kernel<float><<<blocks, 256, 0, ctx->stream>>>((float *)buf, iters * 4, 0.999999f, 1e-9f);
        cudaEventRecord(b, ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) break;
        cudaEventElapsedTime(&ms, a, b);
        tf = 2.0 * 8.0 * iters * 4 * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best32) best32 = tf;
    }
    cudaFree(buf);
    if (e != cudaSuccess) return cuda_fail(e, "cp_measure_peaks");
    *fp64_tflops = best64;
    *fp32_tflops = best32;
    return CP_OK;
}

cp_status cp_score(cp_ctx *ctx, const cp_problem *problem, const double *dof, uint64_t n, double *out_scores) {
    if (!ctx || !dof || !out_scores) return cp_fail(CP_ERR_INVALID_ARG, "cp_score: null argument");
    if (n == 0) return CP_OK;
    KProblem kp;
    cp_status st = make_kproblem(problem, &kp);
    if (st != CP_OK) return st;
    CP_CUDA(cudaSetDevice(ctx->device));
    double *d_dof = nullptr, *d_out = nullptr;
    CP_CUDA(cudaMalloc(&d_dof, sizeof(double) * CP_MAX_DOF * n));
    cudaError_t e = cudaMalloc(&d_out, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_dof, dof, sizeof(double) * CP_MAX_DOF * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const int lanes = effective_lanes(ctx, kp);
        e = cp_launch_score(kp, lanes, d_dof, n, d_out, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_scores, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_dof);
    cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "cp_score");
    return CP_OK;
}

cp_status cp_images(cp_ctx *ctx, const cp_problem *problem, const double *dof, uint64_t n, double *out) {
    if (!ctx || !dof || !out) return cp_fail(CP_ERR_INVALID_ARG, "cp_images: null argument");
    if (n == 0) return CP_OK;
    KProblem kp;
    cp_status st = make_kproblem(problem, &kp);
    if (st != CP_OK) return st;
    CP_CUDA(cudaSetDevice(ctx->device));
    const size_t out_bytes = sizeof(double) * 8 * kp.n_syms * n;
    double *d_dof = nullptr, *d_out = nullptr;
    CP_CUDA(cudaMalloc(&d_dof, sizeof(double) * CP_MAX_DOF * n));
    cudaError_t e = cudaMalloc(&d_out, out_bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_dof, dof, sizeof(double) * CP_MAX_DOF * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cp_launch_images(kp, d_dof, n, d_out, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_dof);
    cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "cp_images");
    return CP_OK;
}

cp_status cp_replay(cp_ctx *ctx, const cp_problem *problem, const double *init_dof, const cp_move *moves,
                    uint64_t n_moves, uint64_t n_replicas, cp_decision *out) {
    return cp_replay_bounded(ctx, problem, init_dof, nullptr, moves, n_moves, n_replicas, out);
}

cp_status cp_replay_bounded(cp_ctx *ctx, const cp_problem *problem, const double *init_dof, const double *bound_dof,
                            const cp_move *moves, uint64_t n_moves, uint64_t n_replicas, cp_decision *out) {
    if (!ctx || !init_dof || !moves || !out) return cp_fail(CP_ERR_INVALID_ARG, "cp_replay: null argument");
    if (n_moves == 0 || n_replicas == 0) return CP_OK;
    KProblem kp;
    cp_status st = make_kproblem(problem, &kp);
    if (st != CP_OK) return st;
    CP_CUDA(cudaSetDevice(ctx->device));
    const size_t total = (size_t)n_moves * n_replicas;
    double *d_init = nullptr;
    cp_move *d_mv = nullptr;
    cp_decision *d_out = nullptr;
    kp.replay_has_bounds = bound_dof ? 1 : 0;
    const size_t dof_bytes = sizeof(double) * CP_MAX_DOF * n_replicas;
    cudaError_t e = cudaMalloc(&d_init, dof_bytes * (bound_dof ? 2 : 1));
    if (e == cudaSuccess) e = cudaMalloc(&d_mv, sizeof(cp_move) * total);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(cp_decision) * total);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_init, init_dof, dof_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && bound_dof)
        e = cudaMemcpyAsync(d_init + CP_MAX_DOF * n_replicas, bound_dof, dof_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_mv, moves, sizeof(cp_move) * total, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const int lanes = effective_lanes(ctx, kp);
        e = cp_launch_replay(kp, lanes, d_init, d_mv, n_moves, n_replicas, d_out, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(cp_decision) * total, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_init);
    cudaFree(d_mv);
    cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "cp_replay");
    return CP_OK;
}

}  // extern "C"
