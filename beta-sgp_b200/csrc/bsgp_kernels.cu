// libbsgp: persistent thread-block-cluster kernels for the SGP / beta-SGP restoration loop on
// B200 (sm_100a) and their C ABI (include/bsgp.h).
//
// Execution model
//   * One thread-block CLUSTER of G CTAs restores one image from start to finish
//     (bsgp_solver.cuh); clusters are persistent and pull image indices from a global queue, so
//     images with very different iteration counts (2..160 observed) never wait for each other and
//     the host is not involved between "inputs resident" and "outputs written".
//   * Scalars travel between the CTAs of a cluster through distributed shared memory: every CTA
//     stores its partial sums into every peer's inbox and one hardware cluster barrier publishes
//     them; all CTAs then add the G partials in the same order, so all controllers agree bit for bit.
//   * Per-image state that does not fit in shared memory lives in a per-CLUSTER scratch area (not
//     per image): with ~18 clusters in flight the whole working set stays resident in the 126 MB L2,
//     HBM only sees each input once and each output once.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "bsgp_plan.h"
#include "bsgp_solver.cuh"

namespace cg = cooperative_groups;
using namespace bsgp;

// ------------------------------------------------------------------------------------------------
// device context
// ------------------------------------------------------------------------------------------------
constexpr int kMaxWarps = 32;
constexpr int kMaxG = 16;
constexpr int kMaxK = 8;

struct SharedCtl {
    double warp_part[kMaxWarps][kMaxK];
    double inbox[2][kMaxG][kMaxK];
    int next_img;
    int pad[3];
};

__device__ __forceinline__ double red_combine(int op, double a, double b) {
    if (op == 0) return a + b;
    if (op == 1) return (b < a) ? b : a;
    return (b > a) ? b : a;
}

struct DeviceCtx {
    int tid, nt, rank, G;
    SharedCtl* sh;
    int parity;

    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void cluster_sync() {
        if (G > 1) cg::this_cluster().sync();
        else __syncthreads();
    }
    __device__ __forceinline__ double now() {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return (double)t * 1e-9;
    }
    // all-reduce of k <= 8 doubles over the whole cluster; every thread of every CTA receives the
    // same bits.  One block barrier + one cluster barrier.
    __device__ __forceinline__ void allreduce(int op, double* v, int k) {
        const int lane = tid & 31, warp = tid >> 5, nwarps = (nt + 31) >> 5;
        for (int j = 0; j < k; ++j) {
            double x = v[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x = red_combine(op, x, __shfl_xor_sync(0xffffffffu, x, o));
            if (lane == 0) sh->warp_part[warp][j] = x;
        }
        __syncthreads();
        if (tid < G * k) {
            const int dst = tid / k, j = tid - dst * k;
            double s = sh->warp_part[0][j];
            for (int w = 1; w < nwarps; ++w) s = red_combine(op, s, sh->warp_part[w][j]);
            double* slot = &sh->inbox[parity][rank][j];
            if (G > 1) slot = cg::this_cluster().map_shared_rank(slot, dst);
            *slot = s;
        }
        cluster_sync();
        for (int j = 0; j < k; ++j) {
            double s = sh->inbox[parity][0][j];
            for (int r = 1; r < G; ++r) s = red_combine(op, s, sh->inbox[parity][r][j]);
            v[j] = s;
        }
        parity ^= 1;
    }
    __device__ __forceinline__ void allreduce_sum(double* v, int k) { allreduce(0, v, k); }
    __device__ __forceinline__ void allreduce_min(double& v) { allreduce(1, &v, 1); }
    __device__ __forceinline__ void allreduce_max(double& v) { allreduce(2, &v, 1); }
};

__device__ __forceinline__ DeviceCtx make_ctx(SharedCtl* sh, int G) {
    DeviceCtx c;
    c.tid = threadIdx.x; c.nt = blockDim.x; c.G = G;
    c.rank = (G > 1) ? (int)cg::this_cluster().block_rank() : 0;
    c.sh = sh; c.parity = 0;
    return c;
}

// next work item for the whole cluster (leader claims it, pushes it into every CTA's shared memory)
__device__ __forceinline__ int next_item(DeviceCtx& ctx, int* queue) {
    if (ctx.rank == 0 && ctx.tid == 0) {
        const int v = atomicAdd(queue, 1);
        if (ctx.G > 1) {
            cg::cluster_group cl = cg::this_cluster();
            for (int r = 0; r < ctx.G; ++r) *cl.map_shared_rank(&ctx.sh->next_img, r) = v;
        } else {
            ctx.sh->next_img = v;
        }
    }
    ctx.cluster_sync();
    const int img = ctx.sh->next_img;
    ctx.cluster_sync();          // nobody may still be reading when the leader claims the next one
    return img;
}

template <typename T> struct SmemLayout {
    SharedCtl* ctl;
    cplx<T>* ws;
    unsigned char* bufs;
};
template <typename T> __device__ __forceinline__ SmemLayout<T> carve(unsigned char* smem, size_t ws_bytes) {
    SmemLayout<T> l;
    l.ctl = reinterpret_cast<SharedCtl*>(smem);
    size_t off = (sizeof(SharedCtl) + 127) & ~(size_t)127;
    l.ws = reinterpret_cast<cplx<T>*>(smem + off);
    off += (ws_bytes + 127) & ~(size_t)127;
    l.bufs = smem + off;
    return l;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
template <typename T, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) bsgp_solve_kernel(const SolveArgs<T> a, const size_t ws_bytes, const size_t tf_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    SmemLayout<T> lay = carve<T>(smem, ws_bytes);
    DeviceCtx ctx = make_ctx(lay.ctl, a.g.G);
    const int cluster_id = blockIdx.x / a.g.G;
    const size_t npix = (size_t)a.g.ny * a.g.nx;
    const size_t nslab = (size_t)a.g.rows_per_cta * a.g.nx;
    T* buf[NBUF];
    size_t soff = 0;
#pragma unroll
    for (int b = 0; b < NBUF; ++b) {
        if (a.resident_mask & (1 << b)) { buf[b] = reinterpret_cast<T*>(lay.bufs + soff); soff += nslab * sizeof(T); }
        else buf[b] = a.work + (size_t)cluster_id * a.work_stride + (size_t)b * npix + (size_t)ctx.rank * nslab;
    }
    cplx<T>* spec = a.spec + (size_t)cluster_id * a.spec_stride;
    for (;;) {
        const int img = next_item(ctx, a.queue);
        if (img >= a.batch) break;
        cplx<T>* tf = a.tf + (a.n_psf > 1 ? (size_t)img * tf_stride : 0);
        solve_image<T>(ctx, a, buf, lay.ws, spec, tf, img);
    }
}

template <typename T> struct ConvArgs {
    ConvGeom g;
    int count;
    const T* in; T* out;                // images / PSFs [count][ny][nx]
    const cplx<T>* twx; const cplx<T>* twy; cplx<T>* tf; int n_psf; size_t tf_stride;
    cplx<T>* spec; size_t spec_stride;
    int mode;                           // CONV_TF / CONV_CTF / CONV_MAKE_TF
    int* queue;
};

// mode CONV_MAKE_TF: tf[p] = fftn(fftshift(psf[p])) in workspace order (sgp.py:109); otherwise
// out[i] = A(in[i]) or A^T(in[i]) (sgp.py:111-120).
template <typename T>
__global__ void __launch_bounds__(512, 1) bsgp_conv_kernel(const ConvArgs<T> a, const size_t ws_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    SmemLayout<T> lay = carve<T>(smem, ws_bytes);
    DeviceCtx ctx = make_ctx(lay.ctl, a.g.G);
    const int cluster_id = blockIdx.x / a.g.G;
    const ConvGeom& g = a.g;
    const size_t npix = (size_t)g.ny * g.nx;
    cplx<T>* spec = a.spec + (size_t)cluster_id * a.spec_stride;
    const int r0 = ctx.rank * g.rows_per_cta;
    for (;;) {
        const int img = next_item(ctx, a.queue);
        if (img >= a.count) break;
        const T* src = a.in + (size_t)img * npix;
        if (a.mode == CONV_MAKE_TF) {
            auto prod = [&](int row, int c) -> T {
                return src[(size_t)((r0 + row + (g.ny >> 1)) & (g.ny - 1)) * g.nx + ((c + (g.nx >> 1)) & (g.nx - 1))];
            };
            auto cons = [&](int, int, T) {};
            conv_image(ctx, g, lay.ws, a.twx, a.twy, spec, a.tf + (size_t)img * a.tf_stride, CONV_MAKE_TF, prod, cons);
        } else {
            T* dst = a.out + (size_t)img * npix;
            auto prod = [&](int row, int c) -> T { return src[(size_t)(r0 + row) * g.nx + c]; };
            auto cons = [&](int row, int c, T v) { dst[(size_t)(r0 + row) * g.nx + c] = v; };
            conv_image(ctx, g, lay.ws, a.twx, a.twy, spec, a.tf + (a.n_psf > 1 ? (size_t)img * a.tf_stride : 0), a.mode, prod, cons);
        }
        ctx.cluster_sync();   // spec is reused by the next item
    }
}

// projectDF (flux_conserve_proj.py:7-144): one CTA per problem, x = clamp((c + lambda) / dia).
__global__ void __launch_bounds__(512, 1) bsgp_project_kernel(const double* __restrict__ b, const double* __restrict__ c,
                                                              const double* __restrict__ dia, int n, int batch, double cap, int has_cap,
                                                              double lambda0, double dlambda0, double tol_lam, int max_projs,
                                                              double* __restrict__ x, int* evals, int* status) {
    __shared__ SharedCtl ctl;
    DeviceCtx ctx = make_ctx(&ctl, 1);
    for (int p = blockIdx.x; p < batch; p += gridDim.x) {
        const double* cp = c + (size_t)p * n;
        const double* dp = dia + (size_t)p * n;
        auto point = [&](int i, double lam) -> double {
            double v = ndiv(nadd(cp[i], lam), dp[i]);
            v = (v <= 0.0) ? 0.0 : v;
            if (has_cap) v = (v >= cap) ? cap : v;
            return v;
        };
        const double target = b[p];
        auto eval = [&](double lam) -> double {
            KSum s; s.clear();
            for (int i = ctx.tid; i < n; i += ctx.nt) s.add(point(i, lam));
            double t = s.value();
            ctx.allreduce_sum(&t, 1);
            return t - target;
        };
        const ProjResult pr = flux_rootfind(eval, target, max_projs, lambda0, dlambda0, tol_lam);
        for (int i = ctx.tid; i < n; i += ctx.nt) x[(size_t)p * n + i] = point(i, pr.lambda);
        if (ctx.tid == 0) { if (evals) evals[p] = pr.evals; if (status) status[p] = pr.status; }
        __syncthreads();
    }
}

// betaDiv partial sums + optional betaDivDeriv (sgp.py:441-495).  part[block][3]
__global__ void __launch_bounds__(256) bsgp_betadiv_kernel(const double* __restrict__ y, const double* __restrict__ x, long long n,
                                                           double beta, double* part, double* deriv) {
    __shared__ SharedCtl ctl;
    DeviceCtx ctx = make_ctx(&ctl, 1);
    DivK<double> dk = make_divk<double>(BSGP_DIV_BETA, beta);
    KSum acc[3];
    acc[0].clear(); acc[1].clear(); acc[2].clear();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        objective_pixel<double>(dk, x[i], y[i], 0.0, true, acc);
        if (deriv) deriv[i] = (dk.kind == 1) ? dbeta_pixel<double>(x[i], y[i], beta) : 0.0;
    }
    double v[3] = {acc[0].value(), acc[1].value(), acc[2].value()};
    ctx.allreduce_sum(v, 3);
    if (threadIdx.x == 0) { part[blockIdx.x * 3 + 0] = v[0]; part[blockIdx.x * 3 + 1] = v[1]; part[blockIdx.x * 3 + 2] = v[2]; }
}

// the two per-pixel pieces of betaDivDerivwrtY (sgp.py:498-499): p1 = den^(beta-1), u = gn * den^(beta-2)
__global__ void __launch_bounds__(256) bsgp_betagrad_kernel(const double* __restrict__ den, const double* __restrict__ gn, long long n,
                                                            double beta, double* __restrict__ p1, double* __restrict__ u) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d = den[i];
        const double p = mpow(d, beta - 1.0);
        p1[i] = p;
        u[i] = nmul(gn[i], ndiv(p, d));
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(expr)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) return fail(BSGP_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

struct bsgp_plan {
    int ny = 0, nx = 0, dtype = 0, device = 0;
    int num_sms = 0, max_smem = 0;
    int want_G = 0, want_threads = 0;
    bool configured = false;
    ConvGeom g{};
    size_t ws_bytes = 0, smem_bytes = 0, elem = 8;
    int threads = 0, minb = 1, num_clusters = 0, resident_mask = 0;
    void* twx = nullptr; void* twy = nullptr;
    void* tf = nullptr; int n_psf = 0; int tf_capacity = 0; size_t tf_stride = 0;
    void* work = nullptr; size_t work_stride = 0;
    void* spec = nullptr; size_t spec_stride = 0;
    int* queue = nullptr;
    size_t workspace_bytes = 0;
};

template <typename T> static const void* solve_kernel_ptr(int threads) {
    if (threads <= 256) return (const void*)bsgp_solve_kernel<T, 256, 2>;
    return (const void*)bsgp_solve_kernel<T, 512, 1>;
}
template <typename T> static const void* conv_kernel_ptr() { return (const void*)bsgp_conv_kernel<T>; }

static int launch_clustered(const void* func, int grid, int block, size_t smem, int G, cudaStream_t st, void** args) {
    // the attribute is per function, not per plan: plans of different shapes share the kernels, so (re)state
    // the dynamic shared-memory limit this launch needs
    CU(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (G > 8) CU(cudaFuncSetAttribute(func, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CU(cudaLaunchKernelExC(&cfg, func, args));
    return BSGP_OK;
}

static int query_clusters(const void* func, int block, size_t smem, int G, int num_sms, int* out) {
    CU(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (G > 8) CU(cudaFuncSetAttribute(func, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (G == 1) {
        int nb = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, func, block, smem));
        *out = nb * num_sms;
        return BSGP_OK;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(G * num_sms); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    CU(cudaOccupancyMaxActiveClusters(&n, func, &cfg));
    *out = n;
    return BSGP_OK;
}

static void plan_free_buffers(bsgp_plan* p) {
    cudaFree(p->twx); cudaFree(p->twy); cudaFree(p->tf); cudaFree(p->work); cudaFree(p->spec); cudaFree(p->queue);
    p->twx = p->twy = p->tf = p->work = p->spec = nullptr; p->queue = nullptr;
    p->tf_capacity = 0; p->n_psf = 0;
}

// Residency priority: the two projection buffers (read E times per iteration), then the arrays with
// the most touches per iteration.
static const int kResidencyOrder[NBUF] = {B_D, B_T1, B_G, B_XTF, B_DTF, B_GN, B_XA, B_XB, B_BKG};

template <typename T> static int plan_setup_t(bsgp_plan* p) {
    const size_t npix = (size_t)p->ny * p->nx;
    p->elem = sizeof(T);
    // cluster size: slabs of <= 64 KB, at most 8 CTAs (portable cluster limit)
    int G = p->want_G;
    if (G <= 0) {
        G = 1;
        while (G < 8 && npix * sizeof(T) / G > 65536 && p->ny % (4 * G) == 0 && (p->nx / 2) % (2 * G) == 0) G *= 2;
    }
    int threads = p->want_threads;
    const size_t nslab = npix / G;
    if (threads <= 0) threads = (nslab <= 2048) ? 256 : 512;
    if (threads != 256 && threads != 512 && threads != 128) return fail(BSGP_E_ARG, "threads must be 128, 256 or 512");
    const size_t ws_limit = 72 * 1024;
    if (!make_geom(p->ny, p->nx, G, sizeof(cplx<T>), ws_limit, &p->g, &p->ws_bytes))
        return fail(BSGP_E_SHAPE, "unsupported shape %dx%d for cluster size %d (power-of-two sides >= 16 required)", p->ny, p->nx, G);
    p->threads = threads;
    const size_t fixed = ((sizeof(SharedCtl) + 127) & ~(size_t)127) + ((p->ws_bytes + 127) & ~(size_t)127);
    // shared-memory budget per CTA: whole SM for the big configuration, a third for stamps
    size_t budget = (size_t)p->max_smem;
    const size_t slab_bytes = nslab * sizeof(T);
    if (threads <= 256) {
        // several CTAs per SM: aim for everything resident, then see how many CTAs fit
        const size_t all = fixed + NBUF * slab_bytes;
        budget = all <= (size_t)p->max_smem ? all : (size_t)p->max_smem;
    }
    size_t used = fixed;
    int mask = 0;
    for (int k = 0; k < NBUF; ++k) {
        if (used + slab_bytes <= budget) { mask |= 1 << kResidencyOrder[k]; used += slab_bytes; }
    }
    p->resident_mask = mask;
    p->smem_bytes = used;
    const void* fn = solve_kernel_ptr<T>(threads);
    int nc = 0;
    int rc = query_clusters(fn, threads, p->smem_bytes, G, p->num_sms, &nc);
    if (rc) return rc;
    if (nc <= 0) return fail(BSGP_E_CUDA, "kernel does not fit: cluster %d, %d threads, %zu B shared memory", G, threads, p->smem_bytes);
    p->num_clusters = nc;
    rc = query_clusters(conv_kernel_ptr<T>(), 512, fixed, G, p->num_sms, &nc);
    if (rc) return rc;

    std::vector<cplx<T>> tw;
    make_twiddles<T>(p->nx, tw);
    CU(cudaMalloc(&p->twx, tw.size() * sizeof(cplx<T>)));
    CU(cudaMemcpy(p->twx, tw.data(), tw.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice));
    make_twiddles<T>(p->ny, tw);
    CU(cudaMalloc(&p->twy, tw.size() * sizeof(cplx<T>)));
    CU(cudaMemcpy(p->twy, tw.data(), tw.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice));
    p->work_stride = NBUF * npix;
    p->spec_stride = (size_t)p->ny * p->g.hx;
    p->tf_stride = (size_t)(p->g.hx + 1) * p->ny;
    CU(cudaMalloc(&p->work, (size_t)p->num_clusters * p->work_stride * sizeof(T)));
    CU(cudaMalloc(&p->spec, (size_t)p->num_clusters * p->spec_stride * sizeof(cplx<T>)));
    CU(cudaMalloc((void**)&p->queue, sizeof(int)));
    p->workspace_bytes = (size_t)p->num_clusters * (p->work_stride * sizeof(T) + p->spec_stride * sizeof(cplx<T>));
    p->configured = true;
    return BSGP_OK;
}

static int plan_setup(bsgp_plan* p) {
    if (p->configured) return BSGP_OK;
    CU(cudaSetDevice(p->device));
    return p->dtype == BSGP_F64 ? plan_setup_t<double>(p) : plan_setup_t<float>(p);
}

template <typename T> static int set_psf_t(bsgp_plan* p, const void* psf_dev, int n_psf, cudaStream_t st) {
    if (n_psf > p->tf_capacity) {
        cudaFree(p->tf); p->tf = nullptr; p->tf_capacity = 0;
        CU(cudaMalloc(&p->tf, (size_t)n_psf * p->tf_stride * sizeof(cplx<T>)));
        p->tf_capacity = n_psf;
    }
    p->n_psf = n_psf;
    ConvArgs<T> a;
    memset(&a, 0, sizeof a);
    a.g = p->g; a.count = n_psf; a.in = (const T*)psf_dev; a.out = nullptr;
    a.twx = (const cplx<T>*)p->twx; a.twy = (const cplx<T>*)p->twy; a.tf = (cplx<T>*)p->tf; a.n_psf = n_psf; a.tf_stride = p->tf_stride;
    a.spec = (cplx<T>*)p->spec; a.spec_stride = p->spec_stride; a.mode = CONV_MAKE_TF; a.queue = p->queue;
    CU(cudaMemsetAsync(p->queue, 0, sizeof(int), st));
    const size_t smem = ((sizeof(SharedCtl) + 127) & ~(size_t)127) + ((p->ws_bytes + 127) & ~(size_t)127);
    size_t wsb = p->ws_bytes;
    void* args[] = {&a, &wsb};
    const int nclu = n_psf < p->num_clusters ? n_psf : p->num_clusters;
    return launch_clustered(conv_kernel_ptr<T>(), nclu * p->g.G, 512, smem, p->g.G, st, args);
}

template <typename T> static int apply_psf_t(bsgp_plan* p, const void* x, void* y, int batch, int adjoint, cudaStream_t st) {
    ConvArgs<T> a;
    memset(&a, 0, sizeof a);
    a.g = p->g; a.count = batch; a.in = (const T*)x; a.out = (T*)y;
    a.twx = (const cplx<T>*)p->twx; a.twy = (const cplx<T>*)p->twy; a.tf = (cplx<T>*)p->tf; a.n_psf = p->n_psf; a.tf_stride = p->tf_stride;
    a.spec = (cplx<T>*)p->spec; a.spec_stride = p->spec_stride; a.mode = adjoint ? CONV_CTF : CONV_TF; a.queue = p->queue;
    CU(cudaMemsetAsync(p->queue, 0, sizeof(int), st));
    const size_t smem = ((sizeof(SharedCtl) + 127) & ~(size_t)127) + ((p->ws_bytes + 127) & ~(size_t)127);
    size_t wsb = p->ws_bytes;
    void* args[] = {&a, &wsb};
    const int nclu = batch < p->num_clusters ? batch : p->num_clusters;
    return launch_clustered(conv_kernel_ptr<T>(), nclu * p->g.G, 512, smem, p->g.G, st, args);
}

template <typename T>
static int solve_t(bsgp_plan* p, const bsgp_params* prm, int batch, const bsgp_inputs* in, const bsgp_outputs* out, cudaStream_t st) {
    SolveArgs<T> a;
    memset(&a, 0, sizeof a);
    a.p = *prm; a.g = p->g; a.batch = batch;
    a.gn = (const T*)in->gn; a.bkg = (const T*)in->bkg; a.bkg_is_image = in->bkg_is_image; a.flux = in->flux; a.beta0 = in->beta0;
    a.x0 = (const T*)in->x0; a.obj = (const T*)in->obj;
    a.twx = (const cplx<T>*)p->twx; a.twy = (const cplx<T>*)p->twy; a.tf = (cplx<T>*)p->tf; a.n_psf = p->n_psf;
    a.work = (T*)p->work; a.work_stride = p->work_stride; a.spec = (cplx<T>*)p->spec; a.spec_stride = p->spec_stride;
    a.resident_mask = p->resident_mask;
    a.x_out = (T*)out->x; a.iters = out->iters; a.status = out->status; a.discr = out->discr; a.times = out->times;
    a.stop_value = out->stop_value; a.err = out->err; a.beta_final = out->beta_final; a.proj_evals = out->proj_evals;
    a.ls_trials = out->ls_trials; a.scalars = out->scalars; a.tr_alpha = out->trace_alpha; a.tr_lambda = out->trace_lambda;
    a.tr_beta = out->trace_beta; a.tr_trials = out->trace_trials; a.tr_evals = out->trace_evals;
    a.queue = p->queue;
    CU(cudaMemsetAsync(p->queue, 0, sizeof(int), st));
    size_t wsb = p->ws_bytes, tfs = p->tf_stride;
    void* args[] = {&a, &wsb, &tfs};
    const int nclu = batch < p->num_clusters ? batch : p->num_clusters;
    return launch_clustered(solve_kernel_ptr<T>(p->threads), nclu * p->g.G, p->threads, p->smem_bytes, p->g.G, st, args);
}

static int check_params(const bsgp_params* q) {
    if (q->divergence != BSGP_DIV_KL && q->divergence != BSGP_DIV_BETA) return fail(BSGP_E_ARG, "bad divergence");
    if (q->init_recon < 0 || q->init_recon > 3) return fail(BSGP_E_ARG, "init_recon must be 0..3");
    if (q->proj_type < 0 || q->proj_type > 1) return fail(BSGP_E_ARG, "proj_type must be 0 or 1");
    if (q->stop_criterion < 0 || q->stop_criterion > 4) return fail(BSGP_E_ARG, "stop_criterion must be 0..4");
    if (q->maxit < 1) return fail(BSGP_E_ARG, "MAXIT must be >= 1");
    if (q->m < 1 || q->m > kMaxMem || q->m_alpha < 1 || q->m_alpha > kMaxMem) return fail(BSGP_E_ARG, "M and M_alpha must be in 1..%d", kMaxMem);
    return BSGP_OK;
}

extern "C" {

int bsgp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
const char* bsgp_last_error_string(void) { return g_err.c_str(); }
const char* bsgp_version(void) { return "libbsgp 0.1 (sm_100a)"; }

int bsgp_plan_create(int ny, int nx, int dtype, int device, bsgp_plan** plan) {
    if (!plan) return fail(BSGP_E_ARG, "plan is NULL");
    *plan = nullptr;
    if (dtype != BSGP_F64 && dtype != BSGP_F32) return fail(BSGP_E_ARG, "dtype must be BSGP_F64 or BSGP_F32");
    if (!is_pow2(ny) || !is_pow2(nx) || ny < 16 || nx < 16 || ny > 4096 || nx > 4096)
        return fail(BSGP_E_SHAPE, "unsupported image shape %dx%d: the cluster solver handles power-of-two sides in [16, 4096]", ny, nx);
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(BSGP_E_CUDA, "device %d not available (%d CUDA devices)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 9) return fail(BSGP_E_CUDA, "thread-block clusters need compute capability >= 9.0 (found %d.%d)", prop.major, prop.minor);
    bsgp_plan* p = new bsgp_plan();
    p->ny = ny; p->nx = nx; p->dtype = dtype; p->device = device;
    p->num_sms = prop.multiProcessorCount;
    p->max_smem = (int)prop.sharedMemPerBlockOptin;
    int rc = plan_setup(p);
    if (rc) { plan_free_buffers(p); delete p; return rc; }
    *plan = p;
    return BSGP_OK;
}

int bsgp_plan_configure(bsgp_plan* p, int cluster_size, int threads) {
    if (!p) return fail(BSGP_E_ARG, "plan is NULL");
    CU(cudaSetDevice(p->device));
    CU(cudaDeviceSynchronize());
    plan_free_buffers(p);
    p->configured = false;
    p->want_G = cluster_size; p->want_threads = threads;
    return plan_setup(p);
}

int bsgp_plan_destroy(bsgp_plan* p) {
    if (!p) return BSGP_OK;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    plan_free_buffers(p);
    delete p;
    return BSGP_OK;
}

int bsgp_plan_get_info(const bsgp_plan* p, bsgp_plan_info* info) {
    if (!p || !info) return fail(BSGP_E_ARG, "NULL argument");
    info->ny = p->ny; info->nx = p->nx; info->dtype = p->dtype; info->device = p->device;
    info->cluster_size = p->g.G; info->num_clusters = p->num_clusters; info->threads = p->threads;
    info->smem_bytes = (int)p->smem_bytes; info->num_sms = p->num_sms; info->resident_mask = p->resident_mask;
    info->workspace_bytes = (long long)p->workspace_bytes;
    return BSGP_OK;
}

int bsgp_set_psf(bsgp_plan* p, const void* psf_dev, int n_psf, void* stream) {
    if (!p || !psf_dev || n_psf < 1) return fail(BSGP_E_ARG, "bad argument");
    CU(cudaSetDevice(p->device));
    return p->dtype == BSGP_F64 ? set_psf_t<double>(p, psf_dev, n_psf, (cudaStream_t)stream)
                                : set_psf_t<float>(p, psf_dev, n_psf, (cudaStream_t)stream);
}

int bsgp_set_psf_host(bsgp_plan* p, const void* psf_host, int n_psf) {
    if (!p || !psf_host || n_psf < 1) return fail(BSGP_E_ARG, "bad argument");
    CU(cudaSetDevice(p->device));
    const size_t bytes = (size_t)n_psf * p->ny * p->nx * p->elem;
    void* d = nullptr;
    CU(cudaMalloc(&d, bytes));
    cudaError_t e = cudaMemcpy(d, psf_host, bytes, cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? bsgp_set_psf(p, d, n_psf, nullptr) : fail(BSGP_E_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
    if (rc == BSGP_OK) { e = cudaDeviceSynchronize(); if (e != cudaSuccess) rc = fail(BSGP_E_CUDA, "PSF spectrum kernel failed: %s", cudaGetErrorString(e)); }
    cudaFree(d);
    return rc;
}

int bsgp_solve_batch(bsgp_plan* p, const bsgp_params* prm, int batch, const bsgp_inputs* in, const bsgp_outputs* out, void* stream) {
    if (!p || !prm || !in || !out) return fail(BSGP_E_ARG, "NULL argument");
    if (batch < 1) return fail(BSGP_E_ARG, "batch must be >= 1");
    int rc = check_params(prm);
    if (rc) return rc;
    if (p->n_psf != 1 && p->n_psf != batch) return fail(BSGP_E_STATE, "bsgp_set_psf was called with %d PSFs; need 1 or batch (%d)", p->n_psf, batch);
    if (!in->gn || !in->bkg || !out->x || !out->iters || !out->status || !out->discr || !out->times) return fail(BSGP_E_ARG, "required pointer is NULL");
    if (prm->has_flux && !in->flux) return fail(BSGP_E_ARG, "has_flux set but inputs.flux is NULL");
    if (prm->divergence == BSGP_DIV_BETA && !in->beta0) return fail(BSGP_E_ARG, "beta-divergence needs inputs.beta0");
    if (prm->init_recon == 1 && !in->x0) return fail(BSGP_E_ARG, "init_recon = 1 needs inputs.x0");
    if (prm->errflag && (!in->obj || !out->err)) return fail(BSGP_E_ARG, "errflag needs inputs.obj and outputs.err");
    CU(cudaSetDevice(p->device));
    return p->dtype == BSGP_F64 ? solve_t<double>(p, prm, batch, in, out, (cudaStream_t)stream)
                                : solve_t<float>(p, prm, batch, in, out, (cudaStream_t)stream);
}

// host staging helpers -----------------------------------------------------------------------
struct DevBuf {
    void* d = nullptr;
    ~DevBuf() { cudaFree(d); }
    int up(const void* h, size_t bytes) {
        if (!h) return BSGP_OK;
        CU(cudaMalloc(&d, bytes));
        CU(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
        return BSGP_OK;
    }
    int alloc(bool want, size_t bytes) {
        if (!want) return BSGP_OK;
        CU(cudaMalloc(&d, bytes));
        CU(cudaMemset(d, 0, bytes));
        return BSGP_OK;
    }
    int down(void* h, size_t bytes) {
        if (!h || !d) return BSGP_OK;
        CU(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
        return BSGP_OK;
    }
};

int bsgp_solve_batch_host(bsgp_plan* p, const bsgp_params* prm, int batch, const bsgp_inputs* in, const bsgp_outputs* out) {
    if (!p || !prm || !in || !out) return fail(BSGP_E_ARG, "NULL argument");
    CU(cudaSetDevice(p->device));
    const size_t img = (size_t)p->ny * p->nx * p->elem, B = (size_t)batch, tr = (size_t)(prm->maxit + 1);
    DevBuf gn, bkg, flux, beta0, x0, obj, x, iters, status, discr, times, stopv, err, bfin, pe, lt, sc, ta, tl, tb, tt, te;
    int rc;
#define TRY(e) do { rc = (e); if (rc) return rc; } while (0)
    TRY(gn.up(in->gn, B * img));
    TRY(bkg.up(in->bkg, in->bkg_is_image ? B * img : B * p->elem));
    TRY(flux.up(in->flux, B * 8)); TRY(beta0.up(in->beta0, B * 8)); TRY(x0.up(in->x0, B * img)); TRY(obj.up(in->obj, B * img));
    TRY(x.alloc(true, B * img)); TRY(iters.alloc(true, B * 4)); TRY(status.alloc(true, B * 4));
    TRY(discr.alloc(true, B * tr * 8)); TRY(times.alloc(true, B * tr * 8));
    TRY(stopv.alloc(out->stop_value, B * tr * 8)); TRY(err.alloc(out->err, B * (tr + 1) * 8)); TRY(bfin.alloc(out->beta_final, B * 8));
    TRY(pe.alloc(out->proj_evals, B * 4)); TRY(lt.alloc(out->ls_trials, B * 4)); TRY(sc.alloc(out->scalars, B * BSGP_NSCALARS * 8));
    TRY(ta.alloc(out->trace_alpha, B * tr * 8)); TRY(tl.alloc(out->trace_lambda, B * tr * 8)); TRY(tb.alloc(out->trace_beta, B * tr * 8));
    TRY(tt.alloc(out->trace_trials, B * tr * 4)); TRY(te.alloc(out->trace_evals, B * tr * 4));
    bsgp_inputs di = {gn.d, bkg.d, in->bkg_is_image, (const double*)flux.d, (const double*)beta0.d, x0.d, obj.d};
    bsgp_outputs dout = {x.d, (int*)iters.d, (int*)status.d, (double*)discr.d, (double*)times.d, (double*)stopv.d, (double*)err.d,
                         (double*)bfin.d, (int*)pe.d, (int*)lt.d, (double*)sc.d, (double*)ta.d, (double*)tl.d, (double*)tb.d,
                         (int*)tt.d, (int*)te.d};
    TRY(bsgp_solve_batch(p, prm, batch, &di, &dout, nullptr));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(BSGP_E_CUDA, "solve kernel failed: %s", cudaGetErrorString(e));
    TRY(x.down(out->x, B * img)); TRY(iters.down(out->iters, B * 4)); TRY(status.down(out->status, B * 4));
    TRY(discr.down(out->discr, B * tr * 8)); TRY(times.down(out->times, B * tr * 8)); TRY(stopv.down(out->stop_value, B * tr * 8));
    TRY(err.down(out->err, B * (tr + 1) * 8)); TRY(bfin.down(out->beta_final, B * 8)); TRY(pe.down(out->proj_evals, B * 4));
    TRY(lt.down(out->ls_trials, B * 4)); TRY(sc.down(out->scalars, B * BSGP_NSCALARS * 8)); TRY(ta.down(out->trace_alpha, B * tr * 8));
    TRY(tl.down(out->trace_lambda, B * tr * 8)); TRY(tb.down(out->trace_beta, B * tr * 8)); TRY(tt.down(out->trace_trials, B * tr * 4));
    TRY(te.down(out->trace_evals, B * tr * 4));
    return BSGP_OK;
}

int bsgp_apply_psf(bsgp_plan* p, const void* x_dev, void* y_dev, int batch, int adjoint, void* stream) {
    if (!p || !x_dev || !y_dev || batch < 1) return fail(BSGP_E_ARG, "bad argument");
    if (p->n_psf < 1) return fail(BSGP_E_STATE, "bsgp_set_psf has not been called");
    if (p->n_psf != 1 && p->n_psf != batch) return fail(BSGP_E_STATE, "%d PSFs set; need 1 or batch (%d)", p->n_psf, batch);
    CU(cudaSetDevice(p->device));
    return p->dtype == BSGP_F64 ? apply_psf_t<double>(p, x_dev, y_dev, batch, adjoint, (cudaStream_t)stream)
                                : apply_psf_t<float>(p, x_dev, y_dev, batch, adjoint, (cudaStream_t)stream);
}

int bsgp_apply_psf_host(bsgp_plan* p, const void* x_host, void* y_host, int batch, int adjoint) {
    if (!p || !x_host || !y_host) return fail(BSGP_E_ARG, "NULL argument");
    CU(cudaSetDevice(p->device));
    const size_t bytes = (size_t)batch * p->ny * p->nx * p->elem;
    DevBuf x, y;
    int rc;
    TRY(x.up(x_host, bytes)); TRY(y.alloc(true, bytes));
    TRY(bsgp_apply_psf(p, x.d, y.d, batch, adjoint, nullptr));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(BSGP_E_CUDA, "convolution kernel failed: %s", cudaGetErrorString(e));
    return y.down(y_host, bytes);
}

int bsgp_project_df(const double* b, const double* c, const double* dia, int n, int batch, double sat_cap, double lambda0,
                    double dlambda0, double tol_lam, int max_projs, double* x, int* evals, int* status, int device, void* stream) {
    if (!b || !c || !dia || !x || n < 1 || batch < 1) return fail(BSGP_E_ARG, "bad argument");
    CU(cudaSetDevice(device));
    const int has_cap = (sat_cap == sat_cap) && sat_cap >= 0.0;
    const int grid = batch < 4096 ? batch : 4096;
    bsgp_project_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(b, c, dia, n, batch, sat_cap, has_cap, lambda0, dlambda0, tol_lam,
                                                                max_projs, x, evals, status);
    CU(cudaGetLastError());
    return BSGP_OK;
}

int bsgp_project_df_host(const double* b, const double* c, const double* dia, int n, int batch, double sat_cap, double lambda0,
                         double dlambda0, double tol_lam, int max_projs, double* x, int* evals, int* status, int device) {
    if (!b || !c || !dia || !x) return fail(BSGP_E_ARG, "NULL argument");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(BSGP_E_CUDA, "device %d not available", device);
    CU(cudaSetDevice(device));
    const size_t vb = (size_t)n * batch * 8;
    DevBuf db, dc, dd, dx, de, ds;
    int rc;
    TRY(db.up(b, (size_t)batch * 8)); TRY(dc.up(c, vb)); TRY(dd.up(dia, vb)); TRY(dx.alloc(true, vb));
    TRY(de.alloc(true, (size_t)batch * 4)); TRY(ds.alloc(true, (size_t)batch * 4));
    TRY(bsgp_project_df((const double*)db.d, (const double*)dc.d, (const double*)dd.d, n, batch, sat_cap, lambda0, dlambda0, tol_lam,
                        max_projs, (double*)dx.d, (int*)de.d, (int*)ds.d, device, nullptr));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(BSGP_E_CUDA, "projection kernel failed: %s", cudaGetErrorString(e));
    TRY(dx.down(x, vb)); TRY(de.down(evals, (size_t)batch * 4)); TRY(ds.down(status, (size_t)batch * 4));
    return BSGP_OK;
}

int bsgp_beta_div_host(const double* y, const double* x, long long n, double beta, double* value, double* deriv, int device) {
    if (!y || !x || !value || n < 1) return fail(BSGP_E_ARG, "bad argument");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(BSGP_E_CUDA, "device %d not available", device);
    CU(cudaSetDevice(device));
    DevBuf dy, dx, dp, dd;
    int rc;
    long long want = (n + 255) / 256;
    const int grid = (int)(want < 592 ? want : 592);
    TRY(dy.up(y, (size_t)n * 8)); TRY(dx.up(x, (size_t)n * 8)); TRY(dp.alloc(true, (size_t)grid * 3 * 8)); TRY(dd.alloc(deriv != nullptr, (size_t)n * 8));
    bsgp_betadiv_kernel<<<grid, 256>>>((const double*)dy.d, (const double*)dx.d, n, beta, (double*)dp.d, (double*)dd.d);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    std::vector<double> part((size_t)grid * 3);
    TRY(dp.down(part.data(), part.size() * 8));
    double acc[3] = {0.0, 0.0, 0.0};
    for (int b = 0; b < grid; ++b) for (int k = 0; k < 3; ++k) acc[k] += part[(size_t)b * 3 + k];
    // combine exactly like the solver (sgp.py:452-458); s1 is acc[0] for the generic case
    DivK<double> dk;
    dk.kind = (beta == 0.0) ? 2 : (beta == 1.0 ? 3 : 1);
    if (dk.kind == 1) *value = (acc[0] + acc[1]) - acc[2];
    else if (dk.kind == 2) *value = (acc[0] - acc[1]) - (double)n;
    else *value = (acc[0] - acc[1]) + acc[2];
    TRY(dd.down(deriv, (size_t)n * 8));
    return BSGP_OK;
}
int bsgp_beta_grad_terms_host(const double* den, const double* gn, long long n, double beta, double* p1, double* u, int device) {
    if (!den || !gn || !p1 || !u || n < 1) return fail(BSGP_E_ARG, "bad argument");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(BSGP_E_CUDA, "device %d not available", device);
    CU(cudaSetDevice(device));
    DevBuf dd, dg, dp, du;
    int rc;
    TRY(dd.up(den, (size_t)n * 8)); TRY(dg.up(gn, (size_t)n * 8)); TRY(dp.alloc(true, (size_t)n * 8)); TRY(du.alloc(true, (size_t)n * 8));
    long long want = (n + 255) / 256;
    bsgp_betagrad_kernel<<<(int)(want < 1184 ? want : 1184), 256>>>((const double*)dd.d, (const double*)dg.d, n, beta, (double*)dp.d, (double*)du.d);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    TRY(dp.down(p1, (size_t)n * 8)); TRY(du.down(u, (size_t)n * 8));
    return BSGP_OK;
}
#undef TRY

}  // extern "C"
