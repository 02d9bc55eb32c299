// Device-side execution context of libbsgp: one thread-block CLUSTER of G CTAs works on one image.
//
//   * Scalars travel between the CTAs of a cluster through distributed shared memory: every CTA
//     stores its partial sums into every peer's inbox and one hardware cluster barrier publishes
//     them; all CTAs then add the G partials in the same order, so all controllers agree bit for bit.
//   * Clusters are persistent and pull image indices from a global queue, so images with very
//     different iteration counts (2..160 observed) never wait for each other and the host is not
//     involved between "inputs resident" and "outputs written".
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "bsgp_math.cuh"

namespace bsgp {

namespace cg = cooperative_groups;

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

// dynamic shared memory layout of the persistent kernels (byte offsets, computed by the host)
struct SmemPlan {
    unsigned off_state;     // ImgState<T>
    unsigned off_twx;       // twiddles of the row transforms (if tw_smem)
    unsigned off_twy;       // twiddles of the column transforms (if tw_smem; may equal off_twx)
    unsigned off_ppx;       // padded-position table of the row transforms (nx x u16)
    unsigned off_ws;        // FFT workspace
    unsigned off_bufs;      // resident slab buffers
    int tw_smem;
};

}  // namespace bsgp
