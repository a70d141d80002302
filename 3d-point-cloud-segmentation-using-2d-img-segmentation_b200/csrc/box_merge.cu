// Kernel (4): batched instance-box intersection + union-find merge.
//   pair predicate   check_intersection, Fusion3DSeg/merge_intersecting_bb.py:44-56 (closed intervals :51-53,
//                    gated by equal category / parent id :49,80)
//   closure          the transitive merge the reference's sequential driver (:103-120) approximates
#include "f3d_common.cuh"
#include "f3d_host.h"

#define BOX_TILE 256

// fp32 interval that contains the fp64 interval (round lo down, hi up): a conservative prefilter at full rate,
// survivors are re-tested with the exact float64 comparisons of merge_intersecting_bb.py:51-53.
__device__ __forceinline__ float f_down(double x) { return __double2float_rd(x); }
__device__ __forceinline__ float f_up(double x) { return __double2float_ru(x); }

__device__ __forceinline__ bool axis_overlap(double lo1, double hi1, double lo2, double hi2) {
    return (lo1 <= lo2 && lo2 <= hi1) || (lo2 <= lo1 && lo1 <= hi2);
}

__global__ void __launch_bounds__(BOX_TILE) box_pairs_kernel(const double* __restrict__ lo, const double* __restrict__ hi,
                                                             const int32_t* __restrict__ group, int B, int ntiles,
                                                             int32_t* __restrict__ edges, long long cap,
                                                             unsigned long long* __restrict__ count) {
    // linear block id -> (ti <= tj) upper-triangular tile pair
    long long t = blockIdx.x;
    int ti = (int)((2.0 * ntiles + 1.0 - sqrt((2.0 * ntiles + 1.0) * (2.0 * ntiles + 1.0) - 8.0 * (double)t)) * 0.5);
    // fix up rounding
    while ((long long)ti * (2LL * ntiles - ti + 1) / 2 > t) --ti;
    while ((long long)(ti + 1) * (2LL * ntiles - ti) / 2 <= t) ++ti;
    int tj = ti + (int)(t - (long long)ti * (2LL * ntiles - ti + 1) / 2);

    __shared__ float s_lo[3][BOX_TILE], s_hi[3][BOX_TILE];
    __shared__ int s_g[BOX_TILE];
    const int tid = threadIdx.x;
    const int j0 = tj * BOX_TILE;
    {
        const int j = j0 + tid;
        if (j < B) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                s_lo[k][tid] = f_down(lo[3 * (size_t)j + k]);
                s_hi[k][tid] = f_up(hi[3 * (size_t)j + k]);
            }
            s_g[tid] = group[j];
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                s_lo[k][tid] = 3.0e38f;
                s_hi[k][tid] = -3.0e38f;
            }
            s_g[tid] = -1;
        }
    }
    __syncthreads();
    const int i = ti * BOX_TILE + tid;
    if (i >= B) return;
    float ilo[3], ihi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ilo[k] = f_down(lo[3 * (size_t)i + k]);
        ihi[k] = f_up(hi[3 * (size_t)i + k]);
    }
    const int gi = group[i];
    const int jn = min(BOX_TILE, B - j0);
    for (int jj = 0; jj < jn; ++jj) {
        const int j = j0 + jj;
        if (j <= i) continue;
        // closed-interval overlap <=> lo1 <= hi2 && lo2 <= hi1 for proper intervals; evaluated conservatively
        bool cand = (s_g[jj] == gi) && (ilo[0] <= s_hi[0][jj]) && (s_lo[0][jj] <= ihi[0]) && (ilo[1] <= s_hi[1][jj]) &&
                    (s_lo[1][jj] <= ihi[1]) && (ilo[2] <= s_hi[2][jj]) && (s_lo[2][jj] <= ihi[2]);
        if (cand) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                ok = ok && axis_overlap(lo[3 * (size_t)i + k], hi[3 * (size_t)i + k], lo[3 * (size_t)j + k], hi[3 * (size_t)j + k]);
            if (ok) {
                const unsigned long long e = atomicAdd(count, 1ULL);
                if ((long long)e < cap) {
                    edges[2 * e] = i;
                    edges[2 * e + 1] = j;
                }
            }
        }
    }
}

// ---- union-find: hook the larger root under the smaller one, so a tree's root is its minimum index -------------------
__device__ __forceinline__ int uf_find(int* parent, int x) {
    // volatile reads bypass the (non-coherent) L1 so that a retry after a failed hook sees other SMs' hooks
    volatile int* vp = parent;
    int p = vp[x];
    while (p != x) {
        const int gp = vp[p];
        if (gp != p) parent[x] = gp;   // path halving; benign race: only ever replaces an ancestor by a farther ancestor
        x = p;
        p = gp;
    }
    return x;
}

__global__ void uf_init_kernel(int* parent, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) parent[i] = i;
}

__global__ void uf_link_kernel(int* parent, const int32_t* __restrict__ edges, long long E) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int a = edges[2 * e], b = edges[2 * e + 1];
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) break;
        if (a < b) {
            int tmp = a;
            a = b;
            b = tmp;
        }
        // a > b: try to hook root a under b
        const int old = atomicCAS(parent + a, a, b);
        if (old == a) break;
    }
}

__global__ void uf_flatten_kernel(int* parent, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    int x = i;
    while (true) {
        const int p = ((volatile int*)parent)[x];
        if (p == x) break;
        x = p;
    }
    // roots never change in this kernel (no hooks), so writing the root is race-free in effect
    parent[i] = x;
}

extern "C" int f3d_box_pairs_aabb(const double* lo, const double* hi, const int32_t* group, int32_t B, int32_t* edges,
                                  int64_t cap, unsigned long long* count, void* stream) {
    if (!lo || !hi || !group || !count || B < 0 || cap < 0 || (cap > 0 && !edges))
        return f3d_fail(F3D_ERR_ARG, "f3d_box_pairs_aabb: bad argument");
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(unsigned long long), (cudaStream_t)stream);
    if (e != cudaSuccess) return f3d_check_launch("f3d_box_pairs_aabb(memset)");
    if (B < 2) return F3D_OK;
    const int ntiles = (B + BOX_TILE - 1) / BOX_TILE;
    const long long nblocks = (long long)ntiles * (ntiles + 1) / 2;
    if (nblocks > 0x7fffffffLL) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_box_pairs_aabb: too many boxes");
    box_pairs_kernel<<<(unsigned)nblocks, BOX_TILE, 0, (cudaStream_t)stream>>>(lo, hi, group, B, ntiles, edges, cap, count);
    return f3d_check_launch("f3d_box_pairs_aabb");
}

extern "C" int f3d_union_find(int32_t B, const int32_t* edges, int64_t E, int32_t* labels, void* stream) {
    if (!labels || B < 0 || E < 0 || (E > 0 && !edges)) return f3d_fail(F3D_ERR_ARG, "f3d_union_find: bad argument");
    if (B == 0) return F3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    uf_init_kernel<<<(B + 255) / 256, 256, 0, s>>>(labels, B);
    if (E > 0) {
        const long long blocks = (E + 255) / 256;
        if (blocks > 0x7fffffffLL) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_union_find: too many edges");
        uf_link_kernel<<<(unsigned)blocks, 256, 0, s>>>(labels, edges, E);
    }
    uf_flatten_kernel<<<(B + 255) / 256, 256, 0, s>>>(labels, B);
    return f3d_check_launch("f3d_union_find");
}
