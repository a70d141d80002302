// Geometry kernels either side of the fused path (SURVEY 8(f) rank 3 and kernel (4)'s broad phase):
//   * fixed-radius neighbour search on a uniform grid -> CSR adjacency: KDTree(points).query_radius(points, r=2*radius),
//     Fusion3DSeg/fusion.py:369-377 (scikit-learn's Euclidean leaf test: reduced distance (dx*dx + dy*dy) + dz*dz <= r*r);
//   * batched oriented-box fit of every instance in one pass over the cloud (segmented mean / covariance, 3x3 Jacobi,
//     projection range): the box behind the o3d.geometry.OrientedBoundingBox.create_from_points call sites
//     merge_intersecting_bb.py:18,75,86,126 and get3DSeg.py:434-436 (box MODEL stated in oracle.fit_box: covariance of all
//     points, or axis aligned -- Open3D's hull-based fit itself is unpinned);
//   * sort-and-sweep broad phase of the closed-interval AABB pair predicate (merge_intersecting_bb.py:49-53) over boxes
//     ordered by (group, lo.x): work proportional to the x-overlaps inside a group instead of B^2 / 2 pair tests.
#include "f3d_common.cuh"
#include "f3d_host.h"

// ---- (A) AABB pairs, sort-and-sweep -----------------------------------------------------------------------------------
__device__ __forceinline__ bool axis_overlap_closed(double lo1, double hi1, double lo2, double hi2) {
    return (lo1 <= lo2 && lo2 <= hi1) || (lo2 <= lo1 && lo1 <= hi2);   // merge_intersecting_bb.py:51-53
}

// order[] sorts the boxes by (group, lo.x).  Thread p owns the box a = order[p] and walks the later boxes b of its group
// while lo.x[b] <= hi.x[a] (then (min1 <= min2 <= max1) can hold on x) or lo.x[b] == lo.x[a] (ties: (min2 <= min1 <= max2)
// can hold even for a degenerate box a); the exact float64 predicate decides.  A pair is found exactly once, from the box
// that comes first in the order.
__global__ void __launch_bounds__(128) box_sweep_kernel(const double* __restrict__ lo, const double* __restrict__ hi,
                                                        const int32_t* __restrict__ group, const int32_t* __restrict__ order, int B,
                                                        int32_t* __restrict__ edges, long long cap, unsigned long long* __restrict__ count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const int a = order[p];
    const int ga = group[a];
    const double alo0 = lo[3 * (size_t)a], alo1 = lo[3 * (size_t)a + 1], alo2 = lo[3 * (size_t)a + 2];
    const double ahi0 = hi[3 * (size_t)a], ahi1 = hi[3 * (size_t)a + 1], ahi2 = hi[3 * (size_t)a + 2];
    for (int q = p + 1; q < B; ++q) {
        const int b = order[q];
        if (group[b] != ga) break;
        const double blo0 = lo[3 * (size_t)b];
        if (!(blo0 <= ahi0) && !(blo0 == alo0)) break;
        if (axis_overlap_closed(alo0, ahi0, blo0, hi[3 * (size_t)b]) &&
            axis_overlap_closed(alo1, ahi1, lo[3 * (size_t)b + 1], hi[3 * (size_t)b + 1]) &&
            axis_overlap_closed(alo2, ahi2, lo[3 * (size_t)b + 2], hi[3 * (size_t)b + 2])) {
            const unsigned long long e = atomicAdd(count, 1ULL);
            if ((long long)e < cap) {
                edges[2 * e] = min(a, b);
                edges[2 * e + 1] = max(a, b);
            }
        }
    }
}

extern "C" int f3d_box_pairs_sweep(const double* lo, const double* hi, const int32_t* group, const int32_t* order, int32_t B,
                                   int32_t* edges, int64_t cap, unsigned long long* count, void* stream) {
    if (!lo || !hi || !group || !order || !count || B < 0 || cap < 0 || (cap > 0 && !edges))
        return f3d_fail(F3D_ERR_ARG, "f3d_box_pairs_sweep: bad argument");
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(unsigned long long), (cudaStream_t)stream);
    if (e != cudaSuccess) return f3d_check_launch("f3d_box_pairs_sweep(memset)");
    if (B < 2) return F3D_OK;
    box_sweep_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(lo, hi, group, order, B, edges, cap, count);
    return f3d_check_launch("f3d_box_pairs_sweep");
}

// ---- (B) batched box fit ----------------------------------------------------------------------------------------------
// workspace per instance slot: [0] count, [1..3] sum / mean, [4..9] covariance sums (xx, xy, xz, yy, yz, zz),
// [10..15] projection min[3] / max[3] as order-preserving int64 keys
#define OBB_WS 16

__device__ __forceinline__ long long dkey(double x) {   // monotone double -> int64
    const long long b = __double_as_longlong(x);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double dunkey(long long k) { return __longlong_as_double(k >= 0 ? k : (k ^ 0x7fffffffffffffffLL)); }

__device__ __forceinline__ int obb_slot(const int64_t* __restrict__ ids, const int32_t* __restrict__ slot_of_id, int64_t nslot, int64_t i) {
    const int64_t id = ids[i];
    return (id >= 0 && id < nslot) ? slot_of_id[id] : -1;
}

// warp-aggregated segmented add: lanes of a warp that carry the same slot are combined first (a Morton-sorted cloud puts an
// instance's points next to each other), one lane per distinct slot issues the atomics
template <int NV>
__device__ __forceinline__ void seg_add(double* __restrict__ ws, int slot, const double (&v)[NV], int base) {
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, slot);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = v[k];
    // tree reduction over the peer set (fixed order for a given lane set)
    for (unsigned rest = peers & ~(1u << leader); rest; rest &= rest - 1u) {
        const int src = __ffs(rest) - 1;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const double o = __shfl_sync(peers, v[k], src);
            if (lane == leader) acc[k] += o;
        }
    }
    if (lane == leader && slot >= 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) atomicAdd(ws + (size_t)slot * OBB_WS + base + k, acc[k]);
    }
}

__global__ void obb_sum_kernel(const double* __restrict__ pts, const int64_t* __restrict__ ids, int64_t N,
                               const int32_t* __restrict__ slot_of_id, int64_t nslot, double* __restrict__ ws) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int slot = -1;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < N) {
        slot = obb_slot(ids, slot_of_id, nslot, i);
        if (slot >= 0) {
            v[0] = 1.0;
            v[1] = pts[3 * i];
            v[2] = pts[3 * i + 1];
            v[3] = pts[3 * i + 2];
        }
    }
    seg_add<4>(ws, slot, v, 0);
}

__global__ void obb_mean_kernel(double* __restrict__ ws, int ninst) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ninst) return;
    double* w = ws + (size_t)s * OBB_WS;
    const double n = w[0];
    if (n > 0.0) {
        w[1] /= n;
        w[2] /= n;
        w[3] /= n;
    }
    for (int k = 0; k < 3; ++k) {
        reinterpret_cast<long long*>(w)[10 + k] = dkey(1.0e300);
        reinterpret_cast<long long*>(w)[13 + k] = dkey(-1.0e300);
    }
}

__global__ void obb_cov_kernel(const double* __restrict__ pts, const int64_t* __restrict__ ids, int64_t N,
                               const int32_t* __restrict__ slot_of_id, int64_t nslot, double* __restrict__ ws) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int slot = -1;
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (i < N) {
        slot = obb_slot(ids, slot_of_id, nslot, i);
        if (slot >= 0) {
            const double* w = ws + (size_t)slot * OBB_WS;
            const double x = pts[3 * i] - w[1], y = pts[3 * i + 1] - w[2], z = pts[3 * i + 2] - w[3];
            v[0] = x * x; v[1] = x * y; v[2] = x * z; v[3] = y * y; v[4] = y * z; v[5] = z * z;
        }
    }
    seg_add<6>(ws, slot, v, 4);
}

// cyclic Jacobi on the 3x3 covariance; axes sorted by descending eigenvalue, third axis = first x second (right handed)
__global__ void obb_axes_kernel(const double* __restrict__ ws, int ninst, int model, double* __restrict__ boxes) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ninst) return;
    const double* w = ws + (size_t)s * OBB_WS;
    double* bx = boxes + (size_t)s * 15;
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (model == 0 && w[0] > 0.0) {
        const double nm1 = fmax(w[0] - 1.0, 1.0);
        double A[3][3] = {{w[4] / nm1, w[5] / nm1, w[6] / nm1}, {w[5] / nm1, w[7] / nm1, w[8] / nm1}, {w[6] / nm1, w[8] / nm1, w[9] / nm1}};
        for (int sweep = 0; sweep < 64; ++sweep) {
            const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
            const double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
            if (off <= 1.0e-60 + 1.0e-32 * diag) break;
            for (int p = 0; p < 2; ++p)
                for (int q = p + 1; q < 3; ++q) {
                    if (A[p][q] == 0.0) continue;
                    const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                    for (int k = 0; k < 3; ++k) {   // A <- A J
                        const double akp = A[k][p], akq = A[k][q];
                        A[k][p] = c * akp - sn * akq;
                        A[k][q] = sn * akp + c * akq;
                    }
                    for (int k = 0; k < 3; ++k) {   // A <- J^T A
                        const double apk = A[p][k], aqk = A[q][k];
                        A[p][k] = c * apk - sn * aqk;
                        A[q][k] = sn * apk + c * aqk;
                    }
                    for (int k = 0; k < 3; ++k) {
                        const double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = c * vkp - sn * vkq;
                        V[k][q] = sn * vkp + c * vkq;
                    }
                }
        }
        int o0 = 0, o1 = 1, o2 = 2;
        double e0 = A[0][0], e1 = A[1][1], e2 = A[2][2];
        if (e1 > e0) { int ti = o0; o0 = o1; o1 = ti; double td = e0; e0 = e1; e1 = td; }
        if (e2 > e0) { int ti = o0; o0 = o2; o2 = ti; double td = e0; e0 = e2; e2 = td; }
        if (e2 > e1) { int ti = o1; o1 = o2; o2 = ti; }
        double R[3][3];
        for (int k = 0; k < 3; ++k) {
            R[k][0] = V[k][o0];
            R[k][1] = V[k][o1];
        }
        R[0][2] = R[1][0] * R[2][1] - R[2][0] * R[1][1];
        R[1][2] = R[2][0] * R[0][1] - R[0][0] * R[2][1];
        R[2][2] = R[0][0] * R[1][1] - R[1][0] * R[0][1];
        for (int r = 0; r < 3; ++r)
            for (int c2 = 0; c2 < 3; ++c2) V[r][c2] = R[r][c2];
    }
    for (int r = 0; r < 3; ++r)
        for (int c2 = 0; c2 < 3; ++c2) bx[3 + 3 * r + c2] = V[r][c2];
}

__global__ void obb_range_kernel(const double* __restrict__ pts, const int64_t* __restrict__ ids, int64_t N,
                                 const int32_t* __restrict__ slot_of_id, int64_t nslot, const double* __restrict__ boxes,
                                 double* __restrict__ ws) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int slot = -1;
    double pr[3] = {0.0, 0.0, 0.0};
    if (i < N) {
        slot = obb_slot(ids, slot_of_id, nslot, i);
        if (slot >= 0) {
            const double* w = ws + (size_t)slot * OBB_WS;
            const double* R = boxes + (size_t)slot * 15 + 3;
            const double x = pts[3 * i] - w[1], y = pts[3 * i + 1] - w[2], z = pts[3 * i + 2] - w[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) pr[k] = (x * R[k] + y * R[3 + k]) + z * R[6 + k];   // q @ R, column k
        }
    }
    // min / max are order independent: warp-aggregate over equal slots, then one 64-bit integer atomic per value
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, slot);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    double mn[3] = {pr[0], pr[1], pr[2]}, mx[3] = {pr[0], pr[1], pr[2]};
    for (unsigned rest = peers & ~(1u << leader); rest; rest &= rest - 1u) {
        const int src = __ffs(rest) - 1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double o = __shfl_sync(peers, pr[k], src);
            mn[k] = fmin(mn[k], o);
            mx[k] = fmax(mx[k], o);
        }
    }
    if (lane == leader && slot >= 0) {
        long long* wk = reinterpret_cast<long long*>(ws + (size_t)slot * OBB_WS);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(wk + 10 + k, dkey(mn[k]));
            atomicMax(wk + 13 + k, dkey(mx[k]));
        }
    }
}

__global__ void obb_finish_kernel(const double* __restrict__ ws, int ninst, double* __restrict__ boxes, int64_t* __restrict__ counts) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ninst) return;
    const double* w = ws + (size_t)s * OBB_WS;
    const long long* wk = reinterpret_cast<const long long*>(w);
    double* bx = boxes + (size_t)s * 15;
    counts[s] = (int64_t)w[0];
    if (w[0] <= 0.0) {
        for (int k = 0; k < 15; ++k) bx[k] = 0.0;
        return;
    }
    double mid[3];
    for (int k = 0; k < 3; ++k) {
        const double mn = dunkey(wk[10 + k]), mx = dunkey(wk[13 + k]);
        mid[k] = (mn + mx) * 0.5;
        bx[12 + k] = mx - mn;
    }
    const double* R = bx + 3;
    for (int r = 0; r < 3; ++r) bx[r] = w[1 + r] + ((R[3 * r] * mid[0] + R[3 * r + 1] * mid[1]) + R[3 * r + 2] * mid[2]);   // mean + R @ mid
}

extern "C" int64_t f3d_obb_fit_workspace_bytes(int32_t ninst) { return (int64_t)(ninst > 0 ? ninst : 1) * OBB_WS * 8; }

extern "C" int f3d_obb_fit(const double* points, const int64_t* ids, int64_t N, const int32_t* slot_of_id, int64_t nslot,
                           int32_t ninst, int32_t model, double* boxes15, int64_t* counts, void* workspace, void* stream) {
    if (!points || !ids || !slot_of_id || !boxes15 || !counts || !workspace || N < 0 || nslot < 0 || ninst < 0 || (model != 0 && model != 1))
        return f3d_fail(F3D_ERR_ARG, "f3d_obb_fit: bad argument");
    if (ninst == 0) return F3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    double* ws = reinterpret_cast<double*>(workspace);
    if (cudaMemsetAsync(ws, 0, (size_t)ninst * OBB_WS * 8, s) != cudaSuccess) return f3d_check_launch("f3d_obb_fit(memset)");
    const unsigned pb = (unsigned)((N + 255) / 256), ib = (unsigned)((ninst + 127) / 128);
    if ((N + 255) / 256 > 0x7fffffffLL) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_obb_fit: too many points");
    if (N > 0) obb_sum_kernel<<<pb, 256, 0, s>>>(points, ids, N, slot_of_id, nslot, ws);
    obb_mean_kernel<<<ib, 128, 0, s>>>(ws, ninst);
    if (N > 0 && model == 0) obb_cov_kernel<<<pb, 256, 0, s>>>(points, ids, N, slot_of_id, nslot, ws);
    obb_axes_kernel<<<ib, 128, 0, s>>>(ws, ninst, model, boxes15);
    if (N > 0) obb_range_kernel<<<pb, 256, 0, s>>>(points, ids, N, slot_of_id, nslot, boxes15, ws);
    obb_finish_kernel<<<ib, 128, 0, s>>>(ws, ninst, boxes15, counts);
    return f3d_check_launch("f3d_obb_fit");
}

// ---- (C) fixed-radius neighbours on a uniform grid -> CSR -------------------------------------------------------------
struct GridParams {
    double org[3];       // component-wise minimum of the cloud
    double cell;         // cell size = r
    long long d1, d2;    // key = ((cx + 1) * d1 + (cy + 1)) * d2 + (cz + 1), d = max cell index + 3 (one guard cell each side)
    double r2;
};

__global__ void grid_keys_kernel(const double* __restrict__ pts, int64_t N, GridParams g, int64_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const long long cx = (long long)floor(xdiv(xsub(pts[3 * i], g.org[0]), g.cell));
    const long long cy = (long long)floor(xdiv(xsub(pts[3 * i + 1], g.org[1]), g.cell));
    const long long cz = (long long)floor(xdiv(xsub(pts[3 * i + 2], g.org[2]), g.cell));
    keys[i] = ((cx + 1) * g.d1 + (cy + 1)) * g.d2 + (cz + 1);
}

__device__ __forceinline__ int64_t lower_bound_key(const int64_t* __restrict__ a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// One thread per query point (sorted position p, point order[p]).  The three z-neighbour cells of a (dx, dy) column have
// consecutive keys, so nine key ranges cover the 27 cells.  FILL = false: count the neighbours; FILL = true: write them
// at indptr[i], then sort the row ascending (rows are a few dozen entries: insertion sort by the owning thread).
template <bool FILL>
__global__ void __launch_bounds__(128) radius_rows_kernel(const double* __restrict__ pts, int64_t N, GridParams g,
                                                          const int64_t* __restrict__ skeys, const int64_t* __restrict__ order,
                                                          int64_t* __restrict__ counts, const int64_t* __restrict__ indptr,
                                                          int64_t* __restrict__ indices) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int64_t i = order[p];
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    const int64_t key = skeys[p];
    int64_t n = 0;
    int64_t* __restrict__ row = FILL ? indices + indptr[i] : nullptr;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy) {
            const int64_t k0 = key + ((int64_t)dx * g.d1 + dy) * g.d2 - 1;
            int64_t a = lower_bound_key(skeys, N, k0);
            for (; a < N; ++a) {
                if (__ldg(skeys + a) > k0 + 2) break;
                const int64_t j = __ldg(order + a);
                const double ex = xsub(x, pts[3 * j]), ey = xsub(y, pts[3 * j + 1]), ez = xsub(z, pts[3 * j + 2]);
                const double d2 = xadd(xadd(xmul(ex, ex), xmul(ey, ey)), xmul(ez, ez));
                if (d2 <= g.r2) {
                    if (FILL) row[n] = j;
                    ++n;
                }
            }
        }
    if (!FILL) {
        counts[i] = n;
    } else {
        for (int64_t a = 1; a < n; ++a) {
            const int64_t v = row[a];
            int64_t b = a - 1;
            while (b >= 0 && row[b] > v) {
                row[b + 1] = row[b];
                --b;
            }
            row[b + 1] = v;
        }
    }
}

static int grid_params(const double* h_org3, const double* h_max3, double r, GridParams& g) {
    if (!h_org3 || !h_max3 || !(r > 0.0)) return f3d_fail(F3D_ERR_ARG, "radius adjacency: bad argument");
    long long d[3];
    for (int k = 0; k < 3; ++k) {
        g.org[k] = h_org3[k];
        volatile double span = h_max3[k] - h_org3[k];
        volatile double q = span / r;
        d[k] = (long long)floor(q) + 3;
    }
    if ((double)d[0] * (double)d[1] * (double)d[2] > 4.0e18) return f3d_fail(F3D_ERR_UNSUPPORTED, "radius adjacency: grid too fine for 63-bit keys");
    g.cell = r;
    g.d1 = d[1];
    g.d2 = d[2];
    volatile double rr = r * r;
    g.r2 = rr;
    return F3D_OK;
}

extern "C" int f3d_radius_grid_keys(const double* points, int64_t N, const double* h_min3, const double* h_max3, double r,
                                    int64_t* keys, void* stream) {
    GridParams g;
    int rc = grid_params(h_min3, h_max3, r, g);
    if (rc) return rc;
    if (!points || !keys || N < 0) return f3d_fail(F3D_ERR_ARG, "f3d_radius_grid_keys: bad argument");
    if (N == 0) return F3D_OK;
    grid_keys_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, N, g, keys);
    return f3d_check_launch("f3d_radius_grid_keys");
}

extern "C" int f3d_radius_adjacency(const double* points, int64_t N, const double* h_min3, const double* h_max3, double r,
                                    const int64_t* sorted_keys, const int64_t* order, int64_t* counts, const int64_t* indptr,
                                    int64_t* indices, void* stream) {
    GridParams g;
    int rc = grid_params(h_min3, h_max3, r, g);
    if (rc) return rc;
    if (!points || !sorted_keys || !order || N < 0 || (!counts && !(indptr && indices)))
        return f3d_fail(F3D_ERR_ARG, "f3d_radius_adjacency: bad argument");
    if (N == 0) return F3D_OK;
    const unsigned blocks = (unsigned)((N + 127) / 128);
    if (indices) radius_rows_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(points, N, g, sorted_keys, order, nullptr, indptr, indices);
    else radius_rows_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(points, N, g, sorted_keys, order, counts, nullptr, nullptr);
    return f3d_check_launch("f3d_radius_adjacency");
}
