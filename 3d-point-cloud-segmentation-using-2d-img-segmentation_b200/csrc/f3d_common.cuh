// Shared device structures and the fp64 "exact" arithmetic used by every libf3d kernel (sm_100a only).
//
// The exact helpers evaluate IEEE-754 binary64 operations one by one with round-to-nearest intrinsics
// (__dmul_rn / __dadd_rn / __ddiv_rn / __dsqrt_rn are never contracted into FMAs by nvcc), in the operation
// order fixed by oracle/f3d_oracle.py, which restates the reference's numpy statements:
//   quaternion sandwich    RTAB_utils/spatQuad.py:16-28
//   points2pixel           Fusion3DSeg/camera_utils.py:21-25
//   frustum set-up         Fusion3DSeg/camera_utils.py:60-171, Fusion3DSeg/fusion.py:119-132,254-258
//   cull                   Fusion3DSeg/intersections.py:157-163
//   depth back-projection  RTAB_utils/ios_rtab.py:164-173,185-191
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f3d.h"

#define F3D_U24 5.9604644775390625e-08f /* 2^-24, unit round-off of binary32 */

// ---- per-frame table ---------------------------------------------------------------------------------------
// One 656-byte record per frame in a caller-provided allocation:
//   FrameFast (128 B)  fp32 projection tile streamed by the fused kernel
//   FrameCull ( 80 B)  fp32 frustum planes for the conservative tile cull
//   FrameExact(448 B)  fp64 pose + planes for the exact path

struct __align__(16) FrameFast {
    float thi[3], ss;      // translation high part; |q|^2
    float tlo[3], far_d;   // translation low part (t - thi); far-plane distance along look-at
    float Mu[3], nu;       // row 0 of K * R^T / ss and its max |entry|
    float Mv[3], nv;       // row 1
    float Mz[3], nz;       // row 2
    float Rx[3], lwx;      // rows of R^T (metric camera axes) ; look-at unit vector (world)
    float Ry[3], lwy;
    float Rz[3], lwz;
};
static_assert(sizeof(FrameFast) == 128, "FrameFast must be 128 bytes");

struct __align__(16) FrameCull {
    float4 pl[5];          // (nx, ny, nz, n.a): dp = n.p - n.a ; inside <=> dp >= 0 for all five
};
static_assert(sizeof(FrameCull) == 80, "FrameCull must be 80 bytes");

struct __align__(16) FrameExact {
    double q[4];           // (w,x,y,z) as given
    double qi[4];          // pyquaternion inverse = conj / sum of squares
    double t[3];
    double ss;
    double plane_pt[5][3]; // 4 x eye, far point
    double plane_n[5][3];  // 4 inward face normals, -lookat
    double lookat[3];
    double pad[11];
};
static_assert(sizeof(FrameExact) == 448, "FrameExact must be 448 bytes");

// one record per frame, records contiguous: the table needs no header and can be sliced by frame index
struct __align__(16) FrameRecord {
    FrameFast fast;
    FrameCull cull;
    FrameExact exact;
};
static_assert(sizeof(FrameRecord) == 656, "FrameRecord must be 656 bytes");
#define F3D_FRAME_BYTES (sizeof(FrameRecord))

// ---- exact fp64 arithmetic -----------------------------------------------------------------------------------
struct D3 {
    double x, y, z;
};

__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }

// (a0*b0 + a1*b1) + a2*b2
__device__ __forceinline__ double ddot3(double a0, double a1, double a2, double b0, double b1, double b2) {
    return xadd(xadd(xmul(a0, b0), xmul(a1, b1)), xmul(a2, b2));
}

// np.cross component formulas
__device__ __forceinline__ D3 dcross(D3 a, D3 b) {
    D3 r;
    r.x = xsub(xmul(a.y, b.z), xmul(a.z, b.y));
    r.y = xsub(xmul(a.z, b.x), xmul(a.x, b.z));
    r.z = xsub(xmul(a.x, b.y), xmul(a.y, b.x));
    return r;
}

// SpatQuadranion.rotate (spatQuad.py:16-28) on un-normalised q = (w,x,y,z)
__device__ __forceinline__ D3 dquat_rotate(const double* q, D3 p) {
    const double rq = q[0];
    D3 v = {q[1], q[2], q[3]};
    D3 n = {-q[1], -q[2], -q[3]};
    double rqp = -ddot3(p.x, p.y, p.z, v.x, v.y, v.z);
    D3 c = dcross(v, p);
    D3 w = {xadd(xmul(rq, p.x), c.x), xadd(xmul(rq, p.y), c.y), xadd(xmul(rq, p.z), c.z)};
    D3 d = dcross(w, n);
    D3 o;
    o.x = xadd(xadd(xmul(rqp, n.x), xmul(rq, w.x)), d.x);
    o.y = xadd(xadd(xmul(rqp, n.y), xmul(rq, w.y)), d.y);
    o.z = xadd(xadd(xmul(rqp, n.z), xmul(rq, w.z)), d.z);
    return o;
}

// rows of K @ P (camera_utils.py:23): (K[i,0]*X + K[i,1]*Y) + K[i,2]*Z
__device__ __forceinline__ D3 dproject_h(const double* K, const double* qi, const double* t, D3 p) {
    D3 d = {xsub(p.x, t[0]), xsub(p.y, t[1]), xsub(p.z, t[2])};
    D3 c = dquat_rotate(qi, d);
    D3 h;
    h.x = xadd(xadd(xmul(K[0], c.x), xmul(K[1], c.y)), xmul(K[2], c.z));
    h.y = xadd(xadd(xmul(K[3], c.x), xmul(K[4], c.y)), xmul(K[5], c.z));
    h.z = xadd(xadd(xmul(K[6], c.x), xmul(K[7], c.y)), xmul(K[8], c.z));
    return h;
}

// point_inside_polyhedra for the five frame planes (intersections.py:157-163)
__device__ __forceinline__ bool dinside_planes(const FrameExact* fe, D3 p) {
    bool in = true;
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        double d0 = xsub(p.x, fe->plane_pt[m][0]);
        double d1 = xsub(p.y, fe->plane_pt[m][1]);
        double d2 = xsub(p.z, fe->plane_pt[m][2]);
        double dp = ddot3(d0, d1, d2, fe->plane_n[m][0], fe->plane_n[m][1], fe->plane_n[m][2]);
        in = in && (dp >= 0.0);
    }
    return in;
}

// numpy float64 -> int32 cast of an already floored value (x86 cvttsd2si semantics: out of range / NaN -> INT_MIN)
__device__ __forceinline__ int d2i_numpy(double x) {
    if (!(x >= -2147483648.0 && x < 2147483648.0)) return INT32_MIN;
    return __double2int_rz(x);
}

// launch-constant intrinsics in both precisions
struct Intrinsics {
    double K[9];
    float cx, cy, inv_fx, inv_fy;  // fast-path back-projection of a depth pixel (no skew assumed there; the
                                   // exact path uses K directly)
};



// ---- label resolve parameters shared by the fused epilogue, the fix-up kernels and the exchange merge ----------------
#define RES_MAXC 256
#define F3D_MAX_RANKS 16
struct FuseResolve {
    int enabled, nfilter;
    int32_t unclassified;
    double threshold;
    int16_t fpos[RES_MAXC];        // column -> first position in the filter list (or column itself), -1 = not considered
    int32_t remap[RES_MAXC];       // arg-max position -> label (sequential remap of voting.py:133-135 composed)
};
// host: composed sequential remap + column -> filter position table (fuse_project_vote.cu)
int f3d_build_resolve(int C1, double threshold, const int32_t* h_filter, int nfilter, int nclasses_id, FuseResolve& rp);

// ---- multi-GPU vote exchange layout constants (see FuseParams in fuse_project_vote.cu and vote_exchange.cu) ----------
#define F3D_XCH_NREG 1024      // record sub-regions per (source, owner): row cursors are spread so warps never queue on one
#define F3D_XCH_NSUB 2048      // (cell, count) sub-queues per (source, owner)
#define F3D_XCH_NSUB_FIX 1776  // the first sub-queues belong to the fix-up kernel's blocks (one each, no global atomics)
#define F3D_XCH_NLEVEL 4       // records per (source, block): one per flush of the byte histogram (235 candidate frames each)

// Eight lanes (sub = 0..7) stream one int32 vote row: all loads of a lane are issued before any is used (8-byte loads
// when the row allows it), partial (total, best, first position) per lane; the caller combines with three xor-shuffles.
__device__ __forceinline__ void row_partial8(const int32_t* __restrict__ r, int C1, const int16_t* __restrict__ s_fpos, int sub,
                                             long long& total, int& best, int& bpos) {
    if ((C1 & 1) == 0 && C1 <= 144 && (reinterpret_cast<uintptr_t>(r) & 7u) == 0) {
        const int n2 = C1 >> 1;
        int2 buf[9];
#pragma unroll
        for (int u = 0; u < 9; ++u) buf[u] = (sub + 8 * u < n2) ? __ldg(reinterpret_cast<const int2*>(r) + sub + 8 * u) : make_int2(0, 0);
#pragma unroll
        for (int u = 0; u < 9; ++u) {
            const int c = 2 * (sub + 8 * u);
            if ((buf[u].x | buf[u].y) == 0) continue;
            total += (long long)buf[u].x + (long long)buf[u].y;
            const int p0 = s_fpos[c], p1 = s_fpos[c + 1];
            if (buf[u].x > 0 && p0 >= 0 && (buf[u].x > best || (buf[u].x == best && p0 < bpos))) {
                best = buf[u].x;
                bpos = p0;
            }
            if (buf[u].y > 0 && p1 >= 0 && (buf[u].y > best || (buf[u].y == best && p1 < bpos))) {
                best = buf[u].y;
                bpos = p1;
            }
        }
    } else {
        for (int c = sub; c < C1; c += 8) {
            const int v = __ldg(r + c);
            total += v;
            const int pos = s_fpos[c];
            if (v > 0 && pos >= 0 && (v > best || (v == best && pos < bpos))) {
                best = v;
                bpos = pos;
            }
        }
    }
}
