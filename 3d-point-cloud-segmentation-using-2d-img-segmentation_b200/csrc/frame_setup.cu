// Per-frame table construction (a-3) and the single-call fp64 operators (a-1, a-4).
//   Fusion._get_frustum_data        Fusion3DSeg/fusion.py:119-132
//   get_camera_frustum              Fusion3DSeg/camera_utils.py:60-93
//   camera2world                    Fusion3DSeg/camera_utils.py:96-132
//   get_frustum_unit_vectors        Fusion3DSeg/camera_utils.py:135-150
//   get_frustum_face_normals        Fusion3DSeg/camera_utils.py:153-171
//   per-frame plane set             Fusion3DSeg/fusion.py:254-258
#include "f3d_common.cuh"
#include "f3d_host.h"

struct SetupParams {
    double K[9];
    double max_depth;
    int W, H, F;
};

__device__ __forceinline__ double dnorm3(D3 v) {
    return __dsqrt_rn(xadd(xadd(xmul(v.x, v.x), xmul(v.y, v.y)), xmul(v.z, v.z)));
}

__global__ void frames_setup_kernel(SetupParams sp, const double* __restrict__ wxyz,
                                    const double* __restrict__ trans, void* table) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= sp.F) return;
    FrameRecord* rec = reinterpret_cast<FrameRecord*>(table) + f;
    FrameFast* ff = &rec->fast;
    FrameCull* fc = &rec->cull;
    FrameExact* fe = &rec->exact;

    const double* K = sp.K;
    double q[4] = {wxyz[4 * f + 0], wxyz[4 * f + 1], wxyz[4 * f + 2], wxyz[4 * f + 3]};
    double t[3] = {trans[3 * f + 0], trans[3 * f + 1], trans[3 * f + 2]};
    // pyquaternion inverse: conj / sum of squares, ss = ((w*w + x*x) + y*y) + z*z
    double ss = xadd(xadd(xadd(xmul(q[0], q[0]), xmul(q[1], q[1])), xmul(q[2], q[2])), xmul(q[3], q[3]));
    double qi[4] = {xdiv(q[0], ss), xdiv(-q[1], ss), xdiv(-q[2], ss), xdiv(-q[3], ss)};

    // closed-form inverse of the upper-triangular intrinsic matrix (oracle.intrinsic_inverse)
    double i00 = xdiv(1.0, K[0]), i11 = xdiv(1.0, K[4]), i22 = xdiv(1.0, K[8]);
    double i01 = -xmul(xmul(K[1], i11), i00);
    double i12 = -xmul(xmul(K[5], i22), i11);
    double i02 = -xmul(xadd(xmul(K[1], i12), xmul(K[2], i22)), i00);
    double Ki[9] = {i00, i01, i02, 0.0, i11, i12, 0.0, 0.0, i22};

    const double w = (double)sp.W, h = (double)sp.H;
    const double pix[6][3] = {{0, 0, 0}, {0, 0, 1}, {w, 0, 1}, {w, h, 1}, {0, h, 1}, {w / 2, h / 2, 1}};
    D3 world[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        D3 c;
        c.x = xadd(xadd(xmul(Ki[0], pix[i][0]), xmul(Ki[1], pix[i][1])), xmul(Ki[2], pix[i][2]));
        c.y = xadd(xadd(xmul(Ki[3], pix[i][0]), xmul(Ki[4], pix[i][1])), xmul(Ki[5], pix[i][2]));
        c.z = xadd(xadd(xmul(Ki[6], pix[i][0]), xmul(Ki[7], pix[i][1])), xmul(Ki[8], pix[i][2]));
        D3 r = dquat_rotate(q, c);                                   // camera_utils.py:128
        world[i].x = xadd(r.x, t[0]);                                // camera_utils.py:129
        world[i].y = xadd(r.y, t[1]);
        world[i].z = xadd(r.z, t[2]);
    }
    D3 eye = world[0];
    D3 dirs[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {                                    // camera_utils.py:147-148
        D3 v = {xsub(world[i + 1].x, eye.x), xsub(world[i + 1].y, eye.y), xsub(world[i + 1].z, eye.z)};
        double n = dnorm3(v);
        dirs[i].x = xdiv(v.x, n);
        dirs[i].y = xdiv(v.y, n);
        dirs[i].z = xdiv(v.z, n);
    }
    D3 look = dirs[4];
    D3 normal[5];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                                    // camera_utils.py:163-170
        D3 a = world[1 + i], b = world[1 + ((i + 1) & 3)];
        D3 ea = {xsub(a.x, eye.x), xsub(a.y, eye.y), xsub(a.z, eye.z)};
        D3 eb = {xsub(b.x, eye.x), xsub(b.y, eye.y), xsub(b.z, eye.z)};
        D3 n = dcross(ea, eb);
        double nn = dnorm3(n);
        normal[i].x = xdiv(n.x, nn);
        normal[i].y = xdiv(n.y, nn);
        normal[i].z = xdiv(n.z, nn);
    }
    normal[4].x = -look.x;                                           // fusion.py:257
    normal[4].y = -look.y;
    normal[4].z = -look.z;
    D3 ppt[5] = {eye, eye, eye, eye,
                 {xadd(eye.x, xmul(sp.max_depth, look.x)), xadd(eye.y, xmul(sp.max_depth, look.y)),
                  xadd(eye.z, xmul(sp.max_depth, look.z))}};           // fusion.py:256

    // ---- exact section
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        fe->q[i] = q[i];
        fe->qi[i] = qi[i];
    }
    fe->t[0] = t[0];
    fe->t[1] = t[1];
    fe->t[2] = t[2];
    fe->ss = ss;
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        fe->plane_pt[m][0] = ppt[m].x;
        fe->plane_pt[m][1] = ppt[m].y;
        fe->plane_pt[m][2] = ppt[m].z;
        fe->plane_n[m][0] = normal[m].x;
        fe->plane_n[m][1] = normal[m].y;
        fe->plane_n[m][2] = normal[m].z;
    }
    fe->lookat[0] = look.x;
    fe->lookat[1] = look.y;
    fe->lookat[2] = look.z;
#pragma unroll
    for (int i = 0; i < 11; ++i) fe->pad[i] = 0.0;

    // ---- cull section: dp = n.p - n.a
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        double c = normal[m].x * ppt[m].x + normal[m].y * ppt[m].y + normal[m].z * ppt[m].z;
        fc->pl[m] = make_float4((float)normal[m].x, (float)normal[m].y, (float)normal[m].z, (float)c);
    }

    // ---- fast section.  Rotation of the normalised quaternion (camera -> world), R^T rows = camera axes.
    double inv = 1.0 / sqrt(ss);
    double qw = q[0] * inv, qx = q[1] * inv, qy = q[2] * inv, qz = q[3] * inv;
    double R[3][3] = {{1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)},
                      {2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)},
                      {2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)}};
    double Rt[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rt[i][j] = R[j][i];
    double M[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            M[i][j] = (K[3 * i + 0] * Rt[0][j] + K[3 * i + 1] * Rt[1][j] + K[3 * i + 2] * Rt[2][j]) / ss;
    for (int j = 0; j < 3; ++j) {
        float hi = (float)t[j];
        ff->thi[j] = hi;
        ff->tlo[j] = (float)(t[j] - (double)hi);
        ff->Mu[j] = (float)M[0][j];
        ff->Mv[j] = (float)M[1][j];
        ff->Mz[j] = (float)M[2][j];
        ff->Rx[j] = (float)Rt[0][j];
        ff->Ry[j] = (float)Rt[1][j];
        ff->Rz[j] = (float)Rt[2][j];
    }
    ff->ss = (float)ss;
    ff->far_d = (float)sp.max_depth;
    ff->nu = fmaxf(fmaxf(fabsf(ff->Mu[0]), fabsf(ff->Mu[1])), fabsf(ff->Mu[2]));
    ff->nv = fmaxf(fmaxf(fabsf(ff->Mv[0]), fabsf(ff->Mv[1])), fabsf(ff->Mv[2]));
    ff->nz = fmaxf(fmaxf(fabsf(ff->Mz[0]), fabsf(ff->Mz[1])), fabsf(ff->Mz[2]));
    ff->lwx = (float)look.x;
    ff->lwy = (float)look.y;
    ff->lwz = (float)look.z;
}

__global__ void frames_export_kernel(const void* table, int F, double* eyes, double* lookats, double* normals) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const FrameExact* fe = &(reinterpret_cast<const FrameRecord*>(table) + f)->exact;
    for (int j = 0; j < 3; ++j) {
        eyes[3 * f + j] = fe->plane_pt[0][j];
        lookats[3 * f + j] = fe->lookat[j];
        for (int m = 0; m < 4; ++m) normals[(4 * f + m) * 3 + j] = fe->plane_n[m][j];
    }
}

// ---- points2pixel, whole array, fp64 (camera_utils.py:9-26) -------------------------------------------------
struct ProjParams {
    double K[9];
    double q[4];
    double t[3];
};

__global__ void project_pixels_kernel(const double* __restrict__ pts, int64_t N, ProjParams pp, int32_t* __restrict__ uv) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    D3 p = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    const double* q = pp.q;   // pyquaternion inverse, same order as frames_setup_kernel
    double ss = xadd(xadd(xadd(xmul(q[0], q[0]), xmul(q[1], q[1])), xmul(q[2], q[2])), xmul(q[3], q[3]));
    double qi[4] = {xdiv(q[0], ss), xdiv(-q[1], ss), xdiv(-q[2], ss), xdiv(-q[3], ss)};
    D3 h = dproject_h(pp.K, qi, pp.t, p);
    uv[i] = d2i_numpy(floor(xdiv(h.x, h.z)));
    uv[N + i] = d2i_numpy(floor(xdiv(h.y, h.z)));
}

// ---- SpatQuadranion.rotate, whole array, fp64 (RTAB_utils/spatQuad.py:6-28) -------------------------------------
struct QuatParam {
    double q[4];
};
__global__ void quat_rotate_kernel(const double* __restrict__ pts, int64_t N, QuatParam qp, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    D3 p = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    D3 r = dquat_rotate(qp.q, p);
    out[3 * i] = r.x;
    out[3 * i + 1] = r.y;
    out[3 * i + 2] = r.z;
}

// ---- point_inside_polyhedra, whole array, fp64 (intersections.py:146-164) ----------------------------------
#define F3D_MAX_PLANES 16
struct PlaneParams {
    double pt[F3D_MAX_PLANES][3];
    double n[F3D_MAX_PLANES][3];
    int M;
};

__global__ void frustum_mask_kernel(const double* __restrict__ pts, int64_t N, PlaneParams pl, uint8_t* __restrict__ inside) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    bool in = true;
    for (int m = 0; m < pl.M; ++m) {
        double dp = ddot3(xsub(x, pl.pt[m][0]), xsub(y, pl.pt[m][1]), xsub(z, pl.pt[m][2]), pl.n[m][0], pl.n[m][1], pl.n[m][2]);
        in = in && (dp >= 0.0);
    }
    inside[i] = in ? 1 : 0;
}

// ---- OBB membership (Open3D rule; merge_intersecting_bb.py:76,87) ------------------------------------------
__global__ void obb_contains_kernel(const double* __restrict__ pts, int64_t N, const double* __restrict__ boxes,
                                    int nboxes, uint8_t* __restrict__ inside) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (i >= N) return;
    const double* bx = boxes + 15 * b;   // centre[3], R[9] row-major, extent[3]
    double d0 = xsub(pts[3 * i], bx[0]), d1 = xsub(pts[3 * i + 1], bx[1]), d2 = xsub(pts[3 * i + 2], bx[2]);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double proj = ddot3(d0, d1, d2, bx[3 + k], bx[3 + 3 + k], bx[3 + 6 + k]);
        ok = ok && (fabs(proj) <= xdiv(bx[12 + k], 2.0));
    }
    inside[(int64_t)b * N + i] = ok ? 1 : 0;
}

// ---- C ABI -----------------------------------------------------------------------------------------------------------
extern "C" int64_t f3d_frame_table_bytes(int32_t nframes) { return (int64_t)nframes * (int64_t)F3D_FRAME_BYTES; }

extern "C" int f3d_frames_setup(const double* h_K9, int32_t W, int32_t H, const double* wxyz, const double* trans,
                                int32_t nframes, double max_depth, void* frame_table, void* stream) {
    if (!h_K9 || !wxyz || !trans || !frame_table || nframes <= 0 || W <= 0 || H <= 0)
        return f3d_fail(F3D_ERR_ARG, "f3d_frames_setup: bad argument");
    if (h_K9[3] != 0.0 || h_K9[6] != 0.0 || h_K9[7] != 0.0)
        return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_frames_setup: intrinsic matrix must be upper triangular");
    SetupParams sp;
    for (int i = 0; i < 9; ++i) sp.K[i] = h_K9[i];
    sp.max_depth = max_depth;
    sp.W = W;
    sp.H = H;
    sp.F = nframes;
    frames_setup_kernel<<<(nframes + 63) / 64, 64, 0, (cudaStream_t)stream>>>(sp, wxyz, trans, frame_table);
    return f3d_check_launch("f3d_frames_setup");
}

extern "C" int f3d_frames_export(const void* frame_table, int32_t nframes, double* eyes, double* lookats,
                                 double* face_normals, void* stream) {
    if (!frame_table || nframes <= 0 || !eyes || !lookats || !face_normals)
        return f3d_fail(F3D_ERR_ARG, "f3d_frames_export: bad argument");
    frames_export_kernel<<<(nframes + 63) / 64, 64, 0, (cudaStream_t)stream>>>(frame_table, nframes, eyes, lookats,
                                                                               face_normals);
    return f3d_check_launch("f3d_frames_export");
}

extern "C" int f3d_project_pixels(const double* points, int64_t N, const double* h_K9, const double* h_wxyz,
                                  const double* h_t, int32_t* uv, void* stream) {
    if (!points || !h_K9 || !h_wxyz || !h_t || !uv || N < 0) return f3d_fail(F3D_ERR_ARG, "f3d_project_pixels: bad argument");
    if (N == 0) return F3D_OK;
    ProjParams pp;
    for (int i = 0; i < 9; ++i) pp.K[i] = h_K9[i];
    for (int i = 0; i < 4; ++i) pp.q[i] = h_wxyz[i];
    for (int i = 0; i < 3; ++i) pp.t[i] = h_t[i];
    unsigned blocks = (unsigned)((N + 255) / 256);
    project_pixels_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(points, N, pp, uv);
    return f3d_check_launch("f3d_project_pixels");
}

extern "C" int f3d_quat_rotate(const double* points, int64_t N, const double* h_wxyz, double* out, void* stream) {
    if (!points || !h_wxyz || !out || N < 0) return f3d_fail(F3D_ERR_ARG, "f3d_quat_rotate: bad argument");
    if (N == 0) return F3D_OK;
    QuatParam qp;
    for (int i = 0; i < 4; ++i) qp.q[i] = h_wxyz[i];
    quat_rotate_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, N, qp, out);
    return f3d_check_launch("f3d_quat_rotate");
}

extern "C" int f3d_frustum_mask(const double* points, int64_t N, const double* h_plane_points, const double* h_normals,
                                int32_t nplanes, uint8_t* inside, void* stream) {
    if (!points || !h_plane_points || !h_normals || !inside || N < 0 || nplanes <= 0)
        return f3d_fail(F3D_ERR_ARG, "f3d_frustum_mask: bad argument");
    if (nplanes > F3D_MAX_PLANES) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_frustum_mask: more than 16 planes");
    if (N == 0) return F3D_OK;
    PlaneParams pl;
    pl.M = nplanes;
    for (int m = 0; m < nplanes; ++m)
        for (int j = 0; j < 3; ++j) {
            pl.pt[m][j] = h_plane_points[3 * m + j];
            pl.n[m][j] = h_normals[3 * m + j];
        }
    unsigned blocks = (unsigned)((N + 255) / 256);
    frustum_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(points, N, pl, inside);
    return f3d_check_launch("f3d_frustum_mask");
}

extern "C" int f3d_obb_contains(const double* points, int64_t N, const double* boxes15, int32_t nboxes, uint8_t* inside,
                                void* stream) {
    if (!points || !boxes15 || !inside || N < 0 || nboxes <= 0 || nboxes > 65535)
        return f3d_fail(F3D_ERR_ARG, "f3d_obb_contains: bad argument");
    if (N == 0) return F3D_OK;
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)nboxes);
    obb_contains_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, N, boxes15, nboxes, inside);
    return f3d_check_launch("f3d_obb_contains");
}
