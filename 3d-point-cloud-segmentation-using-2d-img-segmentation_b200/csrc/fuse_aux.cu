// uv2pt writer (the reference's exchange format) and the point-stationary z-buffer splat: C ABI.  Device code: fuse_kernel.cuh.
#include "fuse_kernel.cuh"

__global__ void zbuf_finalize_kernel(const uint32_t* __restrict__ zbuf, uint16_t* __restrict__ out, int64_t total, int H,
                                     int W, int border) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint32_t z = zbuf[i];
    int pix = (int)(i % ((int64_t)H * W));
    int y = pix / W, x = pix - y * W;
    bool edge = (x < border) || (y < border) || (x >= W - border) || (y >= H - border);
    out[i] = (z == 0xffffffffu || edge) ? (uint16_t)0 : (uint16_t)z;
}

extern "C" int f3d_fuse_uv2pt(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                              int32_t frame_end, const void* depth, int32_t depth_fmt, int32_t H, int32_t W,
                              const double* h_K9, double radius, double zmin, double zmax, int32_t* uv2pt,
                              void* workspace, int64_t workspace_bytes, uint64_t* stats, int32_t flags, void* stream) {
    FuseParams P;
    int rc = fill_common(P, points, N, frame_table, frame_begin, frame_end, depth, depth_fmt, H, W, h_K9, radius, zmin,
                         zmax, stats);
    if (rc) return rc;
    if (!uv2pt || !depth) return f3d_fail(F3D_ERR_ARG, "f3d_fuse_uv2pt: bad argument");
    if (N == 0 || frame_end == frame_begin) return F3D_OK;
    if (N > 0x7fffffff) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_fuse_uv2pt: point index does not fit int32");
    const int audit = flags & 1;
    attach_workspace(P, workspace, workspace_bytes, !audit);
    FuseResolve RP;
    RP.enabled = 0;
    const size_t esz = frame_elem_bytes(depth_fmt);
    for (int fb = frame_begin; fb < frame_end; fb += F3D_MAX_FRAMES_PER_LAUNCH) {
        int fe = frame_end - fb > F3D_MAX_FRAMES_PER_LAUNCH ? fb + F3D_MAX_FRAMES_PER_LAUNCH : frame_end;
        P.f_begin = fb;
        P.f_end = fe;
        P.depth = reinterpret_cast<const char*>(depth) + (size_t)(fb - frame_begin) * (size_t)P.frame_stride * esz;
        P.uv2pt = uv2pt + (size_t)(fb - frame_begin) * H * W;
        switch (depth_fmt) {
            case F3D_DEPTH_U16_MM: rc = launch_fuse<MODE_UV2PT, F3D_DEPTH_U16_MM>(P, RP, audit, (cudaStream_t)stream); break;
            case F3D_DEPTH_F32_M: rc = launch_fuse<MODE_UV2PT, F3D_DEPTH_F32_M>(P, RP, audit, (cudaStream_t)stream); break;
            case F3D_FRAMES_U32: rc = launch_fuse<MODE_UV2PT, F3D_FRAMES_U32>(P, RP, audit, (cudaStream_t)stream); break;
            default: rc = launch_fuse<MODE_UV2PT, F3D_FRAMES_U32_T16>(P, RP, audit, (cudaStream_t)stream); break;
        }
        if (rc) return rc;
    }
    return F3D_OK;
}

extern "C" int f3d_zbuffer_splat(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                                 int32_t frame_end, int32_t H, int32_t W, const double* h_K9, uint32_t* zbuf,
                                 uint16_t* depth_out, int32_t border, void* workspace, int64_t workspace_bytes,
                                 uint64_t* stats, int32_t flags, void* stream) {
    FuseParams P;
    int rc = fill_common(P, points, N, frame_table, frame_begin, frame_end, nullptr, F3D_DEPTH_U16_MM, H, W, h_K9, 0.0,
                         0.0, 0.0, stats);
    if (rc) return rc;
    if (!zbuf || !depth_out || border < 0) return f3d_fail(F3D_ERR_ARG, "f3d_zbuffer_splat: bad argument");
    const int audit = flags & 1;
    attach_workspace(P, workspace, workspace_bytes, !audit && N <= 0x7fffffff);
    const int nf = frame_end - frame_begin;
    if (nf == 0) return F3D_OK;
    const int64_t total = (int64_t)nf * H * W;
    cudaError_t e = cudaMemsetAsync(zbuf, 0xff, (size_t)total * sizeof(uint32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) return f3d_check_launch("f3d_zbuffer_splat(memset)");
    FuseResolve RP;
    RP.enabled = 0;
    if (N > 0) {
        for (int fb = frame_begin; fb < frame_end; fb += F3D_MAX_FRAMES_PER_LAUNCH) {
            int fe = frame_end - fb > F3D_MAX_FRAMES_PER_LAUNCH ? fb + F3D_MAX_FRAMES_PER_LAUNCH : frame_end;
            P.f_begin = fb;
            P.f_end = fe;
            P.zbuf = zbuf + (size_t)(fb - frame_begin) * H * W;
            rc = launch_fuse<MODE_SPLAT, F3D_DEPTH_U16_MM>(P, RP, audit, (cudaStream_t)stream);
            if (rc) return rc;
        }
    }
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffff) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_zbuffer_splat: too many pixels for one launch");
    zbuf_finalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(zbuf, depth_out, total, H, W, border);
    return f3d_check_launch("f3d_zbuffer_splat");
}
