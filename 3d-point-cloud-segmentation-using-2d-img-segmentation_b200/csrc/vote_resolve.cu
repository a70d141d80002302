// Level-V vote (uv2pt + mask -> votes), mask nearest-resize, and kernel (3) label resolve.
//   VotingSegmentation.vote     Fusion3DSeg/segUtils/voting.py:75-104
//   VotingSegmentation.segment  Fusion3DSeg/segUtils/voting.py:106-137
//   cv2.resize(INTER_NEAREST)   call site voting.py:93
#include "f3d_common.cuh"
#include "f3d_host.h"

// ---- level V ----------------------------------------------------------------------------------------------------
// numpy's buffered `votes[idx, cls] += 1` (voting.py:98) counts each distinct (point, class) pair of a frame ONCE.
// A vote cell is (frame_tag << 16 | count): a pixel adds 1 only if the cell's tag is not the current frame's tag.
// Pixels of one warp that carry the same (point, class) key are first aggregated with __match_any_sync so a
// single lane issues the CAS (adjacent pixels usually map to the same fused point).
__global__ void __launch_bounds__(256) vote_uv2pt_kernel(const int32_t* __restrict__ uv2pt, const uint8_t* __restrict__ mask,
                                                         int64_t npix, uint32_t tag, uint32_t* __restrict__ votes, int64_t N,
                                                         int C1) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // warp-uniform trip count so that __match_any_sync sees the whole warp
    const int64_t iters = (npix + stride - 1) / stride;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i = start + it * stride;
        long long key = -1;
        if (i < npix) {
            const int pt = __ldg(uv2pt + i);
            const int cls = __ldg(mask + i);
            if (pt >= 0 && pt < N && cls < C1) key = (long long)pt * C1 + cls;   // valid = uv2pt != -1 (voting.py:95)
        }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const bool leader = (__ffs(peers) - 1) == (int)(threadIdx.x & 31);
        if (key >= 0 && leader) {
            uint32_t* cell = votes + key;
            uint32_t old = *cell;
            while ((old >> 16) != tag) {
                const uint32_t assumed = old;
                old = atomicCAS(cell, assumed, (tag << 16) | ((assumed & 0xffffu) + 1u));
                if (old == assumed) break;
            }
        }
    }
}

__global__ void vote_finalize_kernel(uint32_t* __restrict__ v, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] &= 0xffffu;
}

// ---- cv2.resize(..., INTER_NEAREST): sx = min(floor(dx * (1 / (dst_w / src_w))), src_w - 1), float64 ---------------
__global__ void resize_nearest_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst, int dh, int dw,
                                      double ifx, double ify) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int img = blockIdx.z;
    if (x >= dw) return;
    int sx = (int)floor(xmul((double)x, ifx));
    int sy = (int)floor(xmul((double)y, ify));
    sx = min(sx, sw - 1);
    sy = min(sy, sh - 1);
    dst[((size_t)img * dh + y) * dw + x] = __ldg(src + ((size_t)img * sh + sy) * sw + sx);
}

// ---- frame packing: uint16 depth + uint8 class image -> one uint32 texel (depth | class << 16) per pixel -------------------------
// The fused sweep then needs ONE 32-byte sector per point-view instead of a depth sector and a mask sector (its gathers are
// sector-count bound, not byte bound).  The mask may arrive at its own resolution: cv2.resize(mask, (w, h), INTER_NEAREST)
// (voting.py:93, same float64 index rule as resize_nearest_kernel) is folded into the read.  Layout F3D_FRAMES_U32_T16 stores
// 16 x 16-pixel tiles contiguously (tiles row-major, texels row-major inside a tile, partial tiles zero-padded).
__global__ void __launch_bounds__(256) pack_frames_kernel(const uint16_t* __restrict__ depth, const uint8_t* __restrict__ mask,
                                                          uint32_t* __restrict__ out, int H, int W, int mh, int mw, double ifx,
                                                          double ify, int tiled, int tiles_x, long long frame_texels) {
    // a thread packs FOUR consecutive texels of a row (a 16-byte store; an 8-byte depth and a 4-byte mask load when the
    // images allow it): the pass is a pure stream, 3 bytes in and 4 bytes out per pixel
    const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int f = blockIdx.y;
    if (t >= frame_texels) return;
    int ix, iy;
    if (tiled) {
        const int tile = (int)(t >> 8);
        iy = (tile / tiles_x) * 16 + (int)((t >> 4) & 15);
        ix = (tile % tiles_x) * 16 + (int)(t & 15);
    } else {
        iy = (int)(t / W);
        ix = (int)(t - (long long)iy * W);
    }
    uint32_t v[4] = {0u, 0u, 0u, 0u};
    const bool same = (mw == W && mh == H);
    const bool row4 = iy < H && ix + 3 < W && (tiled || frame_texels - t >= 4) && (W & 3) == 0 && (tiled || (ix & 3) == 0);
    if (row4 && same) {
        const uint2 d = __ldg(reinterpret_cast<const uint2*>(depth + ((size_t)f * H + iy) * W + ix));
        const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(mask + ((size_t)f * H + iy) * W + ix));
        v[0] = (d.x & 0xffffu) | ((m & 0xffu) << 16);
        v[1] = (d.x >> 16) | (((m >> 8) & 0xffu) << 16);
        v[2] = (d.y & 0xffffu) | (((m >> 16) & 0xffu) << 16);
        v[3] = (d.y >> 16) | ((m >> 24) << 16);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int x = ix + k, y = iy;
            if (!tiled) {   // linear layout: the four texels may wrap to the next row
                const long long tt = t + k;
                if (tt >= frame_texels) break;
                y = (int)(tt / W);
                x = (int)(tt - (long long)y * W);
            }
            if (x < W && y < H) {
                int sx = x, sy = y;
                if (!same) {
                    sx = min((int)floor(xmul((double)x, ifx)), mw - 1);
                    sy = min((int)floor(xmul((double)y, ify)), mh - 1);
                }
                v[k] = (uint32_t)__ldg(depth + ((size_t)f * H + y) * W + x) | ((uint32_t)__ldg(mask + ((size_t)f * mh + sy) * mw + sx) << 16);
            }
        }
    }
    uint32_t* o = out + (size_t)f * (size_t)frame_texels + (size_t)t;
    if (frame_texels - t >= 4) *reinterpret_cast<uint4*>(o) = make_uint4(v[0], v[1], v[2], v[3]);
    else
        for (int k = 0; k < (int)(frame_texels - t); ++k) o[k] = v[k];
}

// ---- kernel (3): label resolve --------------------------------------------------------------------------------------------
#define RES_MAX_FILTER 256
struct ResolveParams {
    int nfilter;                       // 0 = all columns
    int16_t fpos[RES_MAX_FILTER];      // column -> first position in the filter list (-1 = not considered)
    int32_t remap[RES_MAX_FILTER + 1]; // composed sequential remap (voting.py:133-135) for arg-max position i
    int32_t unclassified;              // value for "unclassified" after the same remap
    double threshold;
};

// Eight lanes per point row, four rows per warp.  A lane streams its share of the row's int32 counters with 8-byte
// loads (row stride 4*C1 bytes: 8-byte aligned when C1 is even), keeps the row total and the first maximum among
// the considered columns, three xor-shuffles combine the eight partials, and one lane applies the float64 tests
// of voting.py:126-131.  HBM-bound: 4*C1 bytes read + 8 bytes written per point.
__device__ __forceinline__ void res_take(int v, int pos, int& best, int& bpos) {
    // zero cells can never win (max == 0 resolves to "unclassified", voting.py:131), so only v > 0 competes
    if (v > 0 && pos >= 0 && (v > best || (v == best && pos < bpos))) {
        best = v;
        bpos = pos;
    }
}

template <bool FILTERED, bool VEC2, typename VT>
__global__ void __launch_bounds__(256) resolve_kernel(const VT* __restrict__ votes, int64_t N, int C1,
                                                      const ResolveParams rp, int64_t* __restrict__ labels) {
    __shared__ int16_t s_fpos[RES_MAX_FILTER];
    if (FILTERED) {
        for (int c = threadIdx.x; c < RES_MAX_FILTER; c += blockDim.x) s_fpos[c] = rp.fpos[c];
        __syncthreads();
    }
    const int sub = threadIdx.x & 7;
    const int64_t rows_per_pass = ((int64_t)gridDim.x * blockDim.x) >> 3;
    const int64_t row0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int64_t passes = (N + rows_per_pass - 1) / rows_per_pass;   // uniform trip count: shuffles see full warps
    for (int64_t p = 0; p < passes; ++p) {
        const int64_t row = row0 + p * rows_per_pass;
        const bool live = row < N;
        long long total = 0;
        int best = 0, bpos = 0x7fff;
        if (live) {
            const VT* __restrict__ r = votes + row * C1;
            if (VEC2) {
                // two counters per load: an int2 for int32 votes, one 32-bit word for packed uint16 votes
                const int n2 = C1 >> 1;
                int2 buf[4];
                for (int i0 = sub; i0 < n2; i0 += 32) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        buf[u] = make_int2(0, 0);
                        if (i0 + 8 * u < n2) {
                            if (sizeof(VT) == 4) {
                                buf[u] = __ldg(reinterpret_cast<const int2*>(r) + i0 + 8 * u);
                            } else {
                                const unsigned w = __ldg(reinterpret_cast<const unsigned*>(r) + i0 + 8 * u);
                                buf[u] = make_int2((int)(w & 0xffffu), (int)(w >> 16));
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = 2 * (i0 + 8 * u);
                        if ((buf[u].x | buf[u].y) != 0) {
                            total += (long long)buf[u].x + (long long)buf[u].y;
                            res_take(buf[u].x, FILTERED ? (int)s_fpos[c] : c, best, bpos);
                            res_take(buf[u].y, FILTERED ? (int)s_fpos[c + 1] : c + 1, best, bpos);
                        }
                    }
                }
            } else {
                for (int c = sub; c < C1; c += 8) {
                    const int v = (int)__ldg(r + c);
                    total += v;
                    res_take(v, FILTERED ? (int)s_fpos[c] : c, best, bpos);
                }
            }
        }
#pragma unroll
        for (int s = 4; s > 0; s >>= 1) {
            total += __shfl_xor_sync(0xffffffffu, total, s);
            const int ob = __shfl_xor_sync(0xffffffffu, best, s);
            const int op = __shfl_xor_sync(0xffffffffu, bpos, s);
            if (ob > best || (ob == best && op < bpos)) {
                best = ob;
                bpos = op;
            }
        }
        if (live && sub == 0) {
            bool unc = (total <= 0) || (best <= 0);                              // voting.py:126,131
            if (!unc) unc = xdiv((double)best, (double)total) < rp.threshold;     // voting.py:128-130
            labels[row] = (int64_t)(unc ? rp.unclassified : (FILTERED ? rp.remap[bpos] : bpos));
        }
    }
}

// ---- C ABI ---------------------------------------------------------------------------------------------------------------
static unsigned grid_for(int64_t n, int block, int max_blocks) {
    int64_t b = (n + block - 1) / block;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (unsigned)b;
}

extern "C" int f3d_vote_uv2pt(const int32_t* uv2pt, const uint8_t* mask, int32_t nframes, int64_t npix, int32_t first_tag,
                              uint32_t* votes_packed, int64_t N, int32_t C1, void* stream) {
    if (!uv2pt || !mask || !votes_packed || nframes < 0 || npix < 0 || N < 0 || C1 <= 0)
        return f3d_fail(F3D_ERR_ARG, "f3d_vote_uv2pt: bad argument");
    if (first_tag < 1 || (int64_t)first_tag + nframes - 1 > 65535)
        return f3d_fail(F3D_ERR_ARG, "f3d_vote_uv2pt: frame tags must stay within 1..65535");
    if (npix == 0 || N == 0) return F3D_OK;
    // frames are serialised by stream order: a cell's tag identifies the frame that last voted for it
    const unsigned grid = grid_for(npix, 256, 148 * 8);
    for (int f = 0; f < nframes; ++f) {
        vote_uv2pt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(uv2pt + (size_t)f * npix, mask + (size_t)f * npix, npix,
                                                                  (uint32_t)(first_tag + f), votes_packed, N, C1);
    }
    return f3d_check_launch("f3d_vote_uv2pt");
}

extern "C" int f3d_vote_finalize(uint32_t* votes_packed, int64_t ncells, void* stream) {
    if (!votes_packed || ncells < 0) return f3d_fail(F3D_ERR_ARG, "f3d_vote_finalize: bad argument");
    if (ncells == 0) return F3D_OK;
    vote_finalize_kernel<<<grid_for(ncells, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(votes_packed, ncells);
    return f3d_check_launch("f3d_vote_finalize");
}

extern "C" int f3d_resize_nearest_u8(const uint8_t* src, int32_t nimg, int32_t src_h, int32_t src_w, uint8_t* dst,
                                     int32_t dst_h, int32_t dst_w, void* stream) {
    if (!src || !dst || nimg <= 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || nimg > 65535 || dst_h > 65535)
        return f3d_fail(F3D_ERR_ARG, "f3d_resize_nearest_u8: bad argument");
    // OpenCV: inv_scale = dsize / ssize (double), ifx = 1 / inv_scale
    volatile double isx = (double)dst_w / (double)src_w, isy = (double)dst_h / (double)src_h;
    const double ifx = 1.0 / isx, ify = 1.0 / isy;
    dim3 grid((unsigned)((dst_w + 255) / 256), (unsigned)dst_h, (unsigned)nimg);
    resize_nearest_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_h, src_w, dst, dst_h, dst_w, ifx, ify);
    return f3d_check_launch("f3d_resize_nearest_u8");
}

extern "C" int f3d_pack_frames(const uint16_t* depth_mm, const uint8_t* mask, int32_t nframes, int32_t H, int32_t W,
                               int32_t mask_h, int32_t mask_w, int32_t frame_fmt, uint32_t* out, void* stream) {
    if (!depth_mm || !mask || !out || nframes < 0 || nframes > 65535 || H <= 0 || W <= 0 || mask_h <= 0 || mask_w <= 0 ||
        (frame_fmt != F3D_FRAMES_U32 && frame_fmt != F3D_FRAMES_U32_T16))
        return f3d_fail(F3D_ERR_ARG, "f3d_pack_frames: bad argument");
    if (nframes == 0) return F3D_OK;
    const int tiles_x = (W + 15) / 16;
    const long long texels = frame_fmt == F3D_FRAMES_U32_T16 ? (long long)tiles_x * ((H + 15) / 16) * 256 : (long long)H * W;
    volatile double isx = (double)W / (double)mask_w, isy = (double)H / (double)mask_h;   // OpenCV: inv_scale = dsize / ssize
    dim3 grid((unsigned)((texels + 1023) / 1024), (unsigned)nframes);
    pack_frames_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(depth_mm, mask, out, H, W, mask_h, mask_w, 1.0 / isx, 1.0 / isy,
                                                               frame_fmt == F3D_FRAMES_U32_T16, tiles_x, texels);
    return f3d_check_launch("f3d_pack_frames");
}

template <typename VT>
static int resolve_impl(const VT* votes, int64_t N, int32_t C1, double threshold, const int32_t* h_filter, int32_t nfilter,
                        int32_t nclasses_id, int64_t* labels, void* stream) {
    if (!votes || !labels || N < 0 || C1 <= 0 || nfilter < 0 || (nfilter > 0 && !h_filter))
        return f3d_fail(F3D_ERR_ARG, "f3d_resolve_labels: bad argument");
    if (nfilter > RES_MAX_FILTER || (nfilter > 0 && C1 > RES_MAX_FILTER))
        return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_resolve_labels: more than 256 filter classes / columns");
    if (N == 0) return F3D_OK;
    ResolveParams rp;
    rp.nfilter = nfilter;
    rp.threshold = threshold;
    rp.unclassified = nclasses_id;
    for (int c = 0; c < RES_MAX_FILTER; ++c) rp.fpos[c] = -1;
    for (int k = nfilter - 1; k >= 0; --k) {
        if (h_filter[k] < 0 || h_filter[k] >= C1) return f3d_fail(F3D_ERR_ARG, "f3d_resolve_labels: filter class out of range");
        rp.fpos[h_filter[k]] = (int16_t)k;   // the first position of a repeated class wins (argmax over votes[:, filter])
    }
    // compose the sequential remap `for i, cls in enumerate(filter): pc[pc == i] = cls` (voting.py:133-135),
    // including its aliasing, for every start value an arg-max index or the unclassified id can take
    for (int start = 0; start <= nfilter; ++start) {
        int v = start < nfilter ? start : nclasses_id;
        for (int i = 0; i < nfilter; ++i)
            if (v == i) v = h_filter[i];
        if (start < nfilter) rp.remap[start] = v;
        else rp.unclassified = v;
    }
    const unsigned grid = grid_for(N * 8, 256, 148 * 8);
    // rows must start on the vector width: 8 bytes for int32 pairs, 4 bytes for uint16 pairs
    const bool vec2 = ((C1 & 1) == 0) && ((reinterpret_cast<uintptr_t>(votes) & (2 * sizeof(VT) - 1)) == 0);
    cudaStream_t s = (cudaStream_t)stream;
    if (nfilter > 0) {
        if (vec2) resolve_kernel<true, true, VT><<<grid, 256, 0, s>>>(votes, N, C1, rp, labels);
        else resolve_kernel<true, false, VT><<<grid, 256, 0, s>>>(votes, N, C1, rp, labels);
    } else {
        if (vec2) resolve_kernel<false, true, VT><<<grid, 256, 0, s>>>(votes, N, C1, rp, labels);
        else resolve_kernel<false, false, VT><<<grid, 256, 0, s>>>(votes, N, C1, rp, labels);
    }
    return f3d_check_launch("f3d_resolve_labels");
}

extern "C" int f3d_resolve_labels(const int32_t* votes, int64_t N, int32_t C1, double threshold, const int32_t* h_filter,
                                  int32_t nfilter, int32_t nclasses_id, int64_t* labels, void* stream) {
    return resolve_impl<int32_t>(votes, N, C1, threshold, h_filter, nfilter, nclasses_id, labels, stream);
}

extern "C" int f3d_resolve_labels_u16(const uint16_t* votes, int64_t N, int32_t C1, double threshold, const int32_t* h_filter,
                                      int32_t nfilter, int32_t nclasses_id, int64_t* labels, void* stream) {
    return resolve_impl<uint16_t>(votes, N, C1, threshold, h_filter, nfilter, nclasses_id, labels, stream);
}
