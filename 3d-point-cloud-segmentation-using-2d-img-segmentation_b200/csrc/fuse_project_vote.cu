// Kernel (1) fused project + z-test + mask gather + vote, kernel (2) z-buffer splat and the uv2pt writer.
//
// Replaces, for a fixed cloud, the per-frame body of Fusion.fuse (Fusion3DSeg/fusion.py:248-298:
// point_inside_polyhedra intersections.py:146-164 -> points2pixel camera_utils.py:9-26 -> single-pixel
// `criterion` fusion.py:223-228) composed with VotingSegmentation.vote (segUtils/voting.py:89-98).
//
// Design (point-stationary, B200):
//   * one CTA owns a tile of BLOCK consecutive points (float4, coalesced 16 B/thread); a thread keeps its point
//     in registers for the whole launch, so the cloud is streamed from HBM exactly once;
//   * the tile's axis-aligned box is tested conservatively against every frame's five frustum planes (fp32 with
//     an explicit rounding margin); surviving frame ids are compacted into shared memory.  With a spatially
//     sorted cloud this skips ~90 % of the nominal point-views without changing any result;
//   * candidate frames' 128-byte fp32 projection tiles are staged into shared memory in batches and broadcast;
//   * per point-view the fp32 path carries a rigorous rounding bound; any decision (frustum, pixel floor, depth
//     distance) that falls inside its bound is re-evaluated by `exact_eval` in fp64 in the reference's operation
//     order, so every integer outcome is bit-exact against the numpy path.  Both the population of that band
//     and every fp32-vs-fp64 divergence inside it are counted;
//   * votes are accumulated in a per-CTA shared-memory histogram (uint16 [C1][BLOCK], conflict-free: a thread
//     owns its point's column) and written to HBM exactly once, coalesced -- no global atomics, no memset.
#include "f3d_common.cuh"
#include "f3d_host.h"

#define MODE_VOTE 0
#define MODE_SPLAT 1
#define MODE_UV2PT 2

#define FUSE_BLOCK 256
#define FUSE_FCHUNK 1024  // frames culled per pass (candidate list capacity)
#define FUSE_STAGE 16     // FrameFast tiles staged per batch (2 KB)
#define HIST_PAD 2

struct FuseParams {
    const float4* points;
    int64_t N;
    const void* table;
    int f_begin, f_end;
    const void* depth;
    const uint8_t* mask;
    int H, W;
    double K[9];
    float cx, cy, inv_fx, inv_fy;
    float dunit;            // depth sample -> metres (0.001 for uint16 mm, 1 for float32 m)
    float radius;
    double radius_d, zmin, zmax;
    uint32_t d_lo, d_hi;    // uint16 depth: valid <=> d_lo <= d <= d_hi   (fusion.py:62-63 on d/1000)
    int32_t* votes;
    int C1, accumulate;
    int32_t* uv2pt;
    uint32_t* zbuf;
    unsigned long long* stats;
    int audit;
};

struct ExactOut {
    int in, pix, vis, near_edge;
    double zcam;
};

// fp64 evaluation of one point-view in the oracle's operation order (oracle.fuse_frame_visibility).
template <int MODE, int FMT>
__device__ __noinline__ void exact_eval(const FuseParams& P, const FrameExact* __restrict__ fe, int frel, float px,
                                        float py, float pz, ExactOut& o) {
    o.in = 0;
    o.pix = 0;
    o.vis = 0;
    o.near_edge = 0;
    o.zcam = 0.0;
    D3 p = {(double)px, (double)py, (double)pz};
    if (!dinside_planes(fe, p)) return;                              // fusion.py:260
    D3 h = dproject_h(P.K, fe->qi, fe->t, p);                        // camera_utils.py:21-23
    double uf = xdiv(h.x, h.z), vf = xdiv(h.y, h.z);                 // camera_utils.py:24
    double fu = floor(uf), fv = floor(vf);                           // camera_utils.py:25
    int iu = d2i_numpy(fu), iv = d2i_numpy(fv);
    if (iu < 0 || iu >= P.W || iv < 0 || iv >= P.H) return;
    o.in = 1;
    o.pix = iv * P.W + iu;
    o.zcam = h.z;
    double fru = xsub(uf, fu), frv = xsub(vf, fv);
    o.near_edge = (fmin(fru, xsub(1.0, fru)) < 1e-4) || (fmin(frv, xsub(1.0, frv)) < 1e-4);
    if (MODE == MODE_SPLAT) return;
    size_t off = (size_t)frel * (size_t)P.H * (size_t)P.W + (size_t)o.pix;
    double dd;
    bool valid;
    if (FMT == F3D_DEPTH_U16_MM) {
        uint32_t d = __ldg(reinterpret_cast<const uint16_t*>(P.depth) + off);
        valid = (d >= P.d_lo) && (d <= P.d_hi);
        dd = (double)d;
    } else {
        dd = (double)__ldg(reinterpret_cast<const float*>(P.depth) + off);
        valid = (dd > P.zmin) && (dd <= P.zmax);                     // fusion.py:62-63
    }
    if (!valid) return;
    D3 c;                                                            // ios_rtab.py:168-173
    c.x = xmul(xsub((double)iu, P.K[2]), xdiv(dd, P.K[0]));
    c.y = xmul(xsub((double)iv, P.K[5]), xdiv(dd, P.K[4]));
    c.z = dd;
    if (FMT == F3D_DEPTH_U16_MM) {                                   // ios_rtab.py:185
        c.x = xdiv(c.x, 1000.0);
        c.y = xdiv(c.y, 1000.0);
        c.z = xdiv(c.z, 1000.0);
    }
    D3 m = dquat_rotate(fe->q, c);                                   // ios_rtab.py:189-190
    double d0 = xsub(xadd(m.x, fe->t[0]), p.x);
    double d1 = xsub(xadd(m.y, fe->t[1]), p.y);
    double d2 = xsub(xadd(m.z, fe->t[2]), p.z);
    double dist = __dsqrt_rn(xadd(xadd(xmul(d0, d0), xmul(d1, d1)), xmul(d2, d2)));   // fusion.py:224
    o.vis = dist < P.radius_d;                                       // fusion.py:225
}

__device__ __forceinline__ uint32_t quantise_mm(double z) {
    double q = floor(xadd(xmul(z, 1000.0), 0.5));
    q = fmin(fmax(q, 1.0), 65535.0);
    return (uint32_t)q;
}

template <int MODE, int FMT>
__global__ void __launch_bounds__(FUSE_BLOCK, 3) fuse_kernel(const FuseParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [stage FrameFast x FUSE_STAGE][cand u16 x FUSE_FCHUNK][red float x 48 + ints][hist u16 x C1*(BLOCK+PAD)]
    float4* stage = reinterpret_cast<float4*>(smem_raw);
    uint16_t* cand = reinterpret_cast<uint16_t*>(smem_raw + FUSE_STAGE * sizeof(FrameFast));
    float* red = reinterpret_cast<float*>(smem_raw + FUSE_STAGE * sizeof(FrameFast) + FUSE_FCHUNK * sizeof(uint16_t));
    int* ncand_s = reinterpret_cast<int*>(red + 48);
    uint16_t* hist = reinterpret_cast<uint16_t*>(red + 64);
    const int HS = FUSE_BLOCK + HIST_PAD;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t tile_base = (int64_t)blockIdx.x * FUSE_BLOCK;
    const int64_t gi = tile_base + tid;
    const bool active = gi < P.N;
    const int HW = P.H * P.W;

    float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) pt = __ldg(P.points + gi);

    if (MODE == MODE_VOTE) {
        uint32_t* h32 = reinterpret_cast<uint32_t*>(hist);
        const int nwords = (P.C1 * HS + 1) / 2;
        for (int i = tid; i < nwords; i += FUSE_BLOCK) h32[i] = 0u;
    }

    // ---- tile bounding box (exact min / max of the float32 coordinates)
    {
        const float big = 3.0e38f;
        float lo[3] = {active ? pt.x : big, active ? pt.y : big, active ? pt.z : big};
        float hi[3] = {active ? pt.x : -big, active ? pt.y : -big, active ? pt.z : -big};
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], s));
                hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], s));
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                red[warp * 6 + k] = lo[k];
                red[warp * 6 + 3 + k] = hi[k];
            }
        }
    }
    __syncthreads();
    float blo[3], bhi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float l = red[k], h = red[3 + k];
#pragma unroll
        for (int w = 1; w < FUSE_BLOCK / 32; ++w) {
            l = fminf(l, red[w * 6 + k]);
            h = fmaxf(h, red[w * 6 + 3 + k]);
        }
        blo[k] = l;
        bhi[k] = h;
    }
    const float box_mag = fabsf(blo[0]) + fabsf(blo[1]) + fabsf(blo[2]) + fabsf(bhi[0]) + fabsf(bhi[1]) + fabsf(bhi[2]);

    const FrameRecord* __restrict__ frec = reinterpret_cast<const FrameRecord*>(P.table);

    unsigned n_cand = 0, n_exact = 0, n_div = 0, n_edge = 0, n_seen = 0, n_bad = 0;

    for (int cbase = P.f_begin; cbase < P.f_end; cbase += FUSE_FCHUNK) {
        __syncthreads();   // previous chunk's candidate list fully consumed; red[] reads done
        if (tid == 0) *ncand_s = 0;
        __syncthreads();
        const int cend = min(cbase + FUSE_FCHUNK, P.f_end);
        // ---- conservative tile x frustum cull (fp32 + explicit rounding margin; never drops a visible pair)
        for (int f0 = cbase; f0 < cend; f0 += FUSE_BLOCK) {
            const int f = f0 + tid;
            bool keep = false;
            if (f < cend) {
                keep = true;
                const float4* pl = frec[f].cull.pl;
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    const float4 q = __ldg(pl + m);
                    float mx = fmaxf(q.x * blo[0], q.x * bhi[0]) + fmaxf(q.y * blo[1], q.y * bhi[1]) +
                               fmaxf(q.z * blo[2], q.z * bhi[2]) - q.w;
                    float margin = 2.0e-6f * (box_mag + fabsf(q.w)) + 1.0e-7f;
                    keep = keep && (mx >= -margin);
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(ncand_s, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) cand[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)(f - P.f_begin);
        }
        __syncthreads();
        const int ncand = *ncand_s;
        if (active) n_cand += (unsigned)ncand;

        for (int b0 = 0; b0 < ncand; b0 += FUSE_STAGE) {
            const int nb = min(FUSE_STAGE, ncand - b0);
            __syncthreads();   // previous batch consumed
            if (tid < nb * 8) {
                const int k = tid >> 3;
                stage[tid] = __ldg(reinterpret_cast<const float4*>(&frec[P.f_begin + cand[b0 + k]].fast) + (tid & 7));
            }
            __syncthreads();
            if (!active) continue;
            for (int k = 0; k < nb; ++k) {
                const float4* s = stage + k * 8;
                const int frel = cand[b0 + k];
                const float4 A0 = s[0], A1 = s[1];
                const float d0 = (pt.x - A0.x) - A1.x;
                const float d1 = (pt.y - A0.y) - A1.y;
                const float d2 = (pt.z - A0.z) - A1.z;
                const float S = fabsf(d0) + fabsf(d1) + fabsf(d2) + 1.0e-9f;
                const float4 Mz = s[4];
                const float z = fmaf(Mz.x, d0, fmaf(Mz.y, d1, Mz.z * d2));
                const float ez = 8.0f * F3D_U24 * Mz.w * S;

                int st;          // 0 = certainly not seen, 1 = pixel certified, 2 = unsure -> fp64
                int pix = 0;
                int g_in = 0;    // fp32 best guess (for divergence logging)
                float zc = 0.f;
                if (z < -16.0f * ez) {
                    st = 0;
                } else if (z <= 16.0f * ez) {
                    st = 2;
                } else {
                    const float4 Mu = s[2], Mv = s[3];
                    const float a = fmaf(Mu.x, d0, fmaf(Mu.y, d1, Mu.z * d2));
                    const float b = fmaf(Mv.x, d0, fmaf(Mv.y, d1, Mv.z * d2));
                    const float r = __frcp_rn(z);
                    const float u = a * r, v = b * r;
                    const float ea = 8.0f * F3D_U24 * Mu.w * S, eb = 8.0f * F3D_U24 * Mv.w * S;
                    const float eu = 1.125f * (ea + fabsf(u) * ez) * r + 8.0f * F3D_U24 * fabsf(u) + 1.0e-4f;
                    const float ev = 1.125f * (eb + fabsf(v) * ez) * r + 8.0f * F3D_U24 * fabsf(v) + 1.0e-4f;
                    const float4 R0 = s[5], R1 = s[6], R2 = s[7];
                    const float sl = fmaf(R0.w, d0, fmaf(R1.w, d1, R2.w * d2));   // (p - eye) . lookat
                    const float es = 8.0f * F3D_U24 * S + 1.0e-6f * A1.w;
                    const float fW = (float)P.W, fH = (float)P.H;
                    const float fu = floorf(u), fv = floorf(v);
                    if ((u + eu < 0.f) || (u - eu >= fW) || (v + ev < 0.f) || (v - ev >= fH) || (sl - es > A1.w)) {
                        st = 0;
                    } else {
                        const bool cu = (u - fu >= eu) && (fu + 1.0f - u > eu);
                        const bool cv = (v - fv >= ev) && (fv + 1.0f - v > ev);
                        const bool cf = (sl + es < A1.w);
                        g_in = (u >= 0.f) && (fu < fW) && (v >= 0.f) && (fv < fH) && (sl < A1.w);
                        pix = (int)fv * P.W + (int)fu;
                        st = (cu && cv && cf) ? 1 : 2;
                    }
                    zc = z;
                    if (st == 1 && MODE != MODE_SPLAT) {
                        // ---- z-test against the frame's depth (valid range + single-pixel criterion)
                        const size_t off = (size_t)frel * (size_t)HW + (size_t)pix;
                        float dm;   // depth sample in metres
                        bool valid;
                        if (FMT == F3D_DEPTH_U16_MM) {
                            const uint32_t d = __ldg(reinterpret_cast<const uint16_t*>(P.depth) + off);
                            valid = (d >= P.d_lo) && (d <= P.d_hi);
                            dm = (float)d * 0.001f;
                        } else {
                            const float d = __ldg(reinterpret_cast<const float*>(P.depth) + off);
                            valid = ((double)d > P.zmin) && ((double)d <= P.zmax);
                            dm = d;
                        }
                        if (!valid) {
                            st = 0;
                        } else {
                            // metric camera coordinates of the cloud point and of the depth pixel (scaled by |q|^2)
                            const float X = fmaf(R0.x, d0, fmaf(R0.y, d1, R0.z * d2));
                            const float Y = fmaf(R1.x, d0, fmaf(R1.y, d1, R1.z * d2));
                            const float Z = fmaf(R2.x, d0, fmaf(R2.y, d1, R2.z * d2));
                            const float ds = dm * A0.w;
                            const float xn = (fu - P.cx) * P.inv_fx, yn = (fv - P.cy) * P.inv_fy;
                            const float qx = X - xn * ds, qy = Y - yn * ds, qz = Z - ds;
                            const float dist2 = fmaf(qx, qx, fmaf(qy, qy, qz * qz));
                            const float del = 16.0f * F3D_U24 * (S + ds * (1.0f + fabsf(xn) + fabsf(yn)));
                            const float rlo = fmaxf(P.radius - del, 0.f), rhi = P.radius + del;
                            if (dist2 < rlo * rlo * (1.0f - 16.0f * F3D_U24)) {
                                st = 1;
                            } else if (dist2 > rhi * rhi * (1.0f + 16.0f * F3D_U24)) {
                                st = 0;
                            } else {
                                st = 3;   // distance inside its band
                                g_in = dist2 < P.radius * P.radius;
                            }
                        }
                    }
                }

                bool seen = (st == 1);
                uint32_t zq = 0;
                if (MODE == MODE_SPLAT && st == 1) {
                    // quantised camera z: floor(z*1000 + 0.5); certify the floor
                    const float zm = fmaf(zc, 1000.0f, 0.5f);
                    const float fz = floorf(zm);
                    const float eq = 1010.0f * ez + 8.0f * F3D_U24 * zm;
                    if ((zm - fz >= eq) && (fz + 1.0f - zm > eq)) {
                        zq = (uint32_t)fminf(fmaxf(fz, 1.0f), 65535.0f);
                    } else {
                        st = 3;
                        g_in = (int)fminf(fmaxf(fz, 1.0f), 65535.0f);
                    }
                }

                if (st >= 2 || P.audit) {
                    ExactOut eo;
                    exact_eval<MODE, FMT>(P, &frec[P.f_begin + frel].exact, frel, pt.x, pt.y, pt.z, eo);
                    bool e_seen = (MODE == MODE_SPLAT) ? (eo.in != 0) : (eo.vis != 0);
                    uint32_t e_zq = (MODE == MODE_SPLAT && eo.in) ? quantise_mm(eo.zcam) : 0u;
                    if (st >= 2) {
                        ++n_exact;
                        bool diverged;
                        if (st == 2) diverged = (g_in != eo.in) || (eo.in && pix != eo.pix);
                        else if (MODE == MODE_SPLAT) diverged = (!eo.in) || (pix != eo.pix) || ((uint32_t)g_in != e_zq);
                        else diverged = (g_in != eo.vis) || (eo.in && pix != eo.pix);
                        n_div += diverged ? 1u : 0u;
                    } else {
                        // audit: a certified fp32 outcome must equal the fp64 outcome
                        bool bad = (seen != e_seen) || (seen && pix != eo.pix) || (seen && MODE == MODE_SPLAT && zq != e_zq);
                        n_bad += bad ? 1u : 0u;
                    }
                    n_edge += (eo.in && eo.near_edge) ? 1u : 0u;
                    seen = e_seen;
                    pix = eo.pix;
                    zq = e_zq;
                }

                if (seen) {
                    ++n_seen;
                    const size_t off = (size_t)frel * (size_t)HW + (size_t)pix;
                    if (MODE == MODE_VOTE) {
                        const int cls = __ldg(P.mask + off);
                        if (cls < P.C1) hist[cls * HS + tid] += 1;
                    } else if (MODE == MODE_SPLAT) {
                        atomicMin(P.zbuf + off, zq);
                    } else {
                        atomicMax(P.uv2pt + off, (int)gi);
                    }
                }
            }
        }
    }

    // ---- epilogue: histogram -> HBM, written once, coalesced
    if (MODE == MODE_VOTE) {
        __syncthreads();
        const int64_t npts_tile = min((int64_t)FUSE_BLOCK, P.N - tile_base);
        const int total = (int)npts_tile * P.C1;
        int32_t* __restrict__ out = P.votes + tile_base * P.C1;
        int j = tid / P.C1, c = tid - j * P.C1;
        const int dj = FUSE_BLOCK / P.C1, dc = FUSE_BLOCK - dj * P.C1;
        for (int e = tid; e < total; e += FUSE_BLOCK) {
            const int v = hist[c * HS + j];
            // accumulate mode touches only the (sparse) non-zero cells; overwrite mode writes every cell once
            if (!P.accumulate) out[e] = v;
            else if (v) out[e] += v;
            j += dj;
            c += dc;
            if (c >= P.C1) {
                c -= P.C1;
                ++j;
            }
        }
    }

    // ---- statistics
    if (P.stats) {
        unsigned vals[6] = {n_cand, n_exact, n_div, n_edge, n_seen, n_bad};
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            unsigned v = vals[i];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            if (lane == 0 && v) atomicAdd(P.stats + i, (unsigned long long)v);
        }
    }
}

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

// ---- host side ----------------------------------------------------------------------------------------------------
static size_t fuse_smem_bytes(int mode, int C1) {
    size_t b = FUSE_STAGE * sizeof(FrameFast) + FUSE_FCHUNK * sizeof(uint16_t) + 64 * sizeof(float);
    if (mode == MODE_VOTE) b += ((size_t)C1 * (FUSE_BLOCK + HIST_PAD) * sizeof(uint16_t) + 15) & ~(size_t)15;
    return b;
}

template <int MODE, int FMT>
static int launch_fuse(const FuseParams& P, cudaStream_t stream) {
    size_t smem = fuse_smem_bytes(MODE, P.C1);
    if (smem > 227 * 1024) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_fuse: nclasses+1 too large for the shared-memory histogram");
    cudaError_t e = cudaFuncSetAttribute(fuse_kernel<MODE, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return f3d_check_launch("f3d_fuse(cudaFuncSetAttribute)");
    int64_t tiles = (P.N + FUSE_BLOCK - 1) / FUSE_BLOCK;
    if (tiles > 0x7fffffff) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_fuse: too many points for one launch");
    fuse_kernel<MODE, FMT><<<(unsigned)tiles, FUSE_BLOCK, smem, stream>>>(P);
    return f3d_check_launch("f3d_fuse");
}

static int fill_common(FuseParams& P, const void* points, int64_t N, const void* table, int fb, int fe, const void* depth,
                       int fmt, int H, int W, const double* h_K9, double radius, double zmin, double zmax,
                       uint64_t* stats, int flags) {
    if (!points || !table || !h_K9 || N < 0 || fb < 0 || fe < fb || H <= 0 || W <= 0)
        return f3d_fail(F3D_ERR_ARG, "f3d_fuse: bad argument");
    if (fmt != F3D_DEPTH_U16_MM && fmt != F3D_DEPTH_F32_M) return f3d_fail(F3D_ERR_ARG, "f3d_fuse: unknown depth format");
    if ((int64_t)H * W > 0x7fffffff) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_fuse: image too large");
    P.points = reinterpret_cast<const float4*>(points);
    P.N = N;
    P.table = table;
    P.depth = depth;
    P.H = H;
    P.W = W;
    for (int i = 0; i < 9; ++i) P.K[i] = h_K9[i];
    P.cx = (float)h_K9[2];
    P.cy = (float)h_K9[5];
    P.inv_fx = (float)(1.0 / h_K9[0]);
    P.inv_fy = (float)(1.0 / h_K9[4]);
    P.dunit = fmt == F3D_DEPTH_U16_MM ? 0.001f : 1.0f;
    P.radius = (float)radius;
    P.radius_d = radius;
    P.zmin = zmin;
    P.zmax = zmax;
    // valid <=> (d/1000 > zmin) & (d/1000 <= zmax) in float64 (fusion.py:62-63); monotone in d -> integer range
    uint32_t lo = 65536, hi = 0;
    for (uint32_t d = 0; d < 65536; ++d) {
        volatile double z = (double)d / 1000.0;
        if (z > zmin && z <= zmax) {
            if (d < lo) lo = d;
            hi = d;
        }
    }
    P.d_lo = lo;
    P.d_hi = hi;
    P.stats = reinterpret_cast<unsigned long long*>(stats);
    P.audit = flags & 1;
    P.votes = nullptr;
    P.uv2pt = nullptr;
    P.zbuf = nullptr;
    P.mask = nullptr;
    P.C1 = 0;
    P.accumulate = 0;
    return F3D_OK;
}

// frames are processed in launches of at most 65535 (uint16 candidate ids and histogram counters)
#define F3D_MAX_FRAMES_PER_LAUNCH 65535

extern "C" int f3d_fuse_project_vote(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                                     int32_t frame_end, const void* depth, int32_t depth_fmt, const uint8_t* mask,
                                     int32_t H, int32_t W, const double* h_K9, double radius, double zmin, double zmax,
                                     int32_t* votes, int32_t C1, int32_t accumulate, uint64_t* stats, int32_t flags,
                                     void* stream) {
    FuseParams P;
    int rc = fill_common(P, points, N, frame_table, frame_begin, frame_end, depth, depth_fmt, H, W, h_K9, radius, zmin,
                         zmax, stats, flags);
    if (rc) return rc;
    if (!votes || C1 <= 0 || C1 > 256 || (frame_end > frame_begin && (!depth || !mask)))
        return f3d_fail(F3D_ERR_ARG, "f3d_fuse_project_vote: bad argument (votes/mask/depth NULL or C1 not in 1..256)");
    if (N == 0) return F3D_OK;
    P.votes = votes;
    P.C1 = C1;
    const size_t esz = depth_fmt == F3D_DEPTH_U16_MM ? 2 : 4;
    int fb = frame_begin;
    bool first = true;
    do {
        int fe = frame_end - fb > F3D_MAX_FRAMES_PER_LAUNCH ? fb + F3D_MAX_FRAMES_PER_LAUNCH : frame_end;
        P.f_begin = fb;
        P.f_end = fe;
        P.depth = reinterpret_cast<const char*>(depth) + (size_t)(fb - frame_begin) * H * W * esz;
        P.mask = mask + (size_t)(fb - frame_begin) * H * W;
        P.accumulate = (first && !accumulate) ? 0 : 1;
        rc = depth_fmt == F3D_DEPTH_U16_MM ? launch_fuse<MODE_VOTE, F3D_DEPTH_U16_MM>(P, (cudaStream_t)stream)
                                           : launch_fuse<MODE_VOTE, F3D_DEPTH_F32_M>(P, (cudaStream_t)stream);
        if (rc) return rc;
        first = false;
        fb = fe;
    } while (fb < frame_end);
    return F3D_OK;
}

extern "C" int f3d_fuse_uv2pt(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                              int32_t frame_end, const void* depth, int32_t depth_fmt, int32_t H, int32_t W,
                              const double* h_K9, double radius, double zmin, double zmax, int32_t* uv2pt,
                              uint64_t* stats, int32_t flags, void* stream) {
    FuseParams P;
    int rc = fill_common(P, points, N, frame_table, frame_begin, frame_end, depth, depth_fmt, H, W, h_K9, radius, zmin,
                         zmax, stats, flags);
    if (rc) return rc;
    if (!uv2pt || !depth) return f3d_fail(F3D_ERR_ARG, "f3d_fuse_uv2pt: bad argument");
    if (N == 0 || frame_end == frame_begin) return F3D_OK;
    if (N > 0x7fffffff) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_fuse_uv2pt: point index does not fit int32");
    const size_t esz = depth_fmt == F3D_DEPTH_U16_MM ? 2 : 4;
    for (int fb = frame_begin; fb < frame_end; fb += F3D_MAX_FRAMES_PER_LAUNCH) {
        int fe = frame_end - fb > F3D_MAX_FRAMES_PER_LAUNCH ? fb + F3D_MAX_FRAMES_PER_LAUNCH : frame_end;
        P.f_begin = fb;
        P.f_end = fe;
        P.depth = reinterpret_cast<const char*>(depth) + (size_t)(fb - frame_begin) * H * W * esz;
        P.uv2pt = uv2pt + (size_t)(fb - frame_begin) * H * W;
        rc = depth_fmt == F3D_DEPTH_U16_MM ? launch_fuse<MODE_UV2PT, F3D_DEPTH_U16_MM>(P, (cudaStream_t)stream)
                                           : launch_fuse<MODE_UV2PT, F3D_DEPTH_F32_M>(P, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return F3D_OK;
}

extern "C" int f3d_zbuffer_splat(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                                 int32_t frame_end, int32_t H, int32_t W, const double* h_K9, uint32_t* zbuf,
                                 uint16_t* depth_out, int32_t border, uint64_t* stats, int32_t flags, void* stream) {
    FuseParams P;
    int rc = fill_common(P, points, N, frame_table, frame_begin, frame_end, nullptr, F3D_DEPTH_U16_MM, H, W, h_K9, 0.0,
                         0.0, 0.0, stats, flags);
    if (rc) return rc;
    if (!zbuf || !depth_out || border < 0) return f3d_fail(F3D_ERR_ARG, "f3d_zbuffer_splat: bad argument");
    const int nf = frame_end - frame_begin;
    if (nf == 0) return F3D_OK;
    const int64_t total = (int64_t)nf * H * W;
    cudaError_t e = cudaMemsetAsync(zbuf, 0xff, (size_t)total * sizeof(uint32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) return f3d_check_launch("f3d_zbuffer_splat(memset)");
    if (N > 0) {
        for (int fb = frame_begin; fb < frame_end; fb += F3D_MAX_FRAMES_PER_LAUNCH) {
            int fe = frame_end - fb > F3D_MAX_FRAMES_PER_LAUNCH ? fb + F3D_MAX_FRAMES_PER_LAUNCH : frame_end;
            P.f_begin = fb;
            P.f_end = fe;
            P.zbuf = zbuf + (size_t)(fb - frame_begin) * H * W;
            rc = launch_fuse<MODE_SPLAT, F3D_DEPTH_U16_MM>(P, (cudaStream_t)stream);
            if (rc) return rc;
        }
    }
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffff) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_zbuffer_splat: too many pixels for one launch");
    zbuf_finalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(zbuf, depth_out, total, H, W, border);
    return f3d_check_launch("f3d_zbuffer_splat");
}
