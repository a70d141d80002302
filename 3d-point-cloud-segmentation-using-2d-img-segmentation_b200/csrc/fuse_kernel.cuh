// Kernel (1) fused project + z-test + mask gather + vote (+ optional fused label resolve), kernel (2) z-buffer splat
// (point-stationary variant) and the uv2pt writer: device code and launch templates.  The extern "C" entry points live
// in fuse_vote.cu / fuse_aux.cu, which instantiate the modes they need (separate translation units compile in parallel).
//
// Replaces, for a fixed cloud, the per-frame body of Fusion.fuse (Fusion3DSeg/fusion.py:248-298:
// point_inside_polyhedra intersections.py:146-164 -> points2pixel camera_utils.py:9-26 -> single-pixel
// `criterion` fusion.py:223-228) composed with VotingSegmentation.vote (segUtils/voting.py:89-98) and, when asked,
// VotingSegmentation.segment (segUtils/voting.py:106-137).
//
// Design (point-stationary, B200):
//   * one CTA owns a tile of FUSE_BLOCK consecutive points (float4, coalesced 16 B/thread); a thread keeps its point in
//     registers for the whole launch, so the cloud is streamed from HBM exactly once;
//   * cull: supertile_cull_kernel lists the frames that can see each 4096-point super-tile; the CTA then tests every
//     (listed frame, warp box) pair -- one pair per thread -- against the frame's five frustum planes (fp32 with an
//     explicit rounding margin, conservative) and compacts the surviving frame ids with the mask of warps that keep them;
//   * sweep: WARP-AUTONOMOUS and software pipelined.  Each warp walks the candidate list on its own, one candidate per
//     iteration: the 128-byte fp32 projection tile of candidate i+2 and the packed depth|mask texel of candidate i travel
//     to the warp's own shared-memory rings as cp.async copies (LDGSTS, no registers held), candidate i is projected in
//     fp32 and candidate i-2 is depth-tested and voted from its landed texel.  There is no CTA-wide barrier and no shared
//     staging ring inside the sweep (round 1 staged the tiles CTA-wide through TMA + mbarriers and lost 21 % of its warp
//     samples on those barriers), and the loop body is a few hundred instructions (it must live in the instruction cache);
//   * per point-view the fp32 path carries a rigorous rounding bound; any decision (frustum, pixel floor, depth
//     distance) that falls inside its bound is re-evaluated in fp64 in the reference's operation order (`exact_eval`),
//     so every integer outcome is bit-exact against the numpy path.  The population of that band and every
//     fp32-vs-fp64 divergence inside it are counted;
//   * votes are accumulated in a per-CTA shared-memory byte histogram (a thread owns its point's row) and written to
//     HBM exactly once with 32-byte stores -- no global atomics, no memset.  Labels come from the running arg-max.
#pragma once
#include "f3d_common.cuh"
#include "f3d_host.h"

#define MODE_VOTE 0
#define MODE_SPLAT 1
#define MODE_UV2PT 2

#ifndef FUSE_BLOCK
#define FUSE_BLOCK 128
#endif
#define FUSE_NW (FUSE_BLOCK / 32)
#ifndef FUSE_WCAP
#define FUSE_WCAP 128     // candidate frames a warp's own list holds per cull pass (longer frame lists are swept in several passes)
#endif
#ifndef FUSE_MINB
#define FUSE_MINB (512 / FUSE_BLOCK)        // resident CTAs per SM of the uint16-histogram / splat / uv2pt builds (128 registers)
#endif
// Byte-histogram build of the vote kernel (HB = 1): uint8 counters halve the shared-memory footprint, so 24 warps are
// resident per SM and their latency-bound phases (cull, gathers, vote write) overlap.  A counter can hold 255: a warp
// flushes its rows to HBM (first flush overwrites, later ones add) before more than FUSE_LIMIT8 candidates have been
// swept since its last flush, so no count is ever lost.
#ifndef FUSE_MINB8
#define FUSE_MINB8 (FUSE_BLOCK == 128 ? 5 : 2)   // 128-point tiles: 5 CTAs = 20 warps per SM at 96 registers (6 would need 80: spills)
#endif
#ifndef FUSE_ST_POINTS
#define FUSE_ST_POINTS 4096                 // points per super-tile of the first cull level
#endif
#define FUSE_ST_TILES (FUSE_ST_POINTS / FUSE_BLOCK)
#define FUSE_ST_LCAP 1024  // candidate frames a super-tile list holds
#ifndef FUSE_QWARP
#define FUSE_QWARP 20     // deferred entries per warp (12 B each); overflow falls back to inline evaluation
#endif
#define FUSE_LIMIT8 (255 - FUSE_QWARP)
// compiler-level fence between the per-candidate bodies of the unrolled group loops: keeps nvcc from hoisting the ~23
// shared-memory operands of all NB candidates at once (which costs more registers than the 80-register budget has)
#ifndef FUSE_NO_FENCE
#define FUSE_SCHED_FENCE() asm volatile("" ::: "memory")
#else
#define FUSE_SCHED_FENCE()
#endif   // the warp's deferred pass can add up to FUSE_QWARP votes to one cell at the end

// an uncertain point-view handed from the fused sweep to the fix-up kernels through the caller's workspace
struct GEntry {
    int32_t pt;        // global point index (-1: reserved but unused slot)
    uint32_t w;        // frame (relative to f_begin) | st << 16 | seen << 24
    int32_t pix;       // fp32 pixel guess (linear v * W + u)
    uint32_t guess;    // fp32 decision guess (visibility / quantised depth)
};

struct FuseParams {
    const float4* points;
    int64_t N;
    const void* table;
    int f_begin, f_end;
    const void* depth;      // uint16 mm / float32 m images, or packed uint32 texels (depth | class << 16)
    const uint8_t* mask;    // NULL with the packed formats
    int H, W;
    int64_t frame_stride;   // elements per frame of `depth` (H*W, or tiles * 256 for the tiled layout)
    int tiles_x;            // tiled layout: 16-pixel tile columns per row of tiles
    double K[9];
    float cx, cy, inv_fx, inv_fy;
    float radius;
    double radius_d, zmin, zmax;
    uint32_t d_lo, d_hi;    // uint16 depth: valid <=> d_lo <= d <= d_hi   (fusion.py:62-63 on d/1000)
    int32_t* votes;
    uint16_t* votes16;      // alternative packed output (uint16 counters): halves the vote write and the dense multi-GPU exchange
    int C1, RS, accumulate;
    int32_t* uv2pt;
    uint32_t* zbuf;
    int64_t* labels;
    unsigned long long* stats;
    // first cull level (optional, from the workspace): candidate frames of every super-tile, found by
    // supertile_cull_kernel; a tile then tests only its super-tile's list instead of every frame of the launch
    const unsigned* st_count;      // [super-tiles] list length, 0xFFFFFFFF = list overflowed (scan all frames)
    const uint16_t* st_list;       // [super-tiles][FUSE_ST_LCAP] frame ids relative to f_begin
    // compacted launch (exchange mode, optional): the grid covers only the super-tiles some frame of the launch can see;
    // block b works on tile live_list[b / FUSE_ST_TILES] * FUSE_ST_TILES + b % FUSE_ST_TILES (NULL: tile b)
    const unsigned* live_list;
    unsigned* live_scratch;        // [super-tiles] storage of the list + its length at live_count (host side only)
    unsigned* live_count;
    int compact;                   // host side: build the list and launch the compacted grid (one stream synchronisation)
    // fused labels + deferred queue: per-point resolve state (total | best << 24 | first position << 48) written by the sweep,
    // advanced by the fix-up kernel with a compare-and-swap per deferred vote, so the labels of the touched points are
    // re-resolved from 8 bytes instead of re-reading their 4*C1-byte vote rows (NULL: re-read the rows)
    unsigned long long* summ;
    GEntry* gq;                    // workspace queue of deferred point-views (NULL: evaluate them inside the sweep)
    unsigned long long* gq_count;
    unsigned long long gq_cap;
    // ---- multi-GPU vote exchange fused into the kernel (sender side; the owner side is vote_exchange.cu) -------------
    // Rank d owns the points [d * xg_per, (d + 1) * xg_per), xg_per a multiple of the tile, so a CTA has one owner.
    // Slot records: per 32-point block L rows of 64 B, row j = the j-th class (in order of first appearance) of each of the
    // block's points as uint16 class | count << 8 (0 = none), L = the longest list in the block.  The block's warp reserves
    // L rows in one of F3D_XCH_NREG sub-regions of this rank's record region at the owner (an atomic on a local cursor
    // that only ~300 warps share -- a single cursor serialises at ~10 ns per warp in L2 and binds the whole kernel),
    // writes them and the directory entry {row offset, L} straight into the owner's memory over NVLink.  A warp that sweeps
    // more than FUSE_LIMIT8 candidate frames flushes its byte histogram several times: every flush writes its own record,
    // the directory holds F3D_XCH_NLEVEL of them per block.
    // What cannot go into a record (sub-region full, more than F3D_XCH_NLEVEL flushes, and the deferred fp64 votes of the
    // fix-up kernel) is appended as (cell, count) to one of F3D_XCH_NSUB sub-queues:
    // fix-up block b owns sub-queue b (no global atomics at all), spills take the sub-queues above F3D_XCH_NSUB_FIX.
    int xg_G;                                  // 0 = off
    long long xg_per;                          // points per owner shard: owner(p) = p / xg_per
    uint16_t* xg_slots[F3D_MAX_RANKS];         // this rank's record region inside rank d's receive buffer (peer pointers)
    uint2* xg_dir[F3D_MAX_RANKS];              // ... its directory [xg_per / 32][F3D_XCH_NLEVEL] of {row offset, L}
    unsigned long long* xg_queue[F3D_MAX_RANKS];   // ... its (cell, count) queue [F3D_XCH_NSUB][xg_subcap]
    unsigned* xg_rowcur;                       // local [G][F3D_XCH_NREG] row cursors (caller zeroes them per call)
    unsigned* xg_qcur;                         // local [G][F3D_XCH_NSUB] queue cursors (ditto); published to the owners afterwards
    unsigned xg_subrows, xg_subcap;            // rows per record sub-region, entries per sub-queue
    unsigned* xg_overflow;                     // set when a sub-queue is full (entries are then dropped: the owner side raises)
};
#define FUSE_NSLOT 32   // classes per point remembered by cast_vote; longer lists are re-read from the histogram row
#define FUSE_STG_ROWS 8 // record rows per staging chunk (512 B per warp: shared memory taken here is L1 taken from the gathers)

__device__ __forceinline__ unsigned long long summ_pack(int total, int best, int bpos) {
    return (unsigned long long)(unsigned)total | ((unsigned long long)(unsigned)best << 24) | ((unsigned long long)(unsigned)(bpos & 0xffff) << 48);
}

// one (cell, count) entry for owner d through sub-queue `sub`
__device__ __forceinline__ void xg_append(const FuseParams& P, int d, unsigned sub, unsigned at, unsigned long long key, unsigned count) {
    if (at < P.xg_subcap) P.xg_queue[d][(size_t)sub * P.xg_subcap + at] = key | ((unsigned long long)count << 40);
    else atomicExch(P.xg_overflow, 1u);
}

// ---- frame addressing ----------------------------------------------------------------------------------------------
// element offset of pixel (iu, iv) of frame `frel` in P.depth (and P.mask for the two-array formats)
template <int FMT>
__device__ __forceinline__ size_t frame_off(const FuseParams& P, int frel, int iu, int iv) {
    if (FMT == F3D_FRAMES_U32_T16)
        return (size_t)frel * (size_t)P.frame_stride + (size_t)((((iv >> 4) * P.tiles_x + (iu >> 4)) << 8) | ((iv & 15) << 4) | (iu & 15));
    return (size_t)frel * (size_t)P.frame_stride + (size_t)(iv * P.W + iu);
}

// Frame gathers.  A plain global load that misses L1 makes this GPU fill the WHOLE 128-byte line from HBM (measured:
// tools/micro/gather_fetch.cu, 3.9 DRAM sectors per 4-byte gather; cudaLimitMaxL2FetchGranularity changes nothing), i.e. four
// times the bytes a random gather needs.  The .L2::64B qualifier is the smallest fill the ISA offers: 2.2 sectors per gather
// and 27 % less time in the micro-benchmark.
#ifndef FUSE_PLAIN_GATHER
__device__ __forceinline__ uint32_t gather_u32(const uint32_t* p) {
    uint32_t v;
    asm("ld.global.nc.L2::64B.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t gather_u16(const uint16_t* p) {
    uint16_t v;
    asm("ld.global.nc.L2::64B.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t gather_u8(const uint8_t* p) {
    uint32_t v;
    asm("ld.global.nc.L2::64B.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
#else
__device__ __forceinline__ uint32_t gather_u32(const uint32_t* p) { return __ldg(p); }
__device__ __forceinline__ uint32_t gather_u16(const uint16_t* p) { return __ldg(p); }
__device__ __forceinline__ uint32_t gather_u8(const uint8_t* p) { return __ldg(p); }
#endif

// raw depth word (uint16 mm, float32 bits) and -- vote mode only -- the class id of one pixel
template <int MODE, int FMT>
__device__ __forceinline__ void load_pixel(const FuseParams& P, size_t off, uint32_t& draw, uint32_t& cls) {
    cls = 0;
    if (FMT == F3D_DEPTH_U16_MM) {
        draw = __ldg(reinterpret_cast<const uint16_t*>(P.depth) + off);
        if (MODE == MODE_VOTE) cls = __ldg(P.mask + off);
    } else if (FMT == F3D_DEPTH_F32_M) {
        draw = __float_as_uint(__ldg(reinterpret_cast<const float*>(P.depth) + off));
        if (MODE == MODE_VOTE) cls = __ldg(P.mask + off);
    } else {
        const uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(P.depth) + off);
        draw = t & 0xffffu;
        cls = (t >> 16) & 0xffu;
    }
}

__device__ __forceinline__ uint32_t quantise_mm(double z) {
    double q = floor(xadd(xmul(z, 1000.0), 0.5));
    q = fmin(fmax(q, 1.0), 65535.0);
    return (uint32_t)q;
}

// fp64 evaluation of one point-view in the oracle's operation order (oracle.fuse_frame_visibility).  Returns
//   bit 0 in-image, bit 1 visible, bit 2 within 1e-4 px of a pixel edge, bits 8..15 class (vote mode, when visible),
//   bits 16..31 quantised camera z (splat mode), bits 32..63 linear pixel index
// packed into one register pair, so no local of the caller has its address taken (the call sits on a cold path of the
// sweep and must not cost the hot path any stack traffic).
#define EX_IN 1ull
#define EX_VIS 2ull
#define EX_EDGE 4ull
template <int MODE, int FMT>
__device__ __noinline__ unsigned long long exact_eval(const FuseParams& P, const FrameExact* __restrict__ fe, int frel, float px,
                                                      float py, float pz) {
    D3 p = {(double)px, (double)py, (double)pz};
    if (!dinside_planes(fe, p)) return 0ull;                         // fusion.py:260
    D3 h = dproject_h(P.K, fe->qi, fe->t, p);                        // camera_utils.py:21-23
    double uf = xdiv(h.x, h.z), vf = xdiv(h.y, h.z);                 // camera_utils.py:24
    double fu = floor(uf), fv = floor(vf);                           // camera_utils.py:25
    int iu = d2i_numpy(fu), iv = d2i_numpy(fv);
    if (iu < 0 || iu >= P.W || iv < 0 || iv >= P.H) return 0ull;
    unsigned long long r = EX_IN | ((unsigned long long)(unsigned)(iv * P.W + iu) << 32);
    double fru = xsub(uf, fu), frv = xsub(vf, fv);
    if ((fmin(fru, xsub(1.0, fru)) < 1e-4) || (fmin(frv, xsub(1.0, frv)) < 1e-4)) r |= EX_EDGE;
    if (MODE == MODE_SPLAT) return r | ((unsigned long long)quantise_mm(h.z) << 16);
    uint32_t draw, cls;
    load_pixel<MODE, FMT>(P, frame_off<FMT>(P, frel, iu, iv), draw, cls);
    double dd;
    bool valid;
    if (FMT != F3D_DEPTH_F32_M) {
        valid = (draw >= P.d_lo) && (draw <= P.d_hi);
        dd = (double)draw;
    } else {
        dd = (double)__uint_as_float(draw);
        valid = (dd > P.zmin) && (dd <= P.zmax);                     // fusion.py:62-63
    }
    if (!valid) return r;
    D3 c;                                                            // ios_rtab.py:168-173
    c.x = xmul(xsub((double)iu, P.K[2]), xdiv(dd, P.K[0]));
    c.y = xmul(xsub((double)iv, P.K[5]), xdiv(dd, P.K[4]));
    c.z = dd;
    if (FMT != F3D_DEPTH_F32_M) {                                    // ios_rtab.py:185
        c.x = xdiv(c.x, 1000.0);
        c.y = xdiv(c.y, 1000.0);
        c.z = xdiv(c.z, 1000.0);
    }
    D3 m = dquat_rotate(fe->q, c);                                   // ios_rtab.py:189-190
    double d0 = xsub(xadd(m.x, fe->t[0]), p.x);
    double d1 = xsub(xadd(m.y, fe->t[1]), p.y);
    double d2 = xsub(xadd(m.z, fe->t[2]), p.z);
    double dist = __dsqrt_rn(xadd(xadd(xmul(d0, d0), xmul(d1, d1)), xmul(d2, d2)));   // fusion.py:224
    if (dist < P.radius_d) r |= EX_VIS | ((unsigned long long)cls << 8);             // fusion.py:225
    return r;
}

// ---- mbarrier / TMA bulk-copy helpers (cp.async.bulk: SASS UBLKCP), used by the fix-up kernel's frame-record staging
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// classification of one point-view by the fp32 path
//   st: 0 = certainly not seen, 1 = pixel certified (depth test pending), 2 = unsure -> fp64
struct Cls {
    int st, g_in;
    uint32_t puv;   // fu | fv << 16 when st != 0
    float z;
};

__device__ __forceinline__ Cls classify(const float4* __restrict__ s, const float4 pt, const float fW, const float fH) {
    Cls c;
    c.st = 0;
    c.g_in = 0;
    c.puv = 0;
    c.z = 0.f;
    const float4 A0 = s[0], A1 = s[1];
    const float d0 = (pt.x - A0.x) - A1.x;
    const float d1 = (pt.y - A0.y) - A1.y;
    const float d2 = (pt.z - A0.z) - A1.z;
    const float S = fabsf(d0) + fabsf(d1) + fabsf(d2) + 1.0e-9f;
    const float4 Mz = s[4];
    const float z = fmaf(Mz.x, d0, fmaf(Mz.y, d1, Mz.z * d2));
    // rounding bound of a row: 8 u * sum |M_i| |d_i| (M rounded to fp32: 1 u; d: 2 u; product and two fused adds: 3 u; the
    // sum itself is computed to 3 u) -- two to three times tighter than 8 u * max|M_i| * sum|d_i|, which is what decides
    // how many pixel floors have to be re-evaluated in fp64.  The 1e-9 m floor covers the absolute error of d near 0.
    const float ad0 = fabsf(d0), ad1 = fabsf(d1), ad2 = fabsf(d2);
    const float ez = 8.0f * F3D_U24 * fmaf(fabsf(Mz.x), ad0, fmaf(fabsf(Mz.y), ad1, fmaf(fabsf(Mz.z), ad2, Mz.w * 1.0e-9f)));
    c.z = z;
    if (z < -16.0f * ez) return c;
    if (z <= 16.0f * ez) {
        c.st = 2;
        return c;
    }
    const float4 Mu = s[2], Mv = s[3];
    const float a = fmaf(Mu.x, d0, fmaf(Mu.y, d1, Mu.z * d2));
    const float b = fmaf(Mv.x, d0, fmaf(Mv.y, d1, Mv.z * d2));
    // rcp.approx: relative error <= 2^-23 = 2 u (PTX ISA); with the rounding of a * r that is 3 u of the 8 u the bounds below
    // allow for u and v.  z > 16 ez > 0 is a normal number here.
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
    const float u = a * r, v = b * r;
    const float ea = 8.0f * F3D_U24 * fmaf(fabsf(Mu.x), ad0, fmaf(fabsf(Mu.y), ad1, fmaf(fabsf(Mu.z), ad2, Mu.w * 1.0e-9f)));
    const float eb = 8.0f * F3D_U24 * fmaf(fabsf(Mv.x), ad0, fmaf(fabsf(Mv.y), ad1, fmaf(fabsf(Mv.z), ad2, Mv.w * 1.0e-9f)));
    const float eu = 1.125f * (ea + fabsf(u) * ez) * r + 8.0f * F3D_U24 * fabsf(u) + 1.0e-4f;
    const float ev = 1.125f * (eb + fabsf(v) * ez) * r + 8.0f * F3D_U24 * fabsf(v) + 1.0e-4f;
    const float sl = fmaf(s[5].w, d0, fmaf(s[6].w, d1, s[7].w * d2));   // (p - eye) . lookat
    const float es = 8.0f * F3D_U24 * S + 1.0e-6f * A1.w;
    if ((u + eu < 0.f) || (u - eu >= fW) || (v + ev < 0.f) || (v - ev >= fH) || (sl - es > A1.w)) return c;
    const float fu = floorf(u), fv = floorf(v);
    const bool cu = (u - fu >= eu) && (fu + 1.0f - u > eu);
    const bool cv = (v - fv >= ev) && (fv + 1.0f - v > ev);
    const bool cf = (sl + es < A1.w);
    c.g_in = (u >= 0.f) && (fu < fW) && (v >= 0.f) && (fv < fH) && (sl < A1.w);
    const int iu = min(max((int)fu, 0), 65535), iv = min(max((int)fv, 0), 65535);
    c.puv = (uint32_t)iu | ((uint32_t)iv << 16);
    c.st = (cu && cv && cf) ? 1 : 2;
    return c;
}

// fp32 single-pixel criterion ||modPoints[pix] - p|| < radius in metric camera space, with its rounding band.
//   returns 1 = certainly inside, 0 = certainly outside, 3 = inside the band (g = fp32 guess)
__device__ __forceinline__ int distance_test(const float4* __restrict__ s, const float4 pt, const uint32_t puv, const float dm,
                                             const FuseParams& P, int& g) {
    const float4 A0 = s[0], A1 = s[1];
    const float d0 = (pt.x - A0.x) - A1.x;
    const float d1 = (pt.y - A0.y) - A1.y;
    const float d2 = (pt.z - A0.z) - A1.z;
    const float S = fabsf(d0) + fabsf(d1) + fabsf(d2) + 1.0e-9f;
    const float4 R0 = s[5], R1 = s[6], R2 = s[7];
    const float X = fmaf(R0.x, d0, fmaf(R0.y, d1, R0.z * d2));
    const float Y = fmaf(R1.x, d0, fmaf(R1.y, d1, R1.z * d2));
    const float Z = fmaf(R2.x, d0, fmaf(R2.y, d1, R2.z * d2));
    const float ds = dm * A0.w;                                   // depth scaled by |q|^2 (un-normalised rotate)
    const float xn = ((float)(puv & 0xffffu) - P.cx) * P.inv_fx, yn = ((float)(puv >> 16) - P.cy) * P.inv_fy;
    const float qx = X - xn * ds, qy = Y - yn * ds, qz = Z - ds;
    const float dist2 = fmaf(qx, qx, fmaf(qy, qy, qz * qz));
    const float del = 16.0f * F3D_U24 * (S + ds * (1.0f + fabsf(xn) + fabsf(yn)));
    const float rlo = fmaxf(P.radius - del, 0.f), rhi = P.radius + del;
    g = dist2 < P.radius * P.radius;
    if (dist2 < rlo * rlo * (1.0f - 16.0f * F3D_U24)) return 1;
    if (dist2 > rhi * rhi * (1.0f + 16.0f * F3D_U24)) return 0;
    return 3;
}

// One deferred (uncertain) point-view.  Each warp queues the pairs its fp32 sweep could not certify and evaluates
// them in fp64 afterwards with one entry per lane, so the expensive exact path runs on dense warps instead of one
// or two live lanes; no CTA-wide barrier is involved (queue, histogram rows and output rows are all warp-private).
struct Deferred {
    uint32_t w0, w1, w2;   // owner lane | frame (relative) << 16 ; pixel guess ; st | guess << 8
};
struct Tally {
    unsigned n_cand, n_seen;   // hot counters (registers); the rare ones (exact / diverged / near-edge / audit-bad) are
    unsigned* rare;            // shared-memory counters [F3D_STAT_*], bumped with atomics where they occur
    int total, best, bpos;   // running VotingSegmentation.segment state of this thread's point (fused resolve)
    int nlist;               // slot-record mode: distinct classes this point has received since the last flush
    uint8_t* clist;          // ... and their list, element j at clist[j * FUSE_BLOCK] (NULL outside slot-record mode)
};

// a vote for class `cls` of the thread's own point: bump the histogram and keep the running arg-max exact:
// best = max count so far, bpos = smallest filter position among the classes whose count equals best.
template <typename CellT>
__device__ __forceinline__ void cast_vote(CellT* hist, int row_off, int cls, const FuseParams& P, const FuseResolve& RP,
                                          Tally& t) {
    if (cls >= P.C1) return;
    const int v = (int)hist[row_off + cls] + 1;
    hist[row_off + cls] = (CellT)v;
    if (t.clist && v == 1) {   // first vote for this class: remember it, so the record is built without scanning the row
        if (t.nlist < FUSE_NSLOT) t.clist[t.nlist * FUSE_BLOCK] = (uint8_t)cls;
        ++t.nlist;
    }
    if (RP.enabled) {
        ++t.total;
        const int pos = RP.fpos[cls];
        if (pos >= 0 && (v > t.best || (v == t.best && pos < t.bpos))) {
            t.best = v;
            t.bpos = pos;
        }
    }
}

// fp64 decision for one point-view inside the sweep (audit mode, a full warp queue, the warp's own deferred pass when
// there is no workspace queue) and its consequence.  `st` >= 2: the fp32 path was unsure (g_in / pix = its guess);
// st < 2 (audit): the fp32 path certified `fast_seen` / `pix` / `fast_zq`.
template <int MODE, int FMT, typename CellT>
__device__ __forceinline__ void resolve_exact(const FuseParams& P, const FuseResolve& RP, const FrameRecord* __restrict__ frec,
                                              CellT* hist, int RS, int64_t tile_base, int owner_tid, int frel, float px, float py,
                                              float pz, int st, int g_in, int pix, bool fast_seen, uint32_t fast_zq,
                                              bool owner_is_self, unsigned* dirty_w, Tally& t) {
    const unsigned long long eo = exact_eval<MODE, FMT>(P, &frec[P.f_begin + frel].exact, frel, px, py, pz);
    const int e_in = (int)(eo & EX_IN), e_vis = (int)((eo >> 1) & 1ull), e_pix = (int)(eo >> 32);
    const bool e_seen = (MODE == MODE_SPLAT) ? (e_in != 0) : (e_vis != 0);
    const uint32_t e_zq = (MODE == MODE_SPLAT && e_in) ? (uint32_t)((eo >> 16) & 0xffffull) : 0u;
    if (st >= 2) {
        atomicAdd(t.rare + F3D_STAT_EXACT, 1u);
        bool diverged;
        if (st == 2) diverged = (g_in != e_in) || (e_in && pix != e_pix);
        else if (MODE == MODE_SPLAT) diverged = (!e_in) || (pix != e_pix) || ((uint32_t)g_in != e_zq);
        else diverged = (g_in != e_vis) || (e_in && pix != e_pix);
        if (diverged) atomicAdd(t.rare + F3D_STAT_DIVERGED, 1u);
    } else {
        // audit: a certified fp32 outcome must equal the fp64 outcome
        const bool bad = (fast_seen != e_seen) || (fast_seen && pix != e_pix) || (fast_seen && MODE == MODE_SPLAT && fast_zq != e_zq);
        if (bad) atomicAdd(t.rare + F3D_STAT_AUDIT_BAD, 1u);
    }
    if (e_in && (eo & EX_EDGE)) atomicAdd(t.rare + F3D_STAT_NEAR_EDGE, 1u);
    if (e_seen) {
        ++t.n_seen;
        if (MODE == MODE_VOTE) {
            const int cls = (int)((eo >> 8) & 0xffull);
            if (owner_is_self) {
                cast_vote(hist, owner_tid * RS, cls, P, RP, t);
            } else if (cls < P.C1) {
                // another lane's row: 32-bit atomic on the word holding the counter (cannot carry: the counter bound is
                // enforced by the flush rule / the frames-per-launch limit); the owner re-derives its arg-max from the
                // row afterwards (dirty bit)
                const int h = owner_tid * RS + cls;
                if (sizeof(CellT) == 2) atomicAdd(reinterpret_cast<unsigned*>(hist) + (h >> 1), (h & 1) ? 0x10000u : 1u);
                else atomicAdd(reinterpret_cast<unsigned*>(hist) + (h >> 2), 1u << (8 * (h & 3)));
                atomicOr(dirty_w, 1u << (owner_tid & 31));
            }
        } else if (MODE == MODE_SPLAT) {
            atomicMin(P.zbuf + (size_t)frel * (size_t)(P.H * P.W) + (size_t)e_pix, e_zq);
        } else {
            atomicMax(P.uv2pt + (size_t)frel * (size_t)(P.H * P.W) + (size_t)e_pix, (int)(tile_base + owner_tid));
        }
    }
}

template <int HB> struct HistCell { typedef uint16_t T; };
template <> struct HistCell<1> { typedef uint8_t T; };

__device__ __forceinline__ void st_global_v8(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                             uint32_t a6, uint32_t a7) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5),
                 "r"(a6), "r"(a7)
                 : "memory");
}

// ---- byte-histogram flush: the warp's 32 rows (32*C1 contiguous bytes, row stride == C1) -> HBM ------------------
// first flush of a non-accumulating launch overwrites (every cell written exactly once, 32-byte stores: STG.256); later
// flushes add their non-zero cells.  Rows are warp-private, so no CTA barrier and no atomics are involved.
// V8: 32-byte stores.  Only the inlined epilogue instance uses them: inside a non-inlined (ABI) function ptxas 12.9 was seen
// to lower the st.global.v8 of some kernel instantiations to a single 32-bit store.
template <bool V8>
__device__ __forceinline__ void flush8(const FuseParams& P, uint8_t* hist, int warp, int lane, int64_t tile_base, bool add,
                                       bool rezero, int nlist, const uint8_t* clist, uint16_t* stg, int level, bool dirty, uint2* dcache) {
    const int row0 = warp * 32;
    const int nrows = (int)max((int64_t)0, min((int64_t)32, P.N - tile_base - row0));
    const int total = nrows * P.C1;
    const int n4 = total >> 2;
    uint32_t* __restrict__ h32 = reinterpret_cast<uint32_t*>(hist + row0 * P.C1);   // 32*C1 bytes per warp: 32-byte aligned
    const uint8_t* __restrict__ h8 = hist + row0 * P.C1;
    if (P.xg_G > 0) {
        // slot records (see FuseParams)
        const long long p0 = tile_base + row0;
        const int d = (int)(p0 / P.xg_per);
        const bool live = lane < nrows;
        // flush number `level` of this warp writes directory level `level`; flushes beyond F3D_XCH_NLEVEL and rows another
        // lane's deferred pass touched go cell by cell to the owner's queue
        const bool rec = level < F3D_XCH_NLEVEL;
        bool spill = live && (!rec || dirty);
        const int C1 = P.C1;
        const uint8_t* __restrict__ row = h8 + lane * C1;
        int n = (live && !spill) ? min(nlist, C1) : 0;        // classes of this lane's point
        int L = n;
#pragma unroll
        for (int s2 = 16; s2 > 0; s2 >>= 1) L = max(L, __shfl_xor_sync(0xffffffffu, L, s2));
        unsigned off = 0;
        if (rec) {
            const unsigned reg = (blockIdx.x * FUSE_NW + warp) & (F3D_XCH_NREG - 1);
            if (lane == 0 && L > 0) off = atomicAdd(P.xg_rowcur + d * F3D_XCH_NREG + reg, (unsigned)L);
            off = __shfl_sync(0xffffffffu, off, 0);
            if (off + (unsigned)L > P.xg_subrows) {   // sub-region full: everything of this block goes to the queue
                spill = live;
                n = 0;
                L = 0;
            }
            off += reg * P.xg_subrows;
            if (lane == 0) dcache[level] = make_uint2(off, (unsigned)L);
        }
        if (!rezero && nrows > 0) {
            // last flush of the tile: publish all directory levels of the block at once (unused levels are zero)
            __syncwarp();
            if (lane < F3D_XCH_NLEVEL)
                P.xg_dir[d][((p0 - (long long)d * P.xg_per) >> 5) * F3D_XCH_NLEVEL + lane] = dcache[lane];
        }
        uint4* st4 = reinterpret_cast<uint4*>(stg);
        uint4* __restrict__ dst = reinterpret_cast<uint4*>(P.xg_slots[d] + (size_t)off * 32);
        int scan_c = 0;                                           // row-scan position of a long list (n > FUSE_NSLOT)
        for (int j0 = 0; j0 < L; j0 += FUSE_STG_ROWS) {           // FUSE_STG_ROWS rows (512 B) at a time through the staging block
            const int rows_here = min(FUSE_STG_ROWS, L - j0);
            for (int i = lane; i < rows_here * 4; i += 32) st4[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            if (n <= FUSE_NSLOT) {
                const int j1 = min(n, j0 + FUSE_STG_ROWS);
                for (int j = j0; j < j1; ++j) {                   // the classes cast_vote remembered
                    const unsigned cls = clist[j * FUSE_BLOCK];
                    stg[(j - j0) * 32 + lane] = (uint16_t)(cls | ((unsigned)row[cls] << 8));
                }
            } else {
                int j = 0;                                        // long list: next non-zero cells of the row
                while (j < rows_here && scan_c < C1) {
                    const unsigned v = row[scan_c];
                    if (v) {
                        stg[j * 32 + lane] = (uint16_t)((unsigned)scan_c | (v << 8));
                        ++j;
                    }
                    ++scan_c;
                }
            }
            __syncwarp();
            for (int i = lane; i < rows_here * 4; i += 32) dst[j0 * 4 + i] = st4[i];
            __syncwarp();
        }
        if (spill) {
            const unsigned long long key0 = (unsigned long long)(p0 + lane - (long long)d * P.xg_per) * (unsigned long long)C1;
            const unsigned sub = F3D_XCH_NSUB_FIX + blockIdx.x % (F3D_XCH_NSUB - F3D_XCH_NSUB_FIX);
            for (int c = 0; c < C1; ++c) {
                const unsigned v = row[c];
                if (v) xg_append(P, d, sub, atomicAdd(P.xg_qcur + d * F3D_XCH_NSUB + sub, 1u), key0 + c, v);
            }
        }
    } else if (nrows > 0) {
        if (P.votes) {
            int32_t* __restrict__ out = P.votes + (tile_base + row0) * P.C1;
            if (!add) {
                if (V8 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0) {
                    // eight cells per lane and step: LDS.64 -> 8 x PRMT -> one 32-byte store (1 KB contiguous per warp instruction)
                    const uint2* __restrict__ h64 = reinterpret_cast<const uint2*>(h8);
                    const int n8 = total >> 3;
                    for (int i = lane; i < n8; i += 32) {
                        const uint2 w = h64[i];
                        st_global_v8(out + 8 * i, __byte_perm(w.x, 0u, 0x4440), __byte_perm(w.x, 0u, 0x4441), __byte_perm(w.x, 0u, 0x4442),
                                     __byte_perm(w.x, 0u, 0x4443), __byte_perm(w.y, 0u, 0x4440), __byte_perm(w.y, 0u, 0x4441),
                                     __byte_perm(w.y, 0u, 0x4442), __byte_perm(w.y, 0u, 0x4443));
                    }
                    for (int e = (n8 << 3) + lane; e < total; e += 32) out[e] = (int)h8[e];
                } else {
                    for (int i = lane; i < n4; i += 32) {
                        const uint32_t w = h32[i];
                        const int4 v4 = make_int4((int)__byte_perm(w, 0u, 0x4440), (int)__byte_perm(w, 0u, 0x4441),
                                                  (int)__byte_perm(w, 0u, 0x4442), (int)__byte_perm(w, 0u, 0x4443));
                        *reinterpret_cast<int4*>(out + 4 * i) = v4;
                    }
                    for (int e = (n4 << 2) + lane; e < total; e += 32) out[e] = (int)h8[e];
                }
            } else {
                for (int i = lane; i < n4; i += 32) {
                    const uint32_t w = h32[i];
                    if (!w) continue;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int v = (int)((w >> (8 * b)) & 0xffu);
                        if (v) out[4 * i + b] += v;
                    }
                }
                for (int e = (n4 << 2) + lane; e < total; e += 32)
                    if (h8[e]) out[e] += (int)h8[e];
            }
        }
        if (P.votes16) {
            uint16_t* __restrict__ out = P.votes16 + (tile_base + row0) * P.C1;
            if (!add) {
                for (int i = lane; i < n4; i += 32) {
                    const uint32_t w = h32[i];
                    *reinterpret_cast<uint2*>(out + 4 * i) =
                        make_uint2((w & 0xffu) | ((w << 8) & 0xff0000u), ((w >> 16) & 0xffu) | ((w >> 8) & 0xff0000u));
                }
                for (int e = (n4 << 2) + lane; e < total; e += 32) out[e] = (uint16_t)h8[e];
            } else {
                for (int i = lane; i < n4; i += 32) {
                    const uint32_t w = h32[i];
                    if (!w) continue;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const unsigned v = (w >> (8 * b)) & 0xffu;
                        if (v) out[4 * i + b] += (uint16_t)v;
                    }
                }
                for (int e = (n4 << 2) + lane; e < total; e += 32)
                    if (h8[e]) out[e] += (uint16_t)h8[e];
            }
        }
    }
    if (rezero) {
        __syncwarp();
        for (int i = lane; i < 8 * P.C1; i += 32) h32[i] = 0u;
        __syncwarp();
    }
}

// The sweep's own flush (a warp passed FUSE_LIMIT8 candidates): rare, so NOT inlined -- the hot loop must stay small enough
// for the instruction cache (inlined there it made `no_instruction` the third largest stall reason).
static __device__ __noinline__ void flush8_mid(const FuseParams& P, uint8_t* hist, int warp, int lane, int64_t tile_base, bool add,
                                               int nlist, const uint8_t* clist, uint16_t* stg, int level, uint2* dcache) {
    flush8<false>(P, hist, warp, lane, tile_base, add, true, nlist, clist, stg, level, false, dcache);
}

// Outputs of a warp whose 32 points received no vote in this launch, written without touching the histogram: zero rows (or
// zero directory entries in exchange mode; nothing when accumulating), label = unclassified (voting.py:126, total == 0).
__device__ __forceinline__ void write_no_votes(const FuseParams& P, const FuseResolve& RP, int warp, int lane, int64_t tile_base, int64_t gi,
                                               bool active) {
    const int row0 = warp * 32;
    const int nrows = (int)max((int64_t)0, min((int64_t)32, P.N - tile_base - row0));
    if (P.xg_G > 0) {
        if (nrows > 0 && lane < F3D_XCH_NLEVEL) {
            const long long p0 = tile_base + row0;
            const int d = (int)(p0 / P.xg_per);
            P.xg_dir[d][((p0 - (long long)d * P.xg_per) >> 5) * F3D_XCH_NLEVEL + lane] = make_uint2(0u, 0u);
        }
    } else if (!P.accumulate && nrows > 0) {
        if (P.votes) {
            uint4* out = reinterpret_cast<uint4*>(P.votes + (tile_base + row0) * P.C1);   // 128 * C1 bytes per warp: 16-byte multiple
            for (int i = lane; i < nrows * P.C1 / 4; i += 32) out[i] = make_uint4(0u, 0u, 0u, 0u);
            for (int e = (nrows * P.C1 / 4) * 4 + lane; e < nrows * P.C1; e += 32) P.votes[(tile_base + row0) * P.C1 + e] = 0;
        }
        if (P.votes16)
            for (int e = lane; e < nrows * P.C1; e += 32) P.votes16[(tile_base + row0) * P.C1 + e] = 0;
    }
    if (RP.enabled && active) {
        P.labels[gi] = (int64_t)RP.unclassified;
        if (P.summ) P.summ[gi] = summ_pack(0, 0, 0x7fff);
    }
}

// small per-CTA scalars in shared memory
struct __align__(16) FuseShared {
    uint2 dcache[FUSE_NW][F3D_XCH_NLEVEL];     // exchange mode: the warp's directory levels {row offset, L}
    int nq[FUSE_NW];                           // per-warp deferred counts
    unsigned dirty[FUSE_NW];                   // rows another lane's deferred pass touched
    unsigned stat[8];                          // CTA totals of the statistics, flushed once at the end
};

// 16-byte asynchronous global -> shared copy (LDGSTS): the projection tiles of the NEXT group land in the warp's other
// staging buffer while the current group is processed; no registers are tied up and nobody but the warp itself waits
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int NPEND> __device__ __forceinline__ void cp_async_wait_pending() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }
// 4-byte asynchronous gather with the 64-byte L2 fill (see gather_u32)
__device__ __forceinline__ void cp_async4_l2_64(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global.L2::64B [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

__host__ __device__ constexpr size_t fuse_align16(size_t v) { return (v + 15) & ~(size_t)15; }
// shared-memory layout: [stage: NW x FUSE_TR FrameFast][texel rings: NW x FUSE_GR x 32 u32][FuseShared][per-warp candidate lists: NW x FUSE_WCAP u16][deferred queues][hist]
//                       [exchange mode: class lists u8 x FUSE_NSLOT x FUSE_BLOCK][record staging: NW x 512 B]
#ifndef FUSE_PD
#define FUSE_PD 2   // pipeline depth of the sweep: candidate it - PD is consumed while tile it + PD and texel it are started
#endif
#define FUSE_TR 8   // per-warp ring of staged projection tiles (candidates it-PD .. it+PD are live)
#define FUSE_GR 4   // per-lane ring of gathered texels (candidates it-PD .. it)
#define FUSE_OFF_TEXR ((size_t)FUSE_NW * FUSE_TR * sizeof(FrameFast))
#define FUSE_OFF_SHARED (FUSE_OFF_TEXR + (size_t)FUSE_NW * FUSE_GR * 32 * sizeof(uint32_t))
#define FUSE_OFF_WLIST (FUSE_OFF_SHARED + sizeof(FuseShared))
#define FUSE_OFF_QUEUE fuse_align16(FUSE_OFF_WLIST + (size_t)FUSE_NW * FUSE_WCAP * sizeof(uint16_t))
#define FUSE_OFF_HIST fuse_align16(FUSE_OFF_QUEUE + (size_t)FUSE_NW * FUSE_QWARP * sizeof(Deferred))

template <int MODE, int FMT, int HB, bool AUDIT>
__global__ void __launch_bounds__(FUSE_BLOCK, (MODE == MODE_VOTE && HB == 1) ? FUSE_MINB8 : FUSE_MINB)
    fuse_kernel(const __grid_constant__ FuseParams P, const __grid_constant__ FuseResolve RP) {
    typedef typename HistCell<HB>::T CellT;
    static_assert(FUSE_NW <= 8, "cmask holds one bit per warp");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* stage = reinterpret_cast<float4*>(smem_raw);
    FuseShared& sh = *reinterpret_cast<FuseShared*>(smem_raw + FUSE_OFF_SHARED);
    Deferred* queue_all = reinterpret_cast<Deferred*>(smem_raw + FUSE_OFF_QUEUE);
    CellT* hist = reinterpret_cast<CellT*>(smem_raw + FUSE_OFF_HIST);
    const int RS = P.RS;
    // exchange mode only: [class lists u8 x FUSE_NSLOT x FUSE_BLOCK][staging blocks: NW x 512 B] after the histogram
    uint8_t* clist_all = reinterpret_cast<uint8_t*>(hist) + fuse_align16((size_t)FUSE_BLOCK * RS * sizeof(CellT));
    uint16_t* stg_all = reinterpret_cast<uint16_t*>(clist_all + FUSE_NSLOT * FUSE_BLOCK);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    Deferred* queue = queue_all + warp * FUSE_QWARP;
    uint32_t* texring = reinterpret_cast<uint32_t*>(smem_raw + FUSE_OFF_TEXR);
    unsigned tile_id = blockIdx.x;
    if (P.live_list) tile_id = __ldg(P.live_list + blockIdx.x / FUSE_ST_TILES) * FUSE_ST_TILES + blockIdx.x % FUSE_ST_TILES;
    const int64_t tile_base = (int64_t)tile_id * FUSE_BLOCK;
    if (tile_base >= P.N) return;              // tail tiles of the cloud's last (partial) super-tile in a compacted launch
    const int64_t gi = tile_base + tid;
    const bool active = gi < P.N;
    const int HW = P.H * P.W;
    const float fW = (float)P.W, fH = (float)P.H;

    // frames to test: the super-tile's candidate list when the first cull level ran, else every frame of the launch
    const unsigned st_n = P.st_count ? __ldg(P.st_count + tile_id / FUSE_ST_TILES) : 0xffffffffu;
    const bool use_list = st_n != 0xffffffffu;
    const uint16_t* __restrict__ st_list = P.st_list + (size_t)(tile_id / FUSE_ST_TILES) * FUSE_ST_LCAP;
    const int ntest = use_list ? (int)st_n : P.f_end - P.f_begin;
    // the first 32 list entries are fetched before the count is known (the list storage always exists): one dependent
    // memory round trip less on the way to the first vote, which matters for short sweeps (C1: 3.8 candidates per point)
    const int first_frel = P.st_count ? (int)__ldg(st_list + lane) : lane;
    if (ntest == 0) {
        // No frame of this launch can see the tile's super-tile (the common case of a rank whose frames look at another part
        // of the building): nothing is loaded or cleared, the outputs of "no votes" are written straight away.
        if (MODE == MODE_VOTE) write_no_votes(P, RP, warp, lane, tile_base, gi, active);
        return;
    }

    float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) pt = __ldg(P.points + gi);

    if (lane == 0) {
        sh.nq[warp] = 0;
        sh.dirty[warp] = 0u;
    }
    if (tid < 8) sh.stat[tid] = 0u;
    if (lane < F3D_XCH_NLEVEL) sh.dcache[warp][lane] = make_uint2(0u, 0u);

    if (MODE == MODE_VOTE) {   // every warp clears its own 32 rows (a thread's row is tid * RS) while its points are in flight
        uint4* h128 = reinterpret_cast<uint4*>(hist + (size_t)warp * 32 * RS);
        const int n128 = 32 * RS * (int)sizeof(CellT) / 16;
        for (int i = lane; i < n128; i += 32) h128[i] = make_uint4(0u, 0u, 0u, 0u);
    }

    // ---- per-warp bounding box (exact min / max of the float32 coordinates of the warp's 32 points), in every lane
    const float big = 3.0e38f;
    float blo0 = active ? pt.x : big, blo1 = active ? pt.y : big, blo2 = active ? pt.z : big;
    float bhi0 = active ? pt.x : -big, bhi1 = active ? pt.y : -big, bhi2 = active ? pt.z : -big;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        blo0 = fminf(blo0, __shfl_xor_sync(0xffffffffu, blo0, s));
        blo1 = fminf(blo1, __shfl_xor_sync(0xffffffffu, blo1, s));
        blo2 = fminf(blo2, __shfl_xor_sync(0xffffffffu, blo2, s));
        bhi0 = fmaxf(bhi0, __shfl_xor_sync(0xffffffffu, bhi0, s));
        bhi1 = fmaxf(bhi1, __shfl_xor_sync(0xffffffffu, bhi1, s));
        bhi2 = fmaxf(bhi2, __shfl_xor_sync(0xffffffffu, bhi2, s));
    }
    const float bmag = fabsf(blo0) + fabsf(blo1) + fabsf(blo2) + fabsf(bhi0) + fabsf(bhi1) + fabsf(bhi2);
    const bool warp_live = tile_base + warp * 32 < P.N;

    const FrameRecord* __restrict__ frec = reinterpret_cast<const FrameRecord*>(P.table);

    Tally T;
    T.n_cand = T.n_seen = 0u;
    T.rare = sh.stat;
    T.nlist = 0;
    T.clist = (MODE == MODE_VOTE && HB == 1 && P.xg_G > 0) ? clist_all + threadIdx.x : nullptr;
    T.total = 0;
    T.best = 0;
    T.bpos = 0x7fff;
    int since_flush = 0, nflush = 0;   // byte histogram: candidates swept since the warp's last flush, its flushes so far

    __syncthreads();       // the CTA's statistics counters are zeroed; from here on every warp is on its own
    // The warps of a CTA never meet again before the final statistics flush: each one culls the listed frames against
    // ITS OWN box (one frame per lane), keeps the survivors in its own list and sweeps them.  In a spatially sorted cloud a
    // warp's points are a few centimetres apart, so a warp is almost always entirely inside or entirely outside a frustum:
    // the warp-box test removes ~45 % of the (warp, frame) pairs of the frames that touch the tile before any per-point work.
    uint16_t* wlist = reinterpret_cast<uint16_t*>(smem_raw + FUSE_OFF_WLIST) + warp * FUSE_WCAP;
    int tbase = 0;         // next entry of the frame list to cull (warp-uniform)
    bool hist_live = false;   // did any listed frame reach this warp's box?  (if not, its outputs are written without the histogram)
    while (warp_live && tbase < ntest) {
        // ---- conservative (frame, warp box) cull (fp32 + explicit rounding margin; never drops a visible pair)
        int ncand = 0;
        while (tbase < ntest && ncand <= FUSE_WCAP - 32) {
            const int fi = tbase + lane;
            bool kw = fi < ntest;
            int frel = 0;
            if (kw) {
                frel = use_list ? (tbase == 0 ? first_frel : (int)__ldg(st_list + fi)) : fi;
                const float4* pl = frec[P.f_begin + frel].cull.pl;
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    const float4 q = __ldg(pl + m);
                    const float mx = fmaxf(q.x * blo0, q.x * bhi0) + fmaxf(q.y * blo1, q.y * bhi1) + fmaxf(q.z * blo2, q.z * bhi2) - q.w;
                    const float margin = 2.0e-6f * (bmag + fabsf(q.w)) + 1.0e-7f;
                    kw = kw && (mx >= -margin);
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, kw);
            if (kw) wlist[ncand + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)frel;
            ncand += __popc(bal);
            tbase += 32;
        }
        hist_live = hist_live || ncand > 0;
        __syncwarp();

        // ---- warp-autonomous sweep over this warp's candidates of the list: ONE candidate per iteration, software pipelined
        // through shared memory so that the loop body stays a few hundred instructions (the L0 instruction cache holds
        // ~380; the NB-times unrolled body of the first version spent more cycles on instruction fetch than on gathers):
        //   iteration i:  wait until tile i and texel i-2 have landed | start the async copy of tile i+2 |
        //                 project candidate i in fp32, start the async 4-byte copy of its texel into this lane's ring slot |
        //                 depth test + vote of candidate i-2 from its landed texel.
        // All copies are cp.async (LDGSTS): no registers are tied up while they are in flight, two gathers per lane are
        // outstanding across iterations, and nobody but the warp itself ever waits.  The two-array frame formats (a 2-byte
        // depth sample cannot be a cp.async) and the splat consume their candidate in the same iteration instead.
        constexpr bool PIPE = (MODE != MODE_SPLAT) && (FMT >= F3D_FRAMES_U32);
        int li = 0;             // next entry of the warp's candidate list (warp-uniform)
        float4* tiles = stage + warp * (FUSE_TR * 8);
        uint32_t* texr = texring + warp * (FUSE_GR * 32) + lane;
        auto pop = [&]() -> int {   // next candidate frame of this warp, -1 = none (warp-uniform)
            if (li >= ncand) return -1;
            return (int)wlist[li++];
        };
        auto stage_tile = [&](int slot, int f) {   // lanes 0..7 copy the 8 x 16 bytes of the frame's projection tile
            if (f >= 0 && lane < 8) cp_async16(tiles + slot * 8 + lane, reinterpret_cast<const uint4*>(&frec[P.f_begin + f].fast) + lane);
        };
        // depth validity + distance criterion + vote (or splat / uv2pt write) of one classified candidate
        auto consume = [&](const int frel, const float4* __restrict__ s, const uint32_t puv, const unsigned sg, const uint32_t g0,
                           const uint32_t g1, const float zc) {
            int st = (int)(sg & 7u);
            int g_in = (int)((sg >> 3) & 1u);
            const int pix = (int)((puv >> 16) * P.W + (puv & 0xffffu));
            uint32_t zq = 0, cls = 0;
            if (st == 1 && MODE != MODE_SPLAT) {
                float dm;
                bool valid;
                if (FMT != F3D_DEPTH_F32_M) {
                    const uint32_t d = g0 & 0xffffu;
                    valid = (d >= P.d_lo) && (d <= P.d_hi);
                    dm = (float)d * 0.001f;
                } else {
                    dm = __uint_as_float(g0);
                    valid = ((double)dm > P.zmin) && ((double)dm <= P.zmax);
                }
                if (MODE == MODE_VOTE) cls = (FMT < F3D_FRAMES_U32) ? g1 : ((g0 >> 16) & 0xffu);
                st = valid ? distance_test(s, pt, puv, dm, P, g_in) : 0;
            }
            if (MODE == MODE_SPLAT && st == 1) {
                // quantised camera z: floor(z*1000 + 0.5); certify the floor
                const float zm = fmaf(zc, 1000.0f, 0.5f);
                const float fz = floorf(zm);
                const float ez = 8.0f * F3D_U24 * s[4].w *
                                 (fabsf((pt.x - s[0].x) - s[1].x) + fabsf((pt.y - s[0].y) - s[1].y) + fabsf((pt.z - s[0].z) - s[1].z) + 1.0e-9f);
                const float eq = 1010.0f * ez + 8.0f * F3D_U24 * zm;
                if ((zm - fz >= eq) && (fz + 1.0f - zm > eq)) {
                    zq = (uint32_t)fminf(fmaxf(fz, 1.0f), 65535.0f);
                } else {
                    st = 3;
                    g_in = (int)fminf(fmaxf(fz, 1.0f), 65535.0f);
                }
            }
            const bool seen = (st == 1);
            if (st >= 2 || AUDIT) {
                // defer to the warp's dense fp64 pass (queue full or audit sweep: evaluate inline)
                const int slot = (st >= 2 && !AUDIT) ? atomicAdd(sh.nq + warp, 1) : FUSE_QWARP;
                if (slot < FUSE_QWARP) {
                    queue[slot].w0 = (uint32_t)tid | ((uint32_t)frel << 16);
                    queue[slot].w1 = (uint32_t)pix;
                    queue[slot].w2 = (uint32_t)st | ((uint32_t)g_in << 8);
                } else {
                    resolve_exact<MODE, FMT, CellT>(P, RP, frec, hist, RS, tile_base, tid, frel, pt.x, pt.y, pt.z, st, g_in, pix, seen, zq,
                                                    true, sh.dirty + warp, T);
                }
            } else if (seen) {
                ++T.n_seen;
                if (MODE == MODE_VOTE) {
                    cast_vote(hist, tid * RS, (int)cls, P, RP, T);
                } else {
                    const size_t off = (size_t)frel * (size_t)HW + (size_t)pix;
                    if (MODE == MODE_SPLAT) atomicMin(P.zbuf + off, zq);
                    else atomicMax(P.uv2pt + off, (int)gi);
                }
            }
        };

        // Pipeline depth PD (FUSE_PD): tile `it + PD` and texel `it` are started in iteration `it`, candidate `it - PD` is consumed.
        // Frame ids of the candidates in the pipeline live in warp-uniform registers: ahead[k] = candidate `it + k` (k < PD, tiles
        // staged), behind[k] = candidate `it - 1 - k` (texels in flight / landed); -1 = none.
        constexpr int PD = FUSE_PD;
        static_assert(PD >= 1 && 2 * PD + 1 <= FUSE_TR && PD + 1 <= FUSE_GR, "rings too small for this pipeline depth");
        int ahead[PD], behind[PD];
        uint32_t puvq[PD];      // pixel of candidates it-1 .. it-PD
        unsigned sgq[PD];       // their st | g_in << 3
#pragma unroll
        for (int k = 0; k < PD; ++k) {
            ahead[k] = pop();
            stage_tile(k, ahead[k]);
            cp_async_commit();
            behind[k] = -1;
            puvq[k] = 0;
            sgq[k] = 0;
        }
#pragma unroll 1
        for (int it = 0;; ++it) {
            cp_async_wait_pending<PD - 1>();   // every copy of this lane except the newest PD-1 groups: tile `it`, texel `it - PD` have landed
            __syncwarp();                      // ... everybody else's too; every lane is done with iteration it - 1
            const int f0 = ahead[0];
            const int fl = PIPE ? behind[PD - 1] : -1;
            bool drained = f0 < 0;
            if (PIPE) {
#pragma unroll
                for (int k = 0; k < PD; ++k) drained = drained && behind[k] < 0;
            }
            if (drained) break;
            const int fn = pop();
            stage_tile((it + PD) & (FUSE_TR - 1), fn);
#pragma unroll
            for (int k = PD - 1; k > 0; --k) behind[k] = behind[k - 1];
            behind[0] = f0;
#pragma unroll
            for (int k = 0; k + 1 < PD; ++k) ahead[k] = ahead[k + 1];
            ahead[PD - 1] = fn;
            if (MODE == MODE_VOTE && HB == 1 && f0 >= 0) {
                // a byte counter holds 255: flush the warp's rows before this candidate could push a cell past the limit
                if (since_flush + 1 > FUSE_LIMIT8) {
                    flush8_mid(P, reinterpret_cast<uint8_t*>(hist), warp, lane, tile_base, nflush > 0 || P.accumulate, T.nlist, T.clist,
                               stg_all + warp * (FUSE_STG_ROWS * 32), nflush, sh.dcache[warp]);
                    T.nlist = 0;
                    ++nflush;
                    since_flush = PD;   // the candidates still in the pipeline vote after this flush
                }
                ++since_flush;
            }
            // ---- candidate `it`: fp32 projection + certification; its frame gather starts
            uint32_t puv0 = 0;
            unsigned sg0 = 0;
            if (f0 >= 0 && active) {
                ++T.n_cand;
                const float4* s = tiles + (it & (FUSE_TR - 1)) * 8;
                const Cls c = classify(s, pt, fW, fH);
                puv0 = c.puv;
                sg0 = (unsigned)(c.st | (c.g_in << 3));
                if (PIPE) {
                    if (c.st == 1)
                        cp_async4_l2_64(texr + (it & (FUSE_GR - 1)) * 32,
                                        reinterpret_cast<const uint32_t*>(P.depth) + frame_off<FMT>(P, f0, (int)(puv0 & 0xffffu), (int)(puv0 >> 16)));
                } else if (c.st != 0 || AUDIT) {
                    uint32_t g0 = 0, g1 = 0;
                    if (MODE != MODE_SPLAT && c.st == 1) {
                        const size_t off = frame_off<FMT>(P, f0, (int)(puv0 & 0xffffu), (int)(puv0 >> 16));
                        if (FMT == F3D_DEPTH_U16_MM) g0 = gather_u16(reinterpret_cast<const uint16_t*>(P.depth) + off);
                        else g0 = gather_u32(reinterpret_cast<const uint32_t*>(P.depth) + off);
                        if (MODE == MODE_VOTE && FMT < F3D_FRAMES_U32) g1 = gather_u8(P.mask + off);
                    }
                    consume(f0, s, puv0, sg0, g0, g1, c.z);
                }
            }
            cp_async_commit();   // tile it + PD and texel it travel as one group
            // ---- candidate `it - PD`: its texel has landed
            if (PIPE) {
                if (fl >= 0 && active && ((sgq[PD - 1] & 7u) != 0u || AUDIT))
                    consume(fl, tiles + ((it + FUSE_TR - PD) & (FUSE_TR - 1)) * 8, puvq[PD - 1], sgq[PD - 1],
                            texr[((it + FUSE_GR - PD) & (FUSE_GR - 1)) * 32], 0u, 0.f);
#pragma unroll
                for (int k = PD - 1; k > 0; --k) {
                    puvq[k] = puvq[k - 1];
                    sgq[k] = sgq[k - 1];
                }
                puvq[0] = puv0;
                sgq[0] = sg0;
            }
        }
        cp_async_wait_all();
        __syncwarp();          // every lane is done with the list before the next cull pass overwrites it
    }

    if (MODE == MODE_VOTE && !hist_live) {
        // no listed frame reaches this warp's box: no candidate, no deferred pair, no vote
        if (warp_live) write_no_votes(P, RP, warp, lane, tile_base, gi, active);
    } else {
    // ---- warp-private dense fp64 pass over the deferred point-views (one entry per lane)
    __syncwarp();
    {
        const int nq = min(sh.nq[warp], FUSE_QWARP);
        // preferred: hand the uncertain pairs to the dense fix-up kernels through the caller's workspace queue
        bool flushed = false;
        if (P.gq && nq > 0) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.gq_count, (unsigned long long)nq);
            base = __shfl_sync(0xffffffffu, base, 0);
            flushed = base + (unsigned long long)nq <= P.gq_cap;
            if (lane < nq && base + lane < P.gq_cap) {
                const Deferred d = queue[lane];
                GEntry e;
                e.pt = flushed ? (int)(tile_base + (d.w0 & 0xffffu)) : -1;   // -1: slot reserved but unused (queue overflow)
                e.w = (d.w0 >> 16) | ((d.w2 & 0xffu) << 16);
                e.pix = (int)d.w1;
                e.guess = d.w2 >> 8;
                P.gq[base + lane] = e;
            }
        }
        if (!flushed && lane < nq) {
            const Deferred d = queue[lane];
            const int owner = (int)(d.w0 & 0xffffu), frel = (int)(d.w0 >> 16);
            const float4 op = __ldg(P.points + tile_base + owner);
            resolve_exact<MODE, FMT, CellT>(P, RP, frec, hist, RS, tile_base, owner, frel, op.x, op.y, op.z, (int)(d.w2 & 0xffu),
                                            (int)(d.w2 >> 8), (int)d.w1, false, 0u, false, sh.dirty + warp, T);
        }
    }
    __syncwarp();

    // ---- epilogue (warp-private rows): histogram -> HBM, written once; fused label resolve
    if constexpr (MODE == MODE_VOTE && HB == 1) {
        uint8_t* hist8 = reinterpret_cast<uint8_t*>(hist);
        flush8<true>(P, hist8, warp, lane, tile_base, nflush > 0 || P.accumulate, false, T.nlist, T.clist, stg_all + warp * (FUSE_STG_ROWS * 32),
               nflush, ((sh.dirty[warp] >> lane) & 1u) != 0u, sh.dcache[warp]);
        if (RP.enabled && active) {
            // VotingSegmentation.segment (voting.py:120-135).  The running (total, best, bpos) is exact unless another
            // lane's deferred pass added votes to this row (re-derived from the row) or the warp flushed more than
            // once (re-derived from the complete row in HBM, which this warp has just written).
            const bool multi = nflush > 0 && (P.votes || P.votes16);
            if (multi) __syncwarp();
            if (multi || ((sh.dirty[warp] >> lane) & 1u)) {
                T.total = 0;
                T.best = 0;
                T.bpos = 0x7fff;
                for (int c = 0; c < P.C1; ++c) {
                    int v;
                    if (!multi) v = hist8[tid * RS + c];
                    else if (P.votes) v = P.votes[(size_t)gi * P.C1 + c];
                    else v = P.votes16[(size_t)gi * P.C1 + c];
                    T.total += v;
                    const int pos = RP.fpos[c];
                    if (v > 0 && pos >= 0 && (v > T.best || (v == T.best && pos < T.bpos))) {
                        T.best = v;
                        T.bpos = pos;
                    }
                }
            }
            bool unc = (T.total <= 0) || (T.best <= 0);                                  // voting.py:126,131
            if (!unc) unc = xdiv((double)T.best, (double)T.total) < RP.threshold;         // voting.py:128-130
            P.labels[gi] = (int64_t)(unc ? RP.unclassified : RP.remap[T.bpos]);
            if (P.summ) P.summ[gi] = summ_pack(T.total, T.best, T.bpos);
        }
    } else if constexpr (MODE == MODE_VOTE) {
        const int row0 = warp * 32;
        const int nrows = (int)max((int64_t)0, min((int64_t)32, P.N - tile_base - row0));
        if (P.votes16 && nrows > 0) {
            uint16_t* __restrict__ out = P.votes16 + (tile_base + row0) * P.C1;
            for (int j = 0; j < nrows; ++j) {
                for (int c = lane; c < P.C1; c += 32) {
                    const uint16_t v = hist[(row0 + j) * RS + c];
                    if (!P.accumulate) out[j * P.C1 + c] = v;
                    else if (v) out[j * P.C1 + c] += v;
                }
            }
        }
        if (P.votes && nrows > 0) {
            int32_t* __restrict__ out = P.votes + (tile_base + row0) * P.C1;
            for (int j = 0; j < nrows; ++j) {
                for (int c = lane; c < P.C1; c += 32) {
                    const int v = (int)hist[(row0 + j) * RS + c];
                    if (!P.accumulate) out[j * P.C1 + c] = v;   // overwrite mode writes every cell exactly once
                    else if (v) out[j * P.C1 + c] += v;         // accumulate mode touches only the sparse non-zero cells
                }
            }
        }
        if (RP.enabled && active) {
            // VotingSegmentation.segment (voting.py:120-135).  The running (total, best, bpos) is exact unless another
            // lane's deferred pass added votes to this row: then it is re-derived from the row itself.
            if ((sh.dirty[warp] >> lane) & 1u) {
                const uint16_t* __restrict__ row = reinterpret_cast<const uint16_t*>(hist) + tid * RS;
                T.total = 0;
                T.best = 0;
                T.bpos = 0x7fff;
                for (int c = 0; c < P.C1; ++c) {
                    const int v = row[c];
                    T.total += v;
                    const int pos = RP.fpos[c];
                    if (v > 0 && pos >= 0 && (v > T.best || (v == T.best && pos < T.bpos))) {
                        T.best = v;
                        T.bpos = pos;
                    }
                }
            }
            bool unc = (T.total <= 0) || (T.best <= 0);                                  // voting.py:126,131
            if (!unc) unc = xdiv((double)T.best, (double)T.total) < RP.threshold;         // voting.py:128-130
            P.labels[gi] = (int64_t)(unc ? RP.unclassified : RP.remap[T.bpos]);
            if (P.summ) P.summ[gi] = summ_pack(T.total, T.best, T.bpos);
        }
    }

    }   // hist_live

    // ---- statistics: one atomic per counter per CTA
    if (P.stats) {
        const unsigned nc = __reduce_add_sync(0xffffffffu, T.n_cand), ns = __reduce_add_sync(0xffffffffu, T.n_seen);
        if (lane == 0) {
            if (nc) atomicAdd(sh.stat + F3D_STAT_CANDIDATES, nc);
            if (ns) atomicAdd(sh.stat + F3D_STAT_SEEN, ns);
        }
        __syncthreads();
        if (tid < 6 && sh.stat[tid]) atomicAdd(P.stats + tid, (unsigned long long)sh.stat[tid]);
    }
}

// ---- first cull level: candidate frames per super-tile (4096 consecutive points) ---------------------------------------
// One CTA per super-tile: exact box of its points, every frame of the launch tested with the same conservative rule as
// the warp-box test, surviving frame ids appended to the super-tile's list.  With F frames and T tiles this replaces T*F
// plane tests inside the fused kernel by T*F/16 here plus (list length) per tile -- what keeps the cull from dominating
// at thousands of frames (C3 / C4) and keeps the frame table out of the fused kernel's L1.
static __global__ void __launch_bounds__(256) supertile_cull_kernel(const __grid_constant__ FuseParams P, unsigned* __restrict__ st_count,
                                                                    uint16_t* __restrict__ st_list) {
    __shared__ float s_box[8 * 6];
    __shared__ unsigned s_n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p0 = (int64_t)blockIdx.x * FUSE_ST_POINTS;
    const float big = 3.0e38f;
    float lo[3] = {big, big, big}, hi[3] = {-big, -big, -big};
    for (int k = 0; k < FUSE_ST_POINTS / 256; ++k) {
        const int64_t i = p0 + (int64_t)k * 256 + tid;
        if (i < P.N) {
            const float4 q = __ldg(P.points + i);
            lo[0] = fminf(lo[0], q.x); hi[0] = fmaxf(hi[0], q.x);
            lo[1] = fminf(lo[1], q.y); hi[1] = fmaxf(hi[1], q.y);
            lo[2] = fminf(lo[2], q.z); hi[2] = fmaxf(hi[2], q.z);
        }
    }
#pragma unroll
    for (int s2 = 16; s2 > 0; s2 >>= 1)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], s2));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], s2));
        }
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s_box[warp * 6 + k] = lo[k];
            s_box[warp * 6 + 3 + k] = hi[k];
        }
    if (tid == 0) s_n = 0u;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = s_box[k];
        hi[k] = s_box[3 + k];
        for (int w = 1; w < 8; ++w) {
            lo[k] = fminf(lo[k], s_box[w * 6 + k]);
            hi[k] = fmaxf(hi[k], s_box[w * 6 + 3 + k]);
        }
    }
    const float box_mag = fabsf(lo[0]) + fabsf(lo[1]) + fabsf(lo[2]) + fabsf(hi[0]) + fabsf(hi[1]) + fabsf(hi[2]);
    const FrameRecord* __restrict__ frec = reinterpret_cast<const FrameRecord*>(P.table);
    const int nf = P.f_end - P.f_begin;
    uint16_t* __restrict__ list = st_list + (size_t)blockIdx.x * FUSE_ST_LCAP;
    for (int f0 = 0; f0 < nf; f0 += 256) {
        const int fi = f0 + tid;
        bool keep = fi < nf;
        if (keep) {
            const float4* pl = frec[P.f_begin + fi].cull.pl;
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                const float4 q = __ldg(pl + m);
                const float mx = fmaxf(q.x * lo[0], q.x * hi[0]) + fmaxf(q.y * lo[1], q.y * hi[1]) + fmaxf(q.z * lo[2], q.z * hi[2]) - q.w;
                const float margin = 2.0e-6f * (box_mag + fabsf(q.w)) + 1.0e-7f;
                keep = keep && (mx >= -margin);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        unsigned base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_n, (unsigned)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned at = base + __popc(bal & ((1u << lane) - 1u));
        if (keep && at < FUSE_ST_LCAP) list[at] = (uint16_t)fi;
    }
    __syncthreads();
    if (tid == 0) st_count[blockIdx.x] = s_n > FUSE_ST_LCAP ? 0xffffffffu : s_n;
}

// Large scenes (thousands of super-tiles x thousands of frames, C3): the kernel above re-reads the 80-byte plane block
// of every frame once per super-tile and is bound by that L2 traffic (24.4 K x 5000 x 80 B = 9.8 GB at C3).  Here one
// CTA owns FUSE_ST_GROUP consecutive super-tiles -- warp w builds the box of super-tile w -- and a thread tests its
// frame's planes, loaded once, against all of them.  Same rule, same list contents (order within a list is not part of
// the contract: votes commute).
#define FUSE_ST_GROUP 8
static __global__ void __launch_bounds__(256) supertile_cull_group_kernel(const __grid_constant__ FuseParams P, unsigned nst,
                                                                          unsigned* __restrict__ st_count, uint16_t* __restrict__ st_list) {
    __shared__ float s_box[FUSE_ST_GROUP][8];
    __shared__ unsigned s_n[FUSE_ST_GROUP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned st0 = blockIdx.x * FUSE_ST_GROUP;
    {
        const int64_t p0 = (int64_t)(st0 + warp) * FUSE_ST_POINTS;
        const float big = 3.0e38f;
        float lo[3] = {big, big, big}, hi[3] = {-big, -big, -big};
#pragma unroll 4
        for (int k = 0; k < FUSE_ST_POINTS / 32; ++k) {
            const int64_t i = p0 + (int64_t)k * 32 + lane;
            if (i < P.N) {
                const float4 q = __ldg(P.points + i);
                lo[0] = fminf(lo[0], q.x); hi[0] = fmaxf(hi[0], q.x);
                lo[1] = fminf(lo[1], q.y); hi[1] = fmaxf(hi[1], q.y);
                lo[2] = fminf(lo[2], q.z); hi[2] = fmaxf(hi[2], q.z);
            }
        }
#pragma unroll
        for (int s2 = 16; s2 > 0; s2 >>= 1)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], s2));
                hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], s2));
            }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                s_box[warp][k] = lo[k];
                s_box[warp][3 + k] = hi[k];
            }
            s_box[warp][6] = fabsf(lo[0]) + fabsf(lo[1]) + fabsf(lo[2]) + fabsf(hi[0]) + fabsf(hi[1]) + fabsf(hi[2]);
            s_n[warp] = 0u;
        }
    }
    __syncthreads();
    const FrameRecord* __restrict__ frec = reinterpret_cast<const FrameRecord*>(P.table);
    const int nf = P.f_end - P.f_begin;
    const int ns = min((unsigned)FUSE_ST_GROUP, nst - st0);
    for (int f0 = 0; f0 < nf; f0 += 256) {
        const int fi = f0 + tid;
        float4 pl[5];
        if (fi < nf) {
            const float4* src = frec[P.f_begin + fi].cull.pl;
#pragma unroll
            for (int m = 0; m < 5; ++m) pl[m] = __ldg(src + m);
        }
        for (int s = 0; s < ns; ++s) {
            bool keep = fi < nf;
            if (keep) {
                const float l0 = s_box[s][0], l1 = s_box[s][1], l2 = s_box[s][2], h0 = s_box[s][3], h1 = s_box[s][4], h2 = s_box[s][5];
                const float box_mag = s_box[s][6];
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    const float4 q = pl[m];
                    const float mx = fmaxf(q.x * l0, q.x * h0) + fmaxf(q.y * l1, q.y * h1) + fmaxf(q.z * l2, q.z * h2) - q.w;
                    const float margin = 2.0e-6f * (box_mag + fabsf(q.w)) + 1.0e-7f;
                    keep = keep && (mx >= -margin);
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (bal == 0u) continue;
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&s_n[s], (unsigned)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned at = base + __popc(bal & ((1u << lane) - 1u));
            if (keep && at < FUSE_ST_LCAP) st_list[(size_t)(st0 + s) * FUSE_ST_LCAP + at] = (uint16_t)fi;
        }
    }
    __syncthreads();
    if (tid < ns) st_count[st0 + tid] = s_n[tid] > FUSE_ST_LCAP ? 0xffffffffu : s_n[tid];
}

// ---- compacted launch (exchange mode) ------------------------------------------------------------------------------------
// A rank of a frame-sharded job sees a fraction of the building: at 8 ranks 7 of 8 super-tiles have an empty frame list,
// and a CTA that only finds that out holds one of the SM's five slots for a memory round trip (measured: the eight 1/8
// shards of C3 cost 40 ms of kernel time in total, the one 5000-frame launch 25 ms).  One block lists the super-tiles with
// a non-empty list in ascending order; the host reads the count (the one synchronisation of the step) and launches the
// fused kernel over those only.  The directory entries the skipped tiles would have cleared in the owners' memory are
// cleared by dead_directory_kernel with coalesced stores.
static __global__ void __launch_bounds__(1024) supertile_compact_kernel(const unsigned* __restrict__ st_count, unsigned nst,
                                                                        unsigned* __restrict__ live_list, unsigned* __restrict__ live_count) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned base_out = 0;
    for (unsigned base = 0; base < nst; base += 1024) {
        const unsigned s = base + tid;
        const bool live = s < nst && st_count[s] != 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, live);
        if (lane == 0) s_warp[warp] = (unsigned)__popc(bal);
        __syncthreads();
        if (warp == 0) {
            const unsigned v = s_warp[lane];
            unsigned inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            s_warp[lane] = inc - v;
            if (lane == 31) s_total = inc;
        }
        __syncthreads();
        if (live) live_list[base_out + s_warp[warp] + (unsigned)__popc(bal & ((1u << lane) - 1u))] = s;
        base_out += s_total;
        __syncthreads();
    }
    if (tid == 0) *live_count = base_out;
}

static __global__ void __launch_bounds__(256) dead_directory_kernel(const __grid_constant__ FuseParams P, unsigned nst) {
    constexpr unsigned PER_ST = (FUSE_ST_POINTS / 32) * F3D_XCH_NLEVEL;        // directory entries of one super-tile
    const unsigned long long total = (unsigned long long)nst * PER_ST;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned s = (unsigned)(e / PER_ST), r = (unsigned)(e % PER_ST);
        if (__ldg(P.st_count + s) != 0u) continue;
        const long long p0 = (long long)s * FUSE_ST_POINTS + (long long)(r / F3D_XCH_NLEVEL) * 32;
        if (p0 >= P.N) continue;
        const int d = (int)(p0 / P.xg_per);
        P.xg_dir[d][((p0 - (long long)d * P.xg_per) >> 5) * F3D_XCH_NLEVEL + (r % F3D_XCH_NLEVEL)] = make_uint2(0u, 0u);
    }
}

// ---- fix-up kernels: the deferred point-views, one per thread, in fp64 ------------------------------------------------
// one deferred point-view: fp64 evaluation against its frame's exact record `fe`, then the vote / depth sample / index goes
// where the mode wants it (dense votes by atomics, the owner's sub-queue `qsub` in exchange mode, z-buffer, uv2pt)
template <int MODE, int FMT>
__device__ __forceinline__ void fixup_entry(const FuseParams& P, const FuseResolve& RP, const FrameExact* fe, const GEntry& e,
                                            unsigned long long i, unsigned qsub, unsigned* s_qcnt, unsigned& n_exact,
                                            unsigned& n_div, unsigned& n_edge, unsigned& n_seen) {
    const int frel = (int)(e.w & 0xffffu), st = (int)((e.w >> 16) & 0xffu);
    const float4 p = __ldg(P.points + e.pt);
    const unsigned long long eo = exact_eval<MODE, FMT>(P, fe, frel, p.x, p.y, p.z);
    const int e_in = (int)(eo & EX_IN), e_vis = (int)((eo >> 1) & 1ull), e_pix = (int)(eo >> 32);
    const bool e_seen = (MODE == MODE_SPLAT) ? (e_in != 0) : (e_vis != 0);
    const uint32_t e_zq = (MODE == MODE_SPLAT && e_in) ? (uint32_t)((eo >> 16) & 0xffffull) : 0u;
    bool diverged;
    if (st == 2) diverged = ((int)e.guess != e_in) || (e_in && e.pix != e_pix);
    else if (MODE == MODE_SPLAT) diverged = (!e_in) || (e.pix != e_pix) || (e.guess != e_zq);
    else diverged = ((int)e.guess != e_vis) || (e_in && e.pix != e_pix);
    ++n_exact;
    n_div += diverged ? 1u : 0u;
    n_edge += (e_in && (eo & EX_EDGE)) ? 1u : 0u;
    if (!e_seen) return;
    ++n_seen;
    const size_t off = (size_t)frel * (size_t)(P.H * P.W) + (size_t)e_pix;
    if (MODE == MODE_VOTE) {
        const int cls = (int)((eo >> 8) & 0xffull);
        if (cls < P.C1 && P.xg_G > 0) {
            // exchange mode: the vote goes to the owner's queue through a sub-queue this block owns (shared-memory cursor)
            const int d = (int)(e.pt / P.xg_per);
            xg_append(P, d, qsub, atomicAdd(&s_qcnt[d], 1u), (unsigned long long)(e.pt - (long long)d * P.xg_per) * (unsigned long long)P.C1 + (unsigned)cls, 1u);
        } else if (cls < P.C1) {
            if (P.votes16) {
                const size_t cell = (size_t)e.pt * P.C1 + cls;   // 32-bit atomic on the word of the uint16 counter
                atomicAdd(reinterpret_cast<unsigned*>(P.votes16) + (cell >> 1), (cell & 1) ? 0x10000u : 1u);
            } else {
                const int v = atomicAdd(P.votes + (size_t)e.pt * P.C1 + cls, 1) + 1;
                if (P.summ) {
                    // advance the point's resolve state exactly as cast_vote would have: total + 1, and (best, first position)
                    // if this class now leads.  Order of the updates does not matter (the maximum is the maximum).
                    const int pos = RP.fpos[cls];
                    unsigned long long cur = P.summ[e.pt];
                    for (;;) {
                        const int total = (int)(cur & 0xffffffu), best = (int)((cur >> 24) & 0xffffffu), bpos = (int)(cur >> 48);
                        const bool lead = pos >= 0 && (v > best || (v == best && pos < bpos));
                        const unsigned long long nxt = summ_pack(total + 1, lead ? v : best, lead ? pos : bpos);
                        const unsigned long long seen = atomicCAS(P.summ + e.pt, cur, nxt);
                        if (seen == cur) break;
                        cur = seen;
                    }
                }
            }
            P.gq[i].w = e.w | (1u << 24);   // this point's label must be re-resolved
        }
    } else if (MODE == MODE_SPLAT) {
        atomicMin(P.zbuf + off, e_zq);
    } else {
        atomicMax(P.uv2pt + off, e.pt);
    }
}

__device__ __forceinline__ void fixup_stats(const FuseParams& P, unsigned n_exact, unsigned n_div, unsigned n_edge, unsigned n_seen) {
    if (!P.stats) return;
    unsigned vals[4] = {n_exact, n_div, n_edge, n_seen};
    const int idx[4] = {F3D_STAT_EXACT, F3D_STAT_DIVERGED, F3D_STAT_NEAR_EDGE, F3D_STAT_SEEN};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned v = __reduce_add_sync(0xffffffffu, vals[k]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(P.stats + idx[k], (unsigned long long)v);
    }
}

// When the launch has few enough frames, the used part of every frame's exact record (FIXUP_REC_BYTES of 448) fits one
// SM's shared memory: one 1024-thread block per SM stages the whole table on chip with TMA bulk copies (one
// cp.async.bulk per frame record, all completing on one mbarrier) and every thread evaluates its entries against it
// directly -- no per-warp staging round trips, 32 warps per SM in flight.
#define FIXUP_REC_BYTES 368   // q, qi, t, ss, plane_pt, plane_n, lookat = 360 bytes, padded to 16
#define FIXUP_TABLE_THREADS 1024
#define FIXUP_TABLE_QSUBS (F3D_XCH_NSUB_FIX / 148)   // sub-queues a table block owns (warps share them round-robin)
template <int MODE, int FMT>
__global__ void __launch_bounds__(FIXUP_TABLE_THREADS, 1) fixup_apply_table_kernel(const __grid_constant__ FuseParams P,
                                                                                    const __grid_constant__ FuseResolve RP) {
    extern __shared__ __align__(16) unsigned char fx_smem[];
    __shared__ unsigned s_qcnt[FIXUP_TABLE_QSUBS][F3D_MAX_RANKS];
    __shared__ __align__(8) uint64_t s_bar;
    const int nf = P.f_end - P.f_begin;
    const FrameRecord* __restrict__ frec = reinterpret_cast<const FrameRecord*>(P.table);
    for (int i = threadIdx.x; i < FIXUP_TABLE_QSUBS * F3D_MAX_RANKS; i += blockDim.x) (&s_qcnt[0][0])[i] = 0u;
    const unsigned long long n = min(*P.gq_count, P.gq_cap);
    if ((unsigned long long)blockIdx.x * blockDim.x >= n) {   // nothing for this block (grid-stride start past the end)
        if (P.xg_G > 0) {
            __syncthreads();
            for (int i = threadIdx.x; i < FIXUP_TABLE_QSUBS * P.xg_G; i += blockDim.x)
                P.xg_qcur[(i % P.xg_G) * F3D_XCH_NSUB + blockIdx.x * FIXUP_TABLE_QSUBS + i / P.xg_G] = 0u;
        }
        return;
    }
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(&s_bar, (unsigned)(nf * FIXUP_REC_BYTES));
    __syncthreads();
    for (int f = threadIdx.x; f < nf; f += blockDim.x)
        tma_bulk_g2s(fx_smem + (size_t)f * FIXUP_REC_BYTES, &frec[P.f_begin + f].exact, FIXUP_REC_BYTES, &s_bar);
    mbar_wait(&s_bar, 0u);
    const unsigned wsub = (threadIdx.x >> 5) % FIXUP_TABLE_QSUBS;
    unsigned n_exact = 0, n_div = 0, n_edge = 0, n_seen = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const GEntry e = P.gq[i];
        if (e.pt < 0) continue;
        const FrameExact* fe = reinterpret_cast<const FrameExact*>(fx_smem + (size_t)(e.w & 0xffffu) * FIXUP_REC_BYTES);
        fixup_entry<MODE, FMT>(P, RP, fe, e, i, blockIdx.x * FIXUP_TABLE_QSUBS + wsub, s_qcnt[wsub], n_exact, n_div, n_edge, n_seen);
    }
    if (P.xg_G > 0) {
        __syncthreads();
        for (int i = threadIdx.x; i < FIXUP_TABLE_QSUBS * P.xg_G; i += blockDim.x) {
            const int q = i / P.xg_G, d = i % P.xg_G;
            P.xg_qcur[d * F3D_XCH_NSUB + blockIdx.x * FIXUP_TABLE_QSUBS + q] = min(s_qcnt[q][d], P.xg_subcap);
        }
    }
    fixup_stats(P, n_exact, n_div, n_edge, n_seen);
}

#define FIXUP_THREADS 128
template <int MODE, int FMT>
__global__ void __launch_bounds__(FIXUP_THREADS) fixup_apply_kernel(const __grid_constant__ FuseParams P, const __grid_constant__ FuseResolve RP) {
    // Each lane needs the 448-byte fp64 record of ITS frame.  Reading it field by field would be ~50 fully divergent
    // loads per lane; instead the warp copies the 32 records one after the other with coalesced 16-byte loads into
    // shared memory and every lane then evaluates from its own copy.
    extern __shared__ __align__(16) unsigned char fx_smem[];
    __shared__ unsigned s_qcnt[F3D_MAX_RANKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FrameExact* wbuf = reinterpret_cast<FrameExact*>(fx_smem) + warp * 32;
    if (threadIdx.x < F3D_MAX_RANKS) s_qcnt[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned long long n = min(*P.gq_count, P.gq_cap);
    const FrameRecord* __restrict__ frec = reinterpret_cast<const FrameRecord*>(P.table);
    unsigned n_exact = 0, n_div = 0, n_edge = 0, n_seen = 0;
    const unsigned long long warps_total = (unsigned long long)gridDim.x * (FIXUP_THREADS / 32);
    for (unsigned long long base = ((unsigned long long)blockIdx.x * (FIXUP_THREADS / 32) + warp) * 32; base < n;
         base += warps_total * 32) {
        const unsigned long long i = base + lane;
        GEntry e;
        e.pt = -1;
        e.w = 0;
        e.pix = 0;
        e.guess = 0;
        if (i < n) e = P.gq[i];
        const int frel = (int)(e.w & 0xffffu);
        __syncwarp();
        // eight records per round: all eight 16-byte loads of a lane are in flight before the first store (one L2
        // round trip per round instead of one per record)
        for (int l0 = 0; l0 < 32; l0 += 8) {
            uint4 tmp[8];
            int fls[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                fls[k] = __shfl_sync(0xffffffffu, e.pt >= 0 ? frel : -1, l0 + k);
                if (fls[k] >= 0 && lane < (int)(sizeof(FrameExact) / 16))
                    tmp[k] = __ldg(reinterpret_cast<const uint4*>(&frec[P.f_begin + fls[k]].exact) + lane);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (fls[k] >= 0 && lane < (int)(sizeof(FrameExact) / 16)) reinterpret_cast<uint4*>(wbuf + l0 + k)[lane] = tmp[k];
        }
        __syncwarp();
        if (e.pt < 0) continue;
        fixup_entry<MODE, FMT>(P, RP, wbuf + lane, e, i, blockIdx.x, s_qcnt, n_exact, n_div, n_edge, n_seen);
    }
    if (P.xg_G > 0) {
        __syncthreads();
        if ((int)threadIdx.x < P.xg_G) P.xg_qcur[threadIdx.x * F3D_XCH_NSUB + blockIdx.x] = min(s_qcnt[threadIdx.x], P.xg_subcap);
    }
    fixup_stats(P, n_exact, n_div, n_edge, n_seen);
}

// labels of the points whose votes changed, from their 8-byte resolve state (VotingSegmentation.segment, voting.py:120-135)
static __global__ void __launch_bounds__(256) fixup_labels_summary_kernel(const __grid_constant__ FuseParams P, const __grid_constant__ FuseResolve RP) {
    const unsigned long long n = min(*P.gq_count, P.gq_cap);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const GEntry e = P.gq[i];
        if (e.pt < 0 || !((e.w >> 24) & 1u)) continue;
        const unsigned long long cur = P.summ[e.pt];
        const int total = (int)(cur & 0xffffffu), best = (int)((cur >> 24) & 0xffffffu), bpos = (int)(cur >> 48);
        bool unc = (total <= 0) || (best <= 0);
        if (!unc) unc = xdiv((double)best, (double)total) < RP.threshold;
        P.labels[e.pt] = (int64_t)(unc ? RP.unclassified : RP.remap[bpos]);
    }
}

// labels of the points whose votes changed in fixup_apply_kernel (VotingSegmentation.segment, voting.py:120-135);
// eight lanes per entry stream the point's vote row, like resolve_kernel
static __global__ void __launch_bounds__(256) fixup_labels_kernel(const __grid_constant__ FuseParams P, const __grid_constant__ FuseResolve RP) {
    __shared__ int16_t s_fpos[RES_MAXC];
    for (int c = threadIdx.x; c < RES_MAXC; c += blockDim.x) s_fpos[c] = RP.fpos[c];
    __syncthreads();
    const unsigned long long n = min(*P.gq_count, P.gq_cap);
    const int sub = threadIdx.x & 7;
    const unsigned long long per_pass = ((unsigned long long)gridDim.x * blockDim.x) >> 3;
    const unsigned long long passes = (n + per_pass - 1) / per_pass;   // uniform trip count: shuffles see full warps
    unsigned long long i = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    for (unsigned long long p = 0; p < passes; ++p, i += per_pass) {
        GEntry e;
        e.pt = -1;
        e.w = 0;
        if (i < n) e = P.gq[i];
        const bool live = (e.pt >= 0) && ((e.w >> 24) & 1u);
        long long total = 0;
        int best = 0, bpos = 0x7fff;
        if (live && !P.votes16) {
            row_partial8(P.votes + (size_t)e.pt * P.C1, P.C1, s_fpos, sub, total, best, bpos);
        } else if (live) {
            for (int c = sub; c < P.C1; c += 8) {
                const int v = (int)P.votes16[(size_t)e.pt * P.C1 + c];
                total += v;
                const int pos = s_fpos[c];
                if (v > 0 && pos >= 0 && (v > best || (v == best && pos < bpos))) {
                    best = v;
                    bpos = pos;
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
            bool unc = (total <= 0) || (best <= 0);
            if (!unc) unc = xdiv((double)best, (double)total) < RP.threshold;
            P.labels[e.pt] = (int64_t)(unc ? RP.unclassified : RP.remap[bpos]);
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------
static inline int hist_row_stride(int C1) {
    int rs = (C1 + 1) & ~1;            // even number of uint16
    if (((rs / 2) & 1) == 0) rs += 2;  // odd number of 32-bit words per row: lanes hit distinct banks
    return rs;
}

static inline size_t fuse_smem_bytes(int mode, int C1, int hb, bool slots) {
    size_t b = FUSE_OFF_HIST;
    if (mode == MODE_VOTE) b += fuse_align16((size_t)FUSE_BLOCK * (hb == 1 ? (size_t)C1 : hist_row_stride(C1) * sizeof(uint16_t)));
    if (slots) b += (size_t)FUSE_NSLOT * FUSE_BLOCK + (size_t)FUSE_NW * FUSE_STG_ROWS * 32 * sizeof(uint16_t);
    return b;
}

// Timing of the fused kernel alone: the caller hands two of ITS events to f3d_fuse_time_next_call; the next fused launch
// of this host thread records them around fuse_kernel and forgets them (thread-local one-shot hand-over, no other state).
struct FuseTimingSlot {
    cudaEvent_t ev[2];
    bool armed;
};
FuseTimingSlot& f3d_timing_slot();   // fuse_vote.cu

template <int MODE, int FMT, int HB, bool AUDIT>
static int launch_fuse_hb(FuseParams P, const FuseResolve& RP, cudaStream_t stream) {
    if (MODE == MODE_VOTE) P.RS = HB == 1 ? P.C1 : hist_row_stride(P.C1);
    const size_t smem = fuse_smem_bytes(MODE, P.C1, HB, MODE == MODE_VOTE && HB == 1 && P.xg_G > 0);
    if (smem > 227 * 1024) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_fuse: nclasses+1 too large for the shared-memory histogram");
    static thread_local size_t smem_set[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // per device: the attribute is sticky, set it when it grows
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 8 || smem > smem_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fuse_kernel<MODE, FMT, HB, AUDIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return f3d_check_launch("f3d_fuse(cudaFuncSetAttribute)");
        if (dev >= 0 && dev < 8) smem_set[dev] = smem;
    }
    int64_t tiles = (P.N + FUSE_BLOCK - 1) / FUSE_BLOCK;
    if (tiles > 0x7fffffff) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_fuse: too many points for one launch");
    const bool use_queue = P.gq != nullptr;
    if (use_queue) {
        cudaError_t e = cudaMemsetAsync(P.gq_count, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return f3d_check_launch("f3d_fuse(memset)");
    }
    if (P.st_count) {
        if (P.f_end - P.f_begin > 32) {
            const unsigned nst = (unsigned)((P.N + FUSE_ST_POINTS - 1) / FUSE_ST_POINTS);
            // grouped variant once there are enough super-tiles to fill the GPU with groups and enough frames for the
            // plane re-reads to matter
            if (nst >= 148u * 2u * FUSE_ST_GROUP && P.f_end - P.f_begin >= 256)
                supertile_cull_group_kernel<<<(nst + FUSE_ST_GROUP - 1) / FUSE_ST_GROUP, 256, 0, stream>>>(
                    P, nst, const_cast<unsigned*>(P.st_count), const_cast<uint16_t*>(P.st_list));
            else
                supertile_cull_kernel<<<nst, 256, 0, stream>>>(P, const_cast<unsigned*>(P.st_count), const_cast<uint16_t*>(P.st_list));
        } else {
            P.st_count = nullptr;   // few frames: the per-tile scan is cheaper than another launch
        }
    }
    P.live_list = nullptr;
    if (MODE == MODE_VOTE && P.compact && P.st_count && P.xg_G > 0 && P.live_scratch) {
        const unsigned nst = (unsigned)((P.N + FUSE_ST_POINTS - 1) / FUSE_ST_POINTS);
        supertile_compact_kernel<<<1, 1024, 0, stream>>>(P.st_count, nst, P.live_scratch, P.live_count);
        dead_directory_kernel<<<148 * 8, 256, 0, stream>>>(P, nst);
        unsigned h_live = 0;
        cudaError_t e = cudaMemcpyAsync(&h_live, P.live_count, sizeof(unsigned), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return f3d_check_launch("f3d_fuse(compacted launch)");
        if (h_live > nst) return f3d_fail(F3D_ERR_CUDA, "f3d_fuse: corrupt live super-tile count");
        P.live_list = P.live_scratch;
        tiles = (int64_t)h_live * FUSE_ST_TILES;
    }
    FuseTimingSlot& ts = f3d_timing_slot();
    const bool timed = ts.armed;
    ts.armed = false;
    if (timed) cudaEventRecord(ts.ev[0], stream);
    if (tiles > 0) fuse_kernel<MODE, FMT, HB, AUDIT><<<(unsigned)tiles, FUSE_BLOCK, smem, stream>>>(P, RP);
    if (timed) cudaEventRecord(ts.ev[1], stream);
    if (use_queue) {
        // the queue length lives on the device: fixed grids with grid-stride loops, no host synchronisation
        static_assert(F3D_XCH_NSUB_FIX == 148 * 12, "one sub-queue per fix-up block");
        const size_t table_smem = (size_t)(P.f_end - P.f_begin) * FIXUP_REC_BYTES;
        if (table_smem <= 200 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(fixup_apply_table_kernel<MODE, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return f3d_check_launch("f3d_fuse(cudaFuncSetAttribute fixup table)");
            fixup_apply_table_kernel<MODE, FMT><<<148, FIXUP_TABLE_THREADS, table_smem, stream>>>(P, RP);
        } else {
            const int fx_smem = (FIXUP_THREADS / 32) * 32 * (int)sizeof(FrameExact);
            cudaError_t e = cudaFuncSetAttribute(fixup_apply_kernel<MODE, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, fx_smem);
            if (e != cudaSuccess) return f3d_check_launch("f3d_fuse(cudaFuncSetAttribute fixup)");
            fixup_apply_kernel<MODE, FMT><<<F3D_XCH_NSUB_FIX, FIXUP_THREADS, fx_smem, stream>>>(P, RP);
        }
        if (MODE == MODE_VOTE && RP.enabled) {
            if (P.summ) fixup_labels_summary_kernel<<<148 * 8, 256, 0, stream>>>(P, RP);
            else fixup_labels_kernel<<<148 * 8, 256, 0, stream>>>(P, RP);
        }
    }
    return f3d_check_launch("f3d_fuse");
}

// the vote kernel runs on the byte histogram unless the caller wants labels without any vote output from more frames
// than a byte counter can hold between flushes (there is then nowhere to flush to): that case keeps uint16 counters
template <int MODE, int FMT>
static int launch_fuse(const FuseParams& P, const FuseResolve& RP, int audit, cudaStream_t stream) {
    if constexpr (MODE == MODE_VOTE) {
        const bool no_sink = !P.votes && !P.votes16 && P.xg_G == 0;
        if (!(no_sink && P.f_end - P.f_begin > FUSE_LIMIT8))
            return audit ? launch_fuse_hb<MODE, FMT, 1, true>(P, RP, stream) : launch_fuse_hb<MODE, FMT, 1, false>(P, RP, stream);
    }
    return audit ? launch_fuse_hb<MODE, FMT, 2, true>(P, RP, stream) : launch_fuse_hb<MODE, FMT, 2, false>(P, RP, stream);
}

// workspace = [deferred count u64][pad u64][super-tile counts u32 x S, padded to 16 B][super-tile lists u16 x S x FUSE_ST_LCAP]
//             [resolve states u64 x N (fused labels only)][GEntry x cap]; S = super-tiles of the cloud.  The super-tile part is
// attached whenever it fits, the deferred queue only when the caller's mode wants it; returns whether the queue was attached.
static inline int64_t supertile_bytes(int64_t npoints) {
    const int64_t S = (npoints + FUSE_ST_POINTS - 1) / FUSE_ST_POINTS;
    return 2 * ((S * 4 + 15) & ~(int64_t)15) + S * FUSE_ST_LCAP * 2;      // counts, live list, frame lists
}

static inline bool attach_workspace(FuseParams& P, void* workspace, int64_t workspace_bytes, bool want_queue = true, bool want_summ = false) {
    P.summ = nullptr;
    P.gq = nullptr;
    P.gq_count = nullptr;
    P.gq_cap = 0;
    P.st_count = nullptr;
    P.st_list = nullptr;
    P.live_list = nullptr;
    P.live_scratch = nullptr;
    P.live_count = nullptr;
    if (!workspace || workspace_bytes < 16 || (reinterpret_cast<uintptr_t>(workspace) & 15u)) return false;
    char* base = reinterpret_cast<char*>(workspace);
    int64_t off = 16;
    const int64_t stb = supertile_bytes(P.N);
    if (workspace_bytes >= off + stb) {
        const int64_t S = (P.N + FUSE_ST_POINTS - 1) / FUSE_ST_POINTS;
        const int64_t cb = (S * 4 + 15) & ~(int64_t)15;
        P.st_count = reinterpret_cast<const unsigned*>(base + off);
        P.live_scratch = reinterpret_cast<unsigned*>(base + off + cb);
        P.live_count = reinterpret_cast<unsigned*>(base + 8);              // second half of the 16-byte header
        P.st_list = reinterpret_cast<const uint16_t*>(base + off + 2 * cb);
        off += stb;
    }
    if (!want_queue) return false;
    // resolve states only when the queue still gets its share (one entry per 8 points) behind them
    if (want_summ && workspace_bytes - off >= P.N * 8 + (P.N / 8 + 1024) * (int64_t)sizeof(GEntry)) {
        P.summ = reinterpret_cast<unsigned long long*>(base + off);
        off += P.N * 8;
    }
    if (workspace_bytes - off < (int64_t)sizeof(GEntry)) {
        P.summ = nullptr;
        return false;
    }
    P.gq_count = reinterpret_cast<unsigned long long*>(workspace);
    P.gq = reinterpret_cast<GEntry*>(base + off);
    P.gq_cap = (unsigned long long)((workspace_bytes - off) / (int64_t)sizeof(GEntry));
    return true;
}

static inline bool fmt_is_packed(int fmt) { return fmt == F3D_FRAMES_U32 || fmt == F3D_FRAMES_U32_T16; }

static inline int fill_common(FuseParams& P, const void* points, int64_t N, const void* table, int fb, int fe, const void* depth,
                              int fmt, int H, int W, const double* h_K9, double radius, double zmin, double zmax,
                              uint64_t* stats) {
    if (!points || !table || !h_K9 || N < 0 || fb < 0 || fe < fb || H <= 0 || W <= 0)
        return f3d_fail(F3D_ERR_ARG, "f3d_fuse: bad argument");
    if (fmt != F3D_DEPTH_U16_MM && fmt != F3D_DEPTH_F32_M && !fmt_is_packed(fmt)) return f3d_fail(F3D_ERR_ARG, "f3d_fuse: unknown frame format");
    if ((int64_t)H * W > 0x7fffffff || H > 65535 || W > 65535) return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_fuse: image too large");
    P.points = reinterpret_cast<const float4*>(points);
    P.live_list = nullptr;
    P.live_scratch = nullptr;
    P.live_count = nullptr;
    P.compact = 0;
    P.N = N;
    P.table = table;
    P.depth = depth;
    P.H = H;
    P.W = W;
    P.tiles_x = (W + 15) / 16;
    P.frame_stride = fmt == F3D_FRAMES_U32_T16 ? (int64_t)P.tiles_x * ((H + 15) / 16) * 256 : (int64_t)H * W;
    for (int i = 0; i < 9; ++i) P.K[i] = h_K9[i];
    P.cx = (float)h_K9[2];
    P.cy = (float)h_K9[5];
    P.inv_fx = (float)(1.0 / h_K9[0]);
    P.inv_fy = (float)(1.0 / h_K9[4]);
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
    P.votes = nullptr;
    P.votes16 = nullptr;
    P.uv2pt = nullptr;
    P.zbuf = nullptr;
    P.mask = nullptr;
    P.labels = nullptr;
    P.C1 = 0;
    P.RS = 0;
    P.accumulate = 0;
    P.summ = nullptr;
    P.gq = nullptr;
    P.gq_count = nullptr;
    P.gq_cap = 0;
    P.st_count = nullptr;
    P.st_list = nullptr;
    P.xg_G = 0;
    P.xg_per = 1;
    P.xg_rowcur = nullptr;
    P.xg_qcur = nullptr;
    P.xg_subrows = 0;
    P.xg_subcap = 0;
    P.xg_overflow = nullptr;
    for (int i = 0; i < F3D_MAX_RANKS; ++i) {
        P.xg_queue[i] = nullptr;
        P.xg_slots[i] = nullptr;
        P.xg_dir[i] = nullptr;
    }
    return F3D_OK;
}

// frames are processed in launches of at most 65535 - FUSE_QWARP (uint16 candidate ids; a uint16 histogram counter
// must also hold the warp's deferred votes)
#define F3D_MAX_FRAMES_PER_LAUNCH (65535 - FUSE_QWARP)

// byte size of one element of the frame stack `depth` points to
static inline size_t frame_elem_bytes(int fmt) { return fmt == F3D_DEPTH_U16_MM ? 2 : 4; }
