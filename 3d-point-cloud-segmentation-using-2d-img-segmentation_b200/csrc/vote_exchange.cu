// Owner side of the multi-GPU vote exchange (frames sharded over ranks, votes are integer sums over
// frames -- VotingSegmentation.vote, segUtils/voting.py:89-98 -- so any partition of the frames is bit-exact).
//
// Every rank owns a contiguous range of points.  The fused kernel of each source rank writes, straight into the owner's
// memory over NVLink (fuse_project_vote.cu, flush8):
//   * slot records: per (source, 32-point block) L rows of 64 B, row j = the j-th class (order of first appearance) of
//     each of the block's 32 points as uint16 (class | count << 8, 0 = none), L = the longest list in the block; a
//     directory entry (row offset, L) per block and flush level (F3D_XCH_NLEVEL of them: a tile with more than 235
//     candidate frames flushes its byte histogram several times) says where the rows are inside that source's record region (the region is
//     split into F3D_XCH_NREG sub-regions so that the senders' row cursors never become an atomic hot spot);
//   * (cell, count) entries in F3D_XCH_NSUB sub-queues for what does not go into a record (a full record sub-region,
//     later flushes of very dense scans, the deferred fp64 votes of the fix-up pass -- fix-up block b owns sub-queue b).
// Here the owner merges the G records of each of its points into the dense int32 row the reference keeps
// (votes[npts, nclasses + 1], voting.py:34), resolves the label (VotingSegmentation.segment, voting.py:106-137) from
// the on-chip row, then scatter-adds the queue entries and re-resolves the few points they touched.  The merge is a
// streaming pass: 2 B per non-zero (point, class, source) cell (padded to the block's longest list) read and
// 4 * C1 + 8 B written per point -- no dense partial vote tensor ever crosses the fabric or HBM.
#include "f3d_common.cuh"
#include "f3d_host.h"

#define XCH_BLOCK 256
#define XCH_Q 4   // records of a block whose first rows are prefetched together by the merge

struct PeerPtrs {
    unsigned* p[F3D_MAX_RANKS];
};

// all-gather of the labels fused into the kernels that produce them: every rank keeps an int16 array of ALL points in
// symmetric memory and the owner stores each label it resolves into the G copies (2 B per point and peer, coalesced
// 64-byte segments per warp) -- the NVLink traffic rides under the HBM-bound merge instead of following it as a collective
struct PeerLabels {
    int16_t* p[F3D_MAX_RANKS];
    int G;              // 0: no broadcast
    long long first;    // global index of this rank's first point
};

__device__ __forceinline__ void bcast_label(const PeerLabels& PL, long long row, int64_t label) {
    for (int d = 0; d < PL.G; ++d) PL.p[d][PL.first + row] = (int16_t)label;
}

// this rank's queue cursors [G][F3D_XCH_NSUB] -> row [rank] of every owner's count table [G][F3D_XCH_NSUB]
__global__ void exchange_publish_kernel(const unsigned* __restrict__ qcur, PeerPtrs counts, int rank, int G, unsigned subcap) {
    const int d = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < G && i < F3D_XCH_NSUB) counts.p[d][(size_t)rank * F3D_XCH_NSUB + i] = min(qcur[(size_t)d * F3D_XCH_NSUB + i], subcap);
}

// scatter-add every received (cell, count) entry into the dense int32 shard: block (x, y) walks the sub-queues
// x, x + gridDim.x, ... of source y
__global__ void __launch_bounds__(256) queue_accumulate_kernel(const unsigned long long* __restrict__ rx,
                                                               const unsigned* __restrict__ counts, unsigned subcap,
                                                               int32_t* __restrict__ votes, unsigned long long ncells) {
    const int src = blockIdx.y;
    for (int sub = blockIdx.x; sub < F3D_XCH_NSUB; sub += gridDim.x) {
        const unsigned n = min(counts[(size_t)src * F3D_XCH_NSUB + sub], subcap);
        const unsigned long long* __restrict__ seg = rx + ((size_t)src * F3D_XCH_NSUB + sub) * subcap;
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long e = seg[i];                      // cell (40 bits) | count << 40
            const unsigned long long key = e & 0xffffffffffull;
            if (key < ncells) atomicAdd(votes + key, (int)(e >> 40));
        }
    }
}

// labels of the points the queue entries touched (eight lanes stream the dense row with batched loads); a point with
// several entries is re-resolved several times with the same result
__global__ void __launch_bounds__(256) queue_relabel_kernel(const unsigned long long* __restrict__ rx,
                                                            const unsigned* __restrict__ counts, unsigned subcap,
                                                            const int32_t* __restrict__ votes, long long nrows, int C1,
                                                            const __grid_constant__ FuseResolve RP, int64_t* __restrict__ labels,
                                                            const __grid_constant__ PeerLabels PL) {
    __shared__ int16_t s_fpos[RES_MAXC];
    for (int c = threadIdx.x; c < RES_MAXC; c += blockDim.x) s_fpos[c] = RP.fpos[c];
    __syncthreads();
    const int src = blockIdx.y;
    const int sub8 = threadIdx.x & 7;
    for (int sub = blockIdx.x; sub < F3D_XCH_NSUB; sub += gridDim.x) {
        const unsigned n = min(counts[(size_t)src * F3D_XCH_NSUB + sub], subcap);
        const unsigned long long* __restrict__ seg = rx + ((size_t)src * F3D_XCH_NSUB + sub) * subcap;
        for (unsigned i0 = 0; i0 < n; i0 += blockDim.x >> 3) {             // block-uniform trip count
            const unsigned i = i0 + (threadIdx.x >> 3);
            long long pt = -1;
            if (i < n) pt = (long long)((seg[i] & 0xffffffffffull) / (unsigned long long)C1);
            const bool live = pt >= 0 && pt < nrows;
            long long total = 0;
            int best = 0, bpos = 0x7fff;
            if (live) row_partial8(votes + (size_t)pt * C1, C1, s_fpos, sub8, total, best, bpos);
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
            if (live && sub8 == 0) {
                bool unc = (total <= 0) || (best <= 0);                                    // voting.py:126,131
                if (!unc) unc = xdiv((double)best, (double)total) < RP.threshold;          // voting.py:128-130
                const int64_t lab = (int64_t)(unc ? RP.unclassified : RP.remap[bpos]);
                labels[pt] = lab;
                bcast_label(PL, pt, lab);
            }
        }
    }
}

// ---- slot-record merge: one thread per owned point, one warp per 32-point block ------------------------------------------
// shared: uint16 histogram [256][RS] (a thread owns its row; RS/2 odd => conflict-free), written out like the fused
// kernel's epilogue (warp-private rows, 16-byte stores, every cell exactly once -- no memset of the shard needed).
__global__ void __launch_bounds__(XCH_BLOCK, 3) slot_merge_kernel(const uint16_t* __restrict__ slots, const uint2* __restrict__ dir,
                                                                  int G, long long rows_cap, long long blocks_per_src,
                                                                  long long nrows, int C1, int RS,
                                                                  const __grid_constant__ FuseResolve RP,
                                                                  int32_t* __restrict__ votes, int64_t* __restrict__ labels,
                                                                  const __grid_constant__ PeerLabels PL) {
    extern __shared__ __align__(16) unsigned char xs[];
    int16_t* s_fpos = reinterpret_cast<int16_t*>(xs);
    uint16_t* hist = reinterpret_cast<uint16_t*>(xs + RES_MAXC * sizeof(int16_t));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < RES_MAXC; c += XCH_BLOCK) s_fpos[c] = RP.fpos[c];
    {
        uint4* h128 = reinterpret_cast<uint4*>(hist);
        const int n128 = (XCH_BLOCK * RS * 2 + 15) / 16;
        for (int i = tid; i < n128; i += XCH_BLOCK) h128[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const long long blk = (long long)blockIdx.x * (XCH_BLOCK / 32) + warp;   // 32-point block of this warp
    const long long p = blk * 32 + lane;
    uint16_t* __restrict__ row = hist + tid * RS;
    int total = 0, best = 0, bpos = 0x7fff;
    if (blk * 32 < nrows) {
        // the G x F3D_XCH_NLEVEL directory entries of this block, one per lane (two rounds cover 16 ranks)
        uint2 de0 = make_uint2(0u, 0u), de1 = make_uint2(0u, 0u);
        {
            const int s0 = lane / F3D_XCH_NLEVEL, k0 = lane % F3D_XCH_NLEVEL;
            if (s0 < G) de0 = __ldg(dir + ((size_t)s0 * blocks_per_src + blk) * F3D_XCH_NLEVEL + k0);
            if (s0 + 8 < G) de1 = __ldg(dir + ((size_t)(s0 + 8) * blocks_per_src + blk) * F3D_XCH_NLEVEL + k0);
        }
        auto take = [&](unsigned pair) {
            const int cls = (int)(pair & 0xffu), cnt = (int)(pair >> 8);
            if (cnt == 0 || cls >= C1) return;
            const int v = (int)row[cls] + cnt;
            row[cls] = (uint16_t)v;
            total += cnt;
            const int pos = s_fpos[cls];
            if (pos >= 0 && (v > best || (v == best && pos < bpos))) {
                best = v;
                bpos = pos;
            }
        };
        // non-empty records, four at a time: the first four rows of all four are requested before any is used (two HBM
        // round trips for a typical block of an 8-rank job instead of eight); longer records continue row by row
        for (int round = 0; round < 2; ++round) {
            const uint2 de = round ? de1 : de0;
            const int base_s = round * 8;
            unsigned live = __ballot_sync(0xffffffffu, de.y != 0u && base_s + lane / F3D_XCH_NLEVEL < G);
            while (live) {
                unsigned pr[XCH_Q][4];
                int Ls[XCH_Q];
                const uint16_t* recs[XCH_Q];
#pragma unroll
                for (int q = 0; q < XCH_Q; ++q) {
                    const int e = live ? (__ffs(live) - 1) : -1;
                    live &= live - 1u;
                    const unsigned off = __shfl_sync(0xffffffffu, de.x, max(e, 0));
                    const unsigned len = __shfl_sync(0xffffffffu, de.y, max(e, 0));
                    Ls[q] = e < 0 ? 0 : (int)min((unsigned long long)len, (unsigned long long)max(0LL, rows_cap - (long long)off));
                    recs[q] = slots + ((size_t)(base_s + max(e, 0) / F3D_XCH_NLEVEL) * rows_cap + off) * 32 + lane;
#pragma unroll
                    for (int k = 0; k < 4; ++k) pr[q][k] = (k < Ls[q]) ? (unsigned)__ldg(recs[q] + (size_t)k * 32) : 0u;   // 64 contiguous bytes per row
                }
#pragma unroll
                for (int q = 0; q < XCH_Q; ++q) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) take(pr[q][k]);
                    for (int j0 = 4; j0 < Ls[q]; j0 += 4) {
                        unsigned more[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) more[k] = (j0 + k < Ls[q]) ? (unsigned)__ldg(recs[q] + (size_t)(j0 + k) * 32) : 0u;
#pragma unroll
                        for (int k = 0; k < 4; ++k) take(more[k]);
                    }
                }
            }
        }
    }
    __syncwarp();
    const long long row0 = blk * 32;
    const int nr = (int)max(0LL, min(32LL, nrows - row0));
    if (votes && nr > 0) {
        int32_t* __restrict__ out = votes + row0 * C1;
        if (RS == C1) {
            const uint2* __restrict__ h64 = reinterpret_cast<const uint2*>(hist + warp * 32 * RS);
            const int tot = nr * C1, n4 = tot >> 2;
            for (int i = lane; i < n4; i += 32) {
                const uint2 w = h64[i];
                *reinterpret_cast<int4*>(out + 4 * i) =
                    make_int4((int)(w.x & 0xffffu), (int)(w.x >> 16), (int)(w.y & 0xffffu), (int)(w.y >> 16));
            }
            for (int e = (n4 << 2) + lane; e < tot; e += 32) out[e] = (int)hist[warp * 32 * RS + e];
        } else {
            for (int j = 0; j < nr; ++j)
                for (int c = lane; c < C1; c += 32) out[j * C1 + c] = (int)hist[(warp * 32 + j) * RS + c];
        }
    }
    if (labels && p < nrows) {
        bool unc = (total <= 0) || (best <= 0);                                    // voting.py:126,131
        if (!unc) unc = xdiv((double)best, (double)total) < RP.threshold;          // voting.py:128-130
        const int64_t lab = (int64_t)(unc ? RP.unclassified : RP.remap[bpos]);
        labels[p] = lab;
        bcast_label(PL, p, lab);
    }
}

static int fill_peer_labels(PeerLabels& PL, const uint64_t* h_peer_labels16, int32_t nranks, int64_t first_point, const char* who) {
    PL.G = 0;
    PL.first = 0;
    for (int i = 0; i < F3D_MAX_RANKS; ++i) PL.p[i] = nullptr;
    if (!h_peer_labels16) return F3D_OK;
    if (first_point < 0) return f3d_fail(F3D_ERR_ARG, who);
    for (int i = 0; i < nranks; ++i) {
        if (!h_peer_labels16[i]) return f3d_fail(F3D_ERR_ARG, who);
        PL.p[i] = reinterpret_cast<int16_t*>(h_peer_labels16[i]);
    }
    PL.G = nranks;
    PL.first = first_point;
    return F3D_OK;
}

static int xch_row_stride(int C1) {
    int rs = (C1 + 1) & ~1;
    if (((rs / 2) & 1) == 0) rs += 2;
    return rs;
}

extern "C" int f3d_exchange_constants(int32_t* out4) {
    int32_t* out3 = out4;
    if (!out3) return f3d_fail(F3D_ERR_ARG, "f3d_exchange_constants: NULL");
    out3[0] = F3D_XCH_NREG;
    out3[1] = F3D_XCH_NSUB;
    out3[2] = F3D_XCH_NSUB_FIX;
    out3[3] = F3D_XCH_NLEVEL;
    return F3D_OK;
}

extern "C" int f3d_exchange_publish(const uint32_t* cursors, const uint64_t* h_peer_counts, int32_t rank, int32_t nranks,
                                    int64_t sub_cap, void* stream) {
    if (!cursors || !h_peer_counts || nranks < 1 || nranks > F3D_MAX_RANKS || rank < 0 || rank >= nranks || sub_cap <= 0)
        return f3d_fail(F3D_ERR_ARG, "f3d_exchange_publish: bad argument");
    PeerPtrs pp;
    for (int i = 0; i < F3D_MAX_RANKS; ++i) pp.p[i] = i < nranks ? reinterpret_cast<unsigned*>(h_peer_counts[i]) : nullptr;
    dim3 grid(F3D_XCH_NSUB / 256, (unsigned)nranks);
    exchange_publish_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cursors + (size_t)nranks * F3D_XCH_NREG, pp, rank, nranks,
                                                                    (unsigned)sub_cap);
    return f3d_check_launch("f3d_exchange_publish");
}

extern "C" int f3d_exchange_queue_apply(const uint64_t* queue, const uint32_t* counts, int32_t nranks, int64_t sub_cap,
                                        int32_t* votes, int64_t nrows, int32_t C1, double threshold, const int32_t* h_filter,
                                        int32_t nfilter, int32_t nclasses_id, int64_t* labels, const uint64_t* h_peer_labels16,
                                        int64_t first_point, void* stream) {
    if (!queue || !counts || !votes || nranks < 1 || nranks > F3D_MAX_RANKS || sub_cap <= 0 || nrows < 0 || C1 <= 0 || C1 > 256 ||
        nfilter < 0 || (nfilter > 0 && !h_filter))
        return f3d_fail(F3D_ERR_ARG, "f3d_exchange_queue_apply: bad argument");
    if (nrows == 0) return F3D_OK;
    dim3 grid(148 * 4, (unsigned)nranks);
    queue_accumulate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(queue), counts,
                                                                   (unsigned)sub_cap, votes,
                                                                   (unsigned long long)nrows * (unsigned long long)C1);
    if (labels) {
        FuseResolve RP;
        int rc = f3d_build_resolve(C1, threshold, h_filter, nfilter, nclasses_id, RP);
        if (rc) return rc;
        PeerLabels PL;
        rc = fill_peer_labels(PL, h_peer_labels16, nranks, first_point, "f3d_exchange_queue_apply: bad peer label table");
        if (rc) return rc;
        queue_relabel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(queue), counts,
                                                                    (unsigned)sub_cap, votes, (long long)nrows, C1, RP, labels, PL);
    }
    return f3d_check_launch("f3d_exchange_queue_apply");
}

extern "C" int f3d_exchange_merge(const uint16_t* slots, const void* dir, int32_t nranks, int64_t sub_rows,
                                  int64_t points_per_shard, int64_t nrows, int32_t C1, double threshold, const int32_t* h_filter,
                                  int32_t nfilter, int32_t nclasses_id, int32_t* votes, int64_t* labels,
                                  const uint64_t* h_peer_labels16, int64_t first_point, void* stream) {
    if (!slots || !dir || nranks < 1 || nranks > F3D_MAX_RANKS || sub_rows <= 0 || points_per_shard <= 0 ||
        (points_per_shard % XCH_BLOCK) != 0 || nrows < 0 || nrows > points_per_shard || C1 <= 0 || C1 > 256 || (!votes && !labels) ||
        nfilter < 0 || (nfilter > 0 && !h_filter) || (votes && (reinterpret_cast<uintptr_t>(votes) & 15u)))
        return f3d_fail(F3D_ERR_ARG, "f3d_exchange_merge: bad argument");
    if (nrows == 0) return F3D_OK;
    FuseResolve RP;
    int rc = f3d_build_resolve(C1, threshold, h_filter, nfilter, nclasses_id, RP);
    if (rc) return rc;
    PeerLabels PL;
    rc = fill_peer_labels(PL, labels ? h_peer_labels16 : nullptr, nranks, first_point, "f3d_exchange_merge: bad peer label table");
    if (rc) return rc;
    const int RS = xch_row_stride(C1);
    const size_t smem = RES_MAXC * sizeof(int16_t) + (((size_t)XCH_BLOCK * RS * 2 + 15) & ~(size_t)15);
    cudaError_t e = cudaFuncSetAttribute(slot_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return f3d_check_launch("f3d_exchange_merge(cudaFuncSetAttribute)");
    const long long tiles = (nrows + XCH_BLOCK - 1) / XCH_BLOCK;
    slot_merge_kernel<<<(unsigned)tiles, XCH_BLOCK, smem, (cudaStream_t)stream>>>(slots, reinterpret_cast<const uint2*>(dir), nranks,
                                                                                (long long)sub_rows * F3D_XCH_NREG,
                                                                                points_per_shard / 32, (long long)nrows, C1, RS, RP,
                                                                                votes, labels, PL);
    return f3d_check_launch("f3d_exchange_merge");
}
