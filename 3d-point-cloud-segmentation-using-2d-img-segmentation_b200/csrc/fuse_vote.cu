// Kernel (1): C ABI of the fused project + z-test + mask gather + vote path (dense votes, packed uint16 votes, fused label
// resolve, multi-GPU record exchange).  Device code: fuse_kernel.cuh.
#include "fuse_kernel.cuh"

FuseTimingSlot& f3d_timing_slot() {
    static thread_local FuseTimingSlot slot = {{nullptr, nullptr}, false};
    return slot;
}

extern "C" int f3d_fuse_time_next_call(void* event_start, void* event_stop) {
    FuseTimingSlot& s = f3d_timing_slot();
    if (!event_start || !event_stop) {
        s.armed = false;
        return F3D_OK;
    }
    s.ev[0] = (cudaEvent_t)event_start;
    s.ev[1] = (cudaEvent_t)event_stop;
    s.armed = true;
    return F3D_OK;
}

extern "C" int64_t f3d_fuse_workspace_bytes(int64_t npoints) {
    // room for one uncertain point-view per 4 points (measured: ~0.08 per point on the 1920x1440 scene), at least 1 Mi entries,
    // plus the super-tile candidate lists of the first cull level and the 8-byte per-point resolve states
    int64_t cap = npoints / 4;
    if (cap < (1 << 20)) cap = 1 << 20;
    if (npoints < 0) npoints = 0;
    return 16 + supertile_bytes(npoints) + npoints * 8 + cap * (int64_t)sizeof(GEntry);
}

extern "C" int64_t f3d_packed_frame_texels(int32_t H, int32_t W, int32_t frame_fmt) {
    if (H <= 0 || W <= 0) return 0;
    if (frame_fmt == F3D_FRAMES_U32_T16) return (int64_t)((W + 15) / 16) * ((H + 15) / 16) * 256;
    return (int64_t)H * W;
}

// composed sequential remap `for i, cls in enumerate(filter): pc[pc == i] = cls` (voting.py:133-135) and the
// column -> filter position table, shared with f3d_resolve_labels
int f3d_build_resolve(int C1, double threshold, const int32_t* h_filter, int nfilter, int nclasses_id, FuseResolve& rp) {
    if (C1 > RES_MAXC || nfilter > RES_MAXC) return f3d_fail(F3D_ERR_UNSUPPORTED, "label resolve: more than 256 columns / filter classes");
    rp.enabled = 1;
    rp.nfilter = nfilter;
    rp.threshold = threshold;
    for (int c = 0; c < RES_MAXC; ++c) {
        rp.fpos[c] = nfilter > 0 ? (int16_t)-1 : (int16_t)c;
        rp.remap[c] = c;
    }
    for (int k = nfilter - 1; k >= 0; --k) {
        if (h_filter[k] < 0 || h_filter[k] >= C1) return f3d_fail(F3D_ERR_ARG, "label resolve: filter class out of range");
        rp.fpos[h_filter[k]] = (int16_t)k;   // first position wins
    }
    rp.unclassified = nclasses_id;
    for (int start = 0; start <= nfilter; ++start) {
        int v = start < nfilter ? start : nclasses_id;
        for (int i = 0; i < nfilter; ++i)
            if (v == i) v = h_filter[i];
        if (start < nfilter) rp.remap[start] = v;
        else rp.unclassified = v;
    }
    return F3D_OK;
}

static int launch_vote(int fmt, const FuseParams& P, const FuseResolve& RP, int audit, cudaStream_t stream) {
    switch (fmt) {
        case F3D_DEPTH_U16_MM: return launch_fuse<MODE_VOTE, F3D_DEPTH_U16_MM>(P, RP, audit, stream);
        case F3D_DEPTH_F32_M: return launch_fuse<MODE_VOTE, F3D_DEPTH_F32_M>(P, RP, audit, stream);
        case F3D_FRAMES_U32: return launch_fuse<MODE_VOTE, F3D_FRAMES_U32>(P, RP, audit, stream);
        default: return launch_fuse<MODE_VOTE, F3D_FRAMES_U32_T16>(P, RP, audit, stream);
    }
}

static int fuse_vote_impl(const void* points, int64_t N, const void* frame_table, int32_t frame_begin, int32_t frame_end,
                          const void* depth, int32_t depth_fmt, const uint8_t* mask, int32_t H, int32_t W,
                          const double* h_K9, double radius, double zmin, double zmax, int32_t* votes, uint16_t* votes16,
                          int32_t C1, int32_t accumulate, const FuseResolve& RP, int64_t* labels, void* workspace,
                          int64_t workspace_bytes, uint64_t* stats, int32_t flags, void* stream) {
    FuseParams P;
    int rc = fill_common(P, points, N, frame_table, frame_begin, frame_end, depth, depth_fmt, H, W, h_K9, radius, zmin,
                         zmax, stats);
    if (rc) return rc;
    const bool packed = fmt_is_packed(depth_fmt);
    if ((!votes && !votes16 && !labels) || (votes && votes16) || C1 <= 0 || C1 > 256 ||
        (frame_end > frame_begin && (!depth || (!mask && !packed))))
        return f3d_fail(F3D_ERR_ARG, "f3d_fuse_project_vote: bad argument (votes/mask/depth NULL or C1 not in 1..256)");
    if (labels && (accumulate || frame_end - frame_begin > F3D_MAX_FRAMES_PER_LAUNCH))
        return f3d_fail(F3D_ERR_ARG, "f3d_fuse_project_vote_resolve: fused labels need all frames in one non-accumulating launch");
    if ((votes && (reinterpret_cast<uintptr_t>(votes) & 15u)) || (votes16 && (reinterpret_cast<uintptr_t>(votes16) & 15u)))
        return f3d_fail(F3D_ERR_ARG, "f3d_fuse_project_vote: votes must be 16-byte aligned");
    if (N == 0) return F3D_OK;
    const int audit = flags & 1;
    P.votes = votes;
    P.votes16 = votes16;
    P.labels = labels;
    P.C1 = C1;
    attach_workspace(P, workspace, workspace_bytes, (votes || votes16) && !audit && N <= 0x7fffffff,
                     votes && labels && frame_end - frame_begin < (1 << 24));   // labels-only / audit: fp64 inside the sweep
    P.RS = hist_row_stride(C1);
    const size_t esz = frame_elem_bytes(depth_fmt);
    int fb = frame_begin;
    bool first = true;
    do {
        int fe = frame_end - fb > F3D_MAX_FRAMES_PER_LAUNCH ? fb + F3D_MAX_FRAMES_PER_LAUNCH : frame_end;
        P.f_begin = fb;
        P.f_end = fe;
        P.depth = reinterpret_cast<const char*>(depth) + (size_t)(fb - frame_begin) * (size_t)P.frame_stride * esz;
        P.mask = packed ? nullptr : mask + (size_t)(fb - frame_begin) * H * W;
        P.accumulate = (first && !accumulate) ? 0 : 1;
        rc = launch_vote(depth_fmt, P, RP, audit, (cudaStream_t)stream);
        if (rc) return rc;
        first = false;
        fb = fe;
    } while (fb < frame_end);
    return F3D_OK;
}

extern "C" int f3d_fuse_project_vote(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                                     int32_t frame_end, const void* depth, int32_t depth_fmt, const uint8_t* mask,
                                     int32_t H, int32_t W, const double* h_K9, double radius, double zmin, double zmax,
                                     int32_t* votes, int32_t C1, int32_t accumulate, void* workspace,
                                     int64_t workspace_bytes, uint64_t* stats, int32_t flags, void* stream) {
    if (!votes) return f3d_fail(F3D_ERR_ARG, "f3d_fuse_project_vote: votes is NULL");
    FuseResolve RP;
    RP.enabled = 0;
    return fuse_vote_impl(points, N, frame_table, frame_begin, frame_end, depth, depth_fmt, mask, H, W, h_K9, radius, zmin,
                          zmax, votes, nullptr, C1, accumulate, RP, nullptr, workspace, workspace_bytes, stats, flags, stream);
}

extern "C" int f3d_fuse_project_vote_u16(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                                         int32_t frame_end, const void* depth, int32_t depth_fmt, const uint8_t* mask,
                                         int32_t H, int32_t W, const double* h_K9, double radius, double zmin, double zmax,
                                         uint16_t* votes_u16, int32_t C1, int32_t accumulate, void* workspace,
                                         int64_t workspace_bytes, uint64_t* stats, int32_t flags, void* stream) {
    if (!votes_u16) return f3d_fail(F3D_ERR_ARG, "f3d_fuse_project_vote_u16: votes is NULL");
    FuseResolve RP;
    RP.enabled = 0;
    return fuse_vote_impl(points, N, frame_table, frame_begin, frame_end, depth, depth_fmt, mask, H, W, h_K9, radius, zmin,
                          zmax, nullptr, votes_u16, C1, accumulate, RP, nullptr, workspace, workspace_bytes, stats, flags, stream);
}

extern "C" int f3d_fuse_project_vote_resolve(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                                             int32_t frame_end, const void* depth, int32_t depth_fmt, const uint8_t* mask,
                                             int32_t H, int32_t W, const double* h_K9, double radius, double zmin,
                                             double zmax, int32_t* votes, int32_t C1, double threshold,
                                             const int32_t* h_filter, int32_t nfilter, int32_t nclasses_id, int64_t* labels,
                                             void* workspace, int64_t workspace_bytes, uint64_t* stats, int32_t flags,
                                             void* stream) {
    if (!labels || nfilter < 0 || (nfilter > 0 && !h_filter))
        return f3d_fail(F3D_ERR_ARG, "f3d_fuse_project_vote_resolve: bad argument");
    FuseResolve RP;
    int rc = f3d_build_resolve(C1, threshold, h_filter, nfilter, nclasses_id, RP);
    if (rc) return rc;
    return fuse_vote_impl(points, N, frame_table, frame_begin, frame_end, depth, depth_fmt, mask, H, W, h_K9, radius, zmin,
                          zmax, votes, nullptr, C1, 0, RP, labels, workspace, workspace_bytes, stats, flags, stream);
}

// ---- vote exchange over peer memory: sender side (owner side: vote_exchange.cu) ---------------------------------------
extern "C" int f3d_fuse_project_vote_exchange(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                                              int32_t frame_end, const void* depth, int32_t depth_fmt, const uint8_t* mask,
                                              int32_t H, int32_t W, const double* h_K9, double radius, double zmin, double zmax,
                                              int32_t C1, int32_t nranks, int64_t points_per_shard, const uint64_t* h_peer_slots,
                                              const uint64_t* h_peer_dirs, const uint64_t* h_peer_queues, int64_t sub_rows,
                                              int64_t sub_cap, uint32_t* cursors, uint32_t* overflow, void* workspace,
                                              int64_t workspace_bytes, uint64_t* stats, int32_t flags, void* stream) {
    FuseParams P;
    int rc = fill_common(P, points, N, frame_table, frame_begin, frame_end, depth, depth_fmt, H, W, h_K9, radius, zmin,
                         zmax, stats);
    if (rc) return rc;
    const bool packed = fmt_is_packed(depth_fmt);
    if (!h_peer_slots || !h_peer_dirs || !h_peer_queues || nranks < 1 || nranks > F3D_MAX_RANKS || points_per_shard <= 0 ||
        (points_per_shard % 256) != 0 || sub_rows <= 0 || sub_cap <= 0 || sub_rows * F3D_XCH_NREG > 0xffffffffLL ||
        sub_cap > 0x7fffffffLL || !cursors || !overflow || C1 <= 0 || C1 > 256 || (frame_end > frame_begin && (!depth || (!mask && !packed))))
        return f3d_fail(F3D_ERR_ARG, "f3d_fuse_project_vote_exchange: bad argument (points_per_shard must be a multiple of 256)");
    if (frame_end - frame_begin > F3D_MAX_FRAMES_PER_LAUNCH)
        return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_fuse_project_vote_exchange: too many frames per call (limit 65515)");
    if ((int64_t)points_per_shard * C1 >= (1LL << 40))
        return f3d_fail(F3D_ERR_UNSUPPORTED, "f3d_fuse_project_vote_exchange: shard cell index does not fit 40 bits");
    if (points_per_shard * nranks < N)
        return f3d_fail(F3D_ERR_ARG, "f3d_fuse_project_vote_exchange: nranks * points_per_shard does not cover the cloud");
    if (N == 0) return F3D_OK;
    P.C1 = C1;
    P.f_begin = frame_begin;
    P.f_end = frame_end;
    P.mask = packed ? nullptr : mask;
    P.xg_G = nranks;
    P.xg_per = points_per_shard;
    for (int i = 0; i < nranks; ++i) {
        P.xg_slots[i] = reinterpret_cast<uint16_t*>(h_peer_slots[i]);
        P.xg_dir[i] = reinterpret_cast<uint2*>(h_peer_dirs[i]);
        P.xg_queue[i] = reinterpret_cast<unsigned long long*>(h_peer_queues[i]);
    }
    P.xg_rowcur = cursors;
    P.xg_qcur = cursors + (size_t)nranks * F3D_XCH_NREG;
    P.xg_subrows = (unsigned)sub_rows;
    P.xg_subcap = (unsigned)sub_cap;
    P.xg_overflow = overflow;
    // the fix-up kernel's blocks own the first F3D_XCH_NSUB_FIX sub-queues: the deferred queue is mandatory here
    if ((flags & 1) || N > 0x7fffffff || !attach_workspace(P, workspace, workspace_bytes))
        return f3d_fail(F3D_ERR_ARG, "f3d_fuse_project_vote_exchange: needs the workspace of f3d_fuse_workspace_bytes (no audit mode)");
    P.compact = (flags & F3D_FUSE_COMPACT) ? 1 : 0;
    FuseResolve RP;
    RP.enabled = 0;
    return launch_vote(depth_fmt, P, RP, 0, (cudaStream_t)stream);
}
