/*
 * libf3d -- C ABI of the B200-native multi-view 2D->3D label-fusion path.
 *
 * The reference (raviraj988/3D-POINT-CLOUD-SEGMENTATION-USING-2D-IMG-SEGMENTATION) is pure Python and has no
 * FFI; this header is the boundary a maintainer would bind with ctypes.  Each entry point
 * cites the reference statement(s) it replaces (paths relative to the reference repository).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer borrowed from the caller (e.g. a torch tensor) unless its name starts
 *    with `h_` (host).  The library never allocates, frees or retains caller memory and keeps no global
 *    mutable state, so calls on different streams / devices may run concurrently.
 *  - `stream` is a `cudaStream_t` passed as `void*`; all work is enqueued on it, nothing synchronises.
 *  - return value: 0 = OK, < 0 = error; `f3d_last_error()` gives a thread-local message.
 *  - there is no CPU fallback: without a CUDA device every compute call returns F3D_ERR_CUDA.
 */
#ifndef F3D_H_
#define F3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F3D_OK 0
#define F3D_ERR_ARG (-1)
#define F3D_ERR_CUDA (-2)
#define F3D_ERR_UNSUPPORTED (-3)

/* depth image formats (per-frame [H,W] images, frames contiguous) */
#define F3D_DEPTH_U16_MM 0 /* uint16 millimetres, RTAB export convention (RTAB_utils/ios_rtab.py:97-113,185) */
#define F3D_DEPTH_F32_M 1  /* float32 metres (extension: the /1000 step of ios_rtab.py:185 is skipped)      */
/* packed frame formats produced by f3d_pack_frames: one uint32 texel per pixel = uint16 depth mm | class id << 16, so a
 * point-view costs one 32-byte sector instead of a depth and a mask sector.  `depth` then points to the texels and
 * `mask` is ignored (may be NULL).  Device layout only: the on-disk contract (16-bit depth PNG ios_rtab.py:97-113, uint8
 * mask PNG voting.py:66) is unchanged. */
#define F3D_FRAMES_U32 2     /* texels row-major [F,H,W]                                                      */
#define F3D_FRAMES_U32_T16 3 /* 16x16-pixel tiles contiguous, f3d_packed_frame_texels(H,W,3) texels per frame   */

/* indices into the uint64 statistics block written by f3d_fuse_* (F3D_NSTATS entries, accumulated) */
#define F3D_STAT_CANDIDATES 0 /* point-views that survived the conservative tile x frustum cull            */
#define F3D_STAT_EXACT 1      /* point-views re-evaluated in fp64 (inside an fp32 uncertainty band)        */
#define F3D_STAT_DIVERGED 2   /* of those, fp32 best guess != fp64 outcome (logged fp32-vs-fp64 divergence) */
#define F3D_STAT_NEAR_EDGE 3  /* in-bounds point-views whose fp64 u or v is within 1e-4 px of an integer   */
#define F3D_STAT_SEEN 4       /* point-views that passed visibility (votes cast / depth samples written)   */
#define F3D_STAT_AUDIT_BAD 5  /* audit mode only: fp32 path certified an outcome that fp64 contradicts     */
#define F3D_NSTATS 8

const char* f3d_last_error(void);
int f3d_version(void);

/* Timing of the fused kernel alone (not the first cull level or the fix-up kernels of the same call): hands two of the
 * CALLER's cudaEvent_t to the next f3d_fuse_* / f3d_zbuffer_splat call made by THIS host thread, which records them on its
 * stream immediately before and after its fuse_kernel launch and forgets them (thread-local one-shot, like
 * f3d_last_error; nothing else is kept).  NULL, NULL disarms. */
int f3d_fuse_time_next_call(void* event_start, void* event_stop);


/* ---- frame table -------------------------------------------------------------------------------------- */

/* Bytes of packed per-frame data the caller must provide per frame to f3d_frames_setup. */
int64_t f3d_frame_table_bytes(int32_t nframes);

/* Builds the per-frame table on the device: fp64 pose / inverse pose / frustum planes in the reference's
 * operation order (Fusion._get_frustum_data, Fusion3DSeg/fusion.py:119-132; camera_utils.py:60-171; planes
 * fusion.py:254-258) plus fp32 projection tiles for the fast path.
 *   h_K9      host, row-major 3x3 intrinsic (upper triangular)        camera_utils.py:14
 *   wxyz      [F,4] float64 (w,x,y,z), NOT normalised                 fusion.py:71-72
 *   trans     [F,3] float64
 *   max_depth far-plane distance along the look-at ray                fusion.py:256 */
int f3d_frames_setup(const double* h_K9, int32_t W, int32_t H, const double* wxyz, const double* trans,
                     int32_t nframes, double max_depth, void* frame_table, void* stream);

/* Copies the fp64 frustum data out of a frame table for inspection / parity tests:
 * eyes [F,3], lookats [F,3], face_normals [F,4,3] (device float64). */
int f3d_frames_export(const void* frame_table, int32_t nframes, double* eyes, double* lookats,
                      double* face_normals, void* stream);

/* ---- kernel (1): fused project + z-test + mask gather + vote (level P) ---------------------------------- */

/* Optional scratch for the f3d_fuse_* / f3d_zbuffer_splat calls (16-byte aligned device memory, contents
 * irrelevant): point-views whose fp32 decision is uncertain are queued there and re-evaluated in fp64 by dense
 * fix-up kernels after the sweep.  Without it (NULL / 0) they are evaluated inside the sweep -- same results,
 * slower.  A queue that fills up is handled the same way, so the size is a performance knob only. */
int64_t f3d_fuse_workspace_bytes(int64_t npoints);

/* Replaces, for a FIXED cloud, the per-frame body of Fusion.fuse (fusion.py:248-298: point_inside_polyhedra
 * -> points2pixel -> single-pixel criterion) composed with VotingSegmentation.vote (segUtils/voting.py:89-98).
 *   points   [N] float4 (x,y,z,unused) -- float32 values are the contract (widened exactly to fp64)
 *   depth    [F,H,W] uint16 mm or float32 m (depth_fmt); mask [F,H,W] uint8 class ids < C1;
 *            or depth = packed texels of f3d_pack_frames (depth_fmt F3D_FRAMES_U32 / _T16), mask NULL
 *   radius   criterion distance (fusion.py:225, strict <); zmin/zmax valid range (fusion.py:62-63: > / <=)
 *   votes    [N,C1] int32, row-major like the reference's votes[npts, nclasses+1] (voting.py:34).
 *            accumulate = 0: every cell is overwritten (no memset needed); 1: added to.  16-byte aligned.
 *   workspace optional scratch, see f3d_fuse_workspace_bytes
 *   stats    optional uint64[F3D_NSTATS], accumulated with atomics (caller zeroes)
 *   flags    bit 0: audit mode (fp64 for every candidate, counts F3D_STAT_AUDIT_BAD) */
int f3d_fuse_project_vote(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                          int32_t frame_end, const void* depth, int32_t depth_fmt, const uint8_t* mask,
                          int32_t H, int32_t W, const double* h_K9, double radius, double zmin, double zmax,
                          int32_t* votes, int32_t C1, int32_t accumulate, void* workspace, int64_t workspace_bytes,
                          uint64_t* stats, int32_t flags, void* stream);

/* f3d_fuse_project_vote with packed uint16 counters [N,C1] (valid below 65 536 frames per vote tensor).  Used for
 * the multi-GPU exchange: two uint16 counters viewed as one int32 add without carries, so an int32 sum
 * reduce-scatter of the [N,C1/2] view is exact and moves half the bytes. */
int f3d_fuse_project_vote_u16(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                              int32_t frame_end, const void* depth, int32_t depth_fmt, const uint8_t* mask,
                              int32_t H, int32_t W, const double* h_K9, double radius, double zmin, double zmax,
                              uint16_t* votes_u16, int32_t C1, int32_t accumulate, void* workspace,
                              int64_t workspace_bytes, uint64_t* stats, int32_t flags, void* stream);

/* ---- multi-GPU vote exchange over peer memory (NVLink / NVSwitch), fused into the compute kernel ------------------
 * Reference semantics: votes are integer sums over frames (VotingSegmentation.vote, segUtils/voting.py:89-98), so frames
 * can be sharded over ranks and the partial votes added in any order -- bit-exact for every rank count.
 * Votes are ~95 % zeros, so no dense partial vote tensor is written or reduce-scattered.  Rank d owns the points
 * [d * points_per_shard, (d+1) * points_per_shard), points_per_shard a multiple of 256.  f3d_exchange_constants returns
 * {NREG, NSUB, NSUB_FIX, NLEVEL}.  The fused kernel of every source rank writes, straight into the owner's memory through
 * peer-mapped pointers:
 *   - slot records: per (source, 32-point block) L rows of 64 B; row j holds, for each of the block's 32 points, its
 *     j-th class in order of first appearance as uint16 (class | count << 8, 0 = none); L = the longest list in the
 *     block.  The block's warp reserves the rows in one of NREG sub-regions (sub_rows rows each) of its record region
 *     and writes a directory entry {uint32 row offset, uint32 L}; a tile that flushes its byte histogram k times (more than
 *     235 candidate frames each) writes k records, the directory keeps NLEVEL entries per block.
 *     h_peer_slots[d] / h_peer_dirs[d] = device pointers (peer mapped) to THIS rank's record region
 *     [NREG * sub_rows rows] and directory [points_per_shard / 32][NLEVEL] inside rank d's receive buffer.  Every directory
 *     entry is rewritten on every call (no clearing needed);
 *   - (cell, count) entries for everything else (a full sub-region, more than NLEVEL flushes, the deferred fp64
 *     votes): h_peer_queues[d] = peer pointer to THIS rank's queue
 *     [NSUB][sub_cap] uint64 (cell = local_point * C1 + class in the low 40 bits, count above) inside rank d's
 *     buffer; the fix-up kernel's block b owns sub-queue b < NSUB_FIX, the rest take spills.
 * cursors: local uint32 [nranks * (NREG + NSUB)] (row cursors, then queue cursors), zeroed by the caller before the call;
 * *overflow is set to 1 when a sub-queue filled up (entries dropped: the caller must check it and enlarge sub_cap or
 * fall back to the dense exchange).  The deferred-fp64 workspace (f3d_fuse_workspace_bytes) is mandatory here.
 * f3d_exchange_publish then copies the queue cursors into every destination's count table (h_peer_counts[d] = peer
 * pointer to rank d's uint32 [nranks][NSUB] table; row [rank] is written).  After a cross-rank barrier the owner runs
 *   f3d_exchange_merge       (records of all sources -> dense int32 shard rows [nrows, C1], every cell written once, and
 *                             labels: VotingSegmentation.segment, voting.py:106-137), then
 *   f3d_exchange_queue_apply (queue entries scatter-added into the shard, labels of the touched points re-resolved).
 * The label all-gather can ride inside these two kernels: h_peer_labels16 (host array of nranks device pointers, one per
 * rank, each an int16 array of nranks * points_per_shard labels in peer-mapped memory; NULL = off) makes the owner store
 * every label it resolves at index first_point + row of all nranks arrays (labels must fit int16). */
/* f3d_fuse_project_vote_exchange, flags bit 1: launch the fused kernel only over the 4096-point super-tiles that some frame
 * of the call can see (a rank of a frame-sharded job sees a fraction of the cloud).  The call then SYNCHRONISES the stream
 * once (it reads the number of such super-tiles) -- not capturable into a CUDA graph.  Results are identical. */
#define F3D_FUSE_COMPACT 2
int f3d_exchange_constants(int32_t* out4);
int f3d_fuse_project_vote_exchange(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                                   int32_t frame_end, const void* depth, int32_t depth_fmt, const uint8_t* mask,
                                   int32_t H, int32_t W, const double* h_K9, double radius, double zmin, double zmax,
                                   int32_t C1, int32_t nranks, int64_t points_per_shard, const uint64_t* h_peer_slots,
                                   const uint64_t* h_peer_dirs, const uint64_t* h_peer_queues, int64_t sub_rows,
                                   int64_t sub_cap, uint32_t* cursors, uint32_t* overflow, void* workspace,
                                   int64_t workspace_bytes, uint64_t* stats, int32_t flags, void* stream);
int f3d_exchange_publish(const uint32_t* cursors, const uint64_t* h_peer_counts, int32_t rank, int32_t nranks,
                         int64_t sub_cap, void* stream);
int f3d_exchange_merge(const uint16_t* slots, const void* dir, int32_t nranks, int64_t sub_rows, int64_t points_per_shard,
                       int64_t nrows, int32_t C1, double threshold, const int32_t* h_filter, int32_t nfilter,
                       int32_t nclasses_id, int32_t* votes, int64_t* labels, const uint64_t* h_peer_labels16,
                       int64_t first_point, void* stream);
int f3d_exchange_queue_apply(const uint64_t* queue, const uint32_t* counts, int32_t nranks, int64_t sub_cap, int32_t* votes,
                             int64_t nrows, int32_t C1, double threshold, const int32_t* h_filter, int32_t nfilter,
                             int32_t nclasses_id, int64_t* labels, const uint64_t* h_peer_labels16, int64_t first_point,
                             void* stream);

/* f3d_fuse_project_vote with VotingSegmentation.segment (segUtils/voting.py:106-137, see f3d_resolve_labels) fused
 * into the epilogue: labels [N] int64 are resolved straight from the on-chip histograms, so the vote tensor is
 * never re-read.  All frames must be covered by this one call (no accumulation).  votes may be NULL: then only
 * labels are produced and the 4*N*C1-byte vote write is skipped. */
int f3d_fuse_project_vote_resolve(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                                  int32_t frame_end, const void* depth, int32_t depth_fmt, const uint8_t* mask,
                                  int32_t H, int32_t W, const double* h_K9, double radius, double zmin, double zmax,
                                  int32_t* votes, int32_t C1, double threshold, const int32_t* h_filter,
                                  int32_t nfilter, int32_t nclasses_id, int64_t* labels, void* workspace,
                                  int64_t workspace_bytes, uint64_t* stats, int32_t flags, void* stream);

/* Same traversal, but writes the reference's exchange format instead of votes: uv2pt [F,H*W] int32,
 * value = highest cloud-point index seen through the pixel, -1 = none (fusion.py:253,297,322).  The caller
 * fills uv2pt with -1 first. */
int f3d_fuse_uv2pt(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                   int32_t frame_end, const void* depth, int32_t depth_fmt, int32_t H, int32_t W,
                   const double* h_K9, double radius, double zmin, double zmax, int32_t* uv2pt, void* workspace,
                   int64_t workspace_bytes, uint64_t* stats, int32_t flags, void* stream);

/* ---- frame packing (ingest) ---------------------------------------------------------------------------------------- */

/* uint32 texels per frame of a packed frame stack (H*W, or 256 per 16x16 tile for F3D_FRAMES_U32_T16). */
int64_t f3d_packed_frame_texels(int32_t H, int32_t W, int32_t frame_fmt);

/* depth_mm [F,H,W] uint16 + mask [F,mask_h,mask_w] uint8 -> out [F, f3d_packed_frame_texels] uint32 (depth | class << 16).
 * A mask at another resolution is nearest-resized on the fly with OpenCV's INTER_NEAREST index rule
 * (cv2.resize(mask, (w, h), INTER_NEAREST), segUtils/voting.py:93), i.e. f3d_resize_nearest_u8 fused into the pack. */
int f3d_pack_frames(const uint16_t* depth_mm, const uint8_t* mask, int32_t nframes, int32_t H, int32_t W, int32_t mask_h,
                    int32_t mask_w, int32_t frame_fmt, uint32_t* out, void* stream);

/* ---- kernel (2): z-buffer splat ----------------------------------------------------------------------------- */

/* Depth images of the cloud itself: per pixel min over in-frustum points of camera z (the row points2pixel
 * discards, camera_utils.py:23-24), quantised floor(z*1000+0.5) clamped to [1,65535].
 *   zbuf   scratch [F,H*W] uint32 (the call initialises it);  depth_out [F,H,W] uint16, 0 = hole;
 *   border: pixels closer than `border` to the image edge are zeroed (ios_rtab.py:105-109). */
int f3d_zbuffer_splat(const void* points, int64_t N, const void* frame_table, int32_t frame_begin,
                      int32_t frame_end, int32_t H, int32_t W, const double* h_K9, uint32_t* zbuf,
                      uint16_t* depth_out, int32_t border, void* workspace, int64_t workspace_bytes, uint64_t* stats,
                      int32_t flags, void* stream);

/* ---- level V: uv2pt + mask vote --------------------------------------------------------------------------------- */

/* VotingSegmentation.vote for a batch of frames (segUtils/voting.py:89-98): votes[uv2pt[i], mask[i]] += 1 with
 * numpy's per-frame collapse of duplicate (point, class) pairs.  `votes_packed` [N,C1] uint32 holds
 * (frame_tag << 16 | count) while accumulating; frame tags first_tag .. first_tag+F-1 must stay in 1..65535
 * and increase across calls.  Call f3d_vote_finalize to strip the tags. */
int f3d_vote_uv2pt(const int32_t* uv2pt, const uint8_t* mask, int32_t nframes, int64_t npix, int32_t first_tag,
                   uint32_t* votes_packed, int64_t N, int32_t C1, void* stream);
int f3d_vote_finalize(uint32_t* votes_packed, int64_t ncells, void* stream);

/* cv2.resize(mask, (w, h), INTER_NEAREST) (voting.py:93) for a batch of uint8 images. */
int f3d_resize_nearest_u8(const uint8_t* src, int32_t nimg, int32_t src_h, int32_t src_w, uint8_t* dst,
                          int32_t dst_h, int32_t dst_w, void* stream);

/* ---- kernel (3): label resolve ------------------------------------------------------------------------------------ */

/* VotingSegmentation.segment (segUtils/voting.py:106-137): total over all C1 columns, first arg-max over the
 * filter columns (in the given order; NULL/0 = all columns), max/total < threshold (float64) or max == 0 or
 * total == 0 -> nclasses_id, then the sequential index->class remap with its aliasing.
 *   h_filter host int32[nfilter];  labels [N] int64. */
int f3d_resolve_labels(const int32_t* votes, int64_t N, int32_t C1, double threshold, const int32_t* h_filter,
                       int32_t nfilter, int32_t nclasses_id, int64_t* labels, void* stream);

int f3d_resolve_labels_u16(const uint16_t* votes_u16, int64_t N, int32_t C1, double threshold,
                           const int32_t* h_filter, int32_t nfilter, int32_t nclasses_id, int64_t* labels, void* stream);

/* ---- single-call projection / cull operators (a-1, a-4) ----------------------------------------------------------- */

/* points2pixel (Fusion3DSeg/camera_utils.py:9-26): uv [2,N] int32, evaluated in fp64. points [N,3] float64. */
int f3d_project_pixels(const double* points, int64_t N, const double* h_K9, const double* h_wxyz,
                       const double* h_t, int32_t* uv, void* stream);

/* SpatQuadranion.rotate (RTAB_utils/spatQuad.py:6-28): raw Hamilton sandwich q p q* on the UN-normalised h_wxyz, fp64 in the
 * reference's operation order.  points, out [N,3] float64 (may alias). */
int f3d_quat_rotate(const double* points, int64_t N, const double* h_wxyz, double* out, void* stream);

/* point_inside_polyhedra (Fusion3DSeg/intersections.py:146-164): inside [N] uint8. */
int f3d_frustum_mask(const double* points, int64_t N, const double* h_plane_points, const double* h_normals,
                     int32_t nplanes, uint8_t* inside, void* stream);

/* ---- kernel (4): instance-box merge -------------------------------------------------------------------------------- */

/* Closed-interval AABB overlap of check_intersection (merge_intersecting_bb.py:49-53) over all pairs i<j
 * with equal group id.  edges [cap,2] int32, *count = number found (may exceed cap: then re-run larger). */
int f3d_box_pairs_aabb(const double* lo, const double* hi, const int32_t* group, int32_t B, int32_t* edges,
                       int64_t cap, unsigned long long* count, void* stream);

/* Same pair set as f3d_box_pairs_aabb through a sort-and-sweep broad phase: `order` [B] int32 lists the boxes sorted by
 * (group, lo.x) (the caller sorts, e.g. torch); a box is tested only against the later boxes of its group whose lo.x does
 * not exceed its hi.x (or ties its lo.x), with the exact float64 closed-interval predicate (merge_intersecting_bb.py:51-53).
 * Edges are written as (min, max) box indices, unordered. */
int f3d_box_pairs_sweep(const double* lo, const double* hi, const int32_t* group, const int32_t* order, int32_t B,
                        int32_t* edges, int64_t cap, unsigned long long* count, void* stream);

/* Union-find closure: labels[i] = smallest box index of i's connected component. */
int f3d_union_find(int32_t B, const int32_t* edges, int64_t E, int32_t* labels, void* stream);

/* Open3D OrientedBoundingBox.get_point_indices_within_bounding_box rule (merge_intersecting_bb.py:76,87):
 * inside [nboxes, N] uint8 for boxes [nboxes,15] float64 = centre[3], R[9] (row major, columns = axes),
 * extent[3] (device). */
int f3d_obb_contains(const double* points, int64_t N, const double* boxes15, int32_t nboxes,
                     uint8_t* inside, void* stream);

/* Batched oriented-box fit (the box behind the o3d.geometry.OrientedBoundingBox.create_from_points call sites
 * merge_intersecting_bb.py:18,75,86,126 and get3DSeg.py:434-436), every requested instance in one pass over the cloud:
 *   points [N,3] float64, ids [N] int64 instance id per point, slot_of_id [nslot] int32: output slot of instance id
 *   (-1 = not requested; ids outside [0, nslot) are ignored), ninst output slots.
 *   model 0: centre / axes from the mean and covariance of ALL the instance's points (3x3 Jacobi, axes by descending
 *            eigenvalue, third = first x second), extent = range of the projections, centre = mean + R @ mid-range;
 *   model 1: axis-aligned box of the points (R = I).
 * Open3D's own fit (Qhull hull vertices first) is NOT reproduced -- see oracle.fit_box for the stated models.
 *   boxes15 [ninst,15] float64 = centre[3], R[9] row major (columns = axes), extent[3]; counts [ninst] int64 points per slot;
 *   workspace: f3d_obb_fit_workspace_bytes(ninst) bytes of device scratch. */
int64_t f3d_obb_fit_workspace_bytes(int32_t ninst);
int f3d_obb_fit(const double* points, const int64_t* ids, int64_t N, const int32_t* slot_of_id, int64_t nslot, int32_t ninst,
                int32_t model, double* boxes15, int64_t* counts, void* workspace, void* stream);

/* ---- radius adjacency ------------------------------------------------------------------------- */

/* KDTree(points).query_radius(points, r) (Fusion3DSeg/fusion.py:374-375) on a uniform grid of cell size r:
 * membership = scikit-learn's Euclidean leaf test, reduced distance (dx*dx + dy*dy) + dz*dz <= r*r in float64 (a point is
 * its own neighbour).  h_min3 / h_max3: host, component-wise minimum / maximum of the cloud.
 *   f3d_radius_grid_keys   keys [N] int64 = cell key of every point; the caller sorts them (stable) and passes the sorted
 *                          keys and the sorting permutation `order` [N] int64 to
 *   f3d_radius_adjacency   pass 1 (indices NULL): counts [N] int64 = row lengths; the caller builds indptr [N+1] by an
 *                          exclusive scan; pass 2 (indices given): rows written at indptr[i] and sorted ascending. */
int f3d_radius_grid_keys(const double* points, int64_t N, const double* h_min3, const double* h_max3, double r, int64_t* keys,
                         void* stream);
int f3d_radius_adjacency(const double* points, int64_t N, const double* h_min3, const double* h_max3, double r,
                         const int64_t* sorted_keys, const int64_t* order, int64_t* counts, const int64_t* indptr,
                         int64_t* indices, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* F3D_H_ */
