"""Drop-in for the parts of `Fusion3DSeg/fusion.py` that sit on the label-fusion path.

* `parse_rts`, `FrameData.get_valid`, `Fusion.load_data` / `Fusion._get_frustum_data` keep the reference's names,
  arguments and return layouts (reference `fusion.py:50-77,119-132,389-407`).
* `Fusion.fuse` proper (cloud construction by running-mean merging + random patch down-sampling,
  `fusion.py:134-324`) is sequential, order dependent and randomised; it is outside the bit-exact contract
   and is not re-implemented here.
* The GPU path holds the cloud FIXED and offers the same hand-off formats: `Fusion.label_fixed_cloud` produces
  votes/labels directly, `Fusion.write_uv2pt_fixed_cloud` writes `fusion/uv2pt/<frame>.npy` in the reference's
  exchange format so the stock `get3DSeg.segment` flow (and ours) can consume it.
"""
from __future__ import annotations

import pickle
from pathlib import Path

import numpy as np

from .. import engine
from ..fused import FusedLabeler


def parse_rts(rts):
    """Reference `parse_rts` (`fusion.py:67-77`): scaled intrinsics, depth size and poses; the stored quaternions
    are (x, y, z, w) and are re-ordered to (w, x, y, z)."""
    with open(rts, 'rb') as fp:
        rtsdata = pickle.load(fp)
    Ks = rtsdata['intrinsicScaled']
    wxyzs = np.asarray(rtsdata['odo_wxyz'])[:, [3, 0, 1, 2]]
    translations = np.asarray(rtsdata['odo_xyz'])
    h, w, *_ = rtsdata['Depth_res']
    return Ks, w, h, wxyzs, translations


class FrameData:
    """Reference `FrameData` (`fusion.py:17-64`): per-frame pickle loader of an RTAB cache
    (`PointcloudMergeResults/tofsegment_*.pkl` lists the per-frame files) with the valid-range test and the decimation
    lattice.  `__getitem__` returns the reference's 5-tuple; `depth_mm` gives the frame as the fused kernel wants it."""

    def __init__(self, tof, point_range=None, decimation=1, depth_hw=(256, 192)):
        self.point_range = point_range
        self.decimation = decimation
        self.depth_hw = depth_hw
        self.mask = np.ones(depth_hw, bool)
        dirname = Path(str(tof).split('PointcloudMergeResults')[0])
        with open(tof, 'rb') as fp:
            tofdata = pickle.load(fp)
            self.tofcamedata = [dirname / data['fileName'].strip() for data in tofdata]

    def __len__(self):
        return len(self.tofcamedata)

    def _load(self, i):
        with open(self.tofcamedata[i], 'rb') as fp:
            return pickle.load(fp)

    def _valids(self, orgpts):
        if self.point_range is not None:
            valids = self.get_valid(orgpts, self.point_range[0], self.point_range[1])
        else:
            valids = np.ones(len(orgpts), bool)
        if self.decimation > 1:                                   # fusion.py:43-46: only the ::decimation lattice stays valid
            mask = self.mask.copy()
            mask[::self.decimation, ::self.decimation] = False
            valids[mask.reshape(-1)] = False
        return valids

    def __getitem__(self, i):
        data = self._load(i)
        frame_name = str(data['frameNumber'])
        orgpts = np.array(data['orgPoints'])
        return frame_name, np.array(data['modPoints']), np.array(data['modSurfaceNormals']), np.array(data['orgColorPoints']), \
            self._valids(orgpts)

    def depth_mm(self, i):
        """(frame name, uint16 [h,w] depth in millimetres with every pixel `__getitem__` calls invalid set to 0).
        `orgPoints[:, 2]` is depth_png / 1000 in float64 (`ios_rtab.py:185`), so rounding z * 1000 recovers the sensor's
        integer exactly; a cache whose z is not a millimetre multiple is rejected (the kernel contract is integer mm)."""
        data = self._load(i)
        org = np.array(data['orgPoints'])
        z = org[:, 2] * 1000.0
        d = np.rint(z)
        if np.abs(z - d).max(initial=0.0) > 1e-6 or d.min(initial=0) < 0 or d.max(initial=0) > 65535:
            raise ValueError(f"frame {data['frameNumber']}: orgPoints z is not a uint16 millimetre depth")
        d = d.astype(np.uint16)
        d[~self._valids(org)] = 0
        return str(data['frameNumber']), d.reshape(self.depth_hw)

    @staticmethod
    def get_valid(points, mindist, maxdist):
        """Reference `FrameData.get_valid` (`fusion.py:50-64`): camera-space z in (mindist, maxdist]."""
        values = np.asarray(points)[:, 2]
        return (values > mindist) & (values <= maxdist)


class Fusion:
    """Only the static / class-level surface the label-fusion path uses (reference `fusion.py:80-407`)."""

    @staticmethod
    def _get_frustum_data(K, w, h, xyzws, translations, frame_ids=None):
        """Reference `Fusion._get_frustum_data` (`fusion.py:119-132`), computed by `f3d_frames_setup` on the GPU in the
        reference's float64 operation order.  Returns eyes [F,3], lookats [F,3], spoke origins [F,4,3],
        face normals [F,4,3] (numpy float64)."""
        translations = np.asarray(translations, dtype=np.float64).reshape(-1, 3)
        frame_ids = np.arange(len(translations)) if frame_ids is None else np.asarray(frame_ids)
        tab = engine.FrameTable(K, w, h, xyzws, translations, 1.0)
        eyes, look, nrm = [x.cpu().numpy() for x in tab.export()]
        eyes, look, nrm = eyes[frame_ids], look[frame_ids], nrm[frame_ids]
        return eyes, look, np.repeat(eyes[:, None, :], 4, axis=1), nrm

    @classmethod
    def load_data(cls, dirname):
        """Reference `Fusion.load_data` (`fusion.py:389-407`): 8 outputs, `adj` is None when `adj.pkl` is absent."""
        dirname = Path(dirname)
        with open(dirname / 'fusion' / 'fusion_data.pkl', 'rb') as fp:
            data = pickle.load(fp)
        out = [data['points'], data['normals'], data['colors'], data['nmerges'], data['occurences'], data['nframes'],
               data['depth_hw']]
        adjfile = dirname / 'fusion' / 'adj.pkl'
        if adjfile.is_file():
            with open(adjfile, 'rb') as fp:
                adj = pickle.load(fp)
        else:
            adj = None
        out.append(adj)
        return out

    @staticmethod
    def dump_data(dirname, points, normals=None, colors=None, nmerges=None, occurences=None, nframes=0, depth_hw=None,
                  compute_adjacency=False, ds_radius=None):
        """Writes `fusion/fusion_data.pkl` with the reference's keys (`fusion.py:360-368`) and, when asked, `fusion/adj.pkl`
        = `KDTree(points).query_radius(points, r=2*ds_radius)` (`fusion.py:369-377`) from the GPU uniform-grid search
        (`f3d_radius_adjacency`; rows sorted ascending, the reference's are in tree order).  No PLY, no GUI."""
        dirname = Path(dirname)
        (dirname / 'fusion').mkdir(exist_ok=True, parents=True)
        data = {'points': points, 'normals': normals, 'colors': colors, 'nmerges': nmerges, 'occurences': occurences,
                'nframes': nframes, 'depth_hw': depth_hw}
        with (dirname / 'fusion' / 'fusion_data.pkl').open('wb') as fp:
            pickle.dump(data, fp)
        if compute_adjacency:
            if ds_radius is None:
                adj = None
            else:
                indptr, indices = engine.radius_adjacency(engine.as_cuda(np.asarray(points, dtype=np.float64)), 2 * ds_radius)
                ip, ix = indptr.cpu().numpy(), indices.cpu().numpy()
                adj = np.empty(len(ip) - 1, dtype=object)
                for i in range(len(ip) - 1):
                    adj[i] = ix[ip[i]:ip[i + 1]]
            with (dirname / 'fusion' / 'adj.pkl').open('wb') as fp:
                pickle.dump(np.array(adj, dtype=object) if adj is None else adj, fp)

    # ---- fixed-cloud GPU drivers -------------------------------------------------------------------------------------
    @staticmethod
    def label_fixed_cloud(points, K, w, h, wxyzs, translations, depths, masks, point_range=(0.1, 4), radius=0.05,
                          nclasses=133, threshold=0.5, filter_classes=None):
        """Project + z-test + mask gather + vote + resolve for a fixed cloud (SURVEY 8(c) level P).
        depths [F,h,w] uint16 mm (or float32 m), masks [F,h,w] uint8.  Returns (votes float64 [N,nclasses+1],
        classes int64 [N]) -- the pair `get3DSeg.segment` returns (`get3DSeg.py:110`)."""
        fl = FusedLabeler(points, K, w, h, wxyzs, translations, point_range, radius, nclasses)
        fl.vote(depths, masks)
        classes = fl.segment(threshold, filter_classes).cpu().numpy()
        return fl.votes_numpy(), classes

    @staticmethod
    def write_uv2pt_fixed_cloud(dirname, frame_names, points, K, w, h, wxyzs, translations, depths, point_range=(0.1, 4),
                                radius=0.05, chunk=32):
        """Writes `dirname/fusion/uv2pt/<frame>.npy` (int32 [h*w], -1 = none; `fusion.py:253,297,322,326-327`) for a
        fixed cloud, in frame chunks."""
        out_dir = Path(dirname) / 'fusion' / 'uv2pt'
        out_dir.mkdir(exist_ok=True, parents=True)
        fl = FusedLabeler(points, K, w, h, wxyzs, translations, point_range, radius)
        F = len(frame_names)
        for a in range(0, F, chunk):
            b = min(a + chunk, F)
            uv = fl.uv2pt(depths[a:b], frame_begin=a, frame_end=b).cpu().numpy()
            for k in range(a, b):
                np.save(out_dir / f'{frame_names[k]}.npy', uv[k - a])
        return out_dir
