"""CPU-oracle implementation of the backend interface of ygz_slam_b200.vo (TEST INFRASTRUCTURE ONLY, like the rest of
oracle/): the same tracking loop can run on the oracle and on the GPU, which turns trajectory agreement into an
end-to-end parity statement; bench.py times it as the CPU baseline of BASELINE config C5."""
import numpy as np

from ygz_slam_b200 import se3


class OracleBackend:
    def __init__(self, oracle, n_levels=3):
        self.o = oracle
        self.L = n_levels
        self.pyr = {}

    def upload(self, slots, images):
        for s, img in zip(slots, images):
            self.pyr[int(s)] = self.o.build_pyramid(np.ascontiguousarray(img), self.L)

    def detect(self, slots):
        return [self.o.detect(self.pyr[int(s)], n_levels=self.L) for s in slots]

    def sparse_alignment(self, ref_slots, cur_slots, px, depth, T_ref):
        Ts, oks = [], []
        for r, c, p, d, T in zip(ref_slots, cur_slots, px, depth, T_ref):
            ok, Tc = self.o.matcher_sparse_alignment(self.pyr[int(r)], self.pyr[int(c)], 640, 480, self.L, p, d, np.ones(len(d), np.uint8), T, T)
            Ts.append(Tc)
            oks.append(ok)
        return Ts, oks

    def project(self, jobs):
        out_px, out_ok = [], []
        I = np.eye(4)[:3]
        for job in jobs:
            px = np.zeros((len(job["kf"]), 2))
            ok = np.zeros(len(job["kf"]), bool)
            for k, kf in enumerate(job["kfs"]):
                sel = np.nonzero(job["kf"] == k)[0]
                if not len(sel):
                    continue
                n = job["n"][sel]
                p, _, o = self.o.find_direct_projection(self.pyr[int(kf.slot)], self.pyr[int(job["cur_slot"])], 640, 480, self.L, I,
                                                        se3.mul(job["T_cur"], se3.inv(kf.T_cw)), kf.px[n], kf.depth[n], kf.level[n],
                                                        job["init"][sel])
                px[sel] = p
                ok[sel] = o
            out_px.append(px)
            out_ok.append(ok)
        return out_px, out_ok

    def pose_only(self, pts_w, obs, T):
        Ts, inls, cnts = [], [], []
        for p, o, t in zip(pts_w, obs, T):
            if len(p) == 0:
                Ts.append(t); inls.append(np.zeros(0, bool)); cnts.append(0)
                continue
            Tn, inl, _, cnt = self.o.pose_only(p, o, t)
            Ts.append(Tn); inls.append(inl); cnts.append(cnt)
        return Ts, inls, np.array(cnts)

    def local_ba(self, problems):
        out = []
        for poses, fixed, pts, kf_idx, pt_idx, px in problems:
            g2o = np.concatenate([poses[:, 3:], poses[:, :3]], 1)
            P, X, _, _ = self.o.local_ba(g2o, fixed, pts, kf_idx, pt_idx, px)
            out.append((np.concatenate([P[:, 3:], P[:, :3]], 1), X))
        return out
