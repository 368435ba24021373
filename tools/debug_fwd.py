"""Debug aid: which RoIs / bins of the column-table forward kernel differ from the generic kernel."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cpm_r_cnn_b200 import _lib, synthetic
from cpm_r_cnn_b200.roi_align import pooler_forward
import test_gpu_parity as T
SCALES = T.SCALES
for P, C in ((7, 128), (14, 64)):
    for sr, aligned in ((2, False), (1, False), (2, True)):
        gen = torch.Generator().manual_seed(100 + P + C + sr)
        feats = synthetic.pyramid(gen, 2, C, 200, 336)
        rois = T._size_sweep_rois(P * 10 + sr, 2)
        if aligned:
            rois = rois[(rois[:, 3] >= rois[:, 1]) & (rois[:, 4] >= rois[:, 2])]
        for l in (0, 1, 3):
            f = feats[l].cuda().contiguous(memory_format=torch.channels_last)
            a = pooler_forward([f], [SCALES[l]], rois.cuda(), (P, P), sr, aligned, 0, None, impl=_lib.FWD_COLS).cpu().numpy()
            b = pooler_forward([f], [SCALES[l]], rois.cuda(), (P, P), sr, aligned, 0, None, impl=_lib.FWD_GENERIC).cpu().numpy()
            err = np.abs(a - b).max(axis=1)          # (K, P, P)
            tol = 1e-5 * (np.abs(b).max(axis=1) + np.sqrt((b ** 2).mean()))
            bad = np.argwhere(err > tol)
            print("P%d C%d sr%d al%d lvl%d: %d bad bins" % (P, C, sr, aligned, l, len(bad)))
            for k in sorted(set(bad[:, 0].tolist()))[:6]:
                bb = bad[bad[:, 0] == k]
                print("   roi", k, (rois[k, 1:] * SCALES[l]).tolist(), "bins", bb[:, 1:].tolist()[:10], "err", float(err[k].max()))
