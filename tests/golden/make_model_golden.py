"""Generate tests/golden/model.npz: outputs of the UNMODIFIED reference BaseModel on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_model_golden.py

Random-init weights under torch.manual_seed(0) (our BaseModel creates and initialises its
modules in the same order, so the same seed gives the same weights: tests/test_model_mirror.py
checks that against the key / shape / checksum list stored here), seeded synthetic inputs from
model_cases.py, BatchNorm in batch-statistics mode (model_cases.batch_stat_norm explains why).  Stored per mode and frame: the 1/4-res log-DPV at every 2nd pixel,
the refined log-DPV at every 8th pixel, and E[d] of both.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402
import model_cases as MC  # noqa: E402


def main():
    warnings.filterwarnings("ignore")
    make_golden.import_reference()
    import models.models as RM
    import utils.img_utils as IU
    torch.set_num_threads(8)
    out = {}
    for name, (nmode, bn_avg, kind, frames) in MC.MODES.items():
        torch.manual_seed(0)
        model = MC.batch_stat_norm(RM.BaseModel(MC.cfg(name), 0))
        sd = model.state_dict()
        out[name + "_keys"] = np.array(list(sd.keys()))
        out[name + "_shapes"] = np.array([str(tuple(v.shape)) for v in sd.values()])
        out[name + "_sums"] = np.array([float(v.double().sum()) for v in sd.values()])
        prev = None
        for f in range(frames):
            mi = MC.frame_inputs(name, f)
            t = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) and k != "d_candi" else v) for k, v in mi.items()}
            t["prev_output"] = prev
            with torch.no_grad():
                res = model([t])[0]
            bv, refined = res["output"][-1], res["output_refined"][-1]
            if nmode == "default_upsample":
                bv = res["output"][0]
            out["%s_f%d_bv" % (name, f)] = bv[:, :, ::2, ::2].numpy()
            out["%s_f%d_refined" % (name, f)] = refined[:, :, ::8, ::8].numpy()
            out["%s_f%d_depth_q" % (name, f)] = IU.dpv_to_depthmap(bv, MC.D_CANDI, BV_log=True).numpy()
            out["%s_f%d_depth" % (name, f)] = IU.dpv_to_depthmap(refined, MC.D_CANDI, BV_log=True).numpy()
            # feedback hand-off exactly as trainer/default_trainer.py:221-222
            prev = torch.nn.functional.interpolate(refined, scale_factor=0.25, mode="nearest")
            if f + 1 < frames:
                # the next frame's prev_output: stored so that the mirror can be fed the reference's
                # own hand-off (frame-to-frame error does not compound in the comparison)
                out["%s_f%d_prev" % (name, f + 1)] = prev.numpy()
            print(name, f, "bv", tuple(bv.shape), float(bv.min()), "refined", tuple(refined.shape), float(refined.min()),
                  "depth range", float(out["%s_f%d_depth" % (name, f)].min()), float(out["%s_f%d_depth" % (name, f)].max()))
    np.savez_compressed(os.path.join(HERE, "model.npz"), **out)
    print("wrote model.npz", os.path.getsize(os.path.join(HERE, "model.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
