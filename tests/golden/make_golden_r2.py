"""Round-2 goldens, again outputs of the UNMODIFIED reference on seeded inputs (build container only):

    python tests/golden/make_golden_r2.py

sweep_wide.npz  : est_swp_volume_v4 (warping/homography.py:98-135) on images wider than 192 px -- the
                  shapes where our TMA sweep kernel takes its exact-coordinate variant -- including the
                  north-star's literal 256x384 / C=67 / D=64 shape; only the sub-sampled part named by
                  cases.sweep_wide_case(...)["sub"] is stored.
ufield_cfgx.npz : gen_ufield (utils/img_utils.py:268-358) called with cfgx, i.e. the quash_limit branch
                  (:325-332) the ROS caller uses (ros/ros_net.py:279).
conv_refine.npz : the reference BaseModel's own conv0 / conv0_1 / conv0_2 modules + F.log_softmax
                  (models/models.py:456-460,555-560) on a seeded cost volume [2,64,12,20], fp32 CPU, with the
                  module weights stored next to the result.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402
import make_golden  # noqa: E402


def main():
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    torch.set_num_threads(8)
    homography, img_utils, _ = make_golden.import_reference()
    T = torch.from_numpy
    out = {}
    for name in cases.SWEEP_WIDE_CASES:
        c = cases.sweep_wide_case(name)
        cam = {"intrinsic_M_cuda": T(c["K"]), "intrinsic_M": c["K"], "unit_ray_array_2D": T(c["rays"])}
        cv = homography.est_swp_volume_v4(T(c["ref"]), T(c["src"]), c["d_candi"], T(c["R"]), T(c["t"]), cam,
                                          c["sigma"], feat_dist="L2")
        out[name + "_L2"] = np.ascontiguousarray(cases.sub_view(cv.numpy(), c["sub"]))
        print(name, tuple(cv.shape), "->", out[name + "_L2"].shape, float(cv.min()), float(cv.max()))
    np.savez(os.path.join(HERE, "sweep_wide.npz"), **out)

    out = {}
    for name in cases.UFIELD_CFGX_CASES:
        c = cases.ufield_cfgx_case(name)
        ls = torch.nn.functional.log_softmax(T(c["logits"]), dim=1)
        dpv = ls if c["log"] else torch.exp(ls)
        mask = None if c["mask"] is None else T(c["mask"])
        uf, dz = img_utils.gen_ufield(dpv, c["d_candi"], T(c["intr_up"]), BV_log=c["log"], mask=mask,
                                      cfgx=c["cfgx"])
        out[name + "_uf"] = uf.numpy()
        out[name + "_depthzero"] = dz.numpy()
        print(name, "finite UF columns", int(np.isfinite(uf.numpy()).all(1).sum()), "of", uf.shape[-1])
    np.savez(os.path.join(HERE, "ufield_cfgx.npz"), **out)
    # the model's own modules
    sys.path.insert(0, os.path.join(HERE))
    import model_cases as MC
    from oracle import reference_loader
    ref = reference_loader.load()
    torch.manual_seed(0)
    model = ref.models.BaseModel(MC.cfg("default_stereo"), 0)
    cost = torch.from_numpy((4.0 * np.random.RandomState(77).standard_normal((2, 64, 12, 20)) + 10.0).astype(np.float32))
    with torch.no_grad():
        logits = model.conv0_2(model.conv0_1(model.conv0(cost)))
        bv = torch.nn.functional.log_softmax(logits, dim=1)
    mods = (model.conv0[0], model.conv0_1[0], model.conv0_2)
    out = {"cost": cost.numpy(), "logits": logits.numpy(), "bv": bv.numpy(),
           "slope": np.float32(model.conv0[1].negative_slope)}
    for i, m in enumerate(mods):
        out["w%d" % i] = m.weight.detach().numpy()
        out["b%d" % i] = m.bias.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "conv_refine.npz"), **out)
    print("conv_refine logits range", float(logits.min()), float(logits.max()))
    for f in ("sweep_wide.npz", "ufield_cfgx.npz", "conv_refine.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
