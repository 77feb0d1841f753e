"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded inputs.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference is imported in place with the three shims SURVEY.md section 8c
lists (a stub for the un-buildable C++ metric lib it imports at module scope, a
Tensor.repeat wrapper for a call that newer torch rejects, and a no-op
Module.cuda on a CUDA-less host).  Nothing from the reference is copied; only
its outputs are stored.  Inputs come from tests/golden/cases.py.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402

REF = os.environ.get("DPV_REFERENCE", "/root/reference")


def import_reference():
    """(homography, img_utils, correlation_native) of the unmodified reference; the shims live in
    oracle/reference_loader.py (shared with the GPU-side second-oracle tests)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import reference_loader
    os.environ.setdefault("DPV_REFERENCE", REF)
    ref = reference_loader.load()
    return ref.homography, ref.img_utils, ref.correlation_native


class _Cfg:
    """gen_ufield only reads cfg.data.dataset_path (utils/img_utils.py:277)."""
    class data:
        dataset_path = "./kitti/"


def main():
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    torch.set_num_threads(8)
    homography, img_utils, correlation_native = import_reference()
    T = torch.from_numpy

    out = {}
    for name in cases.SWEEP_CASES:
        c = cases.sweep_case(name)
        cam = {"intrinsic_M_cuda": T(c["K"]), "intrinsic_M": c["K"],
               "unit_ray_array_2D": T(c["rays"])}
        for dist in ("L2", "L1"):
            if dist == "L1" and name.endswith("ref_shape"):
                continue
            cv = homography.est_swp_volume_v4(T(c["ref"]), T(c["src"]), c["d_candi"],
                                              T(c["R"]), T(c["t"]), cam, c["sigma"],
                                              feat_dist=dist)
            out["%s_%s" % (name, dist)] = cv.numpy()
    np.savez(os.path.join(HERE, "sweep.npz"), **out)

    out = {}
    for name in cases.WARP_FEATURE_CASES:
        c = cases.warp_feature_case(name)
        cam = {"intrinsic_M_cuda": T(c["K"]), "intrinsic_M": c["K"],
               "unit_ray_array_2D": T(c["rays"])}
        wf = homography.warp_feature(T(c["feat"]), c["d_candi"], T(c["R"]), T(c["t"]), cam).numpy()
        out[name] = wf if name == "small" else wf[:, :, ::4]
    np.savez(os.path.join(HERE, "warp_feature.npz"), **out)

    out = {}
    for name in cases.SOFTMAX_CASES:
        c = cases.softmax_case(name)
        x = T(c["x"])
        ls = torch.nn.functional.log_softmax(x, dim=1)            # models/models.py:351,560
        out[name + "_logdpv"] = ls.numpy()
        out[name + "_depth"] = np.stack(
            [img_utils.dpv_to_depthmap(ls[b:b + 1], c["d_candi"], BV_log=True)[0].numpy()
             for b in range(x.shape[0])])
        var = []
        for b in range(x.shape[0]):                               # default_trainer.py:333-336
            z = torch.exp(ls[b])
            dd = torch.tensor(c["d_candi"]).unsqueeze(1).unsqueeze(1)
            mean = torch.sum(dd * z, dim=0)
            var.append(torch.sum(((dd - mean) ** 2) * z, dim=0).numpy())
        out[name + "_var"] = np.stack(var)
        out[name + "_argmax"] = torch.argmax(ls, dim=1).numpy()
        if x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0:           # default_trainer.py:221-222
            out[name + "_quarter"] = torch.nn.functional.interpolate(
                ls, scale_factor=0.25, mode="nearest").numpy()
    np.savez(os.path.join(HERE, "softmax.npz"), **out)

    out = {}
    for name in cases.FUSE_CASES:
        c = cases.fuse_case(name)
        bv = torch.nn.functional.log_softmax(T(c["bv_logits"]), dim=1)
        prior = img_utils.gen_dpv_withmask(T(c["dmaps"]), T(c["masks"]), c["d_candi"], 0.3)
        fused = torch.exp(bv + torch.log(prior))                  # models/models.py:669-672
        fused = fused / torch.sum(fused, dim=1).unsqueeze(1)
        fused = torch.clamp(fused, img_utils.epsilon, 1.)
        logf = torch.log(fused)
        upd = torch.nn.functional.log_softmax(bv + T(c["resi"]), dim=1)   # models.py:694
        if name == "small":
            out[name + "_prior"] = prior.numpy()
            out[name + "_fused"] = fused.numpy()
            out[name + "_feedback"] = upd.numpy()
        out[name + "_logfused"] = logf.numpy()
    np.savez(os.path.join(HERE, "fuse.npz"), **out)

    out = {}
    for name in cases.UFIELD_CASES:
        c = cases.ufield_case(name)
        ls = torch.nn.functional.log_softmax(T(c["logits"]), dim=1)
        dpv = ls if c["log"] else torch.exp(ls)
        mask = None if c["mask"] is None else T(c["mask"])
        uf, dz = img_utils.gen_ufield(dpv, c["d_candi"], T(c["intr_up"]), BV_log=c["log"],
                                      mask=mask, cfg=_Cfg)
        out[name + "_uf"] = uf.numpy()
        out[name + "_depthzero"] = dz.numpy()
    np.savez(os.path.join(HERE, "ufield.npz"), **out)

    out = {}
    corr = correlation_native.Correlation(max_displacement=4)
    for name in cases.CORR_CASES:
        c = cases.corr_case(name)
        out[name] = corr(T(c["x1"]), T(c["x2"])).numpy()
    np.savez(os.path.join(HERE, "correlation.npz"), **out)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
