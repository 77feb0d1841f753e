"""Drop-in for the reference's models/models.py:BaseModel on the sm_100a hot-path kernels.

Call surface kept (reference models/models.py:440-710, built by models/get_model.py:7-8):

    BaseModel(cfg, id).forward(list[input_dict]) ->
        list[{"output": [...], "output_refined": [...], "flow": None, "flow_refined": None}]

with the same input-dict keys (`rgb`, `d_candi`, `src_cam_poses`, `intrinsics`, `unit_ray`, plus
`prev_output` / `dmaps`, `masks`), the same `cfg.var` keys (`sigma_soft_max`, `feature_dim`, `nmode`,
`ndepth`, `bn_avg`) and the same parameter REGISTRATION ORDER and shapes, because the reference
loads checkpoints by position (trainer/base_trainer.py:83-90): base_encoder.{firstconv, layer1-4,
branch1-4, lastconv} -> base_decoder.{conv0, conv0_1, trans_conv0, conv1, conv1_1, trans_conv1,
conv2, conv2_1, conv2_2} -> conv0, conv0_1, conv0_2 -> (feedback) based_3d.{dres0, classify}.  The
feedback net's residual blocks are kept out of the state_dict exactly as the reference does
(a plain Python list, models/models.py:394-399), and modules are created and initialised in the
reference's order, so that under the same torch seed the random-init weights are identical.

What differs is everything between the CNN blocks.  The dense convolutions stay torch.nn (cuDNN:
out of scope); the depth-probability-volume path runs batched on our kernels:

  reference                                                     here
  per-item Python loop + .cpu().numpy() sync (:528-550)         one dpv_sweep_cost_volume launch
  F.log_softmax + torch.exp (:560,:653)                         one dpv_head launch (logp + prob)
  gen_dpv_withmask + exp/sum/div/clamp/log (:666-672)           one dpv_bayes_fuse launch
  per-item warp_feature loop (:614-627)                         one dpv_warp_feature launch
  log_softmax(BV + resi) + exp (:694,:697)                      one dpv_head launch with addend
  decoder's final F.log_softmax (:351)                          one dpv_head launch
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


# ------------------------------------------------------------------------------ building blocks
def _conv_bn(cin, cout, k, stride, pad, dilation, running):
    """Conv2d(no bias) + BatchNorm2d; BN keeps running statistics only when cfg.var.bn_avg."""
    p = dilation if dilation > 1 else pad
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, p, dilation, bias=False),
                         nn.BatchNorm2d(cout, track_running_stats=running))


def _conv_bn3d(cin, cout, running):
    return nn.Sequential(nn.Conv3d(cin, cout, 3, 1, 1, bias=False),
                         nn.BatchNorm3d(cout, track_running_stats=running))


def _conv_lrelu(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1, bias=True), nn.LeakyReLU())


def _deconv_lrelu(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=True), nn.LeakyReLU())


def _init_weights(m):
    """He-normal convolutions, unit BatchNorm, bilinear-kernel transposed convolutions
    (reference models/models.py:224-238,358-374,478-498)."""
    if isinstance(m, (nn.Conv2d, nn.Conv3d)) and not isinstance(m, nn.ConvTranspose2d):
        fan = m.out_channels
        for k in m.kernel_size:
            fan *= k
        m.weight.data.normal_(0, math.sqrt(2.0 / fan))
    elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
        m.weight.data.fill_(1)
        m.bias.data.zero_()
    elif isinstance(m, nn.Linear):
        m.bias.data.zero_()
    elif isinstance(m, nn.ConvTranspose2d):
        n = m.kernel_size[1]
        factor = (n + 1) // 2
        center = factor - 1 if n % 2 == 1 else factor - 0.5
        og = np.ogrid[:n, :n]
        kern = (1 - abs(og[0] - center) / factor) * (1 - abs(og[1] - center) / factor)
        m.weight.data.copy_(torch.from_numpy(kern))


def _init_2d_only(m):
    """The encoder/decoder/model-level initialisers of the reference do not touch 3-D modules."""
    if not isinstance(m, (nn.Conv3d, nn.BatchNorm3d)):
        _init_weights(m)


class _Residual(nn.Module):
    """Two conv-bn layers with an identity / projected skip (no activation after the sum)."""

    def __init__(self, cin, cout, stride, downsample, pad, dilation, running):
        super().__init__()
        self.conv1 = nn.Sequential(_conv_bn(cin, cout, 3, stride, pad, dilation, running), nn.ReLU(inplace=True))
        self.conv2 = _conv_bn(cout, cout, 3, 1, pad, dilation, running)
        self.downsample = downsample

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + (x if self.downsample is None else self.downsample(x))


class BaseEncoder(nn.Module):
    """PSMNet-style feature extractor with four pooled context branches: image [N,3,H,W] ->
    (1/2-res features, 1/4-res raw features, 1/4-res features [N, feature_dim, H/4, W/4])."""

    POOLS = (64, 32, 16, 8)

    def __init__(self, feature_dim=32, bn_running_avg=False, multi_scale=True):
        super().__init__()
        mul = feature_dim / 64.0
        s0, s1, s2, s3 = (int(v * mul) for v in (16, 32, 64, 128))
        self.multi_scale = multi_scale
        self.bn_ravg = bn_running_avg
        self._cin = s1
        r = bn_running_avg
        stem = []
        for i, (a, b, st) in enumerate(((3, s1, 2), (s1, s1, 1), (s1, s1, 1))):
            stem += [_conv_bn(a, b, 3, st, 1, 1, r), nn.ReLU(inplace=True)]
        self.firstconv = nn.Sequential(*stem)
        self.layer1 = self._stack(s1, 3, 1, 1, 1)
        self.layer2 = self._stack(s2, s0, 2, 1, 1)
        self.layer3 = self._stack(s3, 3, 1, 1, 1)
        self.layer4 = self._stack(s3, 3, 1, 1, 2)
        for i, size in enumerate(self.POOLS):
            setattr(self, "branch%d" % (i + 1),
                    nn.Sequential(nn.AvgPool2d((size, size), stride=(size, size)),
                                  _conv_bn(s3, s1, 1, 1, 0, 1, r), nn.ReLU(inplace=True)))
        self.lastconv = nn.Sequential(_conv_bn(s1 * 4 + s2 + s3, s3, 3, 1, 1, 1, r), nn.ReLU(inplace=True),
                                      nn.Conv2d(s3, feature_dim, kernel_size=1, padding=0, stride=1, bias=False))
        self.apply(_init_2d_only)

    def _stack(self, cout, blocks, stride, pad, dilation):
        proj = None
        if stride != 1 or self._cin != cout:
            proj = nn.Sequential(nn.Conv2d(self._cin, cout, kernel_size=1, stride=stride, bias=False),
                                 nn.BatchNorm2d(cout, track_running_stats=self.bn_ravg))
        layers = [_Residual(self._cin, cout, stride, proj, pad, dilation, self.bn_ravg)]
        self._cin = cout
        layers += [_Residual(cout, cout, 1, None, pad, dilation, self.bn_ravg) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        half = self.layer1(self.firstconv(x))
        raw = self.layer2(half)
        skip = self.layer4(self.layer3(raw))
        size = skip.shape[2:]
        ctx = [F.interpolate(getattr(self, "branch%d" % i)(skip), size, mode="bilinear", align_corners=True)
               for i in (4, 3, 2, 1)]
        feat = self.lastconv(torch.cat([raw, skip] + ctx, 1))
        return (half, raw, feat) if self.multi_scale else feat


class BaseDecoder(nn.Module):
    """DPV refinement: bins as channels, two 2x transposed-conv upsamplings with image-feature
    skips.  `forward` returns the PRE-soft-max logits [N, D, 4h, 4w]; the caller normalises them
    with the fused head kernel (the reference applies F.log_softmax here, models/models.py:351)."""

    def __init__(self, C0, C1, C2, D=64, upsample_D=False):
        super().__init__()
        d0 = 2 * D if upsample_D else D
        d1 = 2 * d0 if upsample_D else D
        cin = D + C0
        self.conv0 = _conv_lrelu(cin, cin)
        self.conv0_1 = _conv_lrelu(cin, cin)
        self.trans_conv0 = _deconv_lrelu(cin, d0)
        self.conv1 = _conv_lrelu(d0 + C1, d0 + C1)
        self.conv1_1 = _conv_lrelu(d0 + C1, d0 + C1)
        self.trans_conv1 = _deconv_lrelu(d0 + C1, d1)
        self.conv2 = _conv_lrelu(d1 + C2, d1 + C2)
        self.conv2_1 = _conv_lrelu(d1 + C2, d1)
        self.conv2_2 = nn.Conv2d(d1, d1, kernel_size=3, stride=1, padding=1, bias=True)
        self.apply(_init_2d_only)

    def forward(self, dpv_raw, img_features):
        x = self.conv0_1(self.conv0(torch.cat([dpv_raw, img_features[0]], 1)))
        x = self.conv1_1(self.conv1(torch.cat([self.trans_conv0(x), img_features[1]], 1)))
        x = self.conv2_1(self.conv2(torch.cat([self.trans_conv1(x), img_features[2]], 1)))
        return self.conv2_2(x)


class Base3D(nn.Module):
    """The feedback net: 3-D convolutions over [N, C, D, h, w] -> residual [N, D, h, w]."""

    def __init__(self, input_volume_channels, feature_dim=32, dres_count=4, bn_running_avg=False, id=0):
        super().__init__()
        r = bn_running_avg
        self.id = id
        self.dres0 = nn.Sequential(_conv_bn3d(input_volume_channels, feature_dim, r), nn.ReLU(),
                                   _conv_bn3d(feature_dim, feature_dim, r), nn.ReLU())
        # deliberately NOT registered (a plain list), as in the reference: these blocks are absent
        # from state_dict() and from the optimiser, and keep torch's default initialisation
        self.dres_modules = [nn.Sequential(_conv_bn3d(feature_dim, feature_dim, r), nn.ReLU(),
                                           _conv_bn3d(feature_dim, feature_dim, r))
                             for _ in range(dres_count)]
        self.classify = nn.Sequential(_conv_bn3d(feature_dim, feature_dim, r), nn.ReLU(),
                                      nn.Conv3d(feature_dim, 1, kernel_size=3, padding=1, stride=1, bias=False))
        self.apply(_init_weights)

    def _tc(self, volume):
        """The tcgen05 path (ops.Base3DConvs): CUDA, no autograd, at most 32 channels; DPV_BASE3D_TC=0 keeps cuDNN.
        Re-packed when a parameter, a running statistic or a BatchNorm's mode changed.  NOTE: BatchNorms that
        normalise with batch statistics do not get their running statistics updated on this path (the reference's
        residual blocks never read theirs)."""
        import os
        if not (volume.is_cuda and not torch.is_grad_enabled() and os.environ.get("DPV_BASE3D_TC", "1") != "0"):
            return None
        convs = [self.dres0[0][0], self.dres0[2][0]] + [b[j][0] for b in self.dres_modules for j in (0, 2)] + \
                [self.classify[0][0], self.classify[2]]
        if any(c.weight.shape[0] > 32 or c.weight.shape[1] > 32 for c in convs) or self.classify[2].weight.shape[0] != 1:
            return None
        for i, blk in enumerate(self.dres_modules):
            if next(blk.parameters()).device != volume.device:
                self.dres_modules[i] = blk.to(volume.device)
        bns = [self.dres0[0][1], self.dres0[2][1]] + [b[j][1] for b in self.dres_modules for j in (0, 2)] + [self.classify[0][1]]
        key = tuple((c.weight.data_ptr(), c.weight._version) for c in convs) + \
            tuple((b.weight.data_ptr(), b.weight._version, b.bias._version, b.training, b.track_running_stats,
                   None if b.running_mean is None else (b.running_mean._version, b.running_var._version)) for b in bns)
        hit = getattr(self, "_tc_net", None)
        if hit is None or hit[0] != key:
            hit = (key, ops.Base3DConvs.from_module(self))
            self._tc_net = hit
        return hit[1](volume)

    def forward(self, volume):
        out = self._tc(volume)
        if out is not None:
            return out
        x = self.dres0(volume.contiguous())
        for i, blk in enumerate(self.dres_modules):
            if next(blk.parameters()).device != x.device:
                self.dres_modules[i] = blk = blk.to(x.device)
            # unregistered, hence never reached by .eval(): they stay in training mode (batch
            # statistics), in the reference and here
            x = blk(x) + x
        return self.classify(x).squeeze(1)


# ------------------------------------------------------------------------------------- the model
class BaseModel(nn.Module):
    def __init__(self, cfg, id):
        super().__init__()
        self.cfg = cfg
        v = cfg.var
        self.sigma_soft_max = v.sigma_soft_max
        self.feature_dim = v.feature_dim
        self.nmode = v.nmode
        self.D = v.ndepth
        self.bn_avg = v.bn_avg
        self.id = id
        self.base_encoder = BaseEncoder(feature_dim=self.feature_dim, multi_scale=True, bn_running_avg=self.bn_avg)
        self.base_decoder = BaseDecoder(int(self.feature_dim), int(self.feature_dim / 2), 3, D=self.D)
        self.conv0 = _conv_lrelu(self.D, self.D)
        self.conv0_1 = _conv_lrelu(self.D, self.D)
        self.conv0_2 = nn.Conv2d(self.D, self.D, kernel_size=3, stride=1, padding=1, bias=True)
        if self.nmode == "default_feedback":
            self.based_3d = Base3D(4, dres_count=2, feature_dim=32, bn_running_avg=self.bn_avg, id=self.id)
        self.apply(_init_2d_only)
        self.viz = None

    def set_viz(self, viz):
        self.viz = viz

    def freeze_weights(self, name):
        for p in getattr(self, name).parameters():
            p.requires_grad = False

    def init_weights(self):
        self.apply(_init_2d_only)

    def _refine_logits(self, cost):
        """conv0 -> conv0_1 -> conv0_2 (reference models/models.py:555-557,632-634).  D = 64 on a CUDA device without
        autograd: the tcgen05 implicit-GEMM kernels at fp32 parity (ops.CostRefine; DPV_REFINE_TC=0 keeps cuDNN);
        otherwise the nn.Conv2d modules themselves (same parameters)."""
        import os
        mods = (self.conv0[0], self.conv0_1[0], self.conv0_2)
        if (cost.is_cuda and self.D == 64 and not torch.is_grad_enabled() and os.environ.get("DPV_REFINE_TC", "1") != "0"
                and all(m.weight.shape == (64, 64, 3, 3) for m in mods)):
            key = tuple((t.data_ptr(), t._version) for m in mods for t in (m.weight, m.bias))
            hit = getattr(self, "_refine_tc", None)
            if hit is None or hit[0] != key:       # (re)pack the weights when the parameters changed
                hit = (key, ops.CostRefine([m.weight for m in mods], [m.bias for m in mods],
                                           slope=self.conv0[1].negative_slope))
                self._refine_tc = hit
            return hit[1](cost, want_logits=True, want_logp=False)
        return self.conv0_2(self.conv0_1(self.conv0(cost)))

    # -- encoder + cost volume + 1/4-res DPV (reference forward_encoder / forward_exp) -----------
    def _encode(self, mi, want_raw):
        rgb = mi["rgb"]
        B, V1 = rgb.shape[:2]
        frames = rgb.reshape((B * V1,) + tuple(rgb.shape[2:]))
        half, raw, feat = self.base_encoder(frames)
        rate = int(frames.shape[3] / feat.shape[3])
        feat_all = torch.cat((feat, F.avg_pool2d(frames, rate)), 1)           # [B*V1, C+3, h, w]
        feat_all = feat_all.reshape((B, V1) + tuple(feat_all.shape[1:])).contiguous()
        half = half.reshape((B, V1) + tuple(half.shape[1:]))
        poses = mi["src_cam_poses"].float()
        # one launch for the whole batch: reference view last, the others are the sources
        cost = ops.sweep_cost_volume(feat_all[:, -1], feat_all[:, :-1], poses[:, :-1].contiguous(),
                                     mi["intrinsics"].float(), mi["unit_ray"].float(), mi["d_candi"],
                                     self.sigma_soft_max, dist="L2")
        logits = self._refine_logits(cost)
        last = [feat_all[:, -1, :-3], half[:, -1]]
        first = [feat_all[:, 0, :-3], half[:, 0]]
        warped = None
        if want_raw:
            raw = raw.reshape((B, V1) + tuple(raw.shape[1:])).contiguous()
            warped = ops.warp_feature(raw, poses.contiguous(), mi["intrinsics"].float(), mi["unit_ray"].float(),
                                      mi["d_candi"])
        return logits, cost, last, first, warped

    def forward_encoder(self, model_input):
        logits, cost, last, first, _ = self._encode(model_input, False)
        BV = ops.head(logits, model_input["d_candi"], logp=True)["logp"]
        return BV, cost, last, first

    def forward_exp(self, model_input):
        logits, cost, last, first, warped = self._encode(model_input, True)
        BV = ops.head(logits, model_input["d_candi"], logp=True)["logp"]
        return BV, cost, last, first, warped

    def _refine(self, dpv, feats, d_candi):
        return ops.head(self.base_decoder(dpv, img_features=feats), d_candi, logp=True)["logp"]

    def forward_int(self, mi):
        d = mi["d_candi"]
        none = {"flow": None, "flow_refined": None}
        if self.nmode == "default":
            logits, _, feats, _, _ = self._encode(mi, False)
            h = ops.head(logits, d, logp=True, prob=True)                       # log-softmax and exp in one pass
            feats.append(mi["rgb"][:, -1])
            return dict(output=[h["logp"]], output_refined=[self._refine(h["prob"], feats, d)], **none)
        if self.nmode == "default_upsample":
            logits, _, feats, _, _ = self._encode(mi, False)
            BV = ops.head(logits, d, logp=True)["logp"]
            feats.append(mi["rgb"][:, -1])
            fused, log_fused = ops.bayes_fuse(BV, d, dmaps=mi["dmaps"].float(), masks=mi["masks"].float(), var=0.3)
            return dict(output=[log_fused, BV], output_refined=[self._refine(fused, feats, d)], **none)
        if self.nmode == "default_feedback":
            logits, _, feats, _, warped = self._encode(mi, True)
            BV = ops.head(logits, d, logp=True)["logp"]
            feats.append(mi["rgb"][:, -1])
            if mi["prev_output"] is None:
                prev = torch.zeros_like(BV).unsqueeze(1) + 1.0 / float(self.D)
            else:
                prev = mi["prev_output"].unsqueeze(1)
            resi = self.based_3d(torch.cat([BV.unsqueeze(1), prev, warped], 1))
            h = ops.head(BV, d, addend=resi.contiguous(), logp=True, prob=True)  # log_softmax(BV + resi), exp
            return dict(output=[BV, h["logp"]], output_refined=[self._refine(h["prob"], feats, d)], **none)
        raise Exception("Nmode wrong")

    def forward(self, inputs):
        return [self.forward_int(x) for x in inputs]
